#!/usr/bin/env python
"""Generate the committed golden fixtures.  Runs ONLY in the authoring container (needs /root/reference).

    python tests/golden/make_golden.py

Writes, next to this file:

* ``weights_2ef43f29.npz``   -- the default shipped checkpoint's ``state_dict`` (fp32, ``model.`` prefix
                                stripped as ``visnet.py:84-86`` does), so GPU-box tests use the real weights
* ``fragments_<prot>.npz``   -- packed FragmentData + protein force map for chig / trpcage / ww / abd
                                (``ai2bmd_b200.pdbfrag`` applied to ``/root/reference/examples/<prot>.pdb``)
* ``fragment_tables.json``   -- the reference's per-residue fragment compositions (``src/utils/reference.py:36-64``,
                                ``fragment_atomic_numbers``) and the residue sequence of each example protein, so
                                the fragmentation can be checked against the reference's own tables on any box
* ``reference_partitions.json`` -- device blocks computed by the reference's OWN work-partition function
                                (``src/Calculators/device_strategy.py:83-127``, extracted with ``ast`` because the module
                                imports ase) for the four proteins at 2/3/4/8 devices
* ``reference_host_logic.npz`` -- outputs of the reference's OWN host classes on the Chignolin fixture: ``FragmentData``
                                slicing / ``scalar_split`` / ``vector_split`` (``src/AIMD/fragment.py:7-47``, imported with a
                                stub for ``ase``) and ``DipeptideBondedCombiner`` (``src/Calculators/combiner.py:11-41``)
                                on seeded random per-fragment energies / forces
* ``reference_nonbonded.npz`` -- energy / forces of the reference's OWN ``MMNonBondedCalculator.__call__`` and pair list
                                ``Protein.initial_mm_adjmatrix`` (``src/Calculators/nonbonded.py:24-63``,
                                ``src/AIMD/protein.py:133-151``; class / method bodies extracted with ``ast`` because
                                the modules import ase / openmm) on Chignolin with synthetic amber-like parameters.
                                The six ``ase.units`` constants are supplied from oracle/nonbonded_ref.py (recalled
                                CODATA-2014 values): the formula, pair list and scatter are pinned, the constants not.
* ``reference_caph.npz``      -- dipeptide coordinates (cap hydrogens placed) from the reference's OWN
                                ``DistanceFragment.get_dipeptide_positions`` body (``src/Fragmentation/distancefrag.py:34-54``,
                                extracted with ``ast``) fed with this repo's recipe indices for Chignolin
* ``reference_caph_lbfgs.npz`` -- the per-step hydrogen refinement on the GLY-centred dipeptide of Chignolin (atoms ordered
                                like ``GG.prmtop`` by name): tables from the reference's OWN ``CTable.from_prmtop`` and filters
                                (``hydrogen/ctable.py``), the five energy terms and the relaxed coordinates from its OWN
                                ``HydrogenOptimizer`` (``hydrogen/energies.py``, decorators stripped with ``ast``; real
                                ``torch.optim.LBFGS``).  Checker of oracle/caph_ref.py.
* ``reference_caph_batch.npz`` -- the same refinement run JOINTLY over the 10 dipeptides of Chignolin (one LBFGS over all 35 added
                                hydrogens, as ``optimize_hydrogen`` does on a ``ProteinDataBatch``): the reference's own
                                CTable filters per dipeptide, concatenated with the offsets of ``ProteinData.__inc__``
                                (``hydrogen/topology.py:108-126``), fed to its own energy functions / optimiser
* ``reference_fragment_index.json`` -- protein-atom membership of every dipeptide / ACE-NME from the reference's OWN
                                ``DipeptideFragment.get_fragments_index`` (``src/Fragmentation/basefrag.py:44-167``,
                                extracted with ``ast``) on the four example proteins
* ``reference_outputs.npz``  -- energies/forces produced by the reference's OWN model source
                                (``/root/reference/src/ViSNet/model``: ``load_model`` -> ``ViSNet.forward``)
                                executed here with the third-party stand-ins of ``oracle/ref_shims.py``,
                                fp32 CPU, plus the fp64 oracle anchor for the same inputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference"
CKPT = f"{REF}/src/ViSNet/checkpoints/visnet-uni-2ef43f29ec78fa5fef0b3de832bfada9.ckpt"

from oracle import ref_shims, visnet_ref as O          # noqa: E402
from ai2bmd_b200.pdbfrag import read_pdb, fragment_protein, single_graph   # noqa: E402


def load_reference_model():
    ref_shims.install()
    sys.path.insert(0, f"{REF}/src")
    torch.jit.script = lambda m, *a, **k: m       # load_model scripts the module (visnet.py:92); eager is equivalent
    from ViSNet.model.visnet import load_model    # the reference's own loader
    return load_model(CKPT).eval()


def ref_eval(model, fd):
    data = dict(z=torch.from_numpy(np.asarray(fd.z, dtype=np.int64)),
                pos=torch.from_numpy(np.asarray(fd.pos, dtype=np.float32)).clone(),
                batch=torch.from_numpy(np.asarray(fd.batch, dtype=np.int64)))
    with torch.set_grad_enabled(True):
        e, f = model(data)
    return e.detach().reshape(-1, 1).numpy(), f.detach().reshape(-1, 3).numpy()


def dense_fragment(seed=0, n=44, box=3.2):
    """Synthetic over-dense fragment: > 32 atoms within 5 A of most atoms (exercises the 32-cap)."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, box, size=(n, 3)).astype(np.float32)
    z = rng.choice([1, 6, 7, 8, 16], size=n).astype(np.int64)
    return single_graph(z, pos)


def write_reference_partitions(frs):
    """Run the reference's own ``DeviceStrategy._set_combined_work_partitions`` (one huge chunk per device, so only
    the per-device block boundaries remain) on the fragment start/end arrays of the example proteins."""
    import ast
    import bisect
    import json
    tree = ast.parse(open(f"{REF}/src/Calculators/device_strategy.py").read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "_set_combined_work_partitions"][0]
    fn.decorator_list = []
    ns = {"bisect": bisect}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_device_strategy", "exec"), ns)

    class Cls:
        _chunk_size = 10 ** 9
        _work_partitions = None

    out = {}
    for name, (fd, _) in frs.items():
        start, end = [int(x) for x in fd.start], [int(x) for x in fd.end]
        for n in (2, 3, 4, 8):
            ns["_set_combined_work_partitions"](Cls, list(range(n)), start, end)
            blocks = []
            for dev in range(n):
                mine = [(a, b) for d, a, b in Cls._work_partitions if d == dev]
                blocks.append([min(a for a, _ in mine), max(b for _, b in mine)] if mine else None)
            out[f"{name}:{n}"] = blocks
    with open(os.path.join(HERE, "reference_partitions.json"), "w") as fh:
        json.dump(out, fh, sort_keys=True)


def write_reference_host_logic(fd, pm):
    """Golden outputs of the reference's FragmentData and DipeptideBondedCombiner on the Chignolin fixture."""
    import importlib.util
    import types
    ref_shims.install()
    if "ase" not in sys.modules:                                   # fragment.py only needs the name
        ase = types.ModuleType("ase")
        ase.Atoms = type("Atoms", (), {"__init__": lambda self, **k: None})
        sys.modules["ase"] = ase

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    rf = load("ref_fragment", f"{REF}/src/AIMD/fragment.py")
    rc = load("ref_combiner", f"{REF}/src/Calculators/combiner.py")
    out = {}
    ref_fd = rf.FragmentData(fd.z, fd.pos, fd.start, fd.end, fd.batch)
    for tag, idx in (("s3_7", slice(3, 7)), ("s0_1", slice(0, 1)), ("i5", 5), ("s10_19", slice(10, 19))):
        sub = ref_fd[idx]
        out[f"{tag}_z"], out[f"{tag}_pos"], out[f"{tag}_start"] = sub.z, sub.pos, sub.start
        out[f"{tag}_end"], out[f"{tag}_batch"] = sub.end, sub.batch
    out["scalar_dip"], out["scalar_an"] = ref_fd.scalar_split()
    out["vector_dip"], out["vector_an"] = ref_fd.vector_split()
    sub = ref_fd[slice(4, 11)]                                     # a chunk that starts on a dipeptide
    out["sub_scalar_dip"], out["sub_scalar_an"] = sub.scalar_split()
    out["sub_vector_dip"], out["sub_vector_an"] = sub.vector_split()
    # combiner: seeded per-fragment energies / per-atom forces, split as DLBondedCalculator.calculate does (bonded.py:91-93)
    rng = np.random.default_rng(0)
    e = rng.standard_normal((len(fd), 1)).astype(np.float32) * 100
    f = rng.standard_normal((len(fd.z), 3)).astype(np.float32)
    sd_, sa_ = out["scalar_dip"], out["scalar_an"]
    vd_, va_ = out["vector_dip"], out["vector_an"]
    order = np.concatenate([np.flatnonzero(vd_), np.flatnonzero(va_)])
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order))
    select, origin = inv[pm.src_atom], pm.dst_atom                 # distancefrag.py:335-353 in this repo's indexing
    out["comb_e_in"], out["comb_f_in"] = e, f
    out["comb_select"], out["comb_origin"] = select, origin
    out["comb_energy"] = rc.DipeptideBondedCombiner.energy_combine(torch.from_numpy(e[sd_]), torch.from_numpy(e[sa_]))
    out["comb_forces"] = rc.DipeptideBondedCombiner.forces_combine(
        pm.n_protein, torch.from_numpy(f[vd_]), torch.from_numpy(f[va_]),
        torch.from_numpy(select.astype(np.int64)), torch.from_numpy(origin.astype(np.int64)))
    np.savez_compressed(os.path.join(HERE, "reference_host_logic.npz"), **out)


def write_reference_nonbonded(fd, pm, prot_pos, prot_z, recipe):
    import ast
    from itertools import product
    from ai2bmd_b200.nonbonded import dipeptide_atom_sets, synthetic_parameters
    from oracle import nonbonded_ref as NB

    def extract(path, name, kind):
        tree = ast.parse(open(path).read())
        node = [n for n in ast.walk(tree) if isinstance(n, kind) and n.name == name][0]
        return ast.Module(body=[node], type_ignores=[])

    ns = {"np": np, "torch": torch, "scatter_add": lambda src, index, dim=0, dim_size=None: ref_shims._scatter(src, index, dim, None, dim_size),
          "C": NB.C, "_eps0": NB._eps0, "kJ": NB.kJ, "mol": NB.mol, "nm": NB.nm, "pi": NB.pi, "Protein": object, "product": product}
    exec(compile(extract(f"{REF}/src/Calculators/nonbonded.py", "MMNonBondedCalculator", ast.ClassDef), "ref_nonbonded", "exec"), ns)
    exec(compile(extract(f"{REF}/src/AIMD/protein.py", "initial_mm_adjmatrix", ast.FunctionDef), "ref_protein", "exec"), ns)

    q, sg, ep = synthetic_parameters(prot_z, seed=1)
    groups = dipeptide_atom_sets(fd, recipe, pm)

    class FakeProtein:                       # the attributes the two reference bodies touch
        def __init__(self):
            self.positions = np.asarray(prot_pos, dtype=np.float64)
            self.sigmas, self.epsilons, self.charges = sg, ep, q
            self.exclude_pair = NB.exclude_pairs_from_groups(groups)      # distancefrag.py:355-361
        initial_mm_adjmatrix = ns["initial_mm_adjmatrix"]

        def get_positions(self):
            return self.positions

        def __len__(self):
            return len(self.positions)

    prot = FakeProtein()
    calc = ns["MMNonBondedCalculator"](device="cpu")
    calc.set_parameters(prot)
    energy, force = calc(prot)
    np.savez_compressed(os.path.join(HERE, "reference_nonbonded.npz"), charges=q, sigmas=sg, epsilons=ep,
                        positions=prot.positions, src=calc.src.numpy(), dst=calc.dst.numpy(), energy=np.float64(energy),
                        forces=force, group_ptr=np.cumsum([0] + [len(g) for g in groups]), group_atoms=np.concatenate(groups))
    print(f"nonbonded reference: {calc.src.numel()} ordered pairs, E = {energy:.6f} eV, max|F| = {np.abs(force).max():.4f} eV/A")


def write_reference_caph(fd, pm, prot_pos, prot_z, recipe):
    import ast
    tree = ast.parse(open(f"{REF}/src/Fragmentation/distancefrag.py").read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "get_dipeptide_positions"][0]
    fn.decorator_list = []
    ns = {"torch": torch, "Protein": object}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_distancefrag", "exec"), ns)
    dip_atoms = np.concatenate([np.arange(fd.start[g], fd.end[g]) for g in range(len(fd)) if pm.frag_sign[g] > 0])
    real, acc, rem, blen = recipe.real[dip_atoms], recipe.acc[dip_atoms], recipe.rem[dip_atoms], recipe.blen[dip_atoms]
    is_real = real >= 0
    out_idx = np.arange(len(dip_atoms))

    class FakeProtein:
        arrays = {"positions": np.asarray(prot_pos, dtype=np.float64)}
        all_dipeptide_index = torch.from_numpy(real[is_real].astype(np.int64))
        all_hydrogen_index = torch.from_numpy(rem[~is_real].astype(np.int64))
        all_acceptor_index = torch.from_numpy(acc[~is_real].astype(np.int64))
        all_hydrogen_radii = torch.from_numpy(blen[~is_real].astype(np.float32))[:, None]
        scatter_original_index = torch.from_numpy(out_idx[is_real].astype(np.int64))[:, None].expand(-1, 3)
        scatter_hydrogen_index = torch.from_numpy(out_idx[~is_real].astype(np.int64))[:, None].expand(-1, 3)
        dipeptides_len = len(dip_atoms)

    pos = ns["get_dipeptide_positions"](FakeProtein, "cpu").numpy()
    np.savez_compressed(os.path.join(HERE, "reference_caph.npz"), dip_atoms=dip_atoms, positions=pos)
    print(f"cap-H reference: {len(dip_atoms)} dipeptide atoms, {int((~is_real).sum())} added hydrogens")


def write_reference_caph_lbfgs():
    import ast
    import glob
    import importlib.util
    import types
    from oracle import caph_ref as CR
    spec = importlib.util.spec_from_file_location("ref_ctable", f"{REF}/src/Fragmentation/hydrogen/ctable.py")
    ctmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ctmod)
    # the restated parser equals the reference's on every shipped table
    fields = ["charge", "atomic_number", "atom_type_idx", "number_excluded_atoms", "nonbonded_parm_index",
              "bond_force_constant", "bond_equil_value", "angle_force_constant", "angle_equil_value",
              "dihedral_force_constant", "dihedral_periodicity", "dihedral_phase", "lennard_jones_acoef",
              "lennard_jones_bcoef", "bonds_inc_hydrogen", "angles_inc_hydrogen", "dihedrals_inc_hydrogen", "excluded_atoms_list"]
    files = sorted(glob.glob(f"{REF}/src/Fragmentation/prmtop/*.prmtop"))
    for path in files:
        ref_t, mine = ctmod.CTable.from_prmtop(path), CR.parse_prmtop(open(path).read())
        for k in ("natom", "ntypes", "numbnd", "numang", "nptra"):
            assert getattr(ref_t, k) == mine[k], (path, k)
        for k in fields:
            assert np.array_equal(getattr(ref_t, k).numpy(), mine[k]), (path, k)
    print(f"prmtop parser: {len(files)} tables identical to the reference's CTable")

    # the reference's energy functions and optimiser (jit decorators removed; ProteinData is only a type hint)
    tree = ast.parse(open(f"{REF}/src/Fragmentation/hydrogen/energies.py").read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))]
    for n in body:
        n.decorator_list = []
    ns = {"torch": torch, "F": torch.nn.functional, "ProteinData": object,
          "scatter_add": lambda src, index, dim=0: ref_shims._scatter(src, index, dim)}
    exec(compile(ast.Module(body=body, type_ignores=[]), "ref_energies", "exec"), ns)

    # GLY-centred dipeptide of Chignolin (residues 7-8-9 of the capped chain) in GG.prmtop atom order, by name
    prot = read_pdb(f"{REF}/examples/chig.pdb")
    P = prot.positions.astype(np.float32)

    def at(res, name):
        return [i for i in range(len(prot)) if prot.resnums[i] == res and prot.names[i] == name][0]

    def cap(acc, rem, blen=np.float32(0.76 + 0.31)):
        d = P[rem] - P[acc]
        return P[acc] + d / np.linalg.norm(d) * blen

    c = 8
    assert prot.resnames[at(c, "CA")] == "GLY"
    lead, trail = c - 1, c + 1
    pos = np.stack([P[at(lead, "HA")], P[at(lead, "CA")], cap(at(lead, "CA"), at(lead, "N")), cap(at(lead, "CA"), at(lead, "CB")),
                    P[at(lead, "C")], P[at(lead, "O")],
                    P[at(c, "N")], P[at(c, "H")], P[at(c, "CA")], P[at(c, "HA2")], P[at(c, "HA3")], P[at(c, "C")], P[at(c, "O")],
                    P[at(trail, "N")], P[at(trail, "H")], P[at(trail, "CA")], P[at(trail, "HA")],
                    cap(at(trail, "CA"), at(trail, "C")), cap(at(trail, "CA"), at(trail, "CB"))]).astype(np.float32)
    atom_idx = np.array([2, 3, 17, 18])
    ct = ctmod.CTable.from_prmtop(f"{REF}/src/Fragmentation/prmtop/GG.prmtop")
    assert ct.atomic_number.tolist() == [1, 6, 1, 1, 6, 8, 7, 1, 6, 1, 1, 6, 8, 7, 1, 6, 1, 1, 1]
    aidx = torch.from_numpy(atom_idx)
    b = types.SimpleNamespace()
    b.pos = torch.from_numpy(pos.copy())
    b.atom_idx = aidx
    b.other_idx = torch.tensor([i for i in range(ct.natom) if i not in atom_idx.tolist()])
    for k in ("charge", "bond_force_constant", "bond_equil_value", "angle_force_constant", "angle_equil_value",
              "dihedral_force_constant", "dihedral_periodicity", "dihedral_phase", "lennard_jones_acoef", "lennard_jones_bcoef"):
        setattr(b, k, getattr(ct, k))
    b.bonds_atom_idx_src, b.bonds_atom_idx_dst, b.bond_idx = ct.filter_bonds(aidx)
    b.angles_atom_idx_i, b.angles_atom_idx_j, b.angles_atom_idx_k, b.angle_idx = ct.filter_angles(aidx)
    (b.dihedrals_atom_idx_i, b.dihedrals_atom_idx_j, b.dihedrals_atom_idx_k, b.dihedrals_atom_idx_l,
     b.dihedral_idx) = ct.filter_dihedrals(aidx)
    b.nonbonded_atom_idx_src, b.nonbonded_atom_idx_dst = ct.gen_nonbonded_pair(aidx)
    b.lj_idx = ct.generate_lj_idx(b.nonbonded_atom_idx_src, b.nonbonded_atom_idx_dst)
    for k, like in (("bond_batch", b.bond_idx), ("angle_batch", b.angle_idx), ("dihedral_batch", b.dihedral_idx), ("nonbonded_batch", b.lj_idx)):
        setattr(b, k, torch.zeros_like(like))
    opt = ns["HydrogenOptimizer"](max_iter=10)                    # DistanceFragment(max_iter=10), distancefrag.py:30-32
    e0 = opt.cal_potential_energy(b).detach().numpy().reshape(-1)
    opt.optimize_hydrogen(b)
    pos1 = b.pos.detach().numpy()
    e1 = opt.cal_potential_energy(b).detach().numpy().reshape(-1)
    out = {"prmtop_text_sha": np.frombuffer(__import__("hashlib").sha256(open(f"{REF}/src/Fragmentation/prmtop/GG.prmtop", "rb").read()).digest(), dtype=np.uint8),
           "atom_idx": atom_idx, "pos0": pos, "pos1": pos1, "energy0": e0, "energy1": e1,
           "bonds": torch.stack([b.bonds_atom_idx_src, b.bonds_atom_idx_dst, b.bond_idx], 1).numpy(),
           "angles": torch.stack([b.angles_atom_idx_i, b.angles_atom_idx_j, b.angles_atom_idx_k, b.angle_idx], 1).numpy(),
           "dihedrals": torch.stack([b.dihedrals_atom_idx_i, b.dihedrals_atom_idx_j, b.dihedrals_atom_idx_k,
                                     b.dihedrals_atom_idx_l, b.dihedral_idx], 1).numpy(),
           "pairs": np.array(sorted(zip(b.nonbonded_atom_idx_src.tolist(), b.nonbonded_atom_idx_dst.tolist())))}
    for k in ("natom", "ntypes", "numbnd", "numang", "nptra"):
        out["t_" + k] = np.int64(getattr(ct, k))
    for k in fields:
        out["t_" + k] = getattr(ct, k).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_caph_lbfgs.npz"), **out)
    print(f"cap-H LBFGS reference: E {e0.sum():.4f} -> {e1.sum():.4f} kcal/mol, max shift {np.abs(pos1 - pos).max():.4f} A, "
          f"{len(out['pairs'])} pairs, {len(out['bonds'])}/{len(out['angles'])}/{len(out['dihedrals'])} bond/angle/dihedral terms")


def write_reference_caph_batch(name="chig"):
    import ast
    import importlib.util
    import types
    from oracle import caph_ref as CR
    spec = importlib.util.spec_from_file_location("ref_ctable", f"{REF}/src/Fragmentation/hydrogen/ctable.py")
    ctmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ctmod)
    tree = ast.parse(open(f"{REF}/src/Fragmentation/hydrogen/energies.py").read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))]
    for n in body:
        n.decorator_list = []
    ns = {"torch": torch, "F": torch.nn.functional, "ProteinData": object,
          "scatter_add": lambda src, index, dim=0: ref_shims._scatter(src, index, dim)}
    exec(compile(ast.Module(body=body, type_ignores=[]), "ref_energies", "exec"), ns)

    prot = read_pdb(f"{REF}/examples/{name}.pdb")
    resn = {int(r): n for r, n in zip(prot.resnums, prot.resnames)}
    R = int(prot.resnums.max())
    cat = {k: [] for k in ("pos", "atom_idx", "other_idx", "charge", "bond_force_constant", "bond_equil_value",
                           "angle_force_constant", "angle_equil_value", "dihedral_force_constant", "dihedral_periodicity",
                           "dihedral_phase", "lennard_jones_acoef", "lennard_jones_bcoef", "bonds_atom_idx_src",
                           "bonds_atom_idx_dst", "bond_idx", "angles_atom_idx_i", "angles_atom_idx_j", "angles_atom_idx_k",
                           "angle_idx", "dihedrals_atom_idx_i", "dihedrals_atom_idx_j", "dihedrals_atom_idx_k",
                           "dihedrals_atom_idx_l", "dihedral_idx", "nonbonded_atom_idx_src", "nonbonded_atom_idx_dst",
                           "lj_idx", "bond_batch", "angle_batch", "dihedral_batch", "nonbonded_batch")}
    out, stems = {}, []
    o_atom = o_bnd = o_ang = o_dih = o_lj = 0
    for g, c in enumerate(range(2, R)):
        stem = CR.PRMTOP_STEM[resn[c]]
        stems.append(stem)
        path = f"{REF}/src/Fragmentation/prmtop/{stem}.prmtop"
        ct = ctmod.CTable.from_prmtop(path)
        pos, aidx_np, src = CR.dipeptide_problem(prot, c, CR.prmtop_atom_names(open(path).read()))
        aidx = torch.from_numpy(aidx_np)
        bs, bd, bi = ct.filter_bonds(aidx)
        ai, aj, ak, an = ct.filter_angles(aidx)
        di, dj, dk, dl, dn = ct.filter_dihedrals(aidx)
        ps, pd = ct.gen_nonbonded_pair(aidx)
        lj = ct.generate_lj_idx(ps, pd)
        other = torch.tensor([i for i in range(ct.natom) if i not in aidx_np.tolist()])
        cat["pos"].append(torch.from_numpy(pos))
        for k in ("charge", "bond_force_constant", "bond_equil_value", "angle_force_constant", "angle_equil_value",
                  "dihedral_force_constant", "dihedral_periodicity", "dihedral_phase", "lennard_jones_acoef", "lennard_jones_bcoef"):
            cat[k].append(getattr(ct, k))
        for k, v in (("atom_idx", aidx), ("other_idx", other), ("bonds_atom_idx_src", bs), ("bonds_atom_idx_dst", bd),
                     ("angles_atom_idx_i", ai), ("angles_atom_idx_j", aj), ("angles_atom_idx_k", ak),
                     ("dihedrals_atom_idx_i", di), ("dihedrals_atom_idx_j", dj), ("dihedrals_atom_idx_k", dk),
                     ("dihedrals_atom_idx_l", dl), ("nonbonded_atom_idx_src", ps), ("nonbonded_atom_idx_dst", pd)):
            cat[k].append(v + o_atom)                               # ProteinData.__inc__: *_atom_idx / other_idx += natom
        cat["bond_idx"].append(bi + o_bnd)
        cat["angle_idx"].append(an + o_ang)
        cat["dihedral_idx"].append(dn + o_dih)
        cat["lj_idx"].append(lj + o_lj)
        for k, like in (("bond_batch", bi), ("angle_batch", an), ("dihedral_batch", dn), ("nonbonded_batch", lj)):
            cat[k].append(torch.full_like(like, g))
        o_atom += ct.natom; o_bnd += ct.numbnd; o_ang += ct.numang; o_dih += ct.nptra
        o_lj += ct.ntypes * (ct.ntypes + 1) // 2
        out[f"g{g}_pos0"], out[f"g{g}_atom_idx"], out[f"g{g}_src"] = pos, aidx_np, src
        out[f"g{g}_names"] = np.array(CR.prmtop_atom_names(open(path).read()))
        for k in ("natom", "ntypes", "numbnd", "numang", "nptra"):
            out[f"g{g}_t_{k}"] = np.int64(getattr(ct, k))
        for k in ("charge", "atomic_number", "atom_type_idx", "number_excluded_atoms", "nonbonded_parm_index",
                  "bond_force_constant", "bond_equil_value", "angle_force_constant", "angle_equil_value",
                  "dihedral_force_constant", "dihedral_periodicity", "dihedral_phase", "lennard_jones_acoef",
                  "lennard_jones_bcoef", "bonds_inc_hydrogen", "angles_inc_hydrogen", "dihedrals_inc_hydrogen", "excluded_atoms_list"):
            out[f"g{g}_t_{k}"] = getattr(ct, k).numpy()
    b = types.SimpleNamespace(**{k: torch.cat(v) for k, v in cat.items()})
    opt = ns["HydrogenOptimizer"](max_iter=10)
    e0 = opt.cal_potential_energy(b).detach().numpy()
    pos0 = b.pos.numpy().copy()
    opt.optimize_hydrogen(b)
    e1 = opt.cal_potential_energy(b).detach().numpy()
    out.update(n_graphs=np.int64(R - 2), stems=np.array(stems), energy0=e0, energy1=e1, pos1=b.pos.detach().numpy())
    np.savez_compressed(os.path.join(HERE, "reference_caph_batch.npz" if name == "chig" else f"reference_caph_batch_{name}.npz"), **out)
    print(f"cap-H joint LBFGS reference ({name}): {R - 2} dipeptides, {len(b.atom_idx)} added H, E {e0.sum():.3f} -> {e1.sum():.3f} kcal/mol, "
          f"max shift {np.abs(b.pos.detach().numpy() - pos0).max():.4f} A")


def write_reference_fragment_index():
    import ast
    import json
    import types
    tree = ast.parse(open(f"{REF}/src/Fragmentation/basefrag.py").read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "get_fragments_index"][0]
    fn.decorator_list = []
    ns = {"Atoms": object, "arguments": types.SimpleNamespace(get=lambda: types.SimpleNamespace(verbose=0))}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_basefrag", "exec"), ns)
    tree_h = ast.parse(open(f"{REF}/src/Fragmentation/distancefrag.py").read())
    fn_h = [n for n in ast.walk(tree_h) if isinstance(n, ast.FunctionDef) and n.name == "get_hydrogen_indices"][0]
    fn_h.decorator_list = []
    ns_h = {"Atoms": object, "np": np}
    exec(compile(ast.Module(body=[fn_h], type_ignores=[]), "ref_distancefrag_h", "exec"), ns_h)
    out = {}
    for name in ("chig", "trpcage", "ww", "abd"):
        prot = read_pdb(f"{REF}/examples/{name}.pdb")

        class FakeAtoms:
            arrays = {"residuenumbers": np.asarray(prot.resnums), "residuenames": np.asarray(prot.resnames),
                      "atomtypes": np.asarray(prot.names)}

            def __len__(self):
                return len(prot)

        dip, an = ns["get_fragments_index"](FakeAtoms())
        # added hydrogens of every dipeptide: (acceptor, removed atom, bond length) from get_hydrogen_indices (distancefrag.py:365-504)
        fake = FakeAtoms()
        fake.atom_masks = {a: np.asarray([n == a for n in prot.names]) for a in ("CA", "N", "C", "CB", "CD")}
        caps = []
        for k, unit in enumerate(dip):
            radii, acc, rem = ns_h["get_hydrogen_indices"](fake, k, unit)
            caps.append([[int(a), int(r), float(b)] for a, r, b in zip(acc, rem, radii.reshape(-1))])
        out[name] = {"dipeptides": [[int(i) for i in u] for u in dip], "acenmes": [[int(i) for i in u] for u in an],
                     "added_hydrogens": caps}
    with open(os.path.join(HERE, "reference_fragment_index.json"), "w") as fh:
        json.dump(out, fh, sort_keys=True)


def synthetic_cyx_protein(prot):
    """The WW-domain example with four SER residues turned into two disulfide-bridged CYX (OG -> SG): the reference ships no
    CYX-containing example; only names / elements change, geometry stays."""
    ser = sorted({int(r) for r, n in zip(prot.resnums, prot.resnames) if n == "SER"})
    pick = ser[:4]
    assert len(pick) == 4, ser
    names, resn, elem = list(prot.names), list(prot.resnames), list(prot.elements)
    for i in range(len(prot)):
        if int(prot.resnums[i]) in pick:
            resn[i] = "CYX"
            if names[i] == "OG":
                names[i], elem[i] = "SG", "S"
    return type(prot)(names, resn, prot.resnums, elem, prot.positions), pick


def write_reference_cyx():
    """Pairing of disulfide-bridged dipeptides by the reference's own ``get_cystine_bonds`` (distancefrag.py:804-844)
    on the synthetic CYX protein, with the dipeptide atom lists of the reference's own ``get_fragments_index``."""
    import ast
    import json
    import types
    prot, pick = synthetic_cyx_protein(read_pdb(f"{REF}/examples/ww.pdb"))
    tree = ast.parse(open(f"{REF}/src/Fragmentation/distancefrag.py").read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "get_cystine_bonds"][0]
    fn.decorator_list = []
    ns = {"np": np, "Protein": object, "arguments": types.SimpleNamespace(get=lambda: types.SimpleNamespace(verbose=0))}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_cyx", "exec"), ns)
    dips = json.load(open(os.path.join(HERE, "reference_fragment_index.json")))["ww"]["dipeptides"]

    class FakeProt:
        arrays = {"atomtypes": np.asarray(prot.names), "residuenames": np.asarray(prot.resnames)}

        def get_positions(self):
            return np.asarray(prot.positions)

    pairs = ns["get_cystine_bonds"](FakeProt(), dips)
    with open(os.path.join(HERE, "reference_cyx.json"), "w") as fh:
        json.dump({"cyx_residues": pick, "pairs": {str(int(k)): int(v) for k, v in pairs.items()}}, fh, sort_keys=True)
    print(f"CYX pairing reference: residues {pick} -> dipeptide pairs {pairs}")


def main():
    sd = O.load_state_dict(CKPT)
    O.save_weights_npz(sd, os.path.join(HERE, "weights_2ef43f29.npz"))

    frs = {}
    for name in ("chig", "trpcage", "ww", "abd"):
        prot = read_pdb(f"{REF}/examples/{name}.pdb")
        fd, pm, rc = fragment_protein(prot, with_recipe=True)
        frs[name] = (fd, pm)
        zmap = {"H": 1, "C": 6, "N": 7, "O": 8, "S": 16}
        np.savez_compressed(os.path.join(HERE, f"fragments_{name}.npz"), z=fd.z, pos=fd.pos, start=fd.start,
                            end=fd.end, batch=fd.batch, n_protein=pm.n_protein, src_atom=pm.src_atom,
                            dst_atom=pm.dst_atom, sign=pm.sign, frag_sign=pm.frag_sign,
                            prot_pos=prot.positions, prot_z=np.array([zmap[e] for e in prot.elements]),
                            rc_real=rc.real, rc_acc=rc.acc, rc_rem=rc.rem, rc_blen=rc.blen,
                            prot_names=np.array(prot.names), prot_resnames=np.array(prot.resnames), prot_resnums=prot.resnums)

    # the reference's own fragment composition tables (numpy-only module) + residue sequences of the examples
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("ref_reference", f"{REF}/src/utils/reference.py")
    refmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refmod)
    tables = {"z_by_residue": {k: [int(x) for x in v] for k, v in refmod.fragment_atomic_numbers.items()}, "sequence": {}}
    for name in ("chig", "trpcage", "ww", "abd"):
        prot = read_pdb(f"{REF}/examples/{name}.pdb")
        seq = {}
        for r, n in zip(prot.resnums, prot.resnames):
            seq[int(r)] = n
        tables["sequence"][name] = [seq[r] for r in sorted(seq)]
    with open(os.path.join(HERE, "fragment_tables.json"), "w") as fh:
        json.dump(tables, fh, indent=0, sort_keys=True)

    write_reference_partitions(frs)
    write_reference_fragment_index()
    write_reference_cyx()
    write_reference_host_logic(*frs["chig"])
    from ai2bmd_b200.fixtures import load_protein
    write_reference_nonbonded(*frs["chig"], *load_protein("chig"))
    write_reference_caph(*frs["chig"], *load_protein("chig"))
    write_reference_caph_lbfgs()
    write_reference_caph_batch("chig")
    write_reference_caph_batch("trpcage")      # PRO / GLY neighbours, 13 residue types

    model = load_reference_model()
    o64 = O.OracleViSNet(sd, torch.float64)
    out = {}
    cases = {
        "chig": frs["chig"][0],
        "trpcage": frs["trpcage"][0],
        # config C1: an ALA-centred dipeptide (ACE-ALA-NME-like, 22 atoms) as ONE graph: trpcage dipeptide 1
        "c1_ala": frs["trpcage"][0][2],
        "dense44": dense_fragment(),
    }
    for key, fd in cases.items():
        e, f = ref_eval(model, fd)
        e64, f64 = o64.energy_and_forces(fd.z, fd.pos, fd.batch)
        slots, deg = O.radius_graph_canonical(fd.pos, fd.batch)
        out[f"{key}_z"], out[f"{key}_pos"], out[f"{key}_batch"] = fd.z, fd.pos, fd.batch
        out[f"{key}_ref_e"], out[f"{key}_ref_f"] = e, f
        out[f"{key}_e64"], out[f"{key}_f64"] = e64.numpy(), f64.numpy()
        out[f"{key}_slots"], out[f"{key}_deg"] = slots, deg
        print(f"{key}: G={len(fd)} N={len(fd.z)} E={int(deg.sum())} maxdeg={int(deg.max())} "
              f"|ref-o64| E {np.abs(e - e64.numpy()).max():.3e} F {np.abs(f - f64.numpy()).max():.3e}")
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)


if __name__ == "__main__":
    main()
