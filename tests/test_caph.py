"""Cap-hydrogen refinement oracle (oracle/caph_ref.py): pinned against outputs of the reference's own
CTable / energy functions / HydrogenOptimizer (tests/golden/reference_caph_lbfgs.npz) and against torch.optim.LBFGS."""
import os

import numpy as np
import pytest
import torch

from oracle import caph_ref as CR


@pytest.fixture(scope="module")
def gold(golden_dir):
    g = np.load(os.path.join(golden_dir, "reference_caph_lbfgs.npz"))
    t = {k[2:]: (int(g[k]) if g[k].ndim == 0 else g[k]) for k in g.files if k.startswith("t_")}
    return g, t


def _rows(a):
    return sorted(map(tuple, np.asarray(a).tolist()))


MINI_PRMTOP = """%VERSION  VERSION_STAMP = V0001.000
%FLAG POINTERS
%FORMAT(10I8)
       3       2       2       0       1       0       0       0       0       0
       3       1       0       0       0       2       1       1       2       0
%FLAG ATOM_NAME
%FORMAT(20a4)
H1  C   H2
%FLAG CHARGE
%FORMAT(5E16.8)
  1.00000000E+00 -2.00000000E+00  1.00000000E+00
%FLAG ATOMIC_NUMBER
%FORMAT(10I8)
       1       6       1
%FLAG ATOM_TYPE_INDEX
%FORMAT(10I8)
       1       2       1
%FLAG NUMBER_EXCLUDED_ATOMS
%FORMAT(10I8)
       2       1       1
%FLAG NONBONDED_PARM_INDEX
%FORMAT(10I8)
       1       2       2       3
%FLAG BOND_FORCE_CONSTANT
%FORMAT(5E16.8)
  3.40000000E+02  3.00000000E+02
%FLAG BOND_EQUIL_VALUE
%FORMAT(5E16.8)
  1.09000000E+00  1.50000000E+00
%FLAG ANGLE_FORCE_CONSTANT
%FORMAT(5E16.8)
  3.50000000E+01
%FLAG ANGLE_EQUIL_VALUE
%FORMAT(5E16.8)
  1.91000000E+00
%FLAG DIHEDRAL_FORCE_CONSTANT
%FORMAT(5E16.8)
  1.50000000E-01
%FLAG DIHEDRAL_PERIODICITY
%FORMAT(5E16.8)
  3.00000000E+00
%FLAG DIHEDRAL_PHASE
%FORMAT(5E16.8)
  0.00000000E+00
%FLAG SCEE_SCALE_FACTOR
%FORMAT(5E16.8)
  1.20000000E+00
%FLAG SCNB_SCALE_FACTOR
%FORMAT(5E16.8)
  2.00000000E+00
%FLAG LENNARD_JONES_ACOEF
%FORMAT(5E16.8)
  1.00000000E+01  2.00000000E+02  3.00000000E+03
%FLAG LENNARD_JONES_BCOEF
%FORMAT(5E16.8)
  1.00000000E+00  2.00000000E+01  3.00000000E+02
%FLAG BONDS_INC_HYDROGEN
%FORMAT(10I8)
       0       3       1       3       6       1
%FLAG ANGLES_INC_HYDROGEN
%FORMAT(10I8)
       0       3       6       1
%FLAG DIHEDRALS_INC_HYDROGEN
%FORMAT(10I8)

%FLAG EXCLUDED_ATOMS_LIST
%FORMAT(10I8)
       2       3       3       0
"""


def test_prmtop_parser_on_a_hand_written_table():
    t = CR.parse_prmtop(MINI_PRMTOP)
    assert (t["natom"], t["ntypes"], t["numbnd"], t["numang"], t["nptra"]) == (3, 2, 2, 1, 1)
    assert t["charge"].tolist() == [1.0, -2.0, 1.0] and t["atom_type_idx"].tolist() == [0, 1, 0]
    assert t["nonbonded_parm_index"].tolist() == [0, 1, 1, 2]
    assert t["bonds_inc_hydrogen"].tolist() == [[0, 1, 0], [1, 2, 0]]         # 3*index -> index, 1-based type -> 0-based
    assert t["angles_inc_hydrogen"].tolist() == [[0, 1, 2, 0]] and t["dihedrals_inc_hydrogen"].shape == (0, 5)
    assert t["excluded_atoms_list"].tolist() == [1, 2, 2, -1]
    terms = CR.hydrogen_terms(t, [0])
    assert terms["bonds"].tolist() == [[0, 1, 0]] and terms["angles"].tolist() == [[0, 1, 2, 0]]
    assert terms["pairs"].shape == (0, 2)                                     # both partners of atom 0 are excluded


def test_tables_and_terms_equal_the_reference_ctable(gold):
    g, t = gold
    assert t["natom"] == 19 and t["atomic_number"].tolist() == [1, 6, 1, 1, 6, 8, 7, 1, 6, 1, 1, 6, 8, 7, 1, 6, 1, 1, 1]
    terms = CR.hydrogen_terms(t, g["atom_idx"])
    for key in ("bonds", "angles", "dihedrals", "pairs"):
        assert _rows(terms[key]) == _rows(g[key]), key
    assert len(terms["pairs"]) == 44 and (terms["lj_idx"] >= 0).all()


def test_energy_terms_equal_the_reference_functions(gold):
    g, t = gold
    terms = CR.hydrogen_terms(t, g["atom_idx"])
    e0 = CR.amber_energy(torch.from_numpy(g["pos0"]), t, terms).numpy()
    assert np.abs(e0 - g["energy0"]).max() <= 2e-6 * np.abs(g["energy0"]).max()
    e1 = CR.amber_energy(torch.from_numpy(g["pos1"]), t, terms).numpy()
    assert np.abs(e1 - g["energy1"]).max() <= 2e-6 * np.abs(g["energy1"]).max()
    assert e1.sum() < e0.sum()


def test_relaxed_hydrogens_equal_the_reference_optimizer(gold):
    g, t = gold
    p1 = CR.optimize_hydrogens(g["pos0"], t, g["atom_idx"], max_iter=10)
    moved = np.abs(g["pos1"] - g["pos0"]).max(axis=1) > 0
    assert moved.nonzero()[0].tolist() == sorted(g["atom_idx"].tolist())     # only the added hydrogens move
    assert np.abs(p1 - g["pos1"]).max() <= 2e-6                                # measured: identical in fp32
    p64 = CR.optimize_hydrogens(g["pos0"], t, g["atom_idx"], max_iter=10, dtype=torch.float64)
    assert np.abs(p64 - g["pos1"]).max() <= 1e-5                               # fp32 trajectory of the reference vs fp64


@pytest.mark.parametrize("scale,max_iter", [(1.0, 10), (5.0, 10), (2.0, 3), (3.0, 25), (0.05, 10), (0.3, 10)])
def test_lbfgs_restatement_equals_torch_lbfgs(scale, max_iter):
    """Same iterates and the same number of closure calls as torch.optim.LBFGS (no line search), including the early
    exits on tolerance_grad / tolerance_change (small scales)."""
    rng = np.random.default_rng(3)
    n = 12
    A = rng.standard_normal((n, n))
    A = A @ A.T / n + np.eye(n) * 0.5
    b = rng.standard_normal(n) * scale
    x0 = rng.standard_normal(n) * scale
    At, bt = torch.tensor(A), torch.tensor(b)

    def f(x):
        return 0.5 * x @ At @ x - bt @ x + 0.1 * torch.sin(x).sum()

    p = torch.nn.Parameter(torch.tensor(x0))
    opt = torch.optim.LBFGS([p], lr=0.1, max_iter=max_iter, tolerance_grad=0.1, tolerance_change=0.01)
    calls = [0]

    def closure():
        opt.zero_grad()
        loss = f(p)
        loss.backward()
        calls[0] += 1
        return loss

    opt.step(closure)

    def fg(x):
        xt = torch.tensor(x, requires_grad=True)
        loss = f(xt)
        (grad,) = torch.autograd.grad(loss, xt)
        return float(loss.detach()), grad.numpy()

    x, evals = CR.lbfgs_fixed_step(fg, x0, max_iter=max_iter)
    assert evals == calls[0]
    assert np.abs(x - p.detach().numpy()).max() <= 1e-12


def test_hand_derived_gradient_equals_autograd(gold):
    """The gradient formulas a device kernel would implement, on a perturbed geometry with every hydrogen selected
    (10 bonds, 21 angles, 25 dihedrals, 79 pairs), against autograd of the restated energy in fp64."""
    g, t = gold
    sel = np.flatnonzero(t["atomic_number"] == 1)
    terms = CR.hydrogen_terms(t, sel)
    assert len(terms["dihedrals"]) >= 20 and len(terms["pairs"]) >= 70
    rng = np.random.default_rng(0)
    pos = g["pos0"].astype(np.float64) + rng.standard_normal((19, 3)) * 0.05
    p = torch.tensor(pos, requires_grad=True)
    e_t = CR.amber_energy(p, t, terms).sum()
    (g_t,) = torch.autograd.grad(e_t, p)
    e, grad = CR.amber_energy_and_grad(pos, t, terms)
    assert abs(float(e) - float(e_t.detach())) <= 1e-10
    assert np.abs(grad - g_t.numpy()).max() <= 1e-10 * max(1.0, np.abs(g_t.numpy()).max())


def test_relaxation_with_the_hand_derived_gradient(gold):
    g, t = gold
    p32 = CR.optimize_hydrogens_analytic(g["pos0"], t, g["atom_idx"], max_iter=10)
    assert p32.dtype == np.float32 and np.abs(p32 - g["pos1"]).max() <= 2e-5      # fp32, different summation order
    p64 = CR.optimize_hydrogens_analytic(g["pos0"].astype(np.float64), t, g["atom_idx"], max_iter=10)
    ref64 = CR.optimize_hydrogens(g["pos0"], t, g["atom_idx"], max_iter=10, dtype=torch.float64)
    assert np.abs(p64 - ref64).max() <= 1e-9


def test_joint_relaxation_equals_the_reference_batch(golden_dir):
    """One LBFGS over the 35 added hydrogens of all 10 Chignolin dipeptides (how the reference runs it, with
    ProteinData.__inc__ offsets) -- coordinates and per-dipeptide energy terms from the reference's own functions."""
    g = np.load(os.path.join(golden_dir, "reference_caph_batch.npz"))
    problems = []
    for k in range(int(g["n_graphs"])):
        pre = f"g{k}_t_"
        t = {n[len(pre):]: (int(g[n]) if g[n].ndim == 0 else g[n]) for n in g.files if n.startswith(pre)}
        problems.append((g[f"g{k}_pos0"], t, g[f"g{k}_atom_idx"]))
    assert sum(len(p[2]) for p in problems) == 35
    e0 = np.stack([CR.amber_energy(torch.from_numpy(p[0]), p[1], CR.hydrogen_terms(p[1], p[2])).numpy() for p in problems])
    assert np.abs(e0 - g["energy0"]).max() <= 4e-6 * np.abs(g["energy0"]).max()
    out = np.concatenate(CR.optimize_hydrogens_batch(problems, max_iter=10))
    pos0 = np.concatenate([p[0] for p in problems])
    assert np.abs(out - g["pos1"]).max() <= 2e-6                            # measured: identical in fp32
    # the joint first step is min(1, 1/|g|_1) * lr over ALL hydrogens: the refinement moves them by millis of an Angstrom
    assert 1e-3 < np.abs(g["pos1"] - pos0).max() < 1e-2


def _batch_problems(golden_dir, name="chig"):
    g = np.load(os.path.join(golden_dir, "reference_caph_batch.npz" if name == "chig" else f"reference_caph_batch_{name}.npz"))
    problems = []
    for k in range(int(g["n_graphs"])):
        pre = f"g{k}_t_"
        t = {n[len(pre):]: (int(g[n]) if g[n].ndim == 0 else g[n]) for n in g.files if n.startswith(pre)}
        problems.append((g[f"g{k}_pos0"], t, g[f"g{k}_atom_idx"]))
    return g, problems


def test_c_restatement_on_flat_arrays(golden_dir, gold):
    """oracle/caph_ref.c (fp32, flat term arrays as a kernel will see them): energy and hydrogen gradient against the
    Python oracle in fp64, relaxed coordinates against the reference's own optimiser output (single and joint)."""
    from oracle import caph_c as CC
    g, problems = _batch_problems(golden_dir)
    f = CC.flatten(problems)
    assert f["pos"].shape == (283, 3) and len(f["h_idx"]) == 35 and len(f["pair_a"]) == 695
    e, grad = CC.energy_grad(f)
    e64, g64 = 0.0, []
    for pos, t, aidx in problems:
        ee, gg = CR.amber_energy_and_grad(pos.astype(np.float64), t, CR.hydrogen_terms(t, aidx))
        e64 += float(ee)
        g64.append(gg[aidx])
    g64 = np.concatenate(g64)
    assert abs(e - e64) <= 2e-5 * max(1.0, abs(e64))
    assert np.abs(grad[f["h_idx"]] - g64).max() <= 1e-5 * np.abs(g64).max()
    x, evals = CC.relax(f)
    assert np.abs(x - g["pos1"]).max() <= 4e-6 and 2 <= evals <= 12          # measured: 1 ulp, 2 evaluations
    g1, t1 = gold
    x1, _ = CC.relax(CC.flatten([(g1["pos0"], t1, g1["atom_idx"])]))
    assert np.abs(x1 - g1["pos1"]).max() <= 4e-6


def test_joint_relaxation_on_trpcage(golden_dir):
    """20 dipeptides, 74 added hydrogens, PRO and GLY neighbours (N-H cap on the N->CD ray, single-hydrogen methyls), ten
    prmtop tables: Python and C restatements against the reference's own optimiser output."""
    from oracle import caph_c as CC
    g, problems = _batch_problems(golden_dir, "trpcage")
    assert len(problems) == 20 and sum(len(p[2]) for p in problems) == 74
    assert {"PP", "GG", "WW", "RR"} <= set(g["stems"].tolist())
    e0 = np.stack([CR.amber_energy(torch.from_numpy(p[0]), p[1], CR.hydrogen_terms(p[1], p[2])).numpy() for p in problems])
    assert np.abs(e0 - g["energy0"]).max() <= 4e-6 * np.abs(g["energy0"]).max()
    out = np.concatenate(CR.optimize_hydrogens_batch(problems, max_iter=10))
    assert np.abs(out - g["pos1"]).max() <= 2e-6
    x, _ = CC.relax(CC.flatten(problems))
    assert np.abs(x - g["pos1"]).max() <= 4e-6


# ---- product-side host logic (ai2bmd_b200/caph.py): flat term arrays over the packed fragment atoms --------------------
@pytest.mark.parametrize("name,n_h,n_pairs", [("chig", 35, 695), ("trpcage", 74, None)])
def test_product_problem_builder_against_the_reference_output(name, n_h, n_pairs):
    """``caph.build_problem`` (name-based table layout, terms, ACE-NME mirrors) feeds the C restatement: relaxing the PACKED
    fragment buffer with its arrays reproduces the coordinates of the reference's own optimiser."""
    from ai2bmd_b200 import caph
    from ai2bmd_b200.fixtures import load_capped_protein, load_caph_tables, load_fragments, load_protein
    from oracle import caph_c as CC
    fd, pm = load_fragments(name)
    _, _, recipe = load_protein(name)
    prot = load_capped_protein(name)
    tables, g = load_caph_tables(name)
    pr = caph.build_problem(prot, fd, recipe, tables)
    assert len(pr.h_idx) == n_h and (n_pairs is None or len(pr.pair_a) == n_pairs)
    # the layout puts every table atom on the fragment atom with the same first-approximation coordinates
    for k, t2f in enumerate(pr.table_to_frag):
        assert np.abs(fd.pos[t2f] - g[f"g{k}_pos0"]).max() <= 2e-6
        assert sorted(t2f.tolist()) == list(range(int(fd.start[2 * k]), int(fd.end[2 * k])))
    # every added hydrogen of an ACE-NME fragment mirrors a dipeptide hydrogen at the same place
    assert len(pr.mirror_dst) == int((recipe.real[np.isin(fd.batch, np.arange(1, len(fd), 2))] < 0).sum())
    assert np.abs(fd.pos[pr.mirror_dst] - fd.pos[pr.mirror_src]).max() <= 1e-5
    flat = {"pos": fd.pos.copy(), "h_idx": pr.h_idx.astype(np.int64)}
    for key in ("bond_ij", "angle_ijk", "dih_ijkl", "pair_ij"):
        flat[key] = np.ascontiguousarray(getattr(pr, key), dtype=np.int64)
    for key in ("bond_k", "bond_r0", "angle_k", "angle_t0", "dih_k", "dih_n", "dih_p", "pair_a", "pair_b", "pair_qq"):
        flat[key] = np.ascontiguousarray(getattr(pr, key), dtype=np.float32)
    x, evals = CC.relax(flat)
    pos1 = g["pos1"]
    off = 0
    for k, t2f in enumerate(pr.table_to_frag):
        assert np.abs(x[t2f] - pos1[off:off + len(t2f)]).max() <= 4e-6
        off += len(t2f)
    assert 2 <= evals <= 12


def test_product_prmtop_parser_equals_the_oracle_parser():
    from ai2bmd_b200 import caph
    a, b = caph.parse_prmtop(MINI_PRMTOP), CR.parse_prmtop(MINI_PRMTOP)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    ta, tb = caph.hydrogen_terms(a, [0]), CR.hydrogen_terms(b, [0])
    for k in ta:
        assert np.array_equal(ta[k], tb[k]), k
