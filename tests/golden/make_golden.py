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
                            rc_real=rc.real, rc_acc=rc.acc, rc_rem=rc.rem, rc_blen=rc.blen)

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
