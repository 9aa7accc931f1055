#!/usr/bin/env python
"""What does the reference's per-step hydrogen refinement change?  CPU only.

Takes the Chignolin fragment fixture (added hydrogens on the acceptor->removed ray, what the device MD loop does) and the
coordinates the reference's own joint LBFGS produces from it (tests/golden/reference_caph_batch.npz), evaluates both with
the CPU ViSNet oracle and reports the difference in whole-protein energy and forces.

    python tools/caph_effect.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ai2bmd_b200.fixtures import WEIGHTS, load_fragments           # noqa: E402
from ai2bmd_b200.parallel import combine_local                     # noqa: E402
from ai2bmd_b200.weights import load_state_dict                    # noqa: E402
from oracle import visnet_ref as O                                 # noqa: E402


def refined_fragment_positions(fd, g):
    """Fragment coordinates with every added hydrogen moved to its refined place (dipeptides from the golden file;
    ACE-NME caps share their hydrogens with the neighbouring dipeptides, distancefrag.py:286-307)."""
    pos0 = np.concatenate([g[f"g{k}_pos0"] for k in range(int(g["n_graphs"]))])
    moved = np.abs(g["pos1"] - pos0).max(1) > 0
    src, dst = pos0[moved], g["pos1"][moved]
    out = fd.pos.copy()
    n_hit = 0
    for a in range(len(out)):
        d = np.abs(src - out[a]).max(1)
        j = int(np.argmin(d))
        if d[j] < 2e-5:                       # an added hydrogen at its first-approximation position
            out[a] = dst[j]
            n_hit += 1
    return out, int(moved.sum()), n_hit


def main():
    fd, pm = load_fragments("chig")
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_caph_batch.npz"))
    pos1, n_moved, n_hit = refined_fragment_positions(fd, g)
    sd = load_state_dict(WEIGHTS)
    oracle = O.OracleViSNet({k: torch.from_numpy(v) for k, v in sd.items()}, torch.float64)
    e0, f0 = oracle.energy_and_forces(fd.z, fd.pos, fd.batch)
    e1, f1 = oracle.energy_and_forces(fd.z, pos1, fd.batch)
    ef0 = combine_local(pm, e0.numpy().reshape(-1), f0.numpy())
    ef1 = combine_local(pm, e1.numpy().reshape(-1), f1.numpy())
    print(f"added hydrogens moved by the reference's joint LBFGS: {n_moved} in dipeptides, {n_hit} fragment atoms updated")
    print(f"max hydrogen shift            {np.abs(pos1 - fd.pos).max():.5f} A")
    print(f"max |dE| per fragment         {np.abs(e1.numpy() - e0.numpy()).max():.3e} eV")
    de = float((pm.frag_sign.astype(np.float64) * (e1.numpy().reshape(-1) - e0.numpy().reshape(-1))).sum())
    print(f"whole-protein energy change   {de:.3e} eV   (fp64 sum; the fp32 buffer resolves 8e-3 eV at -1.3e5 eV)")
    print(f"max |dF| on a protein atom    {np.abs(ef1[:-1] - ef0[:-1]).max():.3e} eV/A   (max |F| {np.abs(ef0[:-1]).max():.2f})")


if __name__ == "__main__":
    main()
