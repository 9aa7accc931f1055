"""CPU restatement of the reference's non-bonded MM term -- TEST INFRASTRUCTURE ONLY (checker of
``ai2bmd_b200/csrc/k_nonbonded.cuh``; never imported by the product path).

Follows ``MMNonBondedCalculator`` line by line (``/root/reference/src/Calculators/nonbonded.py:15-63``) with the
pair list of ``Protein.initial_mm_adjmatrix`` (``src/AIMD/protein.py:133-151``) and the exclusions of
``src/Fragmentation/distancefrag.py:355-363``.  ``torch_scatter.scatter_add`` is restated with ``index_add_``; the
ASE unit constants are ASE 3.22's CODATA-2014 values (recalled: ``ase/units.py``).

Parity status: the formula, the pair list and the scatter are pinned -- ``tests/golden/reference_nonbonded.npz`` holds
the output of the reference's own ``MMNonBondedCalculator.__call__`` / ``Protein.initial_mm_adjmatrix`` bodies executed
in the authoring container (tests/golden/make_golden.py) and this restatement reproduces it to fp32 round-off.  Still
**unpinned**: the six ``ase.units`` constants (ASE is absent; CODATA-2014 values recalled from ``ase/units.py``) and the
force-field parameters (OpenMM amber14 is not installable; tests use synthetic amber-like values).  Further self-checks
(tests/test_nonbonded.py): fp64 forces equal the negative finite-difference gradient of the fp64 energy; a
hand-computed two-atom case with the textbook 1389.35 kJ/mol*A/e^2 Coulomb constant.
"""
import numpy as np
import torch

# ase.units (CODATA 2014)
_c, _mu0 = 299792458.0, 4.0e-7 * np.pi
_eps0 = 1.0 / _mu0 / _c ** 2
_e, _Nav = 1.6021766208e-19, 6.022140857e23
C, mol, kJ, nm, pi = 1.0 / _e, _Nav, 1000.0 / _e, 10.0, np.pi


def pair_list(n_atoms, exclude_pair):
    """All ordered pairs i != j not in ``exclude_pair`` (protein.py:140-151): rows (src, dst)."""
    ex = set(exclude_pair)
    pairs = [(i, j) for i in range(n_atoms) for j in range(n_atoms) if i != j and (i, j) not in ex]
    return torch.tensor(pairs, dtype=torch.long).t().reshape(2, -1)


def exclude_pairs_from_groups(groups):
    """distancefrag.py:355-361."""
    from itertools import combinations
    ex = set()
    for g in groups:
        for x, y in combinations([int(a) for a in g], 2):
            ex.add((x, y))
            ex.add((y, x))
    return ex


def nonbonded(positions, charges, sigmas, epsilons, src, dst, dtype=torch.float32):
    """(energy [eV], forces [n,3] eV/A), nonbonded.py:34-63 (dtype float32 there; float64 for the gradient check)."""
    k = 1 / (4 * pi * _eps0) * 10e6 * mol * C ** (-2)                       # :18
    pos = torch.as_tensor(np.asarray(positions), dtype=dtype)               # :39
    sig = torch.as_tensor(np.asarray(sigmas), dtype=dtype)
    eps = torch.as_tensor(np.asarray(epsilons), dtype=dtype)
    q = torch.as_tensor(np.asarray(charges), dtype=dtype)
    vec = pos[dst] - pos[src]                                               # :41
    d2 = (vec ** 2).sum(-1)
    d = torch.sqrt(d2)
    sigmaij = 0.5 * (sig[src] + sig[dst]) * nm                              # :46
    epsij = torch.sqrt(eps[src] * eps[dst])
    c6 = (sigmaij ** 2 / d2) ** 3
    c12 = c6 ** 2
    energy_lj = 4 * epsij * (c12 - c6)
    force_lj = (24 * epsij * (2 * c12 - c6) / d2).unsqueeze(-1) * vec       # :51
    energy_coulomb = k * q[src] * q[dst] / d                                # :54
    force_coulomb = (energy_coulomb / d2).unsqueeze(-1) * vec
    energy = energy_lj.sum() + energy_coulomb.sum()                         # :58
    force = force_lj + force_coulomb
    out = torch.zeros((len(pos), 3), dtype=dtype)
    out.index_add_(0, dst, force)                                           # scatter_add(force, dst), :60
    return energy.item() * (kJ / mol) / 2, out.numpy() * (kJ / mol)         # :61-62
