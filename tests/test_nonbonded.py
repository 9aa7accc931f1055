"""Non-bonded MM term: oracle self-checks and host tables on CPU, kernel parity on the GPU."""
import os

import numpy as np
import pytest
import torch

from ai2bmd_b200.fixtures import load_fragments, load_protein
from ai2bmd_b200.nonbonded import dipeptide_atom_sets, exclusion_table, synthetic_parameters
from oracle import nonbonded_ref as R


def _chig():
    fd, pm = load_fragments("chig")
    prot_pos, prot_z, recipe = load_protein("chig")
    groups = dipeptide_atom_sets(fd, recipe, pm)
    return fd, pm, prot_pos, prot_z, recipe, groups


def test_oracle_two_atom_case_and_units():
    # two opposite unit charges 3 A apart, no LJ: E = -k/3 kJ/mol with k = 1389.35 kJ/mol*A/e^2, force attractive
    pos = np.array([[0, 0, 0], [3.0, 0, 0]])
    src, dst = torch.tensor([0, 1]), torch.tensor([1, 0])
    e, f = R.nonbonded(pos, [1.0, -1.0], [0.3, 0.3], [0.0, 0.0], src, dst, torch.float64)
    kjmol = 0.010364269574711572
    assert abs(e - (-1389.3545764 / 3.0) * kjmol) < 1e-6
    assert f[0, 0] > 0 and f[1, 0] < 0 and abs(f[0, 0] + f[1, 0]) < 1e-12
    assert abs(f[0, 0] - 1389.3545764 / 9.0 * kjmol) < 1e-6
    # pure LJ at the minimum r = 2^(1/6) sigma: E_pair = -eps, zero force
    r = 2 ** (1 / 6) * 3.0
    e, f = R.nonbonded(np.array([[0, 0, 0], [r, 0, 0]]), [0, 0], [0.3, 0.3], [0.5, 0.5], src, dst, torch.float64)
    assert abs(e - (-0.5) * kjmol) < 1e-9 and np.abs(f).max() < 1e-9


def test_oracle_forces_are_the_energy_gradient_and_exclusions_match():
    fd, pm, prot_pos, prot_z, recipe, groups = _chig()
    n = len(prot_z)
    assert len(groups) == 10 and all(len(g) > 0 for g in groups)
    rowptr, col = exclusion_table(n, groups)
    ex = R.exclude_pairs_from_groups(groups)
    assert rowptr[-1] == len(ex)                                   # ordered pairs, both directions
    listed = {(i, int(j)) for i in range(n) for j in col[rowptr[i]:rowptr[i + 1]]}
    assert listed == ex
    for i in range(n):
        row = col[rowptr[i]:rowptr[i + 1]]
        assert (np.diff(row) > 0).all() and i not in row
    q, sg, ep = synthetic_parameters(prot_z, seed=1)
    src, dst = R.pair_list(n, ex)
    assert src.numel() == n * (n - 1) - len(ex)
    e, f = R.nonbonded(prot_pos, q, sg, ep, src, dst, torch.float64)
    rng = np.random.default_rng(0)
    for _ in range(4):
        a, c = int(rng.integers(n)), int(rng.integers(3))
        h = 1e-5
        p1, p2 = prot_pos.astype(np.float64).copy(), prot_pos.astype(np.float64).copy()
        p1[a, c] += h
        p2[a, c] -= h
        e1, _ = R.nonbonded(p1, q, sg, ep, src, dst, torch.float64)
        e2, _ = R.nonbonded(p2, q, sg, ep, src, dst, torch.float64)
        assert abs(-(e1 - e2) / (2 * h) - f[a, c]) <= 1e-6 * max(1.0, abs(f[a, c]))
    assert np.abs(f.sum(0)).max() < 1e-9                           # pairwise forces cancel


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "reference_nonbonded.npz"))
    groups = [g["group_atoms"][a:b] for a, b in zip(g["group_ptr"][:-1], g["group_ptr"][1:])]
    return g, groups


def test_oracle_equals_the_reference_class_body(golden_dir):
    """tests/golden/reference_nonbonded.npz holds the output of the reference's own ``MMNonBondedCalculator.__call__``
    and ``Protein.initial_mm_adjmatrix`` bodies (make_golden.py): same pair list, same fp32 energy and forces."""
    g, groups = _golden(golden_dir)
    n = len(g["charges"])
    src, dst = R.pair_list(n, R.exclude_pairs_from_groups(groups))
    assert np.array_equal(src.numpy(), g["src"]) and np.array_equal(dst.numpy(), g["dst"])
    e, f = R.nonbonded(g["positions"], g["charges"], g["sigmas"], g["epsilons"], src, dst, torch.float32)
    assert abs(e - float(g["energy"])) <= 1e-6 * abs(float(g["energy"])) + 1e-6
    assert np.abs(f - g["forces"]).max() <= 1e-6 * np.abs(g["forces"]).max()
    rowptr, col = exclusion_table(n, groups)
    assert n * (n - 1) - int(rowptr[-1]) == len(g["src"])


@pytest.mark.gpu
def test_kernel_parity_with_the_reference_golden(real_weights, golden_dir):
    """The CUDA kernel against the reference class body's own output (same tolerance as against the oracle)."""
    from ai2bmd_b200.engine import Engine
    from ai2bmd_b200.nonbonded import MMNonBondedCalculator
    g, groups = _golden(golden_dir)
    n = len(g["charges"])
    calc = MMNonBondedCalculator(Engine(real_weights, 0))
    calc.set_parameters(g["charges"], g["sigmas"], g["epsilons"], *exclusion_table(n, groups))
    e, f = calc(g["positions"])
    assert np.abs(f - g["forces"]).max() <= 4e-5 * np.abs(g["forces"]).max() + 2e-6
    assert abs(e - float(g["energy"])) <= 1e-3 * abs(float(g["energy"])) + 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["chig", "trpcage"])
def test_kernel_parity_with_oracle(real_weights, name):
    """fp32 kernel vs the fp32 restatement of nonbonded.py (and the fp64 anchor): the per-atom sums differ only in
    summation order.  Stated tolerance: |dF| <= 2e-5 * max|F| + 1e-6 eV/A, |dE| <= 2e-6 * sum|E_pair| scale."""
    from ai2bmd_b200.engine import Engine
    from ai2bmd_b200.nonbonded import MMNonBondedCalculator
    fd, pm = load_fragments(name)
    prot_pos, prot_z, recipe = load_protein(name)
    n = len(prot_z)
    groups = dipeptide_atom_sets(fd, recipe, pm)
    rowptr, col = exclusion_table(n, groups)
    q, sg, ep = synthetic_parameters(prot_z, seed=2)
    src, dst = R.pair_list(n, R.exclude_pairs_from_groups(groups))
    e32, f32 = R.nonbonded(prot_pos, q, sg, ep, src, dst, torch.float32)
    e64, f64 = R.nonbonded(prot_pos.astype(np.float32), q, sg, ep, src, dst, torch.float64)
    eng = Engine(real_weights, 0)
    calc = MMNonBondedCalculator(eng)
    calc.set_parameters(q, sg, ep, rowptr, col)
    e, f = calc(prot_pos)
    ftol = 2e-5 * np.abs(f64).max() + 1e-6
    assert np.abs(f - f64).max() <= ftol and np.abs(f - f32).max() <= 2 * ftol
    assert abs(e - e64) <= 2e-5 * abs(e64) + 1e-4 and abs(e - e32) <= 1e-3 * abs(e64) + 1e-3
    # destination slices (the sharded layout) add up to the whole
    parts = np.zeros(3 * n + 1, np.float32)
    for lo, hi in ((0, n // 3), (n // 3, n)):
        calc.set_parameters(q, sg, ep, rowptr, col, lo, hi)
        es, fs = calc(prot_pos)
        parts[:-1] += fs.reshape(-1)
        parts[-1] += es
        assert np.abs(fs[:lo]).max(initial=0.0) == 0 and np.abs(fs[hi:]).max(initial=0.0) == 0
    assert np.abs(parts[:-1].reshape(-1, 3) - f).max() <= 1e-6 and abs(parts[-1] - e) <= 1e-4 * max(1.0, abs(e))


@pytest.mark.gpu
def test_md_step_includes_the_nonbonded_term(real_weights):
    """With vb_set_nonbonded, the device MD step integrates bonded + non-bonded forces
    (FragmentCalculator.calculate, fragment.py:50-68): same trajectory as the host integrator on the summed forces."""
    from ai2bmd_b200.md import BondedForceField, DeviceLangevin, Langevin
    from ai2bmd_b200.nonbonded import MMNonBondedCalculator
    fd, pm, prot_pos, prot_z, recipe, groups = _chig()
    n = len(prot_z)
    rowptr, col = exclusion_table(n, groups)
    q, sg, ep = synthetic_parameters(prot_z, seed=3)
    q *= 0.25                                                   # synthetic charges: keep the toy system gentle
    ff = BondedForceField(real_weights, fd, pm, recipe)
    nb = MMNonBondedCalculator(ff.engine)
    nb.set_parameters(q, sg, ep, rowptr, col)

    def force_fn(x):
        eb, fb = ff(x)
        en, fn = nb(x)
        return eb + en, fb + fn

    host = Langevin(prot_pos, prot_z, force_fn, dt_fs=0.5, friction_per_fs=0.0, seed=4)
    dev = DeviceLangevin(real_weights, fd, pm, recipe, prot_pos, prot_z, dt_fs=0.5, friction_per_fs=0.0, seed=4,
                         velocities=host.v.copy())
    dev.engine.set_nonbonded(q, sg, ep, rowptr, col)
    dev._eval()                                                 # forces of the start positions now include the term
    for _ in range(20):
        host.step()
    dev.run(20)
    x, v, step, _ = dev.state()
    assert step == 20 and np.abs(x - host.x).max() <= 2e-5 and np.abs(v - host.v).max() <= 2e-4
