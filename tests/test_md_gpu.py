"""The loop that drives the hot path: integrator + fragment-position update + device-side reduction."""
import numpy as np
import pytest

from ai2bmd_b200.calculator import ViSNetModel
from ai2bmd_b200.fixtures import load_fragments, load_protein
from ai2bmd_b200.md import BondedForceField, Langevin
from ai2bmd_b200.pdbfrag import single_graph

pytestmark = pytest.mark.gpu


def test_velocity_verlet_conserves_energy_unfragmented(real_weights, reference_outputs):
    """friction = 0 on a single conservative graph (22-atom dipeptide): E_pot + E_kin stays flat, i.e. the analytic
    forces are the gradient of the energy the engine reports (fp32 energies resolve ~1e-3 eV at -9e3 eV)."""
    r = reference_outputs
    z, pos = r["c1_ala_z"], r["c1_ala_pos"].astype(np.float64)
    model = ViSNetModel(real_weights, device="cuda:0")

    def force_fn(x):
        e, f = model.dl_potential_loader(single_graph(z, x.astype(np.float32)))
        return float(e[0, 0]), f.astype(np.float64)

    md = Langevin(pos, z, force_fn, dt_fs=0.5, temperature_K=300.0, friction_per_fs=0.0, seed=1)
    e0 = md.energy + md.kinetic_energy()
    tot = []
    for _ in range(300):
        md.step()
        tot.append(md.energy + md.kinetic_energy())
    tot = np.asarray(tot)
    assert np.abs(tot - e0).max() < 2e-2                 # no blow-up, no drift beyond fp32 energy resolution x O(10)
    assert abs(tot[-50:].mean() - tot[:50].mean()) < 1e-2
    assert md.kinetic_energy() > 0.1                      # the system is really moving (KE ~ 0.4-0.9 eV at 300 K)


def test_langevin_on_fragmented_chignolin_is_stable(real_weights):
    fd, pm = load_fragments("chig")
    prot_pos, prot_z, recipe = load_protein("chig")
    assert np.abs(recipe.positions(prot_pos) - fd.pos).max() < 1e-5
    ff = BondedForceField(real_weights, fd, pm, recipe)
    md = Langevin(prot_pos, prot_z, ff, dt_fs=1.0, temperature_K=300.0, friction_per_fs=0.001, seed=0)
    for _ in range(100):
        md.step()
    assert np.isfinite(md.x).all() and np.isfinite(md.energy)
    assert md.temperature() < 1.5 * 300 + 150             # reference guard: TemperatureRunawayError at 1.5*T0 (utils.py:153-155)
    assert np.abs(md.x - prot_pos).max() < 2.0            # nothing flew away in 100 fs


# Two runs of the same trajectory differ by the fp32 force rounding (atomics order, ~1e-6 eV/A), amplified by the
# dynamics: measured 3e-6 A after 40 steps of 1 fs.  Integrator bugs show up at the 1e-3 A level within a few steps.
X_TOL, V_TOL = 2e-5, 2e-4


def _chig_setup():
    fd, pm = load_fragments("chig")
    prot_pos, prot_z, recipe = load_protein("chig")
    return fd, pm, prot_pos, prot_z, recipe


def test_device_integrator_matches_host_integrator_verlet(real_weights):
    """friction = 0: the device-resident step (kick, cap-H placement, ViSNet, reduction, kick; one CUDA graph per
    step) follows the host restatement of the ASE integrator; differences come only from fp32 force rounding."""
    from ai2bmd_b200.md import DeviceLangevin
    fd, pm, prot_pos, prot_z, recipe = _chig_setup()
    ff = BondedForceField(real_weights, fd, pm, recipe)
    host = Langevin(prot_pos, prot_z, ff, dt_fs=1.0, temperature_K=300.0, friction_per_fs=0.0, seed=3)
    dev = DeviceLangevin(real_weights, fd, pm, recipe, prot_pos, prot_z, dt_fs=1.0, temperature_K=300.0,
                         friction_per_fs=0.0, seed=3, velocities=host.v.copy())
    assert abs(dev.energy - host.energy) <= 2e-2
    n = 40
    host_e = [host.step() for _ in range(n)]
    dev.run(n)
    x, v, step, hist = dev.state(n_hist=n)
    assert step == n
    assert np.abs(x - host.x).max() <= X_TOL and np.abs(v - host.v).max() <= V_TOL
    assert np.abs(hist - np.asarray(host_e)).max() <= 2e-2          # whole-protein energy, fp32 sum of O(2e4 eV) terms


def test_device_integrator_matches_host_integrator_langevin(real_weights):
    """friction > 0 with the device's own Philox stream: the host integrator fed with the host restatement of that
    stream (md.philox_normals) reproduces the trajectory, including the centre-of-mass correction."""
    from ai2bmd_b200.md import DeviceLangevin, philox_normals
    fd, pm, prot_pos, prot_z, recipe = _chig_setup()
    n_prot = len(prot_z)
    seed = 11

    def src(step):
        xi, eta = philox_normals(seed, step, 3 * n_prot)
        return xi.reshape(n_prot, 3), eta.reshape(n_prot, 3)

    ff = BondedForceField(real_weights, fd, pm, recipe)
    host = Langevin(prot_pos, prot_z, ff, dt_fs=1.0, temperature_K=300.0, friction_per_fs=0.01, seed=seed, normal_source=src)
    dev = DeviceLangevin(real_weights, fd, pm, recipe, prot_pos, prot_z, dt_fs=1.0, temperature_K=300.0,
                         friction_per_fs=0.01, seed=seed, velocities=host.v.copy())
    n = 25
    for _ in range(n):
        host.step()
    dev.run(10)
    dev.run(n - 10)                                                  # the step counter carries across calls
    x, v, step, _ = dev.state()
    assert step == n
    assert np.abs(x - host.x).max() <= X_TOL and np.abs(v - host.v).max() <= V_TOL
    m = dev.masses[:, None]
    assert np.abs((m * v).sum(0)).max() <= 1e-9                      # centre of mass at rest
    com0 = (m * prot_pos).sum(0) / m.sum()                           # ... and where it started (ASE fix_com restores the position too)
    assert np.abs((m * x).sum(0) / m.sum() - com0).max() <= 1e-9
    assert np.abs((host.m * host.x).sum(0) / host.m.sum() - com0).max() <= 1e-9
    # externally supplied normals take the same path
    pool = np.stack([np.stack(src(s)) for s in range(n, n + 5)])     # [5, 2, n, 3]
    dev.set_normals(pool)
    # (pool rows are indexed by step % pool_steps: steps 25..29 read rows 0..4)
    for _ in range(5):
        host.step()
    dev.run(5)
    x2, v2, step2, _ = dev.state()
    assert step2 == n + 5 and np.abs(x2 - host.x).max() <= X_TOL


def test_device_integrator_phases_equal_whole_step(real_weights):
    """kick1 / eval / kick2 called one by one (the multi-GPU path, where the caller all-reduces between eval and
    kick2) give the same state as the captured whole-step graph."""
    import torch
    from ai2bmd_b200.md import DeviceLangevin
    fd, pm, prot_pos, prot_z, recipe = _chig_setup()
    a = DeviceLangevin(real_weights, fd, pm, recipe, prot_pos, prot_z, friction_per_fs=0.002, seed=5)
    b = DeviceLangevin(real_weights, fd, pm, recipe, prot_pos, prot_z, friction_per_fs=0.002, seed=5)
    a.run(8)
    sp = torch.cuda.current_stream().cuda_stream
    for _ in range(8):
        b.engine.md_kick1(sp)
        b.engine.md_eval(sp)
        b.engine.md_kick2(sp)
    xa, va, sa, _ = a.state()
    xb, vb, sb, _ = b.state()
    assert sa == sb == 8 and np.abs(xa - xb).max() <= X_TOL and np.abs(va - vb).max() <= V_TOL
    assert 50.0 < a.temperature() < 600.0
