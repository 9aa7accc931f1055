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
