"""Minimal MD driver for the bonded (ViSNet fragment) potential -- the loop that drives the hot path.

The reference runs ASE's ``Langevin`` (``/root/reference/src/AIMD/simulator.py:96-137``: dt = 1 fs, 300 K,
friction 0.001 / fs, Maxwell-Boltzmann start) and asks ``FragmentCalculator`` for forces once per step.  ASE is
not available here, so this module restates that integrator (ASE 3.22 ``ase/md/langevin.py``, recalled; SURVEY
App. C) in numpy around :class:`BondedForceField`, which performs per step exactly what
``DLBondedCalculator.__call__`` does (``src/Calculators/bonded.py:102-123``) minus the cap-hydrogen LBFGS:

  protein positions -> fragment positions (cap hydrogens on the acceptor->removed ray, ``distancefrag.py:34-54``)
  -> engine (energies/forces of every fragment) -> signed reduction to whole-protein energy / forces.

friction = 0 gives velocity Verlet; its energy conservation is a physics check of the analytic forces
(``tests/test_md_gpu.py``).  Units follow ASE: eV, Angstrom, amu, time in Angstrom*sqrt(amu/eV).
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .fragment_data import FragmentData
from .pdbfrag import FragmentRecipe, ProteinMap

FS = 0.09822694788464063        # 1 fs in ASE time units
KB = 8.617330337217213e-05      # eV / K
MASSES = {1: 1.008, 6: 12.011, 7: 14.007, 8: 15.999, 16: 32.06}


class BondedForceField:
    """Whole-protein bonded energy/forces from the fragment batch (device-side reduction)."""

    def __init__(self, state_dict, frags: FragmentData, pm: ProteinMap, recipe: FragmentRecipe, device: int = 0):
        import torch
        self.torch = torch
        self.recipe, self.pm = recipe, pm
        self.engine = Engine(state_dict, device)
        self.engine.set_topology(frags.z, frags.batch, n_graphs=len(frags))
        self.engine.set_protein_map(pm.n_protein, pm.src_atom, pm.dst_atom, pm.sign, pm.frag_sign)
        dev = torch.device("cuda", device)
        self.pos_host = torch.empty((len(frags.z), 3), dtype=torch.float32).pin_memory()
        self.pos_dev = torch.empty((len(frags.z), 3), dtype=torch.float32, device=dev)
        self.ef_dev = torch.empty(3 * pm.n_protein + 1, dtype=torch.float32, device=dev)
        self.ef_host = torch.empty(3 * pm.n_protein + 1, dtype=torch.float32).pin_memory()
        self.stream = torch.cuda.current_stream(dev)

    def __call__(self, prot_pos: np.ndarray):
        """(E [eV], F [n_protein,3] eV/A) for the given protein coordinates."""
        self.pos_host.numpy()[:] = self.recipe.positions(prot_pos)
        self.pos_dev.copy_(self.pos_host, non_blocking=True)
        self.engine.forward_protein_device(self.pos_dev.data_ptr(), self.ef_dev.data_ptr(), self.stream.cuda_stream)
        self.ef_host.copy_(self.ef_dev, non_blocking=True)
        self.stream.synchronize()
        ef = self.ef_host.numpy()
        return float(ef[-1]), ef[:-1].reshape(-1, 3).astype(np.float64)


class Langevin:
    """ASE-style Langevin integrator (``fixcm=True``); ``friction=0`` reduces to velocity Verlet."""

    def __init__(self, positions, numbers, force_fn, dt_fs=1.0, temperature_K=300.0, friction_per_fs=0.001, seed=0):
        self.x = np.array(positions, dtype=np.float64)
        self.m = np.array([MASSES[int(z)] for z in numbers], dtype=np.float64)[:, None]
        self.force_fn = force_fn
        self.dt = dt_fs * FS
        self.T = temperature_K * KB
        self.fr = friction_per_fs / FS
        self.rng = np.random.default_rng(seed)
        # Maxwell-Boltzmann start (simulator.py:96), centre-of-mass motion removed
        self.v = self.rng.standard_normal(self.x.shape) * np.sqrt(self.T / self.m)
        self.v -= (self.v * self.m).sum(0) / self.m.sum()
        self.energy, self.f = force_fn(self.x)
        dt, fr = self.dt, self.fr
        sigma = np.sqrt(2 * self.T * fr / self.m)
        self.c1 = dt / 2.0 - dt * dt * fr / 8.0
        self.c2 = dt * fr / 2 - dt * dt * fr * fr / 8.0
        self.c3 = np.sqrt(dt) * sigma / 2.0 - dt ** 1.5 * fr * sigma / 8.0
        self.c5 = dt ** 1.5 * sigma / (2 * np.sqrt(3))
        self.c4 = fr / 2.0 * self.c5

    def kinetic_energy(self):
        return 0.5 * float((self.m * self.v * self.v).sum())

    def temperature(self):
        return 2.0 * self.kinetic_energy() / (3 * len(self.x)) / KB

    def step(self):
        xi = self.rng.standard_normal(self.x.shape) if self.fr > 0 else 0.0
        eta = self.rng.standard_normal(self.x.shape) if self.fr > 0 else 0.0
        self.v = self.v + (self.c1 * self.f / self.m - self.c2 * self.v + self.c3 * xi - self.c4 * eta)
        x_old = self.x
        self.x = self.x + self.dt * self.v + self.c5 * eta
        self.v = (self.x - x_old - self.c5 * eta) / self.dt
        self.energy, self.f = self.force_fn(self.x)
        self.v = self.v + (self.c1 * self.f / self.m - self.c2 * self.v + self.c3 * xi - self.c4 * eta)
        if self.fr > 0:
            self.v -= (self.v * self.m).sum(0) / self.m.sum()
        return self.energy

    def run(self, n_steps):
        for _ in range(n_steps):
            self.step()
