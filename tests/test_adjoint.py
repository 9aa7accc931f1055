"""The hand-derived reverse sweep (blueprint of the CUDA force kernels) equals autograd in fp64."""
import numpy as np
import pytest
import torch

from oracle import visnet_ref as O
from oracle.adjoint_ref import AdjointViSNet


@pytest.mark.parametrize("weights", ["real", "random"])
def test_adjoint_equals_autograd(real_weights, chig, weights):
    fd, _ = chig
    sub = fd[0:4]
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in real_weights.items()} if weights == "real" \
        else O.random_state_dict(3)
    m = O.OracleViSNet(sd, torch.float64)
    s, d = O.radius_graph_canonical(sub.pos, sub.batch)
    ei = torch.from_numpy(O.slots_to_edge_index(s, d))
    e, f = m.energy_and_forces(sub.z, sub.pos, sub.batch, edge_index=ei)
    E, F, _, _ = AdjointViSNet(m).energy_and_forces(sub.z, sub.pos, sub.batch, ei)
    assert (E - e).abs().max().item() < 1e-9
    assert (F - f).abs().max().item() < 1e-11 * max(1.0, f.abs().max().item()) + 1e-12


def test_adjoint_dense_fragment_with_cap(real_weights, reference_outputs):
    r = reference_outputs
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in real_weights.items()}
    m = O.OracleViSNet(sd, torch.float64)
    ei = torch.from_numpy(O.slots_to_edge_index(r["dense44_slots"], r["dense44_deg"]))
    e, f = m.energy_and_forces(r["dense44_z"], r["dense44_pos"], r["dense44_batch"], edge_index=ei)
    E, F, _, _ = AdjointViSNet(m).energy_and_forces(r["dense44_z"], r["dense44_pos"], r["dense44_batch"], ei)
    assert (F - f).abs().max().item() < 1e-9 * max(1.0, f.abs().max().item())
