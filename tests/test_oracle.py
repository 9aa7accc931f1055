"""The oracle against the reference's own model source (golden vectors) and against physics invariants."""
import numpy as np
import pytest
import torch

from oracle import visnet_ref as O


@pytest.fixture(scope="module")
def sd(real_weights):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in real_weights.items()}


@pytest.mark.parametrize("case", ["c1_ala", "chig", "dense44"])
def test_oracle_matches_reference_model_source(sd, reference_outputs, case):
    """Golden vectors = outputs of /root/reference/src/ViSNet/model (run via oracle/ref_shims.py, fp32 CPU)."""
    r = reference_outputs
    m = O.OracleViSNet(sd, torch.float32)
    e, f = m.energy_and_forces(r[f"{case}_z"], r[f"{case}_pos"], r[f"{case}_batch"])
    scale_f = max(1.0, np.abs(r[f"{case}_ref_f"]).max())
    # same arithmetic, same order of operations: fp32 round-off only
    assert np.abs(e.numpy() - r[f"{case}_ref_e"]).max() <= 4e-3 * max(1.0, np.abs(r[f"{case}_ref_e"]).max() / 2e4)
    assert np.abs(f.numpy() - r[f"{case}_ref_f"]).max() <= 2e-5 * scale_f


def test_fp64_anchor_matches_golden(sd, reference_outputs):
    r = reference_outputs
    m = O.OracleViSNet(sd, torch.float64)
    e, f = m.energy_and_forces(r["c1_ala_z"], r["c1_ala_pos"], r["c1_ala_batch"])
    assert np.abs(e.numpy() - r["c1_ala_e64"]).max() < 1e-9
    assert np.abs(f.numpy() - r["c1_ala_f64"]).max() < 1e-9


def test_neighbour_list_c_vs_numpy_and_golden(reference_outputs):
    r = reference_outputs
    for case in ("chig", "trpcage", "dense44"):
        s1, d1 = O.radius_graph_canonical(r[f"{case}_pos"], r[f"{case}_batch"])
        s2, d2 = O.radius_graph_numpy(r[f"{case}_pos"], r[f"{case}_batch"])
        assert (s1 == s2).all() and (d1 == d2).all()
        assert (s1 == r[f"{case}_slots"]).all() and (d1 == r[f"{case}_deg"]).all()


def test_neighbour_cap_keeps_lowest_indices(reference_outputs):
    r = reference_outputs
    slots, deg = r["dense44_slots"], r["dense44_deg"]
    pos = r["dense44_pos"]
    assert deg.max() == 32
    i = int(np.argmax(deg))
    d = np.linalg.norm(pos.astype(np.float64) - pos[i].astype(np.float64), axis=1)
    inside = np.flatnonzero(d < 5.0 - 1e-4)
    assert len(inside) > 32                       # more candidates than slots ...
    assert (slots[i] == inside[:32]).all()        # ... the first 32 by index survive, not the nearest
    assert i in slots[i] or i > slots[i].max()    # self-loop occupies a slot when its index is early enough


def test_neighbour_edge_cases():
    # single atom: only the self-loop
    s, d = O.radius_graph_canonical(np.zeros((1, 3), np.float32), np.zeros(1, np.int64))
    assert d[0] == 1 and s[0, 0] == 0
    # two graphs never see each other; exactly-at-cutoff pair is excluded (strict <)
    pos = np.array([[0, 0, 0], [5.0, 0, 0], [0, 0, 0], [4.999, 0, 0]], np.float32)
    s, d = O.radius_graph_canonical(pos, np.array([0, 0, 1, 1]))
    assert d.tolist() == [1, 1, 2, 2]
    assert s[2, :2].tolist() == [2, 3]


def _small(reference_outputs):
    r = reference_outputs
    return r["c1_ala_z"], r["c1_ala_pos"].astype(np.float64), r["c1_ala_batch"]


def test_force_is_minus_gradient_fd(sd, reference_outputs):
    z, pos, batch = _small(reference_outputs)
    m = O.OracleViSNet(sd, torch.float64)
    s, d = O.radius_graph_canonical(pos, batch)
    ei = torch.from_numpy(O.slots_to_edge_index(s, d))
    _, f = m.energy_and_forces(z, pos, batch, edge_index=ei)
    zt, bt = torch.as_tensor(z), torch.as_tensor(batch)
    h = 1e-5
    for a, c in [(0, 0), (7, 1), (13, 2), (21, 0)]:
        p = torch.tensor(pos)
        p[a, c] += h
        ep = m.forward(zt, p, bt, edge_index=ei).sum().item()
        p[a, c] -= 2 * h
        em = m.forward(zt, p, bt, edge_index=ei).sum().item()
        assert abs(-(ep - em) / (2 * h) - f[a, c].item()) < 2e-6 * max(1.0, abs(f[a, c].item()))


def test_se3_invariance(sd, reference_outputs):
    z, pos, batch = _small(reference_outputs)
    m = O.OracleViSNet(sd, torch.float64)
    e0, f0 = m.energy_and_forces(z, pos, batch)
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    e1, f1 = m.energy_and_forces(z, pos @ q.T + rng.normal(size=(1, 3)), batch)
    assert abs(e0.item() - e1.item()) < 1e-8
    assert np.abs(f0.numpy() @ q.T - f1.numpy()).max() < 1e-8


def test_batch_composition_independence(sd, chig):
    fd, _ = chig
    m = O.OracleViSNet(sd, torch.float64)
    e_all, f_all = m.energy_and_forces(fd.z, fd.pos, fd.batch)
    sub = fd[3:5]
    e_sub, f_sub = m.energy_and_forces(sub.z, sub.pos, sub.batch)
    a0, a1 = fd.start[3], fd.end[4]
    assert np.abs(e_all.numpy()[3:5] - e_sub.numpy()).max() < 1e-8
    assert np.abs(f_all.numpy()[a0:a1] - f_sub.numpy()).max() < 1e-9


def test_atom_permutation_equivariance(sd, reference_outputs):
    z, pos, batch = _small(reference_outputs)
    m = O.OracleViSNet(sd, torch.float64)
    e0, f0 = m.energy_and_forces(z, pos, batch)
    perm = np.random.default_rng(2).permutation(len(z))
    e1, f1 = m.energy_and_forces(z[perm], pos[perm], batch)
    assert abs(e0.item() - e1.item()) < 1e-8
    assert np.abs(f0.numpy()[perm] - f1.numpy()).max() < 1e-9
