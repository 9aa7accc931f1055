"""CPU oracle for the ViSNet energy/force hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``ai2bmd_b200``) never routes through here
and has no CPU fallback.

What it is: a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference model

* ``src/ViSNet/model/visnet.py:135-166``        ViSNet.forward (energy, autograd force)
* ``src/ViSNet/model/visnet_block.py:103-142``  ViSNetBlock.forward
* ``src/ViSNet/model/visnet_block.py:237-295``  ViS_MP.forward / message / edge_update
* ``src/ViSNet/model/utils.py:10-57,200-341``   cutoff, exp-normal RBF, VecLayerNorm(max_min),
                                                 Distance, NeighborEmbedding, EdgeEmbedding
* ``src/ViSNet/model/output_modules.py:52-62,136-140``  gated equivariant head
* ``src/ViSNet/model/priors.py:86-87``          Atomref

with the un-vendored third-party graph ops replaced by their documented semantics:
``torch_scatter.scatter(reduce='add')`` -> ``index_add_``; PyG ``propagate`` ->
``index_select`` on ``edge_index[0]`` (source j) / ``edge_index[1]`` (target i);
``torch_cluster.radius_graph`` -> :func:`radius_graph_canonical` (rule below).

Parity pinning status: the reference ships no tests or golden vectors for this
path and its third-party graph ops are absent from this image, so the oracle is
pinned two ways (see ``tests/golden/make_golden.py``): (1) the reference's *own*
model source files are imported from ``/root/reference`` with thin shims for the
missing third-party packages and must reproduce this oracle's energies/forces;
(2) physics invariants (force = -dE/dr by fp64 finite differences, SE(3)
equivariance, batch-composition independence).  The neighbour-list rule of
``torch_cluster`` itself stays *recalled, unpinned*.

Canonical neighbour list (bit-exact contract between oracle and CUDA engine):
for every target atom i, scan sources j of the same graph in ascending index
order; accept when d2 = fma(dz,dz, fma(dy,dy, dx*dx)) < cutoff^2 with
dx = pos[j].x - pos[i].x etc., all fp32, strict '<'; self (j == i) is accepted;
stop after ``max_num_neighbors`` hits.  Edge list = target-major, source
ascending; ``edge_index[0] = j`` (source), ``edge_index[1] = i`` (target).
"""
from __future__ import annotations

import ctypes
import math
import os
import re
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------------------
# hyper-parameters of the shipped checkpoints (read back from the ckpt)
# --------------------------------------------------------------------------------------
HP = dict(D=128, L=6, H=8, R=32, cutoff=5.0, max_nbr=32, max_z=100)


# --------------------------------------------------------------------------------------
# neighbour list
# --------------------------------------------------------------------------------------
_radius_lib = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/radius_graph.c -> oracle/_build/libradius_ref.so (gcc, seconds)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "libradius_ref.so")
    src = os.path.join(_HERE, "radius_graph.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src, "-lm"]
        )
    return so


def _load_radius_lib():
    global _radius_lib
    if _radius_lib is None:
        lib = ctypes.CDLL(build_c_oracle())
        lib.radius_graph_ref.restype = ctypes.c_int64
        lib.radius_graph_ref.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
            ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
        ]
        _radius_lib = lib
    return _radius_lib


def radius_graph_canonical(pos: np.ndarray, batch: np.ndarray, cutoff: float = 5.0,
                           max_nbr: int = 32) -> Tuple[np.ndarray, np.ndarray]:
    """Canonical neighbour list via the C restatement (exact fmaf).  Returns (slots[N,K], deg[N])."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    batch = np.ascontiguousarray(batch, dtype=np.int64)
    n = pos.shape[0]
    slots = np.full((n, max_nbr), -1, dtype=np.int32)
    deg = np.zeros((n,), dtype=np.int32)
    lib = _load_radius_lib()
    lib.radius_graph_ref(pos.ctypes.data, batch.ctypes.data, n, np.float32(cutoff),
                         max_nbr, slots.ctypes.data, deg.ctypes.data)
    return slots, deg


def radius_graph_numpy(pos: np.ndarray, batch: np.ndarray, cutoff: float = 5.0,
                       max_nbr: int = 32) -> Tuple[np.ndarray, np.ndarray]:
    """Same rule in numpy; the fp32 FMA chain is emulated through fp64 (cross-check only)."""
    pos = np.asarray(pos, dtype=np.float32)
    batch = np.asarray(batch, dtype=np.int64)
    n = pos.shape[0]
    slots = np.full((n, max_nbr), -1, dtype=np.int32)
    deg = np.zeros((n,), dtype=np.int32)
    r2 = np.float32(cutoff) * np.float32(cutoff)
    for g in np.unique(batch):
        idx = np.flatnonzero(batch == g)
        p = pos[idx]
        d = p[None, :, :] - p[:, None, :]                     # [i, j] = pos[j] - pos[i], fp32
        dx, dy, dz = (d[..., k].astype(np.float64) for k in range(3))
        c0 = (d[..., 0] * d[..., 0]).astype(np.float32)       # fp32 product, rounded
        c1 = (dy * dy + c0.astype(np.float64)).astype(np.float32)
        c2 = (dz * dz + c1.astype(np.float64)).astype(np.float32)
        ok = c2 < r2
        for a, i in enumerate(idx):
            js = idx[np.flatnonzero(ok[a])][:max_nbr]
            deg[i] = len(js)
            slots[i, : len(js)] = js
    return slots, deg


def slots_to_edge_index(slots: np.ndarray, deg: np.ndarray) -> np.ndarray:
    """Compact target-major edge list: row 0 = source j, row 1 = target i."""
    n, k = slots.shape
    mask = np.arange(k)[None, :] < deg[:, None]
    tgt = np.repeat(np.arange(n, dtype=np.int64), deg)
    src = slots[mask].astype(np.int64)
    return np.stack([src, tgt])


# --------------------------------------------------------------------------------------
# checkpoint handling
# --------------------------------------------------------------------------------------
def load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """Read a shipped Lightning checkpoint (``visnet.py:74-87``: strip the leading ``model.``)
    or an ``.npz`` weight fixture written by :func:`save_weights_npz`."""
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k].copy()) for k in z.files}
    ck = torch.load(path, map_location="cpu", weights_only=True)
    sd = {re.sub(r"^model\.", "", k): v.float() for k, v in ck["state_dict"].items()}
    hp = ck.get("hyper_parameters", {})
    if hp:
        assert hp["embedding_dimension"] == HP["D"] and hp["num_layers"] == HP["L"]
        assert hp["num_heads"] == HP["H"] and hp["num_rbf"] == HP["R"] and hp["lmax"] == 1
        assert hp["vecnorm_type"] == "max_min" and hp["rbf_type"] == "expnorm"
        assert float(hp["cutoff"]) == HP["cutoff"] and hp["max_num_neighbors"] == HP["max_nbr"]
    return sd


def save_weights_npz(sd: Dict[str, torch.Tensor], path: str) -> None:
    np.savez(path, **{k: v.detach().cpu().numpy() for k, v in sd.items()})


def random_state_dict(seed: int = 0, atomref: bool = True) -> Dict[str, torch.Tensor]:
    """Random-init weights of the shipped architecture (key set == checkpoint key set)."""
    g = torch.Generator().manual_seed(seed)
    D, L, R = HP["D"], HP["L"], HP["R"]

    def lin(o, i, scale=1.0):
        return (torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6.0 / (i + o)) * scale

    def vec(o, s=0.1):
        return (torch.rand(o, generator=g) * 2 - 1) * s

    sd: Dict[str, torch.Tensor] = {"mean": torch.tensor(0.0), "std": torch.tensor(1.0)}
    rm = "representation_model."
    sd[rm + "embedding.weight"] = torch.randn(100, D, generator=g)
    start = math.exp(-HP["cutoff"])
    sd[rm + "distance_expansion.means"] = torch.linspace(start, 1, R)
    sd[rm + "distance_expansion.betas"] = torch.full((R,), (2 / R * (1 - start)) ** -2)
    sd[rm + "neighbor_embedding.embedding.weight"] = torch.randn(100, D, generator=g)
    sd[rm + "neighbor_embedding.distance_proj.weight"] = lin(D, R)
    sd[rm + "neighbor_embedding.distance_proj.bias"] = vec(D)
    sd[rm + "neighbor_embedding.combine.weight"] = lin(D, 2 * D)
    sd[rm + "neighbor_embedding.combine.bias"] = vec(D)
    sd[rm + "edge_embedding.edge_proj.weight"] = lin(D, R)
    sd[rm + "edge_embedding.edge_proj.bias"] = vec(D)
    for l in range(L):
        p = rm + f"vis_mp_layers.{l}."
        sd[p + "layernorm.weight"] = 1 + vec(D)
        sd[p + "layernorm.bias"] = vec(D)
        sd[p + "vec_layernorm.weight"] = torch.ones(D)
        sd[p + "vec_proj.weight"] = lin(3 * D, D)
        for n in ("q", "k", "v", "dk", "dv"):
            sd[p + f"{n}_proj.weight"] = lin(D, D)
            sd[p + f"{n}_proj.bias"] = vec(D)
        sd[p + "s_proj.weight"] = lin(2 * D, D)
        sd[p + "s_proj.bias"] = vec(2 * D)
        if l < L - 1:
            sd[p + "f_proj.weight"] = lin(D, D)
            sd[p + "f_proj.bias"] = vec(D)
            sd[p + "w_src_proj.weight"] = lin(D, D)
            sd[p + "w_trg_proj.weight"] = lin(D, D)
        sd[p + "o_proj.weight"] = lin(3 * D, D)
        sd[p + "o_proj.bias"] = vec(3 * D)
    sd[rm + "out_norm.weight"] = 1 + vec(D)
    sd[rm + "out_norm.bias"] = vec(D)
    sd[rm + "vec_out_norm.weight"] = torch.ones(D)
    o = "output_model.output_network."
    sd[o + "0.vec1_proj.weight"] = lin(D, D)
    sd[o + "0.vec2_proj.weight"] = lin(D // 2, D)
    sd[o + "0.update_net.0.weight"] = lin(D, 2 * D)
    sd[o + "0.update_net.0.bias"] = vec(D)
    sd[o + "0.update_net.2.weight"] = lin(D, D)
    sd[o + "0.update_net.2.bias"] = vec(D)
    sd[o + "1.vec1_proj.weight"] = lin(D // 2, D // 2)
    sd[o + "1.vec2_proj.weight"] = lin(1, D // 2)
    sd[o + "1.update_net.0.weight"] = lin(D // 2, D)
    sd[o + "1.update_net.0.bias"] = vec(D // 2)
    sd[o + "1.update_net.2.weight"] = lin(2, D // 2)
    sd[o + "1.update_net.2.bias"] = vec(2)
    ar = torch.zeros(100, 1)
    if atomref:
        for zz, e in ((1, -13.554), (6, -1027.537), (7, -1484.846), (8, -2041.727), (16, -10830.209)):
            ar[zz, 0] = e
    sd["prior_model.initial_atomref"] = ar.clone()
    sd["prior_model.atomref.weight"] = ar.clone()
    return sd


# --------------------------------------------------------------------------------------
# model pieces (functional)
# --------------------------------------------------------------------------------------
def cosine_cutoff(r: torch.Tensor, cutoff: float) -> torch.Tensor:
    """utils.py:16-19"""
    c = 0.5 * (torch.cos(r * math.pi / cutoff) + 1.0)
    return c * (r < cutoff).to(r.dtype)


def silu(x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.silu(x)


def layer_norm(x, w, b):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def vec_layer_norm_max_min(vec: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """utils.py:200-228 (max_min).  vec: [N,3,D]."""
    dist = torch.norm(vec, dim=1, keepdim=True)
    if (dist == 0).all():
        return torch.zeros_like(vec) * weight.view(1, 1, -1)
    dist = dist.clamp(min=1e-12)
    direct = vec / dist
    max_val, _ = torch.max(dist, dim=-1)
    min_val, _ = torch.min(dist, dim=-1)
    delta = (max_val - min_val).view(-1)
    delta = torch.where(delta == 0, torch.ones_like(delta), delta)
    dist = (dist - min_val.view(-1, 1, 1)) / delta.view(-1, 1, 1)
    return torch.relu(dist) * direct * weight.view(1, 1, -1)


class OracleViSNet:
    """Functional restatement; weights are held in ``dtype`` (fp32 default, fp64 for anchors)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu"):
        """``device="cuda"`` runs the same eager PyTorch ops on a GPU (the PyTorch-eager proxy for
        "the reference on a modern GPU"); the neighbour list is still the canonical CPU rule."""
        self.dtype = dtype
        self.device = torch.device(device)
        self.sd = {k: v.to(dtype).to(self.device) for k, v in state_dict.items()}
        self.cutoff = HP["cutoff"]
        self.D, self.L, self.H = HP["D"], HP["L"], HP["H"]

    # -- graph + geometry (utils.py:259-276, visnet_block.py:111-117) --------------------
    def geometry(self, pos: torch.Tensor, edge_index: torch.Tensor):
        src, dst = edge_index[0], edge_index[1]
        edge_vec = pos[src] - pos[dst]
        mask = src != dst
        safe = torch.where(mask.unsqueeze(1), edge_vec, torch.ones_like(edge_vec))
        nrm = torch.norm(safe, dim=-1)
        r = torch.where(mask, nrm, torch.zeros_like(nrm))
        d = torch.where(mask.unsqueeze(1), edge_vec / nrm.unsqueeze(1), torch.zeros_like(edge_vec))
        return r, d, mask

    def rbf(self, r: torch.Tensor) -> torch.Tensor:
        """utils.py:53-57, alpha = 5/cutoff."""
        p = "representation_model.distance_expansion."
        means, betas = self.sd[p + "means"], self.sd[p + "betas"]
        alpha = 5.0 / self.cutoff
        rr = r.unsqueeze(-1)
        return cosine_cutoff(rr, self.cutoff) * torch.exp(-betas * (torch.exp(alpha * (-rr)) - means) ** 2)

    # -- forward with optional intermediate capture --------------------------------------
    def forward(self, z: torch.Tensor, pos: torch.Tensor, batch: torch.Tensor,
                edge_index: Optional[torch.Tensor] = None, n_graphs: Optional[int] = None,
                cap: Optional[dict] = None):
        """Return per-graph energies E[G]; ``pos`` may require grad.  ``cap`` (dict) collects
        named intermediates (kept in the autograd graph) for stage-by-stage checks."""
        sd, D, H = self.sd, self.D, self.H
        rm = "representation_model."
        n = z.shape[0]
        if edge_index is None:
            slots, deg = radius_graph_canonical(pos.detach().cpu().numpy().astype(np.float32),
                                                batch.cpu().numpy(), self.cutoff, HP["max_nbr"])
            edge_index = torch.from_numpy(slots_to_edge_index(slots, deg))
        edge_index = edge_index.to(pos.device)
        src, dst = edge_index[0], edge_index[1]
        pos = pos.to(self.dtype)
        dev = pos.device

        def keep(name, t):
            if cap is not None:
                if t.requires_grad:
                    t.retain_grad()
                cap[name] = t
            return t

        r, d, mask = self.geometry(pos, edge_index)
        rbf = self.rbf(r)
        cut = cosine_cutoff(r, self.cutoff)
        keep("r", r), keep("d", d), keep("rbf", rbf)

        # embeddings (visnet_block.py:110; utils.py:296-317)
        x0 = sd[rm + "embedding.weight"][z]
        p = rm + "neighbor_embedding."
        w_e = (rbf @ sd[p + "distance_proj.weight"].T + sd[p + "distance_proj.bias"]) * cut.unsqueeze(1)
        msg = w_e * sd[p + "embedding.weight"][z][src]
        msg = msg * mask.unsqueeze(1).to(msg.dtype)          # self-loops removed (utils.py:298-302)
        agg = torch.zeros(n, D, dtype=self.dtype, device=dev).index_add_(0, dst, msg)
        x = torch.cat([x0, agg], dim=1) @ sd[p + "combine.weight"].T + sd[p + "combine.bias"]
        vec = torch.zeros(n, 3, D, dtype=self.dtype, device=dev)
        p = rm + "edge_embedding."
        f = (x[dst] + x[src]) * (rbf @ sd[p + "edge_proj.weight"].T + sd[p + "edge_proj.bias"])
        keep("x_emb", x), keep("f_emb", f)

        for l in range(self.L):
            last = l == self.L - 1
            p = rm + f"vis_mp_layers.{l}."
            keep(f"x_in{l}", x), keep(f"vec_in{l}", vec), keep(f"f_in{l}", f)
            xn = layer_norm(x, sd[p + "layernorm.weight"], sd[p + "layernorm.bias"])
            vn = vec_layer_norm_max_min(vec, sd[p + "vec_layernorm.weight"])
            q = (xn @ sd[p + "q_proj.weight"].T + sd[p + "q_proj.bias"]).view(n, H, D // H)
            k = (xn @ sd[p + "k_proj.weight"].T + sd[p + "k_proj.bias"]).view(n, H, D // H)
            v = (xn @ sd[p + "v_proj.weight"].T + sd[p + "v_proj.bias"]).view(n, H, D // H)
            dk = silu(f @ sd[p + "dk_proj.weight"].T + sd[p + "dk_proj.bias"]).view(-1, H, D // H)
            dv = silu(f @ sd[p + "dv_proj.weight"].T + sd[p + "dv_proj.bias"]).view(-1, H, D // H)
            vp = vn @ sd[p + "vec_proj.weight"].T
            vec1, vec2, vec3 = torch.split(vp, D, dim=-1)
            vec_dot = (vec1 * vec2).sum(dim=1)
            keep(f"xn{l}", xn), keep(f"vn{l}", vn), keep(f"vec_dot{l}", vec_dot)
            # message (visnet_block.py:276-288): _i = target (edge_index[1]), _j = source
            attn = (q[dst] * k[src] * dk).sum(dim=-1)
            attn = silu(attn) * cut.unsqueeze(1)
            m = (v[src] * dv * attn.unsqueeze(2)).reshape(-1, D)
            s = silu(m @ sd[p + "s_proj.weight"].T + sd[p + "s_proj.bias"])
            s1, s2 = torch.split(s, D, dim=1)
            vmsg = vn[src] * s1.unsqueeze(1) + s2.unsqueeze(1) * d.unsqueeze(2)
            xa = torch.zeros(n, D, dtype=self.dtype, device=dev).index_add_(0, dst, m)
            va = torch.zeros(n, 3, D, dtype=self.dtype, device=dev).index_add_(0, dst, vmsg)
            keep(f"m{l}", m), keep(f"xa{l}", xa), keep(f"va{l}", va)
            if not last:
                # edge_update (visnet_block.py:290-295)
                t = vn @ sd[p + "w_trg_proj.weight"].T
                u = vn @ sd[p + "w_src_proj.weight"].T
                ti, uj = t[dst], u[src]
                w1 = ti - (ti * d.unsqueeze(2)).sum(dim=1, keepdim=True) * d.unsqueeze(2)
                nd = -d
                w2 = uj - (uj * nd.unsqueeze(2)).sum(dim=1, keepdim=True) * nd.unsqueeze(2)
                w_dot = (w1 * w2).sum(dim=1)
                df = silu(f @ sd[p + "f_proj.weight"].T + sd[p + "f_proj.bias"]) * w_dot
            o = xa @ sd[p + "o_proj.weight"].T + sd[p + "o_proj.bias"]
            o1, o2, o3 = torch.split(o, D, dim=1)
            dx = vec_dot * o2 + o3
            dvec = vec3 * o1.unsqueeze(1) + va
            x = x + dx
            vec = vec + dvec
            if not last:
                f = f + df

        keep("x_out", x), keep("vec_out", vec)
        x = layer_norm(x, sd[rm + "out_norm.weight"], sd[rm + "out_norm.bias"])
        vec = vec_layer_norm_max_min(vec, sd[rm + "vec_out_norm.weight"])
        keep("x_normed", x), keep("vec_normed", vec)

        # gated equivariant head (output_modules.py:52-62,136-140)
        for b in range(2):
            p = f"output_model.output_network.{b}."
            oc = sd[p + "vec2_proj.weight"].shape[0]
            vec1 = torch.norm(vec @ sd[p + "vec1_proj.weight"].T, dim=-2)
            vec2 = vec @ sd[p + "vec2_proj.weight"].T
            h = torch.cat([x, vec1], dim=-1)
            h = silu(h @ sd[p + "update_net.0.weight"].T + sd[p + "update_net.0.bias"])
            h = h @ sd[p + "update_net.2.weight"].T + sd[p + "update_net.2.bias"]
            x, g = torch.split(h, oc, dim=-1)
            vec = g.unsqueeze(1) * vec2
            if b == 0:
                x = silu(x)
        x = x + vec.sum() * 0
        x = x * sd["std"]
        x = x + sd["prior_model.atomref.weight"][z]
        keep("e_atom", x)
        g = int(batch.max().item()) + 1 if n_graphs is None else n_graphs
        out = torch.zeros(g, 1, dtype=self.dtype, device=dev).index_add_(0, batch, x)
        out = out + sd["mean"]
        return out

    def energy_and_forces(self, z, pos, batch, edge_index=None, cap=None):
        """visnet.py:135-166: E[G,1], F[N,3] = -dE/dpos (autograd)."""
        z = torch.as_tensor(z, dtype=torch.long).to(self.device)
        batch = torch.as_tensor(batch, dtype=torch.long).to(self.device)
        pos = torch.as_tensor(pos).to(self.dtype).to(self.device).clone().requires_grad_(True)
        with torch.enable_grad():
            out = self.forward(z, pos, batch, edge_index=edge_index, cap=cap)
            (dy,) = torch.autograd.grad([out], [pos], grad_outputs=[torch.ones_like(out)],
                                        retain_graph=cap is not None)
        if cap is not None:
            cap["pos"] = pos
            cap["E"] = out
        return out.detach(), (-dy).detach()


class OracleCalculatorModel:
    """Restatement of ``ViSNetModel`` (``src/Calculators/visnet_calculator.py:22-63``) over the oracle:
    ``dl_potential_loader(FragmentData) -> (e[G,1] f32, f[N,3] f32)`` as numpy arrays."""

    def __init__(self, state_dict, dtype=torch.float32, device="cpu"):
        self.model = OracleViSNet(state_dict, dtype, device)

    def dl_potential_loader(self, frag):
        e, f = self.model.energy_and_forces(torch.from_numpy(np.asarray(frag.z, dtype=np.int64)),
                                            torch.from_numpy(np.asarray(frag.pos, dtype=np.float32)),
                                            torch.from_numpy(np.asarray(frag.batch, dtype=np.int64)))
        return (e.reshape(-1, 1).to(torch.float32).cpu().numpy(), f.reshape(-1, 3).to(torch.float32).cpu().numpy())
