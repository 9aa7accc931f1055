"""VecLayerNorm(max_min) has a derivative kink where two channel norms tie for the maximum (reference
``src/ViSNet/model/utils.py:199-215``): the gradient is routed through the argmax channel, so an fp32-rounding-sized
change of the input flips it.  This is the model's property behind the one-fragment force jump described in DESIGN.md
section 2 (atom 11957 of the 512-fragment batch); the oracle reproduces it on the CPU."""
import numpy as np
import torch

from oracle.visnet_ref import vec_layer_norm_max_min


def _grad(vec, weight, probe):
    v = vec.clone().requires_grad_(True)
    (vec_layer_norm_max_min(v, weight) * probe).sum().backward()
    return v.grad


def test_gradient_jumps_across_an_argmax_tie():
    torch.manual_seed(0)
    D = 128
    vec = torch.randn(1, 3, D, dtype=torch.float64) * 0.3
    weight = torch.rand(D, dtype=torch.float64) + 0.5
    probe = torch.randn(1, 3, D, dtype=torch.float64)
    norms = vec.norm(dim=1)[0]
    a, b = torch.topk(norms, 2).indices.tolist()
    vec[0, :, b] *= norms[a] / norms[b]                      # channel b now ties with the maximum (channel a)
    eps = 1e-7                                               # one fp32 ulp of the norm, relative
    lo, hi = vec.clone(), vec.clone()
    lo[0, :, b] *= 1 - eps                                   # a wins
    hi[0, :, b] *= 1 + eps                                   # b wins
    out_lo, out_hi = vec_layer_norm_max_min(lo, weight), vec_layer_norm_max_min(hi, weight)
    assert (out_lo - out_hi).abs().max() < 1e-5              # the function itself is continuous ...
    g_lo, g_hi = _grad(lo, weight, probe), _grad(hi, weight, probe)
    jump = (g_lo - g_hi).abs().max().item()
    assert jump > 1e-2 * g_lo.abs().max().item()             # ... its gradient is not
    # the jump sits on the two tied channels, the way the argmax routing predicts
    per_channel = (g_lo - g_hi).abs().amax(dim=(0, 1))
    assert set(torch.topk(per_channel, 2).indices.tolist()) == {a, b}
    # away from the tie the same perturbation changes the gradient only by O(eps)
    far_lo, far_hi = lo.clone(), hi.clone()
    far_lo[0, :, b] *= 0.9; far_hi[0, :, b] *= 0.9
    assert (_grad(far_lo, weight, probe) - _grad(far_hi, weight, probe)).abs().max().item() < 1e-4 * g_lo.abs().max().item()
