#!/usr/bin/env python
"""Stage-by-stage parity of the CUDA engine against the hand-adjoint oracle (GPU box diagnostic).

    python tools/stage_check.py [--weights real|random] [--frags chig] [--out gpurun_out/stage_check.txt]

Runs the evaluation one launch at a time (``vb_debug_run``), reads the engine's internal buffers after
each stage and compares them with the tensors of ``oracle/adjoint_ref.py`` (fp64).  The first stage whose
relative error jumps is where a kernel bug lives.  Test infrastructure; not part of the product path.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from oracle import visnet_ref as O                      # noqa: E402
from oracle.adjoint_ref import AdjointViSNet            # noqa: E402
from ai2bmd_b200.engine import Engine                   # noqa: E402

D, L = 128, 6


def stage_report(frags="chig", weights="real", max_frags=0, opts=""):
    """Run the evaluation one launch at a time and compare every buffer a stage produces with the fp64 hand-adjoint
    oracle.  Returns (lines, worst) where worst = [(stage, what, rel)] of the comparisons."""
    g = np.load(os.path.join(ROOT, "tests", "golden", f"fragments_{frags}.npz"))
    z, pos, batch = g["z"], g["pos"], g["batch"]
    if max_frags:
        keep = batch < max_frags
        z, pos, batch = z[keep], pos[keep], batch[keep]
    sd = O.load_state_dict(os.path.join(ROOT, "tests", "golden", "weights_2ef43f29.npz")) if weights == "real" \
        else O.random_state_dict(int(weights) if str(weights).isdigit() else 0)
    slots, deg = O.radius_graph_canonical(pos, batch)
    ei = torch.from_numpy(O.slots_to_edge_index(slots, deg))
    E, N = ei.shape[1], len(z)
    adj = AdjointViSNet(O.OracleViSNet(sd, torch.float64))
    Eo, Fo, S, B = adj.energy_and_forces(z, pos, batch, ei)
    S = {k: v.numpy() for k, v in S.items()}
    B = {k: v.numpy() for k, v in B.items()}

    eng = Engine({k: v.numpy() for k, v in sd.items()}, 0)
    eng.set_topology(z, batch)
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    names = eng.stage_names()
    dpos = torch.from_numpy(pos).cuda()
    lines, worst = [], []

    def report(stage, what, got, ref):
        ref = np.asarray(ref, dtype=np.float64)
        got = np.asarray(got, dtype=np.float64).reshape(ref.shape)
        err = np.abs(got - ref).max() if ref.size else 0.0
        mag = np.abs(ref).max() if ref.size else 0.0
        rel = err / mag if mag > 0 else err
        flag = "  <<<<<<" if (not np.isfinite(err)) or rel > 2e-3 else ""
        lines.append(f"{stage:16s} {what:14s} maxabs {err:10.3e}  ref {mag:10.3e}  rel {rel:9.2e}{flag}")
        worst.append((stage, what, float(rel) if np.isfinite(err) else float("inf")))

    def flagline(stage, text, ok):
        lines.append(f"{stage:16s} {text}: {bool(ok)}")
        worst.append((stage, text, 0.0 if ok else float("inf")))

    def rd(name, layer, shape):
        return eng.debug_read(name, layer, shape)

    def cat(*xs):
        return np.concatenate(xs, axis=-1)

    def node_fwd_checks(st, k):
        if k >= 1:
            report(st, f"x_in{k}", rd("X", k, (N, D)), S[f"x_in{k}"] if k < L else S["x_out"])
            report(st, f"vec_in{k}", rd("V", k, (N, 3, D)), S[f"vec_in{k}"] if k < L else S["vec_out"])
            report(st, f"o{k-1}", rd("O", k - 1, (N, 3 * D)), S[f"o{k-1}"])
        if k < L:
            report(st, "vn", rd("VN", k, (N, 3, D)), S[f"vn{k}"])
            report(st, "qkv", rd("QKV", k, (N, 3 * D)), cat(S[f"q{k}"], S[f"k{k}"], S[f"v{k}"]))
            report(st, "v123", rd("V123", k, (N, 3, 3 * D)), cat(S[f"v1{k}"], S[f"v2{k}"], S[f"v3{k}"]))
            report(st, "vdot", rd("VDOT", k, (N, D)), S[f"vdot{k}"])
            if k < L - 1:
                report(st, "tu", rd("TU", k, (N, 3, 2 * D)), cat(S[f"t{k}"], S[f"u{k}"]))

    def node_bwd_checks(st, k):
        if k <= L - 1:
            report(st, f"gx_in{k}", rd("GX", 0, (N, D)), B[f"gx_in{k}"])
            report(st, f"gvec_in{k}", rd("GVEC", 0, (N, 3, D)), B[f"gvec_in{k}"])
        if k >= 1:
            report(st, f"g_xa{k-1}", rd("GXA", 0, (N, D)), B[f"g_xa{k-1}"])

    def edge_bwd_checks(st, l, suffix=""):
        report(st, f"gf_in{l}", rd("GF", 0, (N * 32, D))[:E], B[f"gf_in{l}"])
        report(st, "g_qkv", rd("GQKV" + suffix, 0, (N, 3 * D)), cat(B[f"g_q{l}"], B[f"g_k{l}"], B[f"g_v{l}"]))
        report(st, "g_vn_msg", rd("GVNMSG" + suffix, 0, (N, 3, D)), B[f"g_vn_msg{l}"])
        if l < L - 1:
            report(st, "g_tu", rd("GTU" + suffix, 0, (N, 3, 2 * D)), cat(B[f"g_t{l}"], B[f"g_u{l}"]))

    for si, st in enumerate(names):
        eng.debug_run(dpos.data_ptr(), si + 1)
        if st == "nbr_build":
            s2, d2 = eng.get_edges()
            flagline(st, "neighbour list identical", (s2 == slots).all() and (d2 == deg).all())
        elif st == "rowptr_scan":
            rp = rd("rowptr", 0, (N + 1,)).view(np.int32)
            flagline(st, "rowptr ok", (rp == np.concatenate([[0], np.cumsum(deg)])).all())
        elif st == "edge_geom":
            ge = rd("geom", 0, (N * 32, 8))[:E]
            report(st, "r", ge[:, 0], S["r"]); report(st, "C", ge[:, 1], S["C"]); report(st, "d", ge[:, 2:5], S["d"])
            report(st, "rbf", rd("rbf", 0, (N * 32, 32))[:E], S["rbf"])
            es, ed = rd("esrc", 0, (N * 32,)).view(np.int32)[:E], rd("edst", 0, (N * 32,)).view(np.int32)[:E]
            flagline(st, "edge_index ok", (es == ei[0].numpy()).all() and (ed == ei[1].numpy()).all())
        elif st == "embed_node":
            report(st, "x_emb", rd("X", 0, (N, D)), S["x_emb"])
        elif st == "embed_edge":
            report(st, "f0", rd("F", 0, (N * 32, D))[:E], S["f_in0"])
        elif st.startswith("node_fwd"):
            node_fwd_checks(st, int(st[8:]))
        elif st.startswith("oproj"):            # tensor-core node stage: O of the previous layer
            k = int(st[5:])
            report(st, f"o{k-1}", rd("O", k - 1, (N, 3 * D)), S[f"o{k-1}"])
        elif st.startswith("norm"):
            k = int(st[4:])
            if k >= 1:
                report(st, f"x_in{k}", rd("X", k, (N, D)), S[f"x_in{k}"] if k < L else S["x_out"])
                report(st, f"vec_in{k}", rd("V", k, (N, 3, D)), S[f"vec_in{k}"] if k < L else S["vec_out"])
                report(st, f"vdot{k-1}", rd("VDOT", k - 1, (N, D)), S[f"vdot{k-1}"])
            if k < L:
                report(st, "vn", rd("VN", k, (N, 3, D)), S[f"vn{k}"])
        elif st.startswith("proj"):
            k = int(st[4:])
            report(st, "qkv", rd("QKV", k, (N, 3 * D)), cat(S[f"q{k}"], S[f"k{k}"], S[f"v{k}"]))
            report(st, "v123", rd("V123", k, (N, 3, 3 * D)), cat(S[f"v1{k}"], S[f"v2{k}"], S[f"v3{k}"]))
            if k < L - 1:
                report(st, "tu", rd("TU", k, (N, 3, 2 * D)), cat(S[f"t{k}"], S[f"u{k}"]))
        elif st.startswith("bnorm"):
            k = int(st[5:])
            if k <= L - 1:
                report(st, f"gx_in{k}", rd("GX", 0, (N, D)), B[f"gx_in{k}"])
                report(st, f"gvec_in{k}", rd("GVEC", 0, (N, 3, D)), B[f"gvec_in{k}"])
        elif st.startswith("bwdB"):
            k = int(st[4:])
            report(st, f"g_xa{k-1}", rd("GXA3", 0, (3, N, D)).sum(0), B[f"g_xa{k-1}"])
        elif st.startswith("bwdA"):
            pass                                 # K-chunk partials only; their sums are checked at bnorm
        elif st.startswith("edge_fwd"):
            l = int(st[8:])
            report(st, "xa", rd("XA", 0, (N, D)), S[f"xa{l}"])
            report(st, "va", rd("VA", 0, (N, 3, D)), S[f"va{l}"])
            if l < L - 1:
                report(st, f"f_in{l+1}", rd("F", l + 1, (N * 32, D))[:E], S[f"f_in{l+1}"])
        elif st.startswith("fwd"):              # fused: edge stage l + node stage l+1 (xa / va are consumed inside)
            l = int(st[3:])
            if l < L - 1:
                report(st, f"f_in{l+1}", rd("F", l + 1, (N * 32, D))[:E], S[f"f_in{l+1}"])
            node_fwd_checks(st, l + 1)
        elif st == "head":
            report(st, "e_atom", rd("eatom", 0, (N,)), S["e_atom"][:, 0])
            report(st, "gx_out", rd("GX", 0, (N, D)), B["gx_out"])
            report(st, "gvec_out", rd("GVEC", 0, (N, 3, D)), B["gvec_out"])
        elif st in ("energy_reduce", "finalize"):
            report(st, "E", rd("energy", 0, (eng.n_graphs,)), S["E"][:, 0])
        elif st.startswith("node_bwd"):
            node_bwd_checks(st, int(st[8:]))
        elif st.startswith("edge_bwd"):
            edge_bwd_checks(st, int(st[8:]))
        elif st.startswith("bwd"):              # fused: node adjoint l+1 + edge adjoint l (accumulator set l & 1)
            l = int(st[3:])
            node_bwd_checks(st, l + 1)
            edge_bwd_checks(st, l, "2" if l & 1 else "")
        elif st == "embed_edge_bwd":
            report(st, "gx_emb", rd("GX", 0, (N, D)), B["gx_emb"])
        elif st == "embed_node_bwd":
            ea = rd("eacc", 0, (N * 32, 4))[:E]
            report(st, "g_d", ea[:, 1:4] * S["mask"][:, None], B["g_d"] * S["mask"][:, None])
            report(st, "forces", rd("forces", 0, (N, 3)), B["forces"])
    # full evaluation through the public host entry
    e, f = eng.forward_host(pos)
    report("forward_host", "E", e, S["E"][:, 0])
    report("forward_host", "forces", f, B["forces"])
    return lines, worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", default="real")
    ap.add_argument("--frags", default="chig")
    ap.add_argument("--max-frags", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--opts", default="", help="comma list key=value for vb_set_option")
    args = ap.parse_args()
    lines, _ = stage_report(args.frags, args.weights, args.max_frags, args.opts)
    text = "\n".join(lines)
    print(text)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as fh:
            fh.write(text + "\n")


if __name__ == "__main__":
    main()
