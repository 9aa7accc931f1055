#!/usr/bin/env python
"""In-kernel timeline of the tcgen05 edge kernels: SM-clock stamps taken by CTA 0 (first tile) at every phase
boundary (engine option "timeline"; slots documented next to TC_TL in csrc/k_edge_tc.cuh).

    python tools/tc_timeline.py [--workload chig] [--layer 2] [--mhz 1920]

Prints, for the forward and the adjoint kernel of one layer, the time of each stamp relative to kernel start and
the delta to the previous stamp -- where a single 128-edge tile spends its ~30-60 us.
"""
import argparse
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ai2bmd_b200.engine import Engine            # noqa: E402
from ai2bmd_b200.fixtures import WEIGHTS, load_fragments   # noqa: E402
from ai2bmd_b200.synth import synthetic_batch    # noqa: E402
from ai2bmd_b200.weights import load_state_dict  # noqa: E402

FWD = {0: "kernel start", 1: "setup done (barriers, TMEM alloc)", 2: "f tile + meta loaded", 3: "A=f in TMEM, go dk/dv",
       4: "dk done seen", 5: "D0 -> tile", 6: "attention weights done", 7: "dv done seen", 8: "D1 -> tile",
       9: "messages m done", 10: "xa aggregation done", 11: "A=m in TMEM, go s1", 12: "edge update (f) done",
       13: "s1 done seen", 14: "D1 -> tile", 15: "s1 aggregation done", 16: "s2 done seen", 17: "D0 -> tile",
       18: "s2 aggregation done", 31: "teardown done"}
BWD = {0: "kernel start", 1: "setup done", 2: "meta loaded", 3: "s1-half SIMT done", 4: "A in TMEM, go g3a",
       5: "s2-half SIMT done", 6: "g3a done seen", 7: "A in TMEM, go g3b", 8: "g3b done seen", 9: "D1 -> tile",
       10: "g_m / g_Pdv SIMT done", 11: "A in TMEM, go g4dv", 12: "g_Pdk SIMT done", 13: "g4dv done seen",
       14: "A in TMEM, go g4dk", 15: "g_q tile + aggregation done", 16: "g_Pf SIMT done", 17: "g4dk done seen",
       18: "A in TMEM, go g4f", 19: "g_wdot tile + aggregation done", 20: "last job done seen", 21: "D0 -> tile",
       22: "g_f written", 31: "teardown done"}


NFWD = {0: "start", 1: "xa rows staged", 2: "o_proj unit done", 4: "K-quarters summed", 5: "per-node phase done",
        7: "projection unit done", 9: "results written", 11: "vec_dot done"}
NBWD = {0: "start", 1: "A rows staged", 3: "adjoint unit done", 5: "per-node phase done", 7: "o_proj adjoint unit done", 9: "g_xa written"}


def show(title, tl, names, mhz, jobs):
    t0 = int(tl[0])
    print(f"--- {title} ---")
    prev = t0
    for k in sorted(names):
        if tl[k] == 0:
            continue
        t = int(tl[k])
        print(f"  [{k:2d}] {names[k]:<36} {(t - t0) / mhz:8.2f} us   (+{(t - prev) / mhz:6.2f})")
        prev = t
    for j in range(jobs):
        a, b = int(tl[32 + 2 * j]), int(tl[33 + 2 * j])
        if a:
            print(f"  issuer job {j}: go seen {(a - t0) / mhz:8.2f} us, MMAs issued {(b - t0) / mhz:8.2f} us (+{(b - a) / mhz:5.2f})")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="chig")
    ap.add_argument("--layer", type=int, default=2)
    ap.add_argument("--mhz", type=float, default=0.0)
    ap.add_argument("--opts", default="", help="comma list key=value for vb_set_option")
    args = ap.parse_args()
    mhz = args.mhz
    if not mhz:
        try:
            mhz = float(subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"]).split()[0])
        except Exception:
            mhz = 1920.0
    if args.workload.isdigit():
        fd = synthetic_batch(int(args.workload), seed=0)
    else:
        fd, _ = load_fragments(args.workload)
    sd = load_state_dict(WEIGHTS)
    eng = Engine(sd, 0)
    eng.set_option("edge_tc", 3)
    eng.set_option("timeline", 1)
    for kv in filter(None, args.opts.split(",")):
        kk, vv = kv.split("=")
        eng.set_option(kk, int(vv))
    eng.set_topology(fd.z, fd.batch, n_graphs=len(fd))
    pos = np.ascontiguousarray(fd.pos, dtype=np.float32)
    for _ in range(3):
        eng.forward_host(pos)
    torch.cuda.synchronize()
    print(f"workload {args.workload}: {len(fd.z)} atoms, SM clock assumed {mhz:.0f} MHz, layer {args.layer}")
    nl = 6
    tf = eng.debug_read("TL", args.layer, (64,), dtype=np.uint64)
    tb = eng.debug_read("TL", nl + args.layer, (64,), dtype=np.uint64)
    show("edge_fwd_tc", tf, FWD, mhz, 5)
    show("edge_bwd_tc", tb, BWD, mhz, 5)
    if eng.get_option("node_tc") == 0 and eng.get_option("fused") == 0:
        EMB = {0: "start", 1: "previous kernel complete (pdl)", 2: "loads issued, rows staged", 3: "aggregation done", 4: "combine done", 5: "x written"}
        for title, idx, names in ((f"node_fwd2 stage {args.layer}", args.layer, NFWD), (f"node_bwd2 stage {args.layer}", nl + 1 + args.layer, NBWD),
                                  ("embed_node_small", 2 * nl + 2, EMB)):
            tl = eng.debug_read("TLN", idx, (16, 16), dtype=np.uint64).astype(np.int64)     # [stamp][warp]
            t0 = tl[0][tl[0] > 0].min()
            print(f"--- {title} (nodes per CTA {eng.get_option('node_nb')}): per stamp, first / last warp to pass it ---")
            for i in sorted(names):
                row = tl[i][tl[i] > 0]
                if len(row):
                    per = " ".join(f"{(x - t0) / mhz:5.1f}" if x > 0 else "    -" for x in tl[i])
                    print(f"  [{i:2d}] {names[i]:<34} {(row.min() - t0) / mhz:7.2f} .. {(row.max() - t0) / mhz:7.2f} us   | {per}")

if __name__ == "__main__":
    main()
