#!/usr/bin/env python
"""Top source lines by warp-stall samples from an ncu report (needs -lineinfo builds, --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [N]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file, hdr = None, None
agg = collections.Counter()
reason = collections.defaultdict(collections.Counter)
src_text = {}
total = 0
for row in csv.reader(raw.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        ix_s = hdr.index("# Samples")
        stall_ix = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or cur_file is None or not row[0].isdigit():
        continue
    if row[2] != "-":     # SASS rows carry an address; CUDA-line rows have "-"
        continue
    try:
        n = int(row[ix_s])
    except Exception:
        continue
    key = (cur_file, int(row[0]))
    agg[key] += n
    total += n
    src_text[key] = row[1].strip()[:110]
    for i, h in stall_ix:
        try:
            reason[key][h] += int(row[i])
        except Exception:
            pass
print(f"total samples {total}")
for key, n in agg.most_common(topn):
    top = ", ".join(f"{k[6:]}={v}" for k, v in reason[key].most_common(2))
    print(f"{100 * n / max(total, 1):5.1f}%  {key[0]}:{key[1]:<4d} {src_text[key]}   [{top}]")
