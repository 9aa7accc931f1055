#!/usr/bin/env python
"""Collect dram bytes per launch of a kernel from an ncu --set full report into profiles/ncu_traffic.json.

    python tools/ncu_traffic.py <workload> <kernel family> <report.ncu-rep>
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
workload, family, rep = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
vals = []
for row in rows[2:]:
    d = dict(zip(hdr, row))
    t = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        t += float(d[k]) * scale[units[hdr.index(k)]]
    vals.append(t)
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[f"{workload}:{family}"] = sum(vals) / len(vals)
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
print(path, data)
