#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

    python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep  profiles/r01_xxx.txt
    python tools/ncu_summary.py list  gpurun_out/launches.csv   profiles/r01_launches_xxx.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary of {rep}\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            fh.write(f"\nkernel: {d.get('Kernel Name')}\n")
            for k in KEEP:
                if k in d:
                    fh.write(f"  {k:85s} {d[k]:>16s} {units[hdr.index(k)]}\n")
            if "dram__bytes_read.sum" in d:
                fh.write("  (traffic = dram__bytes_read.sum + dram__bytes_write.sum)\n")


def launch_list(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "nbr_build" in n]
    with open(out, "w") as fh:
        fh.write(f"# per-launch device time (ncu --metrics gpu__time_duration.sum --clock-control none), {path}\n")
        fh.write("# cold-cache, serialised launches: compare SHARES of the step, not absolutes\n")
        if len(starts) >= 3:
            s, e = starts[-2], starts[-1]
            agg = collections.OrderedDict()
            tot = 0.0
            for r in rows[s:e]:
                k = re.sub(r"<.*", "", r["Kernel Name"].split("(")[0]).replace("vb::", "").replace("void ", "")
                t = float(r["Metric Value"])
                a = agg.setdefault(k, [0.0, 0, r["Grid Size"], r["Block Size"]])
                a[0] += t
                a[1] += 1
                tot += t
            fh.write(f"# one evaluation = launches {s}..{e - 1}; total {tot / 1e3:.1f} us\n")
            for k, v in agg.items():
                fh.write(f"{k:34s} n={v[1]:2d} total={v[0] / 1e3:9.1f} us  share={v[0] / tot * 100:5.1f}%  grid={v[2]} block={v[3]}\n")


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
