#!/usr/bin/env python3
"""Summarise ncu outputs into small text files for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/rNN_full.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
]


def short(name):
    m = re.search(r"(k_\w+)", name)
    return m.group(1) if m else name[:40]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        k = short(r[ik])
        t = float(r[iv].replace(",", "")) / 1e3  # ns -> us
        e = tot.setdefault(k, [0, 0.0, r[ig], r[ib]])
        e[0] += 1
        e[1] += t
    all_us = sum(e[1] for e in tot.values())
    print(f"# ncu launch list: {path}  ({len(rows) - 1} launches, {all_us / 1e3:.2f} ms summed; cold-cache, serialised)")
    print(f"{'kernel':44s} {'n':>5s} {'sum_us':>11s} {'avg_us':>9s} {'share':>7s}  grid block")
    for k, e in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {e[0]:5d} {e[1]:11.1f} {e[1] / e[0]:9.1f} {100 * e[1] / all_us:6.1f}%  {e[2]} {e[3]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full: {path}")
    for r in rows[2:]:
        print(f"\n## {short(r[hdr.index('Kernel Name')])}  id={r[0]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:62s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
