#!/usr/bin/env python3
"""DRAM traffic per packed pair of every pipeline kernel, from one `ncu --set full` capture.
usage: python tools/ncu_traffic.py <report.ncu-rep> <pairs_per_launch yz> <pairs_per_launch x_fwd/unpack> > profiles/rNN_traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum of the first launch of each kernel, divided by the
pairs that launch processed; bench.py multiplies by the pairs per launch of its own run)."""
import csv, io, json, re, subprocess, sys
rep = sys.argv[1]
pairs_yz = float(sys.argv[2]); pairs_xf = float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
res = {}
for r in rows[2:]:
    m = re.search(r"(k_\w+)", r[ik])
    if not m or m.group(1) in res:
        continue
    k = m.group(1)
    b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    pairs = pairs_xf if k in ("k_x_fwd", "k_unpack") else pairs_yz
    res[k] = {"dram_bytes_per_launch": b, "pairs_per_launch": pairs, "dram_bytes_per_pair": b / pairs}
print(json.dumps({"source": rep, "kernels": res}, indent=1))
