#!/usr/bin/env python3
"""Top stalled SASS instructions of one kernel from an ncu report (source page).
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches may match: take the first block
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
body = rows[hdr_i[0] + 1:end]
iS, iN, iSrc, iEx = h.index("Warp Stall Sampling (All Samples)"), h.index("Warp Stall Sampling (Not-issued Samples)"), h.index("Source"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_")]
tot = sum(int(r[iS] or 0) for r in body)
print(f"{rows[0][1][:90]}  total samples {tot}, {len(body)} SASS instructions")
agg = {}
for r in body:
    for i in stall_cols:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print("stall reasons:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][iS] or 0))[:top]:
    why = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:5d} {int(r[iS]):6d} {100*int(r[iS])/max(tot,1):5.1f}%  {r[iSrc].strip()[:70]:70s} {why}")
