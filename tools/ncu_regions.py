#!/usr/bin/env python
"""Per-region view of an ncu report of wt_step_kernel: SASS instructions grouped by their execution count
(each loop nest of the kernel has a distinct count), with instruction share, stall-sample share and the
dominant stall reasons.   python tools/ncu_regions.py gpurun_out/prof.ncu-rep [warps]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
h = None; rows = []
for r in csv.reader(src.splitlines()):
    if len(r) > 5 and r[0] == "Address": h = r; continue
    if not h or len(r) != len(h): continue
    rows.append(r)
ix = {k: i for i, k in enumerate(h)}
W = int(sys.argv[2]) if len(sys.argv) > 2 else max(int(r[ix["Instructions Executed"]]) for r in rows[:3])
grp = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
for r in rows:
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    g = grp[n]; g[0] += 1; g[1] += n; g[2] += s
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k: g[3][k[6:]] += int(r[ix[k]])
ti = sum(g[1] for g in grp.values()); ts = sum(g[2] for g in grp.values())
print(f"warps {W}; warp-instructions per warp {ti / W:.0f}; samples {ts}")
print("exec/warp  ninstr  inst%  samp%  rel.cost  top stalls")
for n, g in sorted(grp.items(), key=lambda kv: -kv[1][2])[:16]:
    st = sorted(g[3].items(), key=lambda kv: -kv[1])[:5]
    print("%8.2f %6d  %5.1f  %5.1f  %5.2f   %s" % (n / W, g[0], 100 * g[1] / ti, 100 * g[2] / ts, (g[2] / ts) / max(g[1] / ti, 1e-9),
                                                  ", ".join("%s %.0f%%" % (k, 100 * v / max(g[2], 1)) for k, v in st)))
