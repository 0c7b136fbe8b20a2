#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_step_kernel.txt "note"
"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
lines = [f"# ncu summary of {rep}", f"# {note}", ""]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for w in WANT:
        if w in d:
            lines.append(f"{w:70s} {d[w]} {units[hdr.index(w)]}")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
h = None
ops, stalls, tot = collections.Counter(), collections.Counter(), 0
for r in csv.reader(src.splitlines()):
    if len(r) > 5 and r[0] == "Address":
        h = r
        continue
    if not h or len(r) != len(h):
        continue
    ix = {k: i for i, k in enumerate(h)}
    n = int(r[ix["Instructions Executed"]])
    t = r[1].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += n
    tot += n
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k:
            stalls[k] += int(r[ix[k]])
if tot:
    lines.append(f"warp-level instructions executed (SASS view): {tot}")
    lines.append("opcode mix (top 16): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(16)))
    S = sum(stalls.values())
    lines.append("warp stall samples: " + ", ".join(f"{k[6:]} {100 * v / S:.1f}%" for k, v in stalls.most_common(8)))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
