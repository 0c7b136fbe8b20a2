#!/usr/bin/env python
"""Static SASS instructions of a kernel per OUTERMOST wt_step_core.h statement (the frame inlined directly into the
kernel), from nvdisasm -gi.  Shows which statement of run() / begin() costs how much code (instruction-fetch bound kernel).
    python tools/sass_by_stmt.py lib.so kernel-substring [bucket]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
lib, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 1
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cub)], capture_output=True, text=True).stdout
infn, chain, cnt, spill = False, [], Counter(), Counter()
pending = []
for line in dis.splitlines():
    if line.startswith(".text."):
        infn = pat in line
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2)), os.path.basename(m.group(3)) if m.group(3) else None))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        if pending:
            chain, pending = pending, []
        key = None
        for f, l, parent in chain:  # outermost core frame = the one inlined into the .cu file
            if f == "wt_step_core.h" and parent == "wt_kernels.cu":
                key = l
        if key is None:
            key = -1 if not chain else (-2 if chain[-1][0] == "wt_kernels.cu" else -3)
        cnt[key // bucket * bucket] += 1
        if "LDL" in line or "STL" in line:
            spill[key // bucket * bucket] += 1
src = open("ics_wt_physicsengine_b200/csrc/wt_step_core.h").read().splitlines()
print("total", sum(cnt.values()))
for k in sorted(cnt):
    text = src[k - 1].strip()[:100] if k > 0 else {-1: "(no line info)", -2: "(kernel body)", -3: "(other)"}.get(k, "")
    print(f"{k:5d} {cnt[k]:5d} {spill[k]:3d}  {text}")
