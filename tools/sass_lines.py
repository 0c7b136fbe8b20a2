#!/usr/bin/env python
"""Static SASS instructions of wt_step_kernel per source line (innermost inlined frame), from nvdisasm -g.
    python tools/sass_lines.py [lib.so] [top]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
lib = sys.argv[1] if len(sys.argv) > 1 else "ics_wt_physicsengine_b200/csrc/libwt_b200.so"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur, infn, cnt = None, False, Counter()
for line in dis.splitlines():
    if line.startswith(".text."):
        infn = "wt_step_kernel" in line
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        cnt[cur] += 1
print("total", sum(cnt.values()))
src = {}
for (f, l), c in cnt.most_common(top):
    path = os.path.join("ics_wt_physicsengine_b200/csrc", f)
    if f not in src and os.path.exists(path):
        src[f] = open(path).read().splitlines()
    text = src[f][l - 1].strip()[:90] if f in src and l - 1 < len(src[f]) else ""
    print(f"{c:5d}  {f}:{l}  {text}")
