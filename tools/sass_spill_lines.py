#!/usr/bin/env python
"""Where a kernel spills: kernel-body source lines (outermost inline frame, nvdisasm -gi) of its STL / LDL instructions.
usage: python tools/sass_spill_lines.py '<substring of the mangled kernel name>'"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "noc_b200", "libnoc_b200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
cur, line, cnt, total = None, None, collections.Counter(), 0
for l in txt.splitlines():
    m = re.match(r"\.text\.(\S+):", l)
    if m: cur = m.group(1); continue
    if not cur or sys.argv[1] not in cur: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(\S+)", l)
    if m:
        total += 1
        op = m.group(1).split(".")[0]
        if op in ("STL", "LDL"): cnt[(line, op)] += 1
print("instructions", total)
for (ln, op), c in sorted(cnt.items(), key=lambda kv: (kv[0][0][1], kv[0][1])): print(ln, op, c)
