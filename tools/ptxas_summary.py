"""Summarise `nvcc -Xptxas -v` output (build log) per kernel: registers, stack frame, spill bytes, shared memory.
usage: python tools/ptxas_summary.py /tmp/build.log"""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
names = re.findall(r"Compiling entry function '(\S+)'", txt)
dem = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))
cur, st = None, ("?",) * 3
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m: cur = m.group(1)
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: st = m.groups()
    m = re.search(r"Used (\d+) registers(?:.*?(\d+) bytes smem)?", line)
    if m and cur:
        d = re.sub(r"\(.*", "", dem[cur]).replace("void ", "")
        print(f"{d:60s} regs {m.group(1):>3s} stack {st[0]:>5s} spill st/ld {st[1]:>5s}/{st[2]:>5s}")
