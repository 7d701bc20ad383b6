#!/usr/bin/env python
"""Per-kernel SASS instruction counts of libnoc_b200.so (cuobjdump -sass): the evidence VERDICT r1 asked for —
TMA bulk copies (UBLKCP), FP64 FMAs (DFMA), tensor-core FP64 (DMMA), local-memory spills (LDL / STL), registers.
    python tools/sass_summary.py > profiles/sass_r02_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "noc_b200", "libnoc_b200.so")
WATCH = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "UBLKCP", "LDGSTS", "SYNCS", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "BRA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            g = lambda key: int((re.search(key + r":(\d+)", line) or [0, 0])[1])
            regs[cur] = (int(m.group(1)), g("SHARED"), g("STACK") + g("LOCAL"))
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter(); order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
    dm = demangle(order)
    print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}  (sm_100a only; counts are static instructions per kernel)")
    print("# kernel | regs | static smem | stack+local bytes/thread | total | " + " | ".join(WATCH))
    for k in sorted(order, key=lambda k: dm[k]):
        r = regs.get(k, ("?", "?", "?"))
        name = re.sub(r"^void ", "", dm[k]).replace("(anonymous namespace)::", ""); name = re.sub(r"\(.*$", "", name)
        c = counts[k]
        print(f"{name} | {r[0]} | {r[1]} | {r[2]} | {c['total']} | " + " | ".join(str(c[w]) for w in WATCH))


if __name__ == "__main__":
    main()
