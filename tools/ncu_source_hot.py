#!/usr/bin/env python
"""Hot CUDA source lines of an .ncu-rep (captured with --import-source on; library built with -lineinfo).
The CSV source page of ncu carries metrics per SASS instruction only, so the line numbers come from
`nvdisasm -g` of the shipped cubin, joined by instruction offset.
usage: python tools/ncu_source_hot.py prof.ncu-rep [top] [lo:hi:label ...] [--innermost]
Lines are those of the kernel body (outermost inline frame) unless --innermost."""
import csv, os, re, subprocess, sys, tempfile

INNERMOST = "--innermost" in sys.argv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def sass_lines(mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "noc_b200", "libnoc_b200.so")], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    # -gi: an instruction is preceded by its inline chain, innermost first; the LAST "//## File" line before it is the
    # line of the kernel body itself (the outermost frame), which is what the phases of a kernel are told apart by
    funcs, cur, line, fresh = {}, None, ("?", 0), True
    for l in txt.splitlines():
        m = re.match(r"\.text\.(\S+):", l)
        if m: cur = m.group(1); funcs[cur] = {}; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if INNERMOST and not fresh: continue
            line = (os.path.basename(m.group(1)), int(m.group(2))); fresh = False; continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
        if m and cur: funcs[cur][int(m.group(1), 16)] = line; fresh = True
    return funcs

def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
    ranges = [a.split(":") for a in sys.argv[2:] if ":" in a]
    sel = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--kernel=")]   # report with several kernels: pick one by name
    out = subprocess.run(["ncu", "-i", path] + (["--kernel-name", "regex:" + sel[0]] if sel else []) + ["--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    nxt = [i for i, r in enumerate(rows) if i > 0 and r and r[0] == "Kernel Name"]   # keep the first kernel's block only
    if nxt: rows = rows[:nxt[0]]
    kname = rows[0][1]
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(body[0][0], 16)
    funcs = sass_lines(None)
    names = list(funcs)
    demangled = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    norm = lambda s: re.sub(r"\((int|bool)\)", "", s).replace("true", "1").replace("false", "0").replace("const", "").replace(" ", "")
    match = [n for n, d in zip(names, demangled) if norm(d) == norm(kname)]
    if not match: sys.exit(f"kernel {kname} not found in the cubin")
    lmap = funcs[match[0]]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda x: float(x.replace(",", "")) if x not in ("", "-") else 0.0
    agg = {}
    for r in body:
        key = lmap.get(int(r[0], 16) - base, ("?", 0))
        a = agg.setdefault(key, [0.0, 0.0, {}, 0.0, 0.0])
        a[0] += num(r[ci["# Samples"]]); a[1] += num(r[ci["Instructions Executed"]])
        a[3] += num(r[ci["L1 Wavefronts Shared"]]); a[4] += num(r[ci["L1 Wavefronts Shared Ideal"]])
        for h in stalls: a[2][h[6:]] = a[2].get(h[6:], 0.0) + num(r[ci[h]])
    tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
    tot_w = sum(a[3] for a in agg.values()) or 1.0
    print(f"kernel {kname}\ntotal samples {tot_s:.0f}, warp instructions {tot_i:.0f}, shared-memory wavefronts {tot_w:.0f} "
          f"(ideal {sum(a[4] for a in agg.values()):.0f})")
    src = {}
    for lo, hi, label in ranges:
        lo, hi = int(lo), int(hi)
        sel = [a for (f, ln), a in agg.items() if lo <= ln < hi]
        s, i, w = sum(a[0] for a in sel), sum(a[1] for a in sel), sum(a[3] for a in sel)
        print(f"[{lo:4d},{hi:4d}) {label:34s} samples {100*s/tot_s:5.1f}%  instructions {100*i/tot_i:5.1f}%  smem wavefronts {100*w/tot_w:5.1f}%")
    order = 3 if "--by-wavefronts" in sys.argv else 0
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][order])[:top]:
        if f not in src:
            p = os.path.join(ROOT, "noc_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ""
        st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
        print(f"{f[:16]:16s}{ln:4d} {100*a[0]/tot_s:5.1f}% s {100*a[1]/tot_i:5.1f}% i {100*a[3]/tot_w:5.1f}% w (x{a[3]/max(a[4],1):.1f})  {text:80s} " +
              " ".join(f"{n}:{100*v/max(a[0],1):.0f}%" for n, v in st))

if __name__ == "__main__":
    main()
