#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: python tools/launch_summary.py gpurun_out/launches.csv [first_n_rows_to_print]"""
import collections, csv, sys
def main(path, show=0):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            h, start = r, i + 1
            break
    ix = {k: j for j, k in enumerate(h)}
    seq = []
    for r in rows[start:]:
        if len(r) < len(h):
            continue
        try:
            v = float(r[ix['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        u = r[ix['Metric Unit']]
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
        seq.append((r[ix['Kernel Name']].split('(')[0][:44], r[ix['Grid Size']], v))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, v in seq:
        agg[n][0] += 1; agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(seq)} launches, {tot:.1f} us")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:46s} n={v[0]:4d} total_us={v[1]:10.1f} avg_us={v[1]/v[0]:8.1f} share={v[1]/tot:.3f}")
    for s in seq[:show]:
        print(s)
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
