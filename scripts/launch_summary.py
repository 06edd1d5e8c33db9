#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections, csv, re, sys
def main(path, tail=0):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0]); order = []
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        if 'spin_kernel' in name: continue   # the bench's host-decoupling spin, not part of a step
        v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
        agg[name][0] += 1; agg[name][1] += v; order.append((name, row['Grid Size'], v))
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{k[:58]:58s} n={v[0]:4d} total={v[1]/1e3:8.2f} ms share={100*v[1]/tot:5.1f}% avg={v[1]/v[0]:8.1f} us")
    print(f"total {tot/1e3:.2f} ms over {len(order)} launches")
    for name, grid, v in order[-tail:] if tail else []:
        print(f"   {name[:44]:44s} {grid:20s} {v:9.1f} us")
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
