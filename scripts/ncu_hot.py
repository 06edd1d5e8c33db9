#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples for the first kernel in an .ncu-rep (source page)."""
import csv, subprocess, sys
def main(path, top=25, kernel_idx=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': [], 'hdr': None}; blocks.append(cur); continue
        if cur is None or not r: continue
        if r[0] == 'Address': cur['hdr'] = r; continue
        cur['rows'].append(r)
    b = blocks[kernel_idx]; h = {n: i for i, n in enumerate(b['hdr'])}
    tot = sum(int(r[h['# Samples']] or 0) for r in b['rows'])
    print(b['name'], 'total samples', tot, 'instructions', len(b['rows']))
    stall_cols = [n for n in b['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
    agg = {n: sum(int(r[h[n]] or 0) for r in b['rows']) for n in stall_cols}
    print('  stall mix:', ', '.join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ranked = sorted(enumerate(b['rows']), key=lambda ir: -int(ir[1][h['# Samples']] or 0))[:top]
    for i, r in sorted(ranked):
        s = int(r[h['# Samples']] or 0)
        why = max(stall_cols, key=lambda n: int(r[h[n]] or 0))
        print(f"  [{i:5d}] {100*s/max(tot,1):5.1f}%  {why[6:]:12s} {r[h['Source']].strip()[:100]}")
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
