#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` (per kernel)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    want = sys.argv[3] if len(sys.argv) > 3 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    for k in kernels:
        if want and want not in k['name']:
            continue
        h = {n: i for i, n in enumerate(k['hdr'])}
        tot = sum(int(r[h['# Samples']] or 0) for r in k['rows'])
        toti = sum(int(r[h['Instructions Executed']] or 0) for r in k['rows'])
        print('==', k['name'][:110], 'samples', tot, 'warp-instr', toti, 'sass lines', len(k['rows']))
        stalls = [n for n in k['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
        agg = {s: sum(int(r[h[s]] or 0) for r in k['rows']) for s in stalls}
        print('   stall totals:', ', '.join('%s=%.1f%%' % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        order = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][h['# Samples']] or 0))[:top]
        for i in sorted(order):
            r = k['rows'][i]
            s = int(r[h['# Samples']] or 0)
            best = max(stalls, key=lambda st: int(r[h[st]] or 0))
            print('   %5d %5.1f%% exec=%-9s %-12s %s' % (i, 100.0 * s / max(tot, 1), r[h['Instructions Executed']], best[6:], r[h['Source']][:100]))


if __name__ == '__main__':
    main()
