#!/usr/bin/env python
"""Instruction / stall-sample share per SASS region (split where the per-instruction execution count changes scale)."""
import csv, math, subprocess, sys
rep, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': [], 'hdr': None}; ks.append(cur)
    elif cur is not None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
seen = set()
for k in ks:
    if want not in k['name'] or k['name'] in seen:
        continue
    seen.add(k['name'])
    h = {n: i for i, n in enumerate(k['hdr'])}
    seg, prev, acc, n, start, samp = [], None, 0, 0, 0, 0
    for i, r in enumerate(k['rows']):
        e = int(r[h['Instructions Executed']] or 0); s = int(r[h['# Samples']] or 0)
        b = round(math.log2(e + 1) * 2)
        if prev is not None and abs(b - prev) >= 1 and n > 12:
            seg.append((start, i - 1, acc, samp, n)); acc = 0; n = 0; start = i; samp = 0
        prev = b; acc += e; n += 1; samp += s
    seg.append((start, len(k['rows']) - 1, acc, samp, n))
    T = sum(s[2] for s in seg); TS = sum(s[3] for s in seg)
    print(k['name'][22:90], 'total instr', T, 'samples', TS)
    for s in seg:
        if s[2] > 0.004 * T or s[3] > 0.01 * TS:
            r0 = k['rows'][s[0]]
            print('  sass %4d-%4d (%3d)  exec/instr %9d  instr %5.1f%%  samples %5.1f%%   %s' % (s[0], s[1], s[4], s[2] // max(s[4], 1), 100 * s[2] / T, 100 * s[3] / TS, r0[h['Source']][:50]))
