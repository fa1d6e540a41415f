#!/usr/bin/env python
"""Instruction mix / phase profile of one kernel from an ncu report's source page:
    ncu -i X.ncu-rep --page source --csv --kernel-id :::N > src.csv ; python tools/ncu_source_mix.py src.csv [bin]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 6 and r[0].startswith('0x')]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
ex = [int(r[5] or 0) for r in data]
sm = [int(r[4] or 0) for r in data]
tot = sum(ex)
print('kernel:', rows[0][1][:120])
print('SASS instructions', len(data), 'warp-instructions executed', tot, 'samples', sum(sm))
byop = collections.Counter()
for r, e in zip(data, ex):
    t = r[1].split()
    op = t[1] if t and t[0].startswith('@') else (t[0] if t else '?')
    byop[op.split('.')[0]] += e
print('by opcode:', ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in byop.most_common(22)))
for i in range(0, len(data), B):
    print(f'{i:5d} exec {sum(ex[i:i + B]):9d} ({100 * sum(ex[i:i + B]) / tot:4.1f}%) samples {sum(sm[i:i + B]):6d}   {data[i][1].strip()[:60]}')
