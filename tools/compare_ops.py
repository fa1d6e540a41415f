#!/usr/bin/env python
"""Side-by-side of two tools/profile_ops.py outputs:  python tools/compare_ops.py old.txt new.txt [min_us]"""
import re
import sys


def load(path):
    d = {}
    for line in open(path):
        m = re.match(r'op\s+(\d+) L\s*(\d+) (\w+)\s+k(\d) s(\d) cin\s+(\d+) cout\s+(\d+) hw\s+(\d+)\s+([\d.]+) us', line)
        if m:
            d[int(m.group(1))] = (m.group(3), int(m.group(2)), int(m.group(4)), int(m.group(5)), int(m.group(6)), int(m.group(7)),
                                  int(m.group(8)), float(m.group(9)))
    return d


a, b = load(sys.argv[1]), load(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
B = 64
ta = tb = 0.0
for i in sorted(b):
    kind, L, k, s, cin, cout, hw, us = b[i]
    old = a.get(i, (None,) * 7 + (float('nan'),))[7]
    ta += old
    tb += us
    if max(us, old) < thr:
        continue
    extra = ''
    if kind in ('conv', 'detect'):
        ho = hw // s
        fl = 2.0 * cin * cout * k * k * ho * ho * B
        by = B * (hw * hw * cin + ho * ho * cout) * 2
        extra = f'  {fl / us / 1e6:7.1f} TF/s  {by / us / 1e3:7.1f} GB/s  floor {max(fl / 1406.6e6, by / 6464.9e3):6.1f} us'
    print(f'op {i:3d} L{L:2d} {kind:10s} k{k} s{s} {cin:4d}->{cout:4d} @{hw:3d}  {old:8.1f} -> {us:8.1f} us{extra}')
print(f'total {ta / 1e3:.3f} -> {tb / 1e3:.3f} ms')
