#!/usr/bin/env python
"""Top stall locations of an ncu report's source page:  python tools/ncu_hot.py rep.ncu-rep [N] [result-index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = heads[which]
end = heads[which + 1] if which + 1 < len(heads) else len(rows)
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
samp = col['# Samples']
tot = sum(int(r[samp] or 0) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('total samples', tot, 'results in report', len(heads))
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print('by reason:', ', '.join(f'{k[6:]} {100*v/max(tot,1):.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[samp] or 0))[:N]:
    n = int(r[samp] or 0)
    top = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f'{100*n/max(tot,1):5.1f}%  {r[col["Source"]][:90]:90s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}  exec={r[col["Instructions Executed"]]}')
