#!/usr/bin/env python
"""ncu csv (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per conv launch of ONE forward pass)
-> profiles/conv_traffic.json (read by bench.py for roofline.traffic) and a per-launch table.
    python tools/conv_traffic.py gpurun_out/conv_traffic.csv 64 640 profiles/conv_traffic.json profiles/<tag>_conv_traffic.txt [launches-per-pass]"""
import csv
import json
import sys

path, B, S, out_json, out_txt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
limit = int(sys.argv[6]) if len(sys.argv) > 6 else 0      # keep only the first N launches (one forward pass)
rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 10]
hdr = rows[0]
idc, mn, mu, mv = hdr.index('ID'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}
per = {}
for r in rows[1:]:
    d = per.setdefault(int(r[idc]), {})
    d[r[mn]] = float(r[mv].replace(',', '')) * scale.get(r[mu], 1)
ids = sorted(per)
if limit:
    ids = ids[:limit]
tot_r = sum(per[i].get('dram__bytes_read.sum', 0) for i in ids)
tot_w = sum(per[i].get('dram__bytes_write.sum', 0) for i in ids)
tot_t = sum(per[i].get('gpu__time_duration.sum', 0) for i in ids)
n = len(ids)
json.dump({'batch': B, 'size': S, 'launches': n, 'dram_bytes_per_launch': (tot_r + tot_w) / n, 'dram_read_bytes_total': tot_r,
           'dram_write_bytes_total': tot_w, 'time_us_total_under_ncu': tot_t,
           'source': f'ncu dram__bytes_read.sum + dram__bytes_write.sum over the {n} conv_umma_kernel / conv_chain_kernel launches of one forward pass '
                     f'(batch {B} @ {S}), tools/conv_traffic.py'}, open(out_json, 'w'), indent=1)
with open(out_txt, 'w') as f:
    f.write(f'# conv_umma_kernel launches of one forward pass, batch {B} @ {S}: DRAM bytes (ncu) and duration under ncu\n')
    f.write('launch  read_MB  write_MB  time_us\n')
    for k, i in enumerate(ids):
        d = per[i]
        f.write(f'{k:4d} {d.get("dram__bytes_read.sum", 0) / 1e6:9.1f} {d.get("dram__bytes_write.sum", 0) / 1e6:9.1f} {d.get("gpu__time_duration.sum", 0):9.1f}\n')
    f.write(f'total {tot_r / 1e6:9.1f} {tot_w / 1e6:9.1f} {tot_t:9.1f}\n')
print(open(out_json).read())
