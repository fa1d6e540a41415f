#!/usr/bin/env python
"""Summarise an ncu report (--set full) into a small text file for profiles/:
    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/X_ncu_full.txt
and an ncu launch list CSV (--metrics gpu__time_duration.sum) into per-kernel totals / shares:
    python tools/ncu_summary.py --launches gpurun_out/launches_X.csv profiles/X_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tc.sum', 'sm__inst_executed_pipe_tma.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']


def full(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get('Kernel Name')
    with open(out, 'w') as f:
        f.write(f'# ncu --set full --clock-control none; source report {rep}; one column per captured launch\n')
        f.write('kernel: ' + ' | '.join(r[name_i] for r in data) + '\n')
        for h in hdr:
            short = h.split('.', 2)[-1] if h.split('.')[0] in ('LTS', 'SM_C', 'TPC', 'SM_A', 'SM_B') else h
            if any(short == k or h == k for k in KEYS):
                f.write(f'{h} [{units[col[h]]}]: ' + ' | '.join(r[col[h]] for r in data) + '\n')
    print(open(out).read())


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot = OrderedDict()
    n = 0
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        u = r[ui]
        us = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
        k = r[ki].split('(')[0].split('::')[-1]
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        n += 1
    total = sum(a[1] for a in tot.values())
    with open(out, 'w') as f:
        f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none; {n} launches from {path}; cold-cache serialised '
                f'times: compare SHARES\n')
        f.write(f'{"kernel":48s} {"launches":>8s} {"total_us":>12s} {"share":>8s}\n')
        for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k:48s} {c:8d} {us:12.1f} {100 * us / total:7.2f}%\n')
        f.write(f'{"TOTAL":48s} {n:8d} {total:12.1f}\n')
    print(open(out).read())


if __name__ == '__main__':
    if sys.argv[1] == '--launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
