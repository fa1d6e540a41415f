#!/usr/bin/env python
"""Tensor-pipe utilisation of the conv family from ncu counters.

    ncu --metrics sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,\
gpu__time_duration.sum --clock-control none -k regex:"conv_umma_kernel|conv_chain_kernel" -s 218 -c 109 --csv \
        --log-file conv_tensor.csv python tools/profile_ops.py --batch 64 --reps 1
    python tools/tensor_util.py conv_tensor.csv ops_b64.txt [out.txt]

`sm__inst_executed_pipe_tensor_subpipe_hmma` counts UTCHMMA (tcgen05.mma) instructions: on the first 3x3 conv of L1 it reads
1 382 400 = 51 200 pixel tiles x 27 (9 taps x 3 K steps), exactly what the kernel issues.  One tcgen05.mma kind::f16 of shape
M=128 x N x K=16 occupies the pipe for N/2 cycles (4096 MAC / clk / SM, microarchitecture guide "tcgen05 floor"), so
    pipe-active share of a launch = UTCHMMA count x N/2 / (SMs x sm__cycles_elapsed.avg)
and with N taken from the useful work (2*MAC of the layer = what bench.py counts) it is a LOWER bound (padded N columns of the
18-channel heads and zero K steps are not credited):
    useful utilisation = layer FLOP / (2 x 4096 x SMs x elapsed cycles).
The op list (tools/profile_ops.py output) supplies the layer shapes in launch order.
"""
import csv
import re
import sys
from collections import OrderedDict

SMS, MAC_PER_CLK = 148, 4096


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 8]
    ci = {h: i for i, h in enumerate(rows[0])}
    per = OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[ci['ID']], {'kernel': r[ci['Kernel Name']]})[r[ci['Metric Name']]] = r[ci['Metric Value']]
    ops = []
    for ln in open(sys.argv[2]):
        m = re.match(r'op\s+(\d+) L\s*(\d+) (\w+)\s+k(\d) s(\d) cin\s+(\d+) cout\s+(\d+) hw\s+(\d+)\s+([\d.]+) us', ln)
        if m and m.group(3) in ('conv', 'chain', 'detect'):
            ops.append(dict(op=int(m.group(1)), layer=int(m.group(2)), kind=m.group(3), k=int(m.group(4)), cin=int(m.group(6)),
                            cout=int(m.group(7)), hw=int(m.group(8)), us=float(m.group(9))))
    launches = list(per.values())
    assert len(launches) == len(ops), (len(launches), len(ops))
    B = 64
    out = ['# tensor-pipe utilisation of the conv family, one forward pass at batch 64 @ 640 under ncu (clocks as ncu left them)',
           '# op layer kind shape | time us | SM clock MHz | UTCHMMA instr | useful GFLOP | useful % of pipe peak (8192 FLOP/clk/SM)']
    tot_f = tot_c = 0.0
    cls = {}
    for o, l in zip(ops, launches):
        cyc = float(l['sm__cycles_elapsed.avg'])
        mhz = float(l['sm__cycles_elapsed.avg.per_second']) / 1e6
        inst = int(float(l['sm__inst_executed_pipe_tensor_subpipe_hmma.sum']))
        fl = 2.0 * o['cout'] * o['cin'] * o['k'] ** 2 * o['hw'] ** 2 * B if o['kind'] != 'chain' else None
        if fl is None:                     # fused chain: stage shapes are not in the op line; credit the issued MMAs at N = cout
            fl = inst * 2.0 * 128 * o['cout'] * 16
        util = fl / (2.0 * MAC_PER_CLK * SMS * cyc)
        tot_f += fl
        tot_c += cyc
        c = cls.setdefault('3x3 cout>=128' if o['k'] == 3 and o['cout'] >= 128 else ('3x3 cout<128 + chains' if o['k'] == 3 else ('1x1' if o['kind'] == 'conv' else 'detect')), [0.0, 0.0])
        c[0] += fl
        c[1] += cyc
        out.append(f"op {o['op']:3d} L{o['layer']:2d} {o['kind']:6s} k{o['k']} {o['cin']:4d}->{o['cout']:4d} @{o['hw']:3d} | {float(l['gpu__time_duration.sum']) / 1e3:7.1f} | "
                   f"{mhz:6.0f} | {inst:8d} | {fl / 1e9:8.2f} | {100 * util:5.1f}")
    out.append(f'# conv family total: {tot_f / 1e12:.3f} TFLOP in {tot_c:.0f} SM cycles -> {100 * tot_f / (2.0 * MAC_PER_CLK * SMS * tot_c):.1f} % of the tensor-pipe peak at the clock of each launch')
    for k, (f, c) in cls.items():
        out.append(f'#   {k:24s}: {100 * f / (2.0 * MAC_PER_CLK * SMS * c):5.1f} %  ({f / 1e12:.3f} TFLOP, {100 * c / tot_c:4.1f} % of the cycles)')
    txt = '\n'.join(out)
    print(txt)
    if len(sys.argv) > 3:
        open(sys.argv[3], 'w').write(txt + '\n')


if __name__ == '__main__':
    main()
