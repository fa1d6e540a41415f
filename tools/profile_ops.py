#!/usr/bin/env python
"""Per-op CUDA-event breakdown of one forward pass (uses ry_plan_set_profiling).  Run on the GPU box:
    python tools/profile_ops.py [--batch 64] [--size 640] [--reps 5] [--out gpurun_out/ops.txt]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import repyolo_b200 as R  # noqa: E402
from oracle import repyolo_oracle as O  # noqa: E402  (weights generator only)

KIND = {1: 'stem', 2: 'conv', 3: 'dw5', 4: 'maxpool2', 5: 'spp', 6: 'upsample2', 7: 'ca', 8: 'attn_qk', 9: 'crisscross',
        10: 'vertical', 11: 'detect', 12: 'chain'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--size', type=int, default=640)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    layers, save, sd, fz = O.make_model(0, 'calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    dev = torch.device('cuda:0')
    x = torch.rand(a.batch, 3, a.size, a.size, device=dev)
    eng = m.engine(dev)
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    ops = eng.plan_ir.ops
    acc = [0.0] * len(ops)
    for _ in range(a.reps):
        m(x)
        torch.cuda.synchronize()
        for j, t in enumerate(eng.op_times_ms()):
            acc[j] += t / a.reps
    lines = []
    tot = sum(acc)
    bykind, bylayer = {}, {}
    for j, d in enumerate(ops):
        lvl = eng.plan_ir.tensors[d.in0.tensor].level
        hw = (a.size >> lvl) // max(d.stride, 1)
        fl = 2.0 * d.cout * d.cin * d.ksize ** 2 * hw * hw * a.batch if d.kind in (1, 2, 11) else 0.0
        tf = fl / (acc[j] * 1e-3) / 1e12 if acc[j] > 0 else 0
        lines.append(f'op {j:3d} L{d.layer:2d} {KIND[d.kind]:10s} k{d.ksize} s{d.stride} cin {d.cin:4d} cout {d.cout:4d} hw {hw:3d}  {acc[j]*1e3:8.1f} us  {tf:7.1f} TF/s')
        bykind[KIND[d.kind]] = bykind.get(KIND[d.kind], 0.0) + acc[j]
        bylayer[d.layer] = bylayer.get(d.layer, 0.0) + acc[j]
    lines.append(f'TOTAL {tot:.3f} ms for batch {a.batch} @ {a.size}')
    lines.append('by kind: ' + ', '.join(f'{k} {v:.3f} ms ({100*v/tot:.1f}%)' for k, v in sorted(bykind.items(), key=lambda kv: -kv[1])))
    lines.append('by layer: ' + ', '.join(f'L{k} {v:.3f}' for k, v in sorted(bylayer.items())))
    # NMS
    pred, _ = m(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for conf, iou in ((0.25, 0.45), (0.001, 0.65)):
        R.nms_padded(pred, conf, iou)
        e0.record()
        for _ in range(a.reps):
            out, cnt = R.nms_padded(pred, conf, iou)
        e1.record()
        torch.cuda.synchronize()
        lines.append(f'nms conf {conf} iou {iou}: {e0.elapsed_time(e1)/a.reps:.3f} ms; candidates>conf {(pred[..., 4] > conf).sum().item()/a.batch:.0f}/img; dets {cnt.float().mean().item():.0f}/img')
    txt = '\n'.join(lines)
    print(txt)
    if a.out:
        open(a.out, 'w').write(txt + '\n')


if __name__ == '__main__':
    main()
