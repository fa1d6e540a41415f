#!/usr/bin/env python
"""Timings of the other BASELINE.json configurations (parity-test cases, not bench lines) on one GPU:
    configs[3]  Rep-YOLO fused bf16 at 1280x1280, batch 16: forward + NMS(0.25, 0.45)
    configs[4]  NMS stress: test.py settings (conf 0.001, iou 0.65, multi_label, max_det 300) on 25200 candidates x batch 256
                (nc = 1 as the Rep-YOLO head, where general.py:970 forces multi_label off) and the nc = 80 variant x batch 32
    python tools/bench_configs.py [--out gpurun_out/configs.txt]
"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import repyolo_b200 as R  # noqa: E402
from oracle import repyolo_oracle as O  # noqa: E402  (weights generator only)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def stress_pred(B, N, nc, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    p = torch.empty(B, N, 5 + nc, device='cuda')
    p[..., 0:2] = torch.rand(B, N, 2, device='cuda', generator=g) * 640.0
    p[..., 2:4] = torch.exp(torch.rand(B, N, 2, device='cuda', generator=g) * (math.log(320.0) - math.log(8.0)) + math.log(8.0))
    p[..., 4:] = torch.rand(B, N, 1 + nc, device='cuda', generator=g)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    lines = []
    layers, save, sd, fz = O.make_model(0, 'calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(16, 3, 1280, 1280, device='cuda')

    def step():
        pred, _ = m(x)
        return R.nms_padded(pred, 0.25, 0.45)
    ms = timed(step, 5)
    fwd = timed(lambda: m(x), 5)
    lines.append(f'config 4: 1280x1280 batch 16, forward + NMS(0.25, 0.45): {ms:.2f} ms/step = {16 / ms * 1e3:.0f} images/s (forward {fwd:.2f} ms, '
                 f'{275.5 * 16 / fwd:.0f} TFLOP/s over all convs)')
    del x
    p1 = stress_pred(256, 25200, 1)
    ms1 = timed(lambda: R.nms_padded(p1, 0.001, 0.65, multi_label=True), 5)
    lines.append(f'config 5: NMS stress nc=1, batch 256 x 25200 candidates, conf 0.001 iou 0.65: {ms1:.2f} ms = {256 / ms1 * 1e3:.0f} images/s, '
                 f'{256 * 25200 * 6 * 4 / ms1 / 1e6:.1f} GB/s of candidate rows')
    del p1
    p80 = stress_pred(32, 25200, 80)
    ms80 = timed(lambda: R.nms_padded(p80, 0.001, 0.65, multi_label=True), 3)
    lines.append(f'config 5 (nc=80 multi-label variant), batch 32 x 25200 x 85: {ms80:.2f} ms = {32 / ms80 * 1e3:.0f} images/s')
    txt = '\n'.join(lines)
    print(txt)
    if a.out:
        open(a.out, 'w').write(txt + '\n')


if __name__ == '__main__':
    main()
