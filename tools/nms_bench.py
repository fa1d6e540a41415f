#!/usr/bin/env python
"""NMS timing legs on the GPU box (CUDA events, 20 reps): model-produced candidates of both inits at batch 64 @ 640,
detect.py and test.py thresholds, plain (pred -> ry_nms) and fused-filter (ry_decode_filter mask -> ry_nms_filtered) front ends,
and BASELINE configs[4] (256 x 25200 synthetic candidates, conf 0.001 / iou 0.65).
    python tools/nms_bench.py [--out gpurun_out/nms_bench.txt]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import repyolo_b200 as R  # noqa: E402
from oracle import repyolo_oracle as O  # noqa: E402  (weights generator only)


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    lines = []
    x = torch.rand(64, 3, 640, 640, device='cuda')
    for init in ('calibrated', 'default'):
        _, _, sd, _ = O.make_model(0, init)
        m = R.Model()
        m.load_state_dict(sd, strict=True)
        m.fuse()
        for conf, iou in ((0.25, 0.45), (0.001, 0.65)):
            m.decode_filter = None
            plain, _ = m(x)
            m.decode_filter = conf
            filt, _ = m(x)
            ncand = float((plain[..., 4] > conf).sum()) / 64
            tp = timed(lambda: R.nms_padded(plain, conf, iou))
            tf = timed(lambda: R.nms_padded(filt, conf, iou))
            o1, c1 = R.nms_padded(plain, conf, iou)
            o2, c2 = R.nms_padded(filt, conf, iou)
            same = torch.equal(c1, c2) and all(torch.equal(o1[i, :c], o2[i, :c]) for i, c in enumerate(c1.tolist()))
            lines.append(f'{init:10s} conf {conf:5.3f} iou {iou:.2f}: candidates/img {ncand:8.1f} dets/img {float(c1.float().mean()):6.1f} '
                         f'plain {tp * 1e3:7.1f} us  fused-filter {tf * 1e3:7.1f} us  identical {same}')
        del m
    g = torch.Generator().manual_seed(0)
    B, N = 256, 25200
    cxy = torch.rand(B, N, 2, generator=g) * 640
    wh = torch.exp(torch.empty(B, N, 2).uniform_(np.log(8.0), np.log(320.0), generator=g))
    pred = torch.cat([cxy, wh, torch.rand(B, N, 1, generator=g), torch.rand(B, N, 1, generator=g)], 2).cuda()
    t = timed(lambda: R.nms_padded(pred, 0.001, 0.65, multi_label=True), reps=10)
    lines.append(f'configs[4] 256 x 25200 synthetic, conf 0.001 iou 0.65: {t:7.3f} ms = {B / t * 1e3:9.0f} images/s, '
                 f'{pred.numel() * 4 / t / 1e6:7.1f} GB/s of candidate rows')
    txt = '\n'.join(lines)
    print(txt)
    if a.out:
        open(a.out, 'w').write(txt + '\n')


if __name__ == '__main__':
    main()
