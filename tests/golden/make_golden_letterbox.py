"""Golden fixtures for the pre-/post-processing row, minted from the REAL reference.

Run in the build container only (needs /root/reference and its cv2 dependency; neither is read on the GPU box):

    python tests/golden/make_golden_letterbox.py

For seeded synthetic BGR images (regenerated from the seed by the tests, not stored) it stores what the reference computes:
  * ``utils.datasets.letterbox`` (cv2.resize INTER_LINEAR + cv2.copyMakeBorder) followed by the BGR->RGB / HWC->CHW packing of
    ``LoadImages.__next__`` (datasets.py:191-195), for several (shape, img_size, auto, scaleFill, scaleup) cases
  * ``utils.general.scale_coords`` (+ clip_coords, then ``.round()`` as detect.py:114) on seeded fp32 boxes (CPU torch)
-> letterbox_cases.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'

CASES = [
    # (H0, W0), new_shape, auto, scaleFill, scaleup
    ((37, 53), 64, True, False, True),
    ((120, 200), 96, True, False, True),
    ((200, 120), 96, True, False, True),
    ((48, 64), 128, True, False, True),           # up-scaling
    ((48, 64), 128, True, False, False),          # scaleup=False: pad only
    ((90, 160), (96, 160), False, False, True),   # test.py style: fixed rectangular shape, auto=False
    ((100, 75), 64, False, True, True),           # scaleFill
    ((128, 256), 128, True, False, True),         # exact 2x down-scale (cv2 switches to its INTER_AREA fast path)
    ((271, 333), 160, True, False, True),
]


def image(i, shape):
    return np.random.default_rng(1000 + i).integers(0, 256, (shape[0], shape[1], 3), dtype=np.uint8)


def boxes(i, n=40):
    g = torch.Generator().manual_seed(2000 + i)
    c = torch.rand(n, 2, generator=g) * 200.0 - 20.0
    wh = torch.rand(n, 2, generator=g) * 120.0
    return torch.cat([c - wh / 2, c + wh / 2], 1)


def main():
    for name in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):        # plotting-only imports of the reference
        mod = types.ModuleType(name)
        mod.rc = mod.use = lambda *a, **k: None
        sys.modules[name] = mod
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.path.insert(0, REF)
    from utils.datasets import letterbox
    from utils.general import scale_coords
    out = {}
    for i, (shape, new_shape, auto, fill, up) in enumerate(CASES):
        img0 = image(i, shape)
        img, ratio, pad = letterbox(img0, new_shape, auto=auto, scaleFill=fill, scaleup=up, stride=32)
        chw = np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))
        out[f'img_{i}'] = chw
        out[f'ratio_pad_{i}'] = np.array([ratio[0], ratio[1], pad[0], pad[1]], np.float64)
        b = boxes(i)
        sc = scale_coords(chw.shape[1:], b.clone(), img0.shape)
        out[f'scaled_{i}'] = sc.numpy()
        out[f'scaled_round_{i}'] = sc.round().numpy()
        sc2 = scale_coords(chw.shape[1:], b.clone(), img0.shape, ratio_pad=(ratio, pad))      # test.py:141 passes shapes[i][1]
        out[f'scaled_rp_{i}'] = sc2.numpy()
    np.savez_compressed(os.path.join(HERE, 'letterbox_cases.npz'), **out)
    print('wrote letterbox_cases.npz', sum(v.nbytes for v in out.values()), 'bytes (uncompressed)')


if __name__ == '__main__':
    main()
