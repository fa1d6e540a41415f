"""Generates the golden fixtures in this directory by running the REAL reference (DrLSB/Rep-YOLO).

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
  1. builds the reference ``models.yolo.Model('cfg/training/Rep-YOLO.yaml')`` (matplotlib/seaborn shimmed, they are
     plotting-only imports), and records its state_dict key/shape list  -> state_keys.json
  2. loads the oracle's synthetic 'calibrated' weights (oracle/repyolo_oracle.py:synth_state_dict + calibrate_bn_)
     into that reference model (strict=True), calls the reference ``Model.fuse()`` and records a digest of all fused
     weights/biases                                                     -> fold_digest.npz
  3. runs the reference fused forward on seeded 64x64 and 64x96 inputs and stores every layer's fp32 output,
     the decoded prediction and the raw head tensors                    -> layers_64.npz, layers_64x96.npz
  4. runs the reference ``utils.general.non_max_suppression`` on synthetic candidate sets (ties, zero-area boxes,
     empty images, multi-label, agnostic, class filter, > max_nms rows)   -> nms_cases.npz
  5. runs the reference head with ``include_nms`` (IDetect.convert(), models/yolo.py:189-199) and ``end2end`` set on
     the same 64x64 input                                               -> convert_64.npz
     (``python tests/golden/make_golden.py --only-convert`` regenerates just this file)
  6. runs the reference ``non_max_suppression(..., labels=[...])`` (apriori labels of test.py --save-hybrid, general.py:981-987)
                                                                        -> nms_labels.npz   (``--only-labels``)
"""
import contextlib
import io
import json
import logging
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)


def import_reference():
    for name in ('matplotlib', 'matplotlib.pyplot', 'seaborn'):        # plotting-only imports of utils/plots.py:11-25
        mod = types.ModuleType(name)
        mod.rc = mod.use = lambda *a, **k: None
        sys.modules[name] = mod
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.path.insert(0, REF)
    logging.disable(logging.CRITICAL)
    from models.yolo import Model                                        # noqa
    from utils.general import non_max_suppression                        # noqa
    return Model, non_max_suppression


def digest(t: torch.Tensor):
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, 8).long()
    return np.concatenate([[f.sum().item(), f.abs().sum().item(), (f * f).sum().item()], f[idx].numpy()])


def nms_cases():
    """name -> (pred [B,N,5+nc] float32, kwargs)"""
    g = torch.Generator().manual_seed(2024)

    def boxes(B, N, nc, span=640.0, clusters=0, quant=None):
        if clusters:
            ctr = torch.rand(B, clusters, 2, generator=g) * span
            pick = torch.randint(0, clusters, (B, N), generator=g)
            cxy = torch.gather(ctr, 1, pick[..., None].expand(B, N, 2)) + torch.randn(B, N, 2, generator=g) * 12.0
        else:
            cxy = torch.rand(B, N, 2, generator=g) * span
        wh = torch.exp(torch.empty(B, N, 2).uniform_(np.log(8.0), np.log(320.0), generator=g))
        obj = torch.rand(B, N, 1, generator=g)
        cls = torch.rand(B, N, nc, generator=g)
        p = torch.cat([cxy, wh, obj, cls], 2)
        if quant:
            p[..., :4] = (p[..., :4] / quant).round() * quant
        return p

    cases = {}
    cases['nc1_default'] = (boxes(2, 600, 1, clusters=12), dict(conf_thres=0.25, iou_thres=0.45))
    cases['nc1_testpy_maxdet'] = (boxes(2, 1500, 1), dict(conf_thres=0.001, iou_thres=0.65, multi_label=True))
    cases['nc4_multilabel'] = (boxes(2, 500, 4, clusters=10), dict(conf_thres=0.3, iou_thres=0.6, multi_label=True))
    cases['nc4_bestclass'] = (boxes(2, 500, 4, clusters=10), dict(conf_thres=0.25, iou_thres=0.45))
    cases['nc4_agnostic'] = (boxes(1, 500, 4, clusters=10), dict(conf_thres=0.25, iou_thres=0.45, agnostic=True))
    cases['nc4_classes'] = (boxes(1, 500, 4, clusters=10), dict(conf_thres=0.25, iou_thres=0.45, classes=[1, 3]))
    p = boxes(2, 800, 1, clusters=8, quant=4.0)
    p[..., 4] = (p[..., 4] * 20).round() / 20                          # heavy exact score ties
    cases['nc1_ties'] = (p, dict(conf_thres=0.25, iou_thres=0.5))
    p = boxes(3, 300, 1, clusters=5)
    p[1, :, 4] = 0.1                                                   # image 1: nothing passes -> (0,6)
    p[2, :50, 2] = 0.0                                                 # zero-area boxes (NaN IoU among themselves)
    p[2, :50, 3] = 0.0
    p[2, :50, :2] = 100.0
    cases['nc1_empty_zeroarea'] = (p, dict(conf_thres=0.25, iou_thres=0.45))
    cases['nc20_over_maxnms'] = (boxes(1, 2500, 20), dict(conf_thres=0.001, iou_thres=0.65, multi_label=True))
    # IoU exactly at the threshold: unit-offset squares, IoU = 0.6 in fp32 for these coordinates (SURVEY.md 8c probe)
    p = torch.zeros(1, 4, 6)
    p[0, :, 2:4] = 4.0
    p[0, :, 0] = torch.tensor([10.0, 11.0, 30.0, 31.0])                # IoU(4x4 shifted by 1) = 12/20 = 0.6
    p[0, :, 1] = 10.0
    p[0, :, 4] = torch.tensor([0.9, 0.8, 0.7, 0.6])
    p[0, :, 5] = 1.0
    cases['nc1_iou_at_threshold'] = (p, dict(conf_thres=0.25, iou_thres=0.6))
    return cases


def main():
    from oracle import repyolo_oracle as O
    Model, ref_nms = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)
        m = Model(os.path.join(REF, 'cfg/training/Rep-YOLO.yaml'), ch=3).eval()
    ref_sd = m.state_dict()
    keys = {'unfused': [[k, list(v.shape)] for k, v in ref_sd.items()], 'save': sorted(set(m.save)),
            'stride': m.stride.tolist()}

    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated')
    m.load_state_dict(sd, strict=True)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        m.fuse()
    fsd = m.state_dict()
    keys['fused'] = [[k, list(v.shape)] for k, v in fsd.items()]
    with open(os.path.join(HERE, 'state_keys.json'), 'w') as f:
        json.dump(keys, f)
    np.savez_compressed(os.path.join(HERE, 'fold_digest.npz'),
                        **{k: digest(v) for k, v in fsd.items() if v.dtype.is_floating_point})

    # 6. apriori labels (test.py --save-hybrid autolabelling, general.py:981-987): label rows are appended AFTER the confidence filter
    if '--only-labels' in sys.argv or '--only-convert' not in sys.argv:
        g = torch.Generator().manual_seed(77)
        B, N, nc = 3, 400, 4
        cxy = torch.rand(B, N, 2, generator=g) * 320
        wh = torch.rand(B, N, 2, generator=g) * 60 + 8
        lp = torch.cat([cxy, wh, torch.rand(B, N, 1, generator=g), torch.rand(B, N, nc, generator=g)], 2)
        labels = []
        for b, nl in enumerate((5, 0, 2)):                                 # image 1 has no labels
            l = torch.cat([torch.randint(0, nc, (nl, 1), generator=g).float(), lp[b, :nl, :4] + 3.0], 1)   # overlapping real candidates
            labels.append(l)
        blob = {'pred': lp.numpy().astype(np.float32)}
        for ci, kw in enumerate((dict(conf_thres=0.25, iou_thres=0.45), dict(conf_thres=0.1, iou_thres=0.6, multi_label=True))):
            outs = []
            for b in range(B):
                outs += ref_nms(lp[b:b + 1].clone(), labels=[labels[b]], **kw)
            blob[f'c{ci}.counts'] = np.array([o.shape[0] for o in outs], dtype=np.int32)
            blob[f'c{ci}.out'] = torch.cat(outs, 0).numpy().astype(np.float32)
            blob[f'c{ci}.kw'] = np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)
        for b in range(B):
            blob[f'labels{b}'] = labels[b].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, 'nms_labels.npz'), **blob)
        print('nms_labels.npz', [blob[f'c{c}.counts'].tolist() for c in range(2)])
        if '--only-labels' in sys.argv:
            return

    # 5. alternative output contracts of IDetect.fuseforward (yolo.py:158-166) on the tag-'64' input
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(5))
    det = m.model[-1]
    with torch.no_grad():
        det.include_nms = True
        (box, score), = m(x)
        det.include_nms, det.end2end = False, True
        e2e = m(x)
        det.end2end = False
        pred_plain, _ = m(x)
    np.savez_compressed(os.path.join(HERE, 'convert_64.npz'), x=x.numpy(), box=box.numpy(), score=score.numpy(), end2end=e2e.numpy(),
                        pred=pred_plain.numpy())
    print('convert_64.npz', tuple(box.shape), tuple(score.shape), tuple(e2e.shape))
    if '--only-convert' in sys.argv:
        return

    for tag, (h, w), seed in (('64', (64, 64), 5), ('64x96', (64, 96), 6)):
        x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(seed))
        taps = {}
        hooks = [mod.register_forward_hook(lambda mod, i, o, idx=idx: taps.__setitem__(idx, o))
                 for idx, mod in enumerate(m.model)]
        with torch.no_grad():
            pred, raws = m(x)
        for hk in hooks:
            hk.remove()
        blob = {'x': x.numpy(), 'pred': pred.numpy()}
        for j, r in enumerate(raws):
            blob[f'raw{j}'] = r.numpy()
        for i in range(len(m.model) - 1):
            if tag == '64' or layers[i]['kind'] in ('CCVA', 'RepConv'):
                blob[f'layer{i}'] = taps[i].numpy()
        np.savez_compressed(os.path.join(HERE, f'layers_{tag}.npz'), **blob)
        if tag == '64':
            model_pred = pred

    blob, meta = {}, {}
    cases = nms_cases()
    cases['model_pred_default'] = (model_pred, dict(conf_thres=0.25, iou_thres=0.45))
    cases['model_pred_testpy'] = (model_pred, dict(conf_thres=0.001, iou_thres=0.65, multi_label=True))
    for name, (p, kw) in cases.items():
        outs = []
        for b in range(p.shape[0]):                                    # one image per call: the 10 s time_limit never fires
            outs += ref_nms(p[b:b + 1].clone(), **kw)
        blob[f'{name}.pred'] = p.numpy().astype(np.float32)
        blob[f'{name}.counts'] = np.array([o.shape[0] for o in outs], dtype=np.int32)
        blob[f'{name}.out'] = torch.cat(outs, 0).numpy().astype(np.float32)
        meta[name] = kw
        print(name, 'counts', blob[f'{name}.counts'].tolist())
    blob['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, 'nms_cases.npz'), **blob)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == '__main__':
    main()
