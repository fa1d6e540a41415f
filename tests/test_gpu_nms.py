"""-m gpu: ry_nms (through the Python mirror of non_max_suppression) must be BIT-EXACT vs the CPU oracle / reference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import nms_oracle

pytestmark = pytest.mark.gpu


def _same(a_list, b_list, tag):
    assert len(a_list) == len(b_list)
    for i, (a, b) in enumerate(zip(a_list, b_list)):
        a = a.cpu().numpy()
        b = b.cpu().numpy()
        assert a.shape == b.shape, (tag, i, a.shape, b.shape)
        assert a.tobytes() == b.tobytes(), (tag, i)


def test_golden_reference_outputs():
    import repyolo_b200 as R
    g = np.load(os.path.join(GOLDEN, 'nms_cases.npz'))
    meta = json.loads(bytes(g['meta']).decode())
    for name, kw in meta.items():
        pred = torch.from_numpy(g[f'{name}.pred']).cuda()
        outs = R.non_max_suppression(pred, **kw)
        counts = g[f'{name}.counts']
        assert [o.shape[0] for o in outs] == counts.tolist(), name
        got = torch.cat(outs, 0).cpu().numpy()
        assert got.tobytes() == g[f'{name}.out'].tobytes(), name


def _synthetic(B, N, nc, seed, tie=False, span=640.0):
    g = torch.Generator().manual_seed(seed)
    cxy = torch.rand(B, N, 2, generator=g) * span
    wh = torch.exp(torch.empty(B, N, 2).uniform_(np.log(8.0), np.log(320.0), generator=g))
    obj = torch.rand(B, N, 1, generator=g)
    cls = torch.rand(B, N, nc, generator=g)
    if tie:
        obj = (obj * 20).round() / 20
        cxy = (cxy / 16).round() * 16
        wh = (wh / 16).round() * 16 + 16
    return torch.cat([cxy, wh, obj, cls], 2)


@pytest.mark.parametrize('N,nc,conf,iou,ml,tie', [
    (25200, 1, 0.25, 0.45, False, False),      # detect.py settings at the 640x640 candidate count
    (25200, 1, 0.001, 0.65, True, False),      # test.py settings (config 5 of BASELINE.json)
    (25200, 1, 0.001, 0.65, True, True),       # tie-heavy variant (scores quantised to 1/20)
    (6000, 4, 0.1, 0.6, True, False),          # multi-label
    (6000, 4, 0.25, 0.45, False, True),
    (5000, 1, 0.5, 0.5, False, False),
    (1, 1, 0.25, 0.45, False, False),
])
def test_synthetic_vs_oracle(N, nc, conf, iou, ml, tie):
    import repyolo_b200 as R
    pred = _synthetic(3, N, nc, seed=N + nc, tie=tie)
    ref = nms_oracle.non_max_suppression(pred, conf, iou, multi_label=ml)
    got = R.non_max_suppression(pred.cuda(), conf, iou, multi_label=ml)
    _same(got, ref, (N, nc, conf, iou, ml, tie))


def test_over_max_nms_multilabel():
    import repyolo_b200 as R
    pred = _synthetic(2, 4000, 20, seed=9)     # up to 80000 rows per image > max_nms = 30000
    ref = nms_oracle.non_max_suppression(pred, 0.001, 0.65, multi_label=True)
    got = R.non_max_suppression(pred.cuda(), 0.001, 0.65, multi_label=True)
    _same(got, ref, 'over_max_nms')


def test_agnostic_classes_empty():
    import repyolo_b200 as R
    pred = _synthetic(2, 3000, 4, seed=4)
    pred[1, :, 4] = 0.0                        # image 1 has no candidate -> (0, 6)
    for kw in (dict(agnostic=True), dict(classes=[1, 3]), dict(classes=[2], agnostic=True)):
        ref = nms_oracle.non_max_suppression(pred, 0.25, 0.45, **kw)
        got = R.non_max_suppression(pred.cuda(), 0.25, 0.45, **kw)
        _same(got, ref, kw)
        assert got[1].shape == (0, 6)


def test_full_size_properties():
    """BASELINE config 5 size (256 x 25200): size-independent properties + oracle check on a sample of images."""
    import repyolo_b200 as R
    pred = _synthetic(256, 25200, 1, seed=0)
    dets = R.non_max_suppression(pred.cuda(), 0.001, 0.65, multi_label=True)
    assert len(dets) == 256
    for d in dets:
        d = d.cpu()
        assert d.shape[0] <= 300 and d.shape[1] == 6
        assert bool((d[1:, 4] <= d[:-1, 4]).all())                      # sorted by confidence
        assert bool((d[:, 2] >= d[:, 0]).all() and (d[:, 3] >= d[:, 1]).all())
    idx = [0, 17, 100, 255]
    ref = nms_oracle.non_max_suppression(pred[idx], 0.001, 0.65, multi_label=True)
    _same([dets[i] for i in idx], ref, 'config5 sample')
    # idempotence: NMS of its own survivors (as cx,cy,w,h rows) keeps them all
    d = dets[0]
    rows = torch.stack([(d[:, 0] + d[:, 2]) / 2, (d[:, 1] + d[:, 3]) / 2, d[:, 2] - d[:, 0], d[:, 3] - d[:, 1], d[:, 4], d[:, 4]], 1)
    again = R.non_max_suppression(rows[None].contiguous(), 0.001, 0.65)
    assert again[0].shape[0] >= d.shape[0] - 2                          # re-derived xyxy can move by an ulp


def test_empty_class_list_matches_nothing():
    """general.py:1012-1013: classes=[] keeps no row (classes=None means no filter)."""
    import repyolo_b200 as R
    pred = torch.rand(2, 500, 6, generator=torch.Generator().manual_seed(1)).cuda()
    pred[..., :4] *= 300
    assert all(d.shape == (0, 6) for d in R.non_max_suppression(pred, 0.25, 0.45, classes=[]))
    assert any(d.shape[0] > 0 for d in R.non_max_suppression(pred, 0.25, 0.45, classes=None))


# ---- fused decode + confidence filter (ry_decode_filter / ry_nms_filtered, north_star (c)) ----
def _host_mask(pred, conf):
    """The candidate mask the Detect epilogue would leave for `pred`: bit (i & 31) of word (i >> 5) = pred[b, i, 4] > conf."""
    B, N, _ = pred.shape
    words = (N + 31) // 32
    xc = (pred[..., 4] > np.float32(conf)).cpu().numpy()
    bits = np.zeros((B, words * 32), dtype=np.uint8)
    bits[:, :N] = xc
    packed = np.packbits(bits.reshape(B, words, 32), axis=2, bitorder='little').view(np.uint32).reshape(B, words)
    return torch.from_numpy(packed.view(np.int32).copy())


def _with_mask(pred, conf, mask_conf=None):
    pred = pred.cuda().contiguous()
    pred._ry_cand = (_host_mask(pred, conf if mask_conf is None else mask_conf).cuda(), float(conf if mask_conf is None else mask_conf),
                     pred.data_ptr())
    return pred


def test_filtered_front_end_matches_plain_on_every_case():
    """ry_nms_filtered (ordered compaction from the mask words) == ry_nms (every row tested): same bytes on the reference-minted
    golden cases and on the synthetic / multi-label / class-filter / agnostic / > max_nms cases above."""
    import repyolo_b200 as R
    g = np.load(os.path.join(GOLDEN, 'nms_cases.npz'))
    meta = json.loads(bytes(g['meta']).decode())
    for name, kw in meta.items():
        pred = torch.from_numpy(g[f'{name}.pred'])
        conf = kw.get('conf_thres', 0.25)
        outs = R.non_max_suppression(_with_mask(pred, conf), **kw)
        got = torch.cat(outs, 0).cpu().numpy()
        assert [o.shape[0] for o in outs] == g[f'{name}.counts'].tolist(), name
        assert got.tobytes() == g[f'{name}.out'].tobytes(), name
    for N, nc, conf, iou, kw in [(25200, 1, 0.25, 0.45, {}), (25200, 1, 0.001, 0.65, dict(multi_label=True)),
                                 (6000, 4, 0.1, 0.6, dict(multi_label=True)), (6000, 4, 0.25, 0.45, dict(classes=[1, 3])),
                                 (3000, 4, 0.25, 0.45, dict(agnostic=True)), (4000, 20, 0.001, 0.65, dict(multi_label=True)),
                                 (37, 1, 0.25, 0.45, {}), (1, 1, 0.25, 0.45, {})]:
        pred = _synthetic(3, N, nc, seed=7 * N + nc, tie=(N == 6000))
        pred[1, :, 4] *= 0.0 if N == 3000 else 1.0
        plain = R.non_max_suppression(pred.cuda(), conf, iou, **kw)
        _same(R.non_max_suppression(_with_mask(pred, conf), conf, iou, **kw), plain, ('filtered', N, nc, conf))
        # a mask made with a LOWER threshold is a superset: still exact
        _same(R.non_max_suppression(_with_mask(pred, conf, mask_conf=conf * 0.5), conf, iou, **kw), plain, ('superset', N, nc, conf))


@pytest.mark.parametrize('B,H,W,nc,conf', [(2, 640, 640, 1, 0.25), (3, 96, 160, 1, 0.001), (5, 64, 64, 1, 0.25), (2, 128, 96, 2, 0.1),
                                           (64, 640, 640, 1, 0.25)])
def test_decode_filter_mask_and_detections(B, H, W, nc, conf):
    """Model.decode_filter: same pred as the plain forward, mask == (pred[..., 4] > conf) bit for bit (image boundaries inside
    a warp at the small maps, generic head nc = 2, BASELINE batch 64), detections byte-identical to pred -> ry_nms."""
    import repyolo_b200 as R
    from oracle import repyolo_oracle as O
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated', nc=nc)
    m = R.Model(nc=nc)
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(B * H + W)).cuda()
    plain, _ = m(x)
    assert not hasattr(plain, '_ry_cand')
    m.decode_filter = conf
    pred, _ = m(x)
    m.decode_filter = None
    assert torch.equal(pred, plain)
    mask, mconf, ptr = pred._ry_cand
    assert mconf == conf and ptr == pred.data_ptr()
    assert torch.equal(mask.cpu(), _host_mask(pred, conf))
    assert int((pred[..., 4] > conf).sum()) > 0
    for c2, iou in ((conf, 0.45), (max(conf, 0.3), 0.65)):
        _same(R.non_max_suppression(pred, c2, iou), R.non_max_suppression(plain, c2, iou), ('decode_filter', B, H, W, nc, c2))


def test_apriori_labels_match_reference():
    """non_max_suppression(labels=[...]) (test.py --save-hybrid, general.py:981-987) against the reference-minted fixture."""
    import repyolo_b200 as R
    g = np.load(os.path.join(GOLDEN, 'nms_labels.npz'))
    pred = torch.from_numpy(g['pred']).cuda()
    labels = [torch.from_numpy(g[f'labels{b}']).cuda() for b in range(pred.shape[0])]
    for c in range(2):
        kw = json.loads(bytes(g[f'c{c}.kw']).decode())
        outs = R.non_max_suppression(pred, labels=labels, **kw)
        assert [o.shape[0] for o in outs] == g[f'c{c}.counts'].tolist()
        assert torch.cat(outs, 0).cpu().numpy().tobytes() == g[f'c{c}.out'].tobytes()
