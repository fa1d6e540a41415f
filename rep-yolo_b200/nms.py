"""``non_max_suppression`` with the reference signature (utils/general.py:953), executed by the CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as N

MAX_DET = 300      # utils/general.py:966
MAX_NMS = 30000    # utils/general.py:967
_ws_cache = {}


def _workspace(device, B, n, nc, multi_label):
    key = (device, B, n, nc, multi_label)
    ws = _ws_cache.get(key)
    if ws is None:
        nbytes = C.c_size_t()
        N.check(N.lib().ry_nms_workspace_bytes(B, n, nc, int(multi_label), C.byref(nbytes)), 'ry_nms_workspace_bytes')
        if len(_ws_cache) > 8:
            _ws_cache.clear()
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def nms_launch_count(B, n, nc, multi_label=False):
    """kernel launches of one nms_padded / non_max_suppression call (bench.py's gpu_launches claim)"""
    k = C.c_int()
    N.check(N.lib().ry_nms_launch_count(B, n, nc, int(bool(multi_label) and nc > 1), C.byref(k)), 'ry_nms_launch_count')
    return k.value


def nms_padded(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
               max_det=MAX_DET, max_nms=MAX_NMS, out=None, counts=None):
    """Device-side result without any host sync: (out [B, max_det, 6] fp32, counts [B] int32).
    ``out`` / ``counts``: optional preallocated contiguous destinations (e.g. views of a gather payload)."""
    if not prediction.is_cuda:
        raise N.NativeError('non_max_suppression: prediction must be a CUDA tensor (no CPU fallback on this path)')
    # candidate mask left by the fused decode + filter (Model.decode_filter -> ry_decode_filter): usable when it was produced
    # with a threshold <= this call's (the mask is then a superset of `obj > conf_thres`; every row is re-tested exactly)
    cand = getattr(prediction, '_ry_cand', None)
    p = prediction.detach()
    if p.dtype != torch.float32 or not p.is_contiguous():
        p = p.float().contiguous()
        cand = None
    B, n, no = p.shape
    if cand is not None and not (cand[0].shape == (B, (n + 31) // 32) and cand[0].device == p.device and
                                 float(np.float32(cand[1])) <= float(np.float32(conf_thres)) and cand[2] == p.data_ptr()):
        cand = None
    nc = no - 5
    multi_label = bool(multi_label) and nc > 1
    if out is None:
        out = torch.empty((B, max_det, 6), dtype=torch.float32, device=p.device)
    if counts is None:                                  # ry_nms writes every entry: no fill kernel on the hot path
        counts = torch.empty((B,), dtype=torch.int32, device=p.device)
    if (tuple(out.shape) != (B, max_det, 6) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != p.device or
            tuple(counts.shape) != (B,) or counts.dtype != torch.int32 or not counts.is_contiguous() or counts.device != p.device):
        raise ValueError('nms_padded: out must be contiguous fp32 [B, max_det, 6] and counts int32 [B] on the prediction device')
    if B == 0 or n == 0:
        counts.zero_()
        return out, counts
    if classes is not None and len(classes) == 0:
        counts.zero_()
        return out, counts          # general.py:1012-1013: an empty class list matches no row (None = no filter)
    ws = _workspace(p.device, B, n, nc, multi_label)
    cls = np.ascontiguousarray(np.asarray(classes if classes is not None else [], dtype=np.int32).reshape(-1))
    with torch.cuda.device(p.device):
        st = torch.cuda.current_stream(p.device).cuda_stream
        tail = (C.c_float(float(np.float32(conf_thres))), C.c_double(float(iou_thres)), cls.ctypes.data if cls.size else None,
                int(cls.size), int(bool(agnostic)), int(multi_label), int(max_det), int(max_nms), out.data_ptr(), counts.data_ptr(),
                ws.data_ptr(), ws.numel(), C.c_void_p(st))
        if cand is not None:
            N.check(N.lib().ry_nms_filtered(p.data_ptr(), cand[0].data_ptr(), B, n, nc, *tail), 'ry_nms_filtered')
        else:
            N.check(N.lib().ry_nms(p.data_ptr(), B, n, nc, *tail), 'ry_nms')
    return out, counts


def _with_apriori_labels(prediction, labels):
    B, n, no = prediction.shape
    if len(labels) != B:
        raise ValueError('labels: one (n_i, 5) tensor per image')
    lmax = max(len(l) for l in labels)
    extra = torch.zeros((B, lmax, no), dtype=torch.float32, device=prediction.device)
    extra[:, :, 4] = float('-inf')
    for b, l in enumerate(labels):
        l = torch.as_tensor(l, dtype=torch.float32, device=prediction.device).reshape(-1, 5)
        if len(l):
            extra[b, :len(l), :4] = l[:, 1:5]
            extra[b, :len(l), 4] = 1.0
            extra[b, torch.arange(len(l), device=l.device), l[:, 0].long() + 5] = 1.0
    return torch.cat([prediction.detach().float(), extra], 1)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=()):
    """Runs NMS on inference results; returns a list of (n, 6) tensors [xyxy, conf, cls] per image, like the reference.

    Differences from the reference, by design: no 10 s ``time_limit`` break, and the > 30000-row cut is the stable one.
    ``labels`` (test.py --save-hybrid autolabelling, general.py:981-987): per image the label rows (cls, x, y, w, h) are appended
    behind the candidates as [box, conf = 1, one-hot class] rows of the dense tensor (unused rows: obj = -inf, never pass).
    """
    if labels is not None and len(labels) and any(len(l) for l in labels):
        prediction = _with_apriori_labels(prediction, labels)
    out, counts = nms_padded(prediction, conf_thres, iou_thres, classes, agnostic, multi_label)
    cnt = counts.cpu().tolist()                       # the one device->host read of this call
    return [out[i, :c] for i, c in enumerate(cnt)]
