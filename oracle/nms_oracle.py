"""ctypes front-end of oracle/nms_oracle.c (CPU oracle for non_max_suppression) -- TEST INFRASTRUCTURE ONLY.

Mirrors the reference signature ``non_max_suppression(prediction, conf_thres, iou_thres, classes, agnostic,
multi_label, labels)`` (utils/general.py:953) and returns the same ``list`` of ``(n_i, 6)`` float32 tensors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libnms_oracle.so')
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'nms_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B', '_build/libnms_oracle.so'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ry_oracle_greedy_nms.restype = ctypes.c_int
        _lib.ry_oracle_greedy_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_void_p]
        _lib.ry_oracle_nms_batch.restype = None
        _lib.ry_oracle_nms_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def greedy_nms(boxes, scores, iou_thres: float, max_keep: int | None = None) -> np.ndarray:
    """torchvision.ops.nms restatement: kept indices in score order (int64)."""
    boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float32).reshape(-1, 4))
    scores = np.ascontiguousarray(np.asarray(scores, dtype=np.float32).reshape(-1))
    n = scores.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int32)
    k = lib().ry_oracle_greedy_nms(boxes.ctypes.data, scores.ctypes.data, n, float(iou_thres),
                                   n if max_keep is None else int(max_keep), keep.ctypes.data)
    return keep[:k].astype(np.int64)


def with_apriori_labels(prediction, labels):
    """general.py:981-987: per image the label rows (cls, x, y, w, h) become candidates [box, conf = 1, one-hot class] appended
    BEHIND the rows that pass the confidence filter.  Restated as a dense tensor: Lmax extra rows per image behind the N
    candidates; unused rows carry obj = -inf, which never passes `obj > conf_thres`."""
    if labels is None or not len(labels) or not any(len(l) for l in labels):
        return prediction
    B, n, no = prediction.shape
    lmax = max(len(l) for l in labels)
    extra = torch.zeros((B, lmax, no), dtype=prediction.dtype)
    extra[:, :, 4] = float('-inf')
    for b, l in enumerate(labels):
        l = torch.as_tensor(l, dtype=torch.float32).reshape(-1, 5)
        if len(l):
            extra[b, :len(l), :4] = l[:, 1:5]
            extra[b, :len(l), 4] = 1.0
            extra[b, torch.arange(len(l)), l[:, 0].long() + 5] = 1.0
    return torch.cat([prediction, extra], 1)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, max_nms=30000):
    prediction = with_apriori_labels(prediction.detach().cpu().float(), labels)
    p = np.ascontiguousarray(prediction.numpy().astype(np.float32, copy=False))
    B, n, no = p.shape
    out = np.zeros((B, max_det, 6), dtype=np.float32)
    counts = np.zeros(B, dtype=np.int32)
    cls = np.ascontiguousarray(np.asarray(classes if classes is not None else [], dtype=np.int32))
    lib().ry_oracle_nms_batch(p.ctypes.data, B, n, no - 5, float(np.float32(conf_thres)), float(iou_thres),
                              cls.ctypes.data if cls.size else None, int(cls.size), int(bool(agnostic)),
                              int(bool(multi_label)), max_det, max_nms, out.ctypes.data, counts.ctypes.data)
    return [torch.from_numpy(out[b, :counts[b]].copy()) for b in range(B)]
