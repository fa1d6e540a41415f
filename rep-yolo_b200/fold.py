"""Host-side fold pass: the arithmetic of the reference's ``Model.fuse()`` (models/yolo.py:681-704) in fp32.

  Conv+BN      utils/torch_utils.py:181-201   W' = diag(g/sqrt(eps+var)) W,  b' = beta - g*mu/sqrt(var+eps)
  RepS_Block   models/common.py:3462-3517     sum of 3x3 branches + padded 1x1 scale branch + BN-only skip
  RepConv      models/common.py:597-657       3x3 + padded 1x1 (+ identity BN when c1 == c2 and s == 1)
  IDetect      models/yolo.py:170-182         b <- (b + W.ia) * im ;  W <- W * im
BN eps is 1e-3 everywhere (utils/torch_utils.py:150).  Output keys are the reference's *fused* state_dict names.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

EPS = 1e-3


def _affine(sd, p):
    inv = torch.rsqrt(sd[f'{p}.running_var'].double() + EPS)
    scale = sd[f'{p}.weight'].double() * inv
    return scale, sd[f'{p}.bias'].double() - sd[f'{p}.running_mean'].double() * scale


def bn_affine(sd, p):
    """Stand-alone eval BatchNorm as (scale, shift) in fp32."""
    s, t = _affine(sd, p)
    return s.float(), t.float()


def _scaled(sd, wkey, bnp):
    s, t = _affine(sd, bnp)
    return sd[wkey].double() * s.view(-1, 1, 1, 1), t


def _eye(c, k):
    w = torch.zeros(c, c, k, k, dtype=torch.float64)
    i = torch.arange(c)
    w[i, i, k // 2, k // 2] = 1.0
    return w


def _reps(sd, p):
    W, b, j = 0.0, 0.0, 0
    while f'{p}.rbr_conv.{j}.conv.weight' in sd:
        w, t = _scaled(sd, f'{p}.rbr_conv.{j}.conv.weight', f'{p}.rbr_conv.{j}.bn')
        W, b, j = W + w, b + t, j + 1
    k = W.shape[-1]
    if f'{p}.rbr_scale.conv.weight' in sd:
        w, t = _scaled(sd, f'{p}.rbr_scale.conv.weight', f'{p}.rbr_scale.bn')
        W, b = W + F.pad(w, [k // 2] * 4), b + t
    if f'{p}.rbr_skip.weight' in sd:
        s, t = _affine(sd, f'{p}.rbr_skip')
        W, b = W + _eye(W.shape[0], k) * s.view(-1, 1, 1, 1), b + t
    return W.float(), b.float()


def _repconv(sd, p):
    w3, b3 = _scaled(sd, f'{p}.rbr_dense.0.weight', f'{p}.rbr_dense.1')
    w1, b1 = _scaled(sd, f'{p}.rbr_1x1.0.weight', f'{p}.rbr_1x1.1')
    W, b = w3 + F.pad(w1, [1] * 4), b3 + b1
    if f'{p}.rbr_identity.weight' in sd:
        s, t = _affine(sd, f'{p}.rbr_identity')
        W, b = W + _eye(W.shape[0], 3) * s.view(-1, 1, 1, 1), b + t
    return W.float(), b.float()


def fold_state_dict(sd, layers) -> OrderedDict:
    """unfused reference state_dict -> fused tensors (fp32, CPU)."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    out = OrderedDict()
    for key in sd:                       # every plain Conv (conv + bn); RepS/RepConv branches are handled below
        if key.endswith('.conv.weight') and '.rbr_' not in key and f'{key[:-12]}.bn.weight' in sd:
            p = key[:-12]
            w, b = _scaled(sd, key, f'{p}.bn')
            out[f'{p}.conv.weight'], out[f'{p}.conv.bias'] = w.float(), b.float()
    for L in layers:
        p = f'model.{L.i}'
        if L.kind == 'RepS_Block':
            out[f'{p}.reparam_conv.weight'], out[f'{p}.reparam_conv.bias'] = _reps(sd, p)
        elif L.kind == 'DER_Block':
            for s in range(1, 7):
                q = f'{p}.stage{s}.0'
                out[f'{q}.reparam_conv.weight'], out[f'{q}.reparam_conv.bias'] = _reps(sd, q)
        elif L.kind == 'RepConv':
            out[f'{p}.rbr_reparam.weight'], out[f'{p}.rbr_reparam.bias'] = _repconv(sd, p)
        elif L.kind == 'CA':
            out[f'{p}.f1.weight'], out[f'{p}.f2.weight'] = sd[f'{p}.f1.weight'].float(), sd[f'{p}.f2.weight'].float()
        elif L.kind == 'CCVA':
            for m in ('m', 'm1'):
                out[f'{p}.{m}.gamma'] = sd[f'{p}.{m}.gamma'].float()
                for bn in ('bn', 'bn1'):
                    for leaf in ('weight', 'bias', 'running_mean', 'running_var'):
                        out[f'{p}.{m}.{bn}.{leaf}'] = sd[f'{p}.{m}.{bn}.{leaf}'].float()
        elif L.kind == 'IDetect':
            for j in range(len(L.args[2])):
                W, b = sd[f'{p}.m.{j}.weight'].double(), sd[f'{p}.m.{j}.bias'].double()
                ia, im = sd[f'{p}.ia.{j}.implicit'].double().view(-1), sd[f'{p}.im.{j}.implicit'].double().view(-1)
                out[f'{p}.m.{j}.weight'] = (W * im.view(-1, 1, 1, 1)).float()
                out[f'{p}.m.{j}.bias'] = ((b + W.view(W.shape[0], -1) @ ia) * im).float()
            out[f'{p}.anchors'], out[f'{p}.anchor_grid'] = sd[f'{p}.anchors'].float(), sd[f'{p}.anchor_grid'].float()
    return out
