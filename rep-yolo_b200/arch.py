"""Rep-YOLO architecture description in the reference's model-yaml schema and its channel bookkeeping.

Host-side mirror of ``parse_model`` (reference models/yolo.py:730-836) for exactly the module kinds that
cfg/training/Rep-YOLO.yaml instantiates; anything else is rejected (the drop-in covers this one deploy path).
"""
from __future__ import annotations

import math
from collections import OrderedDict

STRIDES = (8.0, 16.0, 32.0)
ANCHORS = ((31, 30, 31, 37, 24, 61), (33, 63, 42, 56, 32, 111), (44, 114, 48, 172, 80, 112))

_SUPPORTED = ('RepS_Block', 'DER_Block', 'MP', 'SPPCSPC', 'GSConv', 'Upsample', 'Concat', 'VoVGSCSP', 'Conv', 'CA',
              'CCVA', 'ADD', 'RepConv', 'IDetect')


def rep_yolo_cfg(nc: int = 1) -> dict:
    """Same content as the reference's cfg/training/Rep-YOLO.yaml ({nc, anchors, backbone, head})."""
    rows, add = [], lambda f, m, a: rows.append([f, 1, m, list(a)])
    add(-1, 'RepS_Block', (48, 3, 2, 1))
    for c in (48, 128, 256, 512):
        add(-1, 'DER_Block', (c, 1, 2))
        add(-1, 'MP', ())
    n_backbone = len(rows)

    def attention_stage(mid, out):
        add(-1, 'Conv', (mid, 1, 1)); add(-1, 'CA', (mid,)); add(-2, 'CCVA', (mid,)); add([-1, -2], 'ADD', ())
        add(-1, 'Conv', (out, 1, 1))

    def gs_down(c, skip):
        add(-1, 'MP', ()); add(-1, 'GSConv', (c, 1, 1)); add(-3, 'GSConv', (c, 1, 1)); add(-1, 'GSConv', (c, 3, 2))
        add([-1, -3, skip], 'Concat', (1,))

    add(-1, 'SPPCSPC', (512,))
    add(-1, 'GSConv', (128, 1, 1)); add(-1, 'nn.Upsample', (None, 2, 'nearest')); add(6, 'GSConv', (256, 1, 1))
    add([-1, -2], 'Concat', (1,)); add(-1, 'VoVGSCSP', (256,))
    add(-1, 'GSConv', (128, 1, 1)); add(-1, 'nn.Upsample', (None, 2, 'nearest')); add(4, 'GSConv', (128, 1, 1))
    add([-1, -2], 'Concat', (1,))
    attention_stage(128, 256); add(-1, 'VoVGSCSP', (128,)); attention_stage(64, 128)
    gs_down(128, 14)
    attention_stage(256, 512); add(-1, 'VoVGSCSP', (256,)); attention_stage(128, 256)
    gs_down(256, 9)
    attention_stage(512, 1024); add(-1, 'VoVGSCSP', (512,)); attention_stage(256, 512)
    add(29, 'RepConv', (256, 3, 1)); add(45, 'RepConv', (512, 3, 1)); add(61, 'RepConv', (1024, 3, 1))
    add([62, 63, 64], 'IDetect', ('nc', 'anchors'))
    return {'nc': nc, 'depth_multiple': 1.0, 'width_multiple': 1.0, 'anchors': [list(a) for a in ANCHORS],
            'backbone': rows[:n_backbone], 'head': rows[n_backbone:]}


class Layer:
    __slots__ = ('i', 'f', 'kind', 'c1', 'c2', 'args')

    def __init__(self, i, f, kind, c1, c2, args):
        self.i, self.f, self.kind, self.c1, self.c2, self.args = i, f, kind, c1, c2, args

    def sources(self):
        """absolute indices of the producing layers (-1 = the image for layer 0)"""
        fs = [self.f] if isinstance(self.f, int) else list(self.f)
        return [(self.i + j if j < 0 else j) for j in fs]


def parse(cfg: dict, ch: int = 3):
    nc, anchors = cfg['nc'], cfg['anchors']
    if cfg.get('depth_multiple', 1.0) != 1.0 or cfg.get('width_multiple', 1.0) != 1.0:
        raise ValueError('only depth_multiple = width_multiple = 1.0 is supported')
    na = len(anchors[0]) // 2
    no = na * (nc + 5)
    layers, save, chs = [], set(), []
    for i, (f, n, kind, args) in enumerate(cfg['backbone'] + cfg['head']):
        kind = kind.replace('nn.', '') if isinstance(kind, str) else kind.__name__
        if kind not in _SUPPORTED or n != 1:
            raise ValueError(f'layer {i}: module {kind} (n={n}) is not part of the Rep-YOLO deploy path')
        args = list(args)
        cin = (lambda j: ch if i == 0 else chs[j])
        if kind in ('Conv', 'RepConv', 'SPPCSPC', 'GSConv', 'VoVGSCSP', 'CCVA'):
            c1, c2 = cin(f), args[0]
            if c2 != no:
                c2 = int(math.ceil(c2 / 8) * 8)
            args = [c1, c2] + args[1:]
        elif kind in ('RepS_Block', 'DER_Block'):
            c1, c2 = cin(f), args[0]
            args = [c1, c2] + args[1:]
        elif kind == 'Concat':
            c1, c2 = None, sum(chs[x] for x in f)
        elif kind == 'ADD':
            c1, c2 = None, chs[f[0]]
        elif kind == 'IDetect':
            c1, c2 = [chs[x] for x in f], None
            args = [nc, anchors, c1]
        else:
            c1 = c2 = cin(f)
        layers.append(Layer(i, f, kind, c1, c2, args))
        save.update(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        chs.append(c2)
    return layers, sorted(save)


# ---- unfused parameter inventory (names = the reference checkpoint's state_dict keys) ----
def _bn(out, p, c):
    for leaf, shape in (('weight', (c,)), ('bias', (c,)), ('running_mean', (c,)), ('running_var', (c,)),
                        ('num_batches_tracked', ())):
        out[f'{p}.{leaf}'] = shape


def _cb(out, p, c1, c2, k, g=1, conv='conv', bn='bn'):
    out[f'{p}.{conv}.weight'] = (c2, c1 // g, k, k)
    _bn(out, f'{p}.{bn}', c2)


def _reps(out, p, c1, c2, k, s, branches):
    if c1 == c2 and s == 1:
        _bn(out, f'{p}.rbr_skip', c1)
    for b in range(branches):
        _cb(out, f'{p}.rbr_conv.{b}', c1, c2, k)
    if k > 1:
        _cb(out, f'{p}.rbr_scale', c1, c2, 1)


def _gs(out, p, c1, c2, k):
    _cb(out, f'{p}.cv1', c1, c2 // 2, k)
    _cb(out, f'{p}.cv2', c2 // 2, c2 // 2, 5, g=c2 // 2)


def _attn(out, p, c):
    for name, co, g in (('query_conv', c // 8, c // 8), ('key_conv', c // 8, c // 8), ('value_conv', c, c)):
        _cb(out, f'{p}.{name}', c, co, 1, g=g)
    out[f'{p}.gamma'] = (1,)
    _bn(out, f'{p}.bn', c // 8)
    _bn(out, f'{p}.bn1', c)


def state_shapes(layers) -> OrderedDict:
    out = OrderedDict()
    for L in layers:
        p, a = f'model.{L.i}', L.args
        if L.kind == 'RepS_Block':
            _reps(out, p, a[0], a[1], a[2], a[3], 1)
        elif L.kind == 'DER_Block':
            c1, c2, br = a[0], a[1], a[3]
            _cb(out, f'{p}.cv1', 3 * c1, c2, 1)
            for j in range(4):
                _cb(out, f'{p}.cv{j}_1', c1, c1 // 2, 1)
                _cb(out, f'{p}.cv{j}_2', c1 // 2, c1, 1)
            for s in range(1, 7):
                c = c1 if s <= 3 else c1 // 2
                _reps(out, f'{p}.stage{s}.0', c, c, 3, 1, br)
        elif L.kind == 'SPPCSPC':
            c1, c2 = a[0], a[1]
            c_ = c2
            for name, ci, co, k in (('cv1', c1, c_, 1), ('cv2', c1, c_, 1), ('cv3', c_, c_, 3), ('cv4', c_, c_, 1),
                                    ('cv5', 4 * c_, c_, 1), ('cv6', c_, c_, 3), ('cv7', 2 * c_, c2, 1)):
                _cb(out, f'{p}.{name}', ci, co, k)
        elif L.kind == 'GSConv':
            _gs(out, p, a[0], a[1], a[2])
        elif L.kind == 'VoVGSCSP':
            c1, c2 = a[0], a[1]
            c_ = c2 // 2
            _cb(out, f'{p}.cv1', c1, c_, 1)
            _cb(out, f'{p}.cv2', c1, c_, 1)
            _gs(out, f'{p}.gsb.0.conv_lighting.0', c_, c_, 1)
            _gs(out, f'{p}.gsb.0.conv_lighting.1', c_, c_, 3)
            _cb(out, f'{p}.gsb.0.shortcut', c_, c_, 1)
            _cb(out, f'{p}.res', c_, c_, 3)
            _cb(out, f'{p}.cv3', 2 * c_, c2, 1)
        elif L.kind == 'Conv':
            _cb(out, p, a[0], a[1], a[2])
        elif L.kind == 'CA':
            out[f'{p}.f1.weight'] = (a[0] // 16, a[0], 1, 1)
            out[f'{p}.f2.weight'] = (a[0], a[0] // 16, 1, 1)
        elif L.kind == 'CCVA':
            c1, c2 = a[0], a[1]
            c_ = c2 // 2
            _cb(out, f'{p}.cv1', c1, c_, 1)
            _cb(out, f'{p}.cv2', c1, c_, 1)
            _cb(out, f'{p}.cv3', 2 * c_, c2, 1)
            _attn(out, f'{p}.m', c_)
            _attn(out, f'{p}.m1', c_)
        elif L.kind == 'RepConv':
            c1, c2, s = a[0], a[1], a[3]
            if c1 == c2 and s == 1:
                _bn(out, f'{p}.rbr_identity', c1)
            _cb(out, f'{p}.rbr_dense', c1, c2, 3, conv='0', bn='1')
            _cb(out, f'{p}.rbr_1x1', c1, c2, 1, conv='0', bn='1')
        elif L.kind == 'IDetect':
            nc, anchors, chs = a
            na, no = len(anchors[0]) // 2, nc + 5
            out[f'{p}.anchors'] = (len(anchors), na, 2)
            out[f'{p}.anchor_grid'] = (len(anchors), 1, na, 1, 1, 2)
            for j, c in enumerate(chs):
                out[f'{p}.m.{j}.weight'] = (no * na, c, 1, 1)
                out[f'{p}.m.{j}.bias'] = (no * na,)
            for j, c in enumerate(chs):
                out[f'{p}.ia.{j}.implicit'] = (1, c, 1, 1)
            for j, c in enumerate(chs):
                out[f'{p}.im.{j}.implicit'] = (1, no * na, 1, 1)
    return out
