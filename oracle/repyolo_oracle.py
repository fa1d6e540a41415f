"""CPU oracle for the Rep-YOLO deployed inference path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-fp32 *restatement* of what the reference computes on the path
    Model.fuse() -> Model.forward() -> IDetect.fuseforward()
It is never imported by the product package (``rep-yolo_b200/``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference from
/root/reference (in the build container), loads the synthetic weights produced here into the
reference ``models.yolo.Model``, and stores the reference's own outputs (state-dict key list, fused
weights digest, per-layer fp32 activations, decoded predictions) under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this restatement against those fixtures.

Every function cites the reference file:line (relative to the reference repo root) it restates.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # utils/torch_utils.py:150  (initialize_weights sets eps=1e-3 on every BatchNorm2d)

# --------------------------------------------------------------------------------------
# Architecture (restates cfg/training/Rep-YOLO.yaml:1-117 in the reference's own schema)
# --------------------------------------------------------------------------------------
ANCHORS = [[31, 30, 31, 37, 24, 61], [33, 63, 42, 56, 32, 111], [44, 114, 48, 172, 80, 112]]


def default_cfg(nc: int = 1) -> dict:
    """The Rep-YOLO model dictionary ({nc, anchors, backbone, head}); rows are [from, number, module, args]."""
    def attn_stage(c_mid, c_out):  # Conv -> (CA || CCVA) -> ADD -> Conv   (yaml lines 55-59 and repeats)
        return [[-1, 1, 'Conv', [c_mid, 1, 1]], [-1, 1, 'CA', [c_mid]], [-2, 1, 'CCVA', [c_mid]],
                [[-1, -2], 1, 'ADD', []], [-1, 1, 'Conv', [c_out, 1, 1]]]

    backbone = [[-1, 1, 'RepS_Block', [48, 3, 2, 1]]]
    for c in (48, 128, 256, 512):
        backbone += [[-1, 1, 'DER_Block', [c, 1, 2]], [-1, 1, 'MP', []]]
    head = [[-1, 1, 'SPPCSPC', [512]],                                   # 9
            [-1, 1, 'GSConv', [128, 1, 1]],                              # 10
            [-1, 1, 'nn.Upsample', [None, 2, 'nearest']],                # 11
            [6, 1, 'GSConv', [256, 1, 1]],                               # 12
            [[-1, -2], 1, 'Concat', [1]],                                # 13
            [-1, 1, 'VoVGSCSP', [256]],                                  # 14
            [-1, 1, 'GSConv', [128, 1, 1]],                              # 15
            [-1, 1, 'nn.Upsample', [None, 2, 'nearest']],                # 16
            [4, 1, 'GSConv', [128, 1, 1]],                               # 17
            [[-1, -2], 1, 'Concat', [1]]]                                # 18
    head += attn_stage(128, 256)                                         # 19-23
    head += [[-1, 1, 'VoVGSCSP', [128]]]                                 # 24
    head += attn_stage(64, 128)                                          # 25-29
    head += [[-1, 1, 'MP', []], [-1, 1, 'GSConv', [128, 1, 1]], [-3, 1, 'GSConv', [128, 1, 1]],
             [-1, 1, 'GSConv', [128, 3, 2]], [[-1, -3, 14], 1, 'Concat', [1]]]   # 30-34
    head += attn_stage(256, 512)                                         # 35-39
    head += [[-1, 1, 'VoVGSCSP', [256]]]                                 # 40
    head += attn_stage(128, 256)                                         # 41-45
    head += [[-1, 1, 'MP', []], [-1, 1, 'GSConv', [256, 1, 1]], [-3, 1, 'GSConv', [256, 1, 1]],
             [-1, 1, 'GSConv', [256, 3, 2]], [[-1, -3, 9], 1, 'Concat', [1]]]    # 46-50
    head += attn_stage(512, 1024)                                        # 51-55
    head += [[-1, 1, 'VoVGSCSP', [512]]]                                 # 56
    head += attn_stage(256, 512)                                         # 57-61
    head += [[29, 1, 'RepConv', [256, 3, 1]], [45, 1, 'RepConv', [512, 3, 1]], [61, 1, 'RepConv', [1024, 3, 1]],
             [[62, 63, 64], 1, 'IDetect', ['nc', 'anchors']]]
    return {'nc': nc, 'depth_multiple': 1.0, 'width_multiple': 1.0, 'anchors': ANCHORS,
            'backbone': backbone, 'head': head}


def build_graph(cfg: dict | None = None, ch: int = 3):
    """Channel bookkeeping of parse_model (models/yolo.py:730-836) for the module kinds Rep-YOLO uses.

    Returns (layers, save): layers[i] = dict(i, f, kind, c1, c2, args); save = sorted skip-list indices.
    """
    cfg = cfg or default_cfg()
    nc, anchors = cfg['nc'], cfg['anchors']
    na = len(anchors[0]) // 2
    no = na * (nc + 5)
    chs, layers, save = [ch], [], []
    for i, (f, n, kind, args) in enumerate(cfg['backbone'] + cfg['head']):
        kind = kind.replace('nn.', '')
        args = list(args)
        src = (lambda j: chs[j] if i > 0 else ch)
        if kind in ('Conv', 'RepConv', 'SPPCSPC', 'GSConv', 'VoVGSCSP', 'CCVA'):        # yolo.py:746-766
            c1, c2 = src(f), args[0]
            if c2 != no:
                c2 = int(math.ceil(c2 / 8) * 8)                                        # make_divisible(c2*gw, 8), gw=1
            args = [c1, c2] + args[1:]
        elif kind in ('RepS_Block', 'DER_Block'):                                       # yolo.py:788-790
            c1, c2 = src(f), args[0]
            args = [c1, c2] + args[1:]
        elif kind == 'Concat':                                                          # yolo.py:784
            c1, c2 = None, sum(chs[x] for x in f)
        elif kind == 'ADD':                                                             # yolo.py:805
            c1, c2 = None, chs[f[0]]
        elif kind == 'IDetect':                                                         # yolo.py:795-798
            c1, c2 = [chs[x] for x in f], None
            args = [nc, anchors, c1]
        else:                                                                           # MP, Upsample, CA: yolo.py:819-820
            c1 = c2 = src(f)
        layers.append(dict(i=i, f=f, kind=kind, c1=c1, c2=c2, args=args))
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)       # yolo.py:831
        if i == 0:
            chs = []
        chs.append(c2)
    return layers, sorted(set(save))


# --------------------------------------------------------------------------------------
# Parameter inventory (names follow the reference's state_dict so weights can be exchanged)
# --------------------------------------------------------------------------------------
def _conv_bn_names(out, prefix, c1, c2, k, g=1, conv='conv', bn='bn'):
    out[f'{prefix}.{conv}.weight'] = (c2, c1 // g, k, k)
    _bn_names(out, f'{prefix}.{bn}', c2)


def _bn_names(out, prefix, c):
    out[f'{prefix}.weight'] = (c,)
    out[f'{prefix}.bias'] = (c,)
    out[f'{prefix}.running_mean'] = (c,)
    out[f'{prefix}.running_var'] = (c,)
    out[f'{prefix}.num_batches_tracked'] = ()


def _reps_names(out, p, c1, c2, k, s, branches):
    """RepS_Block.__init__ (models/common.py:3376-3410): skip-BN first, then conv branches, then 1x1 scale."""
    if c1 == c2 and s == 1:
        _bn_names(out, f'{p}.rbr_skip', c1)
    for b in range(branches):
        _conv_bn_names(out, f'{p}.rbr_conv.{b}', c1, c2, k)
    if k > 1:
        _conv_bn_names(out, f'{p}.rbr_scale', c1, c2, 1)


def _gsconv_names(out, p, c1, c2, k):
    c_ = c2 // 2
    _conv_bn_names(out, f'{p}.cv1', c1, c_, k)
    _conv_bn_names(out, f'{p}.cv2', c_, c_, 5, g=c_)


def _attn_names(out, p, c):
    """CrissCrossAttention / VerticalAttention.__init__ (models/common.py:3677-3687, 3734-3745)."""
    cq = c // 8
    _conv_bn_names(out, f'{p}.query_conv', c, cq, 1, g=math.gcd(c, cq))
    _conv_bn_names(out, f'{p}.key_conv', c, cq, 1, g=math.gcd(c, cq))
    _conv_bn_names(out, f'{p}.value_conv', c, c, 1, g=c)
    out[f'{p}.gamma'] = (1,)
    _bn_names(out, f'{p}.bn', cq)
    _bn_names(out, f'{p}.bn1', c)


def param_shapes(layers) -> OrderedDict:
    """name -> shape for every entry of the unfused reference state_dict, in registration order."""
    out = OrderedDict()
    for L in layers:
        p, kind, a = f"model.{L['i']}", L['kind'], L['args']
        if kind == 'RepS_Block':                       # args = [c1, c2, k, s, pad]; num_conv_branches default 1
            _reps_names(out, p, a[0], a[1], a[2], a[3], 1)
        elif kind == 'DER_Block':                      # models/common.py:3533-3563; args=[c1,c2,num_blocks,branches]
            c1, c2, br = a[0], a[1], a[3]
            _conv_bn_names(out, f'{p}.cv1', 3 * c1, c2, 1)
            for j in range(4):                         # cv3_* are dead weights but exist in the state_dict
                _conv_bn_names(out, f'{p}.cv{j}_1', c1, c1 // 2, 1)
                _conv_bn_names(out, f'{p}.cv{j}_2', c1 // 2, c1, 1)
            for s in range(1, 7):
                c = c1 if s <= 3 else c1 // 2
                _reps_names(out, f'{p}.stage{s}.0', c, c, 3, 1, br)
        elif kind == 'SPPCSPC':                        # models/common.py:272-282 (e=0.5 -> c_ = c2)
            c1, c2 = a[0], a[1]
            c_ = int(2 * c2 * 0.5)
            for name, ci, co, k in (('cv1', c1, c_, 1), ('cv2', c1, c_, 1), ('cv3', c_, c_, 3), ('cv4', c_, c_, 1),
                                    ('cv5', 4 * c_, c_, 1), ('cv6', c_, c_, 3), ('cv7', 2 * c_, c2, 1)):
                _conv_bn_names(out, f'{p}.{name}', ci, co, k)
        elif kind == 'GSConv':                         # models/common.py:3809-3813
            _gsconv_names(out, p, a[0], a[1], a[2])
        elif kind == 'VoVGSCSP':                       # models/common.py:3848-3856
            c1, c2 = a[0], a[1]
            c_ = int(c2 * 0.5)
            _conv_bn_names(out, f'{p}.cv1', c1, c_, 1)
            _conv_bn_names(out, f'{p}.cv2', c1, c_, 1)
            _gsconv_names(out, f'{p}.gsb.0.conv_lighting.0', c_, c_, 1)       # GSBottleneck(c_, c_, e=1.0)
            _gsconv_names(out, f'{p}.gsb.0.conv_lighting.1', c_, c_, 3)
            _conv_bn_names(out, f'{p}.gsb.0.shortcut', c_, c_, 1)
            _conv_bn_names(out, f'{p}.res', c_, c_, 3)                        # dead weights
            _conv_bn_names(out, f'{p}.cv3', 2 * c_, c2, 1)
        elif kind == 'Conv':
            _conv_bn_names(out, p, a[0], a[1], a[2])
        elif kind == 'CA':                             # models/common.py:3789-3795  (ratio 16, bias=False)
            c = a[0]
            out[f'{p}.f1.weight'] = (c // 16, c, 1, 1)
            out[f'{p}.f2.weight'] = (c, c // 16, 1, 1)
        elif kind == 'CCVA':                           # C3.__init__ then CCVA.__init__ (common.py:2646-2651, 3782-3786)
            c1, c2 = a[0], a[1]
            c_ = int(c2 * 0.5)
            _conv_bn_names(out, f'{p}.cv1', c1, c_, 1)
            _conv_bn_names(out, f'{p}.cv2', c1, c_, 1)
            _conv_bn_names(out, f'{p}.cv3', 2 * c_, c2, 1)
            _attn_names(out, f'{p}.m', c_)
            _attn_names(out, f'{p}.m1', c_)
        elif kind == 'RepConv':                        # models/common.py:480-509 (identity only if c1==c2 and s==1)
            c1, c2, s = a[0], a[1], a[3]
            if c1 == c2 and s == 1:
                _bn_names(out, f'{p}.rbr_identity', c1)
            _conv_bn_names(out, f'{p}.rbr_dense', c1, c2, 3, conv='0', bn='1')
            _conv_bn_names(out, f'{p}.rbr_1x1', c1, c2, 1, conv='0', bn='1')
        elif kind == 'IDetect':                        # models/yolo.py:99-113
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


STRIDES = (8.0, 16.0, 32.0)  # models/yolo.py:530-533 (measured by the reference with a 256x256 probe)


def _gen(name: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_state_dict(layers, seed: int = 0, mode: str = 'calibrated') -> OrderedDict:
    """Deterministic, name-keyed synthetic weights with the reference's state_dict names and shapes.

    mode='calibrated' follows SURVEY.md Appendix C/D (BN gamma~U(.5,1.5), beta~N(0,.2), ImplicitM~N(1,.02),
    attention gammas 0.5 / 1e-3, head W~N(0,2/Cin), b~N(0,.5) with obj bias -2); call ``calibrate_bn_`` afterwards.
    mode='default' mimics what ``Model(cfg)`` produces statistically (identity BN stats, ImplicitM~N(0,.02),
    attention gamma 0, reference head-bias prior of models/yolo.py:621-629): the collapsed-activation regime.
    Conv weights are U(+-1/sqrt(fan_in)) in both modes (torch's default Conv2d init).
    """
    assert mode in ('calibrated', 'default')
    sd = OrderedDict()
    det = next(L for L in layers if L['kind'] == 'IDetect')
    nc, anchors, _ = det['args']
    na, no = len(anchors[0]) // 2, nc + 5
    for name, shape in param_shapes(layers).items():
        g = _gen(name, seed)
        leaf = name.rsplit('.', 1)[-1]
        if name.endswith('.anchor_grid'):
            t = torch.tensor(anchors, dtype=torch.float32).view(len(anchors), 1, na, 1, 1, 2)
        elif name.endswith('.anchors'):                   # models/yolo.py:531 anchors /= stride
            t = torch.tensor(anchors, dtype=torch.float32).view(len(anchors), na, 2) / torch.tensor(STRIDES).view(-1, 1, 1)
        elif leaf == 'num_batches_tracked':
            t = torch.zeros((), dtype=torch.long)
        elif leaf == 'running_mean':
            t = torch.zeros(shape)
        elif leaf == 'running_var':
            t = torch.ones(shape)
        elif leaf == 'gamma':
            t = torch.full(shape, 0.0 if mode == 'default' else (0.5 if name.endswith('.m.gamma') else 1e-3))
        elif leaf == 'implicit':
            is_m = '.im.' in name
            mean = 1.0 if (is_m and mode == 'calibrated') else 0.0
            t = torch.empty(shape).normal_(mean, 0.02, generator=g)
        elif f"model.{det['i']}.m." in name:              # head 1x1 convs
            lvl = int(name.split('.')[3])
            if leaf == 'weight':
                if mode == 'calibrated':
                    t = torch.empty(shape).normal_(0.0, math.sqrt(2.0 / shape[1]), generator=g)
                else:
                    b = 1.0 / math.sqrt(shape[1])
                    t = torch.empty(shape).uniform_(-b, b, generator=g)
            else:
                if mode == 'calibrated':
                    t = torch.empty(shape).normal_(0.0, 0.5, generator=g).view(na, no)
                    t[:, 4] -= 2.0
                else:                                     # models/yolo.py:621-629 _initialize_biases
                    cin = param_shapes(layers)[name.replace('bias', 'weight')][1]
                    b = 1.0 / math.sqrt(cin)
                    t = torch.empty(shape).uniform_(-b, b, generator=g).view(na, no)
                    t[:, 4] += math.log(8 / (640 / STRIDES[lvl]) ** 2)
                    t[:, 5:] += math.log(0.6 / (nc - 0.99))
                t = t.reshape(shape).contiguous()
        elif len(shape) == 4:                             # conv weights
            b = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            t = torch.empty(shape).uniform_(-b, b, generator=g)
        elif leaf == 'weight':                            # BN gamma
            t = torch.ones(shape) if mode == 'default' else torch.empty(shape).uniform_(0.5, 1.5, generator=g)
        elif leaf == 'bias':                              # BN beta
            t = torch.zeros(shape) if mode == 'default' else torch.empty(shape).normal_(0.0, 0.2, generator=g)
        else:
            raise KeyError(name)
        sd[name] = t
    return sd


# --------------------------------------------------------------------------------------
# Unfused forward (eval BN, or "calibrate": batch statistics written back as running stats)
# --------------------------------------------------------------------------------------
class _Unfused:
    def __init__(self, sd, calibrate=False):
        self.sd, self.calibrate = sd, calibrate

    def bn(self, p, x):
        sd = self.sd
        if self.calibrate:      # == train-mode BatchNorm2d with momentum=1
            dims = (0, 2, 3)
            mean = x.mean(dims)
            n = x.numel() // x.shape[1]
            var_b = x.var(dims, unbiased=False)
            sd[f'{p}.running_mean'] = mean.clone()
            sd[f'{p}.running_var'] = (var_b * (n / max(n - 1, 1))).clone()
            return F.batch_norm(x, None, None, sd[f'{p}.weight'], sd[f'{p}.bias'], True, 0.0, BN_EPS)
        return F.batch_norm(x, sd[f'{p}.running_mean'], sd[f'{p}.running_var'], sd[f'{p}.weight'],
                            sd[f'{p}.bias'], False, 0.0, BN_EPS)

    def conv_bn(self, p, x, s=1, act=True, conv='conv', bn='bn'):
        """Conv.forward (models/common.py:109-113): act(bn(conv(x))), autopad, groups inferred from the weight."""
        w = self.sd[f'{p}.{conv}.weight']
        g = x.shape[1] // w.shape[1]
        y = self.bn(f'{p}.{bn}', F.conv2d(x, w, None, s, w.shape[-1] // 2, 1, g))
        return F.silu(y) if act else y

    def reps(self, p, x, s=1):
        """RepS_Block.forward, multi-branch path (models/common.py:3418-3434)."""
        sd = self.sd
        out = self.conv_bn(f'{p}.rbr_scale', x, s, act=False) if f'{p}.rbr_scale.conv.weight' in sd else 0
        if f'{p}.rbr_skip.weight' in sd:
            out = out + self.bn(f'{p}.rbr_skip', x)
        b = 0
        while f'{p}.rbr_conv.{b}.conv.weight' in sd:
            out = out + self.conv_bn(f'{p}.rbr_conv.{b}', x, s, act=False)
            b += 1
        return F.silu(out)

    def der(self, p, x):
        """DER_Block.forward (models/common.py:3644-3654); Dropout is identity (eval / p=0 while calibrating)."""
        x1 = self.reps(f'{p}.stage1.0', x)
        x2 = self.reps(f'{p}.stage2.0', x1)
        x3 = self.reps(f'{p}.stage3.0', x2)
        x41 = self.conv_bn(f'{p}.cv0_2', self.reps(f'{p}.stage4.0', self.conv_bn(f'{p}.cv0_1', x3)))
        x42 = self.conv_bn(f'{p}.cv1_2', self.reps(f'{p}.stage5.0', self.conv_bn(f'{p}.cv1_1', x41)))
        x43 = self.conv_bn(f'{p}.cv2_2', self.reps(f'{p}.stage6.0', self.conv_bn(f'{p}.cv2_1', x42)))
        return self.conv_bn(f'{p}.cv1', torch.cat([x1, x41, x43], 1))

    def sppcspc(self, p, x):
        """SPPCSPC.forward (models/common.py:284-290)."""
        x1 = self.conv_bn(f'{p}.cv4', self.conv_bn(f'{p}.cv3', self.conv_bn(f'{p}.cv1', x)))
        pools = [F.max_pool2d(x1, k, 1, k // 2) for k in (5, 9, 13)]
        y1 = self.conv_bn(f'{p}.cv6', self.conv_bn(f'{p}.cv5', torch.cat([x1] + pools, 1)))
        y2 = self.conv_bn(f'{p}.cv2', x)
        return self.conv_bn(f'{p}.cv7', torch.cat((y1, y2), 1))

    def gsconv(self, p, x, s=1, act=True):
        """GSConv.forward (models/common.py:3815-3825) with the literal reshape/permute shuffle."""
        x1 = self.conv_bn(f'{p}.cv1', x, s, act)
        x2 = torch.cat((x1, self.conv_bn(f'{p}.cv2', x1, 1, act)), 1)
        b, n, h, w = x2.shape
        y = x2.reshape(b * n // 2, 2, h * w).permute(1, 0, 2).reshape(2, -1, n // 2, h, w)
        return torch.cat((y[0], y[1]), 1)

    def vov(self, p, x):
        """VoVGSCSP.forward / GSBottleneck.forward (models/common.py:3858-3861, 3837-3838)."""
        t = self.conv_bn(f'{p}.cv1', x)
        g = self.gsconv(f'{p}.gsb.0.conv_lighting.1', self.gsconv(f'{p}.gsb.0.conv_lighting.0', t), 1, act=False)
        x1 = g + self.conv_bn(f'{p}.gsb.0.shortcut', t, act=False)
        return self.conv_bn(f'{p}.cv3', torch.cat((self.conv_bn(f'{p}.cv2', x), x1), 1))

    def ca(self, p, x):
        """CA.forward (models/common.py:3797-3802): the input is overwritten by its global average."""
        x = x.mean((2, 3), keepdim=True)
        a = torch.sigmoid(F.conv2d(F.relu(F.conv2d(x, self.sd[f'{p}.f1.weight'])), self.sd[f'{p}.f2.weight']))
        return x * a + x

    def _qkv(self, p, x):
        q = F.relu6(self.bn(f'{p}.bn', self.conv_bn(f'{p}.query_conv', x)))
        k = F.relu6(self.bn(f'{p}.bn', self.conv_bn(f'{p}.key_conv', x)))      # q and k share one BN (3696, 3701)
        v = F.relu6(self.bn(f'{p}.bn1', self.conv_bn(f'{p}.value_conv', x)))
        return q, k, v

    def crisscross(self, p, x):
        return criss_cross(x, *self._qkv(p, x), self.sd[f'{p}.gamma'])

    def vertical(self, p, x):
        return vertical_attention(x, *self._qkv(p, x), self.sd[f'{p}.gamma'])

    def ccva(self, p, x):
        """C3.forward with m=CrissCrossAttention, m1=VerticalAttention (models/common.py:2654-2655, 3781-3786)."""
        y = self.vertical(f'{p}.m1', self.crisscross(f'{p}.m', self.conv_bn(f'{p}.cv1', x)))
        return self.conv_bn(f'{p}.cv3', torch.cat((y, self.conv_bn(f'{p}.cv2', x)), 1))

    def repconv(self, p, x, s=1):
        """RepConv.forward, train-time branches (models/common.py:511-520)."""
        out = self.conv_bn(f'{p}.rbr_dense', x, s, False, '0', '1') + self.conv_bn(f'{p}.rbr_1x1', x, s, False, '0', '1')
        if f'{p}.rbr_identity.weight' in self.sd:
            out = out + self.bn(f'{p}.rbr_identity', x)
        return F.silu(out)

    def idetect_convs(self, p, xs):
        """IDetect.forward conv part (models/yolo.py:119-121): im * conv(ia + x)."""
        sd = self.sd
        return [sd[f'{p}.im.{j}.implicit'] * F.conv2d(sd[f'{p}.ia.{j}.implicit'] + x, sd[f'{p}.m.{j}.weight'],
                                                      sd[f'{p}.m.{j}.bias']) for j, x in enumerate(xs)]


def criss_cross(x, q, k, v, gamma):
    """CrissCrossAttention.forward after q/k/v (models/common.py:3704-3726), written as einsums.

    eH[b,h,w,g] = sum_d q[b,d,h,w] k[b,d,g,w];  eW[b,h,w,g] = sum_d q[b,d,h,w] k[b,d,h,g];  softmax over (H+W);
    out = sum_g v[b,c,g,w] aH[b,h,w,g] + sum_g v[b,c,h,g] aW[b,h,w,g];  no -inf diagonal;  gamma*out + x.
    """
    H = x.shape[2]
    eH = torch.einsum('bdhw,bdgw->bhwg', q, k)
    eW = torch.einsum('bdhw,bdhg->bhwg', q, k)
    att = torch.softmax(torch.cat([eH, eW], 3), 3)
    out = torch.einsum('bcgw,bhwg->bchw', v, att[..., :H]) + torch.einsum('bchg,bhwg->bchw', v, att[..., H:])
    return gamma * out + x


def vertical_attention(x, q, k, v, gamma):
    """VerticalAttention.forward after q/k/v (models/common.py:3763-3778), following the literal view chain.

    The softmax at :3770 is computed and discarded by the reference; the raw column energies are re-viewed
    after a permute (:3767, :3772), which for H != W scrambles indices -- so the literal chain is kept here.
    """
    B, C, H, W = x.shape
    qH = q.permute(0, 3, 1, 2).contiguous().view(B * W, -1, H).permute(0, 2, 1)
    kH = k.permute(0, 3, 1, 2).contiguous().view(B * W, -1, H)
    vH = v.permute(0, 3, 1, 2).contiguous().view(B * W, -1, H)
    eH = torch.bmm(qH, kH).view(B, W, H, H).permute(0, 2, 1, 3)
    attH = eH.contiguous().view(B * W, H, H)
    out = torch.bmm(vH, attH.permute(0, 2, 1)).view(B, W, -1, H).permute(0, 2, 3, 1)
    return gamma * out + x


def _run_graph(layers, save, x, run_layer):
    """Model.forward_once (models/yolo.py:587-619): sequential walk with the skip list."""
    ys, outs = [], []
    for L in layers:
        f = L['f']
        if f != -1:
            x = ys[f] if isinstance(f, int) else [x if j == -1 else ys[j] for j in f]
        x = run_layer(L, x)
        ys.append(x if L['i'] in save else None)
        outs.append(x)
    return outs


def forward_unfused(sd, layers, save, x, calibrate=False):
    """Per-layer outputs of the unfused model; the last entry is the list of 3 raw head conv maps [B,18,ny,nx]."""
    U = _Unfused(sd, calibrate)

    def run(L, x):
        p, kind, a = f"model.{L['i']}", L['kind'], L['args']
        if kind == 'RepS_Block':
            return U.reps(p, x, a[3])
        if kind == 'DER_Block':
            return U.der(p, x)
        if kind == 'MP':
            return F.max_pool2d(x, 2, 2)
        if kind == 'SPPCSPC':
            return U.sppcspc(p, x)
        if kind == 'GSConv':
            return U.gsconv(p, x, a[3])
        if kind == 'Upsample':
            return F.interpolate(x, scale_factor=2, mode='nearest')
        if kind == 'Concat':
            return torch.cat(x, 1)
        if kind == 'VoVGSCSP':
            return U.vov(p, x)
        if kind == 'Conv':
            return U.conv_bn(p, x, a[3])
        if kind == 'CA':
            return U.ca(p, x)
        if kind == 'CCVA':
            return U.ccva(p, x)
        if kind == 'ADD':
            return x[0] + x[1]
        if kind == 'RepConv':
            return U.repconv(p, x, a[3])
        if kind == 'IDetect':
            return U.idetect_convs(p, x)
        raise KeyError(kind)

    with torch.no_grad():
        return _run_graph(layers, save, x, run)


def calibrate_bn_(sd, layers, save, size=640, batch=1, seed=1234):
    """One batch-statistics pass: every BN's running stats := stats of a seeded batch."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, size, size, generator=g)
    forward_unfused(sd, layers, save, x, calibrate=True)
    return sd


# --------------------------------------------------------------------------------------
# Folding  (Model.fuse, models/yolo.py:681-704)
# --------------------------------------------------------------------------------------
def _bn_scale_shift(sd, p):
    std = (sd[f'{p}.running_var'] + BN_EPS).sqrt()
    t = sd[f'{p}.weight'] / std
    return t, sd[f'{p}.bias'] - sd[f'{p}.running_mean'] * sd[f'{p}.weight'] / std


def fold_conv_bn(sd, p, conv='conv', bn='bn'):
    """fuse_conv_and_bn (utils/torch_utils.py:181-201): W' = diag(g/sqrt(eps+var)) W ; b' = beta - g*mu/sqrt(var+eps)."""
    w = sd[f'{p}.{conv}.weight']
    t = sd[f'{p}.{bn}.weight'] / torch.sqrt(BN_EPS + sd[f'{p}.{bn}.running_var'])
    b = sd[f'{p}.{bn}.bias'] - sd[f'{p}.{bn}.weight'] * sd[f'{p}.{bn}.running_mean'] / torch.sqrt(
        sd[f'{p}.{bn}.running_var'] + BN_EPS)
    return w * t.view(-1, 1, 1, 1), b


def _identity_kernel(c, k):
    w = torch.zeros(c, c, k, k)
    w[torch.arange(c), torch.arange(c), k // 2, k // 2] = 1.0
    return w


def fold_reps(sd, p):
    """RepS_Block._get_kernel_bias (models/common.py:3462-3517): sum of conv branches + padded 1x1 scale + BN skip."""
    W, b, j = 0, 0, 0
    while f'{p}.rbr_conv.{j}.conv.weight' in sd:
        t, sh = _bn_scale_shift(sd, f'{p}.rbr_conv.{j}.bn')
        W = W + sd[f'{p}.rbr_conv.{j}.conv.weight'] * t.view(-1, 1, 1, 1)
        b = b + sh
        j += 1
    k = W.shape[-1]
    if f'{p}.rbr_scale.conv.weight' in sd:
        t, sh = _bn_scale_shift(sd, f'{p}.rbr_scale.bn')
        W = W + F.pad(sd[f'{p}.rbr_scale.conv.weight'] * t.view(-1, 1, 1, 1), [k // 2] * 4)
        b = b + sh
    if f'{p}.rbr_skip.weight' in sd:
        t, sh = _bn_scale_shift(sd, f'{p}.rbr_skip')
        W = W + _identity_kernel(W.shape[0], k) * t.view(-1, 1, 1, 1)
        b = b + sh
    return W, b


def fold_repconv(sd, p):
    """RepConv.fuse_repvgg_block (models/common.py:597-657): 3x3 + pad(1x1) (+ identity BN when c1==c2, s==1)."""
    t3, b3 = _bn_scale_shift(sd, f'{p}.rbr_dense.1')
    t1, b1 = _bn_scale_shift(sd, f'{p}.rbr_1x1.1')
    W = sd[f'{p}.rbr_dense.0.weight'] * t3.view(-1, 1, 1, 1) + F.pad(sd[f'{p}.rbr_1x1.0.weight'] * t1.view(-1, 1, 1, 1), [1] * 4)
    b = b3 + b1
    if f'{p}.rbr_identity.weight' in sd:
        ti, bi = _bn_scale_shift(sd, f'{p}.rbr_identity')
        W = W + _identity_kernel(W.shape[0], 3) * ti.view(-1, 1, 1, 1)
        b = b + bi
    return W, b


def fold_idetect(sd, p, j):
    """IDetect.fuse (models/yolo.py:170-182): b <- (b + W.ia) * im ; W <- W * im."""
    W, b = sd[f'{p}.m.{j}.weight'], sd[f'{p}.m.{j}.bias']
    ia, im = sd[f'{p}.ia.{j}.implicit'].view(-1), sd[f'{p}.im.{j}.implicit'].view(-1)
    b = (b + W.view(W.shape[0], -1) @ ia) * im
    return W * im.view(-1, 1, 1, 1), b


def fold(sd, layers) -> OrderedDict:
    """Restates Model.fuse(): returns {fused_name: tensor} using the reference's *fused* state_dict names
    (``…reparam_conv``, ``…rbr_reparam``, ``….conv.{weight,bias}``); stand-alone attention BNs, CA convs,
    gamma and anchors are carried over unchanged (they are not folded by the reference)."""
    fz = OrderedDict()
    for name in sd:
        if name.endswith('.conv.weight') and f"{name[:-12]}.bn.weight" in sd and '.rbr_' not in name:
            p = name[:-12]
            fz[f'{p}.conv.weight'], fz[f'{p}.conv.bias'] = fold_conv_bn(sd, p)
    for L in layers:
        p, kind = f"model.{L['i']}", L['kind']
        if kind == 'RepS_Block':
            fz[f'{p}.reparam_conv.weight'], fz[f'{p}.reparam_conv.bias'] = fold_reps(sd, p)
        elif kind == 'DER_Block':
            for s in range(1, 7):
                q = f'{p}.stage{s}.0'
                fz[f'{q}.reparam_conv.weight'], fz[f'{q}.reparam_conv.bias'] = fold_reps(sd, q)
        elif kind == 'RepConv':
            fz[f'{p}.rbr_reparam.weight'], fz[f'{p}.rbr_reparam.bias'] = fold_repconv(sd, p)
        elif kind == 'IDetect':
            for j in range(len(L['args'][2])):
                fz[f'{p}.m.{j}.weight'], fz[f'{p}.m.{j}.bias'] = fold_idetect(sd, p, j)
            fz[f'{p}.anchors'], fz[f'{p}.anchor_grid'] = sd[f'{p}.anchors'], sd[f'{p}.anchor_grid']
        elif kind == 'CA':
            fz[f'{p}.f1.weight'], fz[f'{p}.f2.weight'] = sd[f'{p}.f1.weight'], sd[f'{p}.f2.weight']
        elif kind == 'CCVA':
            for m in ('m', 'm1'):
                fz[f'{p}.{m}.gamma'] = sd[f'{p}.{m}.gamma']
                for bn in ('bn', 'bn1'):
                    for leaf in ('weight', 'bias', 'running_mean', 'running_var'):
                        fz[f'{p}.{m}.{bn}.{leaf}'] = sd[f'{p}.{m}.{bn}.{leaf}']
    return fz


# --------------------------------------------------------------------------------------
# Fused (deploy) forward + decode
# --------------------------------------------------------------------------------------
class _Fused:
    def __init__(self, fz):
        self.fz = fz

    def conv(self, p, x, s=1, act=True, name='conv'):
        """Conv.fuseforward (models/common.py:115-116) / RepS_Block deploy branch (3412-3416) / RepConv deploy (511-513)."""
        w, b = self.fz[f'{p}.{name}.weight'], self.fz[f'{p}.{name}.bias']
        y = F.conv2d(x, w, b, s, w.shape[-1] // 2, 1, x.shape[1] // w.shape[1])
        return F.silu(y) if act else y

    def bn(self, p, x):
        fz = self.fz
        return F.batch_norm(x, fz[f'{p}.running_mean'], fz[f'{p}.running_var'], fz[f'{p}.weight'], fz[f'{p}.bias'],
                            False, 0.0, BN_EPS)

    def reps(self, p, x, s=1):
        return self.conv(p, x, s, True, 'reparam_conv')

    def der(self, p, x, taps=None):
        x1 = self.reps(f'{p}.stage1.0', x)
        x2 = self.reps(f'{p}.stage2.0', x1)
        x3 = self.reps(f'{p}.stage3.0', x2)
        x41 = self.conv(f'{p}.cv0_2', self.reps(f'{p}.stage4.0', self.conv(f'{p}.cv0_1', x3)))
        x42 = self.conv(f'{p}.cv1_2', self.reps(f'{p}.stage5.0', self.conv(f'{p}.cv1_1', x41)))
        x43 = self.conv(f'{p}.cv2_2', self.reps(f'{p}.stage6.0', self.conv(f'{p}.cv2_1', x42)))
        return self.conv(f'{p}.cv1', torch.cat([x1, x41, x43], 1))

    def sppcspc(self, p, x):
        x1 = self.conv(f'{p}.cv4', self.conv(f'{p}.cv3', self.conv(f'{p}.cv1', x)))
        pools = [F.max_pool2d(x1, k, 1, k // 2) for k in (5, 9, 13)]
        y1 = self.conv(f'{p}.cv6', self.conv(f'{p}.cv5', torch.cat([x1] + pools, 1)))
        return self.conv(f'{p}.cv7', torch.cat((y1, self.conv(f'{p}.cv2', x)), 1))

    def gsconv(self, p, x, s=1, act=True):
        x1 = self.conv(f'{p}.cv1', x, s, act)
        x2 = torch.cat((x1, self.conv(f'{p}.cv2', x1, 1, act)), 1)
        return torch.cat((x2[:, 0::2], x2[:, 1::2]), 1)          # == the reshape/permute of common.py:3819-3825

    def vov(self, p, x):
        t = self.conv(f'{p}.cv1', x)
        g = self.gsconv(f'{p}.gsb.0.conv_lighting.1', self.gsconv(f'{p}.gsb.0.conv_lighting.0', t), 1, act=False)
        x1 = g + self.conv(f'{p}.gsb.0.shortcut', t, act=False)
        return self.conv(f'{p}.cv3', torch.cat((self.conv(f'{p}.cv2', x), x1), 1))

    def ca(self, p, x):
        x = x.mean((2, 3), keepdim=True)
        a = torch.sigmoid(F.conv2d(F.relu(F.conv2d(x, self.fz[f'{p}.f1.weight'])), self.fz[f'{p}.f2.weight']))
        return x * a + x

    def qkv(self, p, x):
        q = F.relu6(self.bn(f'{p}.bn', self.conv(f'{p}.query_conv', x)))
        k = F.relu6(self.bn(f'{p}.bn', self.conv(f'{p}.key_conv', x)))
        v = F.relu6(self.bn(f'{p}.bn1', self.conv(f'{p}.value_conv', x)))
        return q, k, v

    def ccva(self, p, x):
        t = self.conv(f'{p}.cv1', x)
        t = criss_cross(t, *self.qkv(f'{p}.m', t), self.fz[f'{p}.m.gamma'])
        t = vertical_attention(t, *self.qkv(f'{p}.m1', t), self.fz[f'{p}.m1.gamma'])
        return self.conv(f'{p}.cv3', torch.cat((t, self.conv(f'{p}.cv2', x)), 1))

    def head_convs(self, p, xs):
        return [F.conv2d(x, self.fz[f'{p}.m.{j}.weight'], self.fz[f'{p}.m.{j}.bias']) for j, x in enumerate(xs)]


def run_fused_layer(fz, L, x):
    """One top-level layer of the deploy graph on fp32 input(s) (used for teacher-forced module parity)."""
    M = _Fused(fz)
    p, kind, a = f"model.{L['i']}", L['kind'], L['args']
    if kind == 'RepS_Block':
        return M.reps(p, x, a[3])
    if kind == 'DER_Block':
        return M.der(p, x)
    if kind == 'MP':
        return F.max_pool2d(x, 2, 2)
    if kind == 'SPPCSPC':
        return M.sppcspc(p, x)
    if kind == 'GSConv':
        return M.gsconv(p, x, a[3])
    if kind == 'Upsample':
        return F.interpolate(x, scale_factor=2, mode='nearest')
    if kind == 'Concat':
        return torch.cat(x, 1)
    if kind == 'VoVGSCSP':
        return M.vov(p, x)
    if kind == 'Conv':
        return M.conv(p, x, a[3])
    if kind == 'CA':
        return M.ca(p, x)
    if kind == 'CCVA':
        return M.ccva(p, x)
    if kind == 'ADD':
        return x[0] + x[1]
    if kind == 'RepConv':
        return M.conv(p, x, a[3], True, 'rbr_reparam')
    if kind == 'IDetect':
        return M.head_convs(p, x)
    raise KeyError(kind)


def decode_heads(head_maps, anchor_grid, strides=STRIDES, na=3):
    """IDetect.fuseforward decode (models/yolo.py:139-168).

    head_maps: list of [B, na*no, ny, nx] fp32.  Returns (pred [B, sum(na*ny*nx), no], raw list [B,na,ny,nx,no]).
    Row order: level -> anchor -> y -> x; columns [cx, cy, w, h, obj, cls...]; same op order as the reference.
    """
    z, raws = [], []
    for i, h in enumerate(head_maps):
        bs, _, ny, nx = h.shape
        no = h.shape[1] // na
        r = h.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing='ij')
        grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
        y = r.sigmoid()
        y[..., 0:2] = (y[..., 0:2] * 2. - 0.5 + grid) * strides[i]
        y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * anchor_grid[i].view(1, na, 1, 1, 2)
        z.append(y.view(bs, -1, no))
        raws.append(r)
    return torch.cat(z, 1), raws


def convert(pred):
    """IDetect.convert (models/yolo.py:189-199): the ``include_nms`` output contract -- boxes xywh -> xyxy through the
    4x4 matrix, score = cls * obj.  pred: [B, N, 5+nc] (the concatenated decode).  Returns (box [B,N,4], score [B,N,nc])."""
    box, conf, score = pred[:, :, :4], pred[:, :, 4:5], pred[:, :, 5:]
    score = score * conf
    m = torch.tensor([[1, 0, 1, 0], [0, 1, 0, 1], [-0.5, 0, 0.5, 0], [0, -0.5, 0, 0.5]], dtype=torch.float32)
    return box @ m, score


def forward_fused(fz, layers, save, x):
    """Deploy forward.  Returns (outs, pred, raws): outs[i] = layer i output (outs[-1] = 3 raw head conv maps)."""
    with torch.no_grad():
        outs = _run_graph(layers, save, x, lambda L, t: run_fused_layer(fz, L, t))
        det = layers[-1]
        pred, raws = decode_heads(outs[-1], fz[f"model.{det['i']}.anchor_grid"], na=len(det['args'][1][0]) // 2)
    return outs, pred, raws


def layer_inputs(layers, outs, x0, i):
    """The fp32 input(s) the reference would feed to layer i, given all layer outputs (teacher forcing)."""
    f = layers[i]['f']
    prev = x0 if i == 0 else outs[i - 1]
    at = lambda j: outs[j if j >= 0 else i + j]          # negative indices are relative to layer i (yolo.py:590)
    if f == -1:
        return prev
    if isinstance(f, int):
        return at(f)
    return [prev if j == -1 else at(j) for j in f]


_CACHE = {}


def make_model(seed: int = 0, mode: str = 'calibrated', nc: int = 1):
    """(layers, save, unfused state_dict, fused dict) for the synthetic Rep-YOLO used by tests and the benchmark."""
    key = (seed, mode, nc)
    if key not in _CACHE:
        layers, save = build_graph(default_cfg(nc))
        sd = synth_state_dict(layers, seed, mode)
        if mode == 'calibrated':
            calibrate_bn_(sd, layers, save)
        _CACHE[key] = (layers, save, sd, fold(sd, layers))
    return _CACHE[key]
