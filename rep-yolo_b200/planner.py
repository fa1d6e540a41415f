"""Lowers the fused Rep-YOLO graph to the flat op list of include/repyolo_b200.h.

The walk mirrors ``Model.forward_once`` (reference models/yolo.py:587-619) and the module forwards it calls, with every
``torch.cat`` removed: producers write straight into a channel range of the consumer's input tensor.  GSConv's channel
shuffle (models/common.py:3819-3825, == cat[x2[:, 0::2], x2[:, 1::2]]) is folded into the output-channel order of its
two convolutions, and ``ADD`` of the CA vector (common.py:3345-3349) into the epilogue of CCVA.cv3.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as N
from .fold import bn_affine


class _Blob:
    """fp32 host weight blob; returns byte offsets."""

    def __init__(self):
        self.parts, self.size = [], 0

    def add(self, t) -> int:
        a = np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32, copy=False)).reshape(-1)
        off = self.size
        self.parts.append(a)
        self.size += a.nbytes
        pad = (-self.size) % 16
        if pad:
            self.parts.append(np.zeros(pad // 4, dtype=np.float32))
            self.size += pad
        return off

    def bytes(self) -> np.ndarray:
        return np.concatenate(self.parts) if self.parts else np.zeros(4, dtype=np.float32)


class Group:
    """A set of consecutive reference layers lowered together (what a teacher-forced parity test can isolate)."""

    def __init__(self, layers, first_op):
        self.layers, self.first_op, self.last_op = layers, first_op, first_op
        self.inputs = []        # [(reference layer index or -1 for the image, view)]
        self.output = None      # view of the group's result (None for IDetect)
        self.out_layer = layers[-1]


class Plan:
    def __init__(self):
        self.tensors, self.ops, self.blob, self.groups = [], [], _Blob(), []
        self.layer_out = {}     # reference layer index -> view (tensor, c_off, c_len)

    # ---- tensors / views ----
    def tensor(self, channels, level, dtype=N.RY_BF16, kind=N.T_MAP, slot=0):
        self.tensors.append(N.TensorDesc(kind, dtype, channels, level, slot, 0))
        return len(self.tensors) - 1

    @staticmethod
    def view(t, off, n):
        return (t, off, n)

    def full(self, t):
        return (t, 0, self.tensors[t].channels)

    def level(self, v):
        return self.tensors[v[0]].level

    # ---- ops ----
    def op(self, kind, layer, in0=None, in1=None, in2=None, out0=None, out1=None, out2=None, **kw):
        d = N.OpDesc()
        d.kind, d.layer = kind, layer
        none = (-1, 0, 0)
        for name, v in (('in0', in0), ('in1', in1), ('in2', in2), ('out0', out0), ('out1', out1), ('out2', out2)):
            t, o, n = v if v is not None else none
            setattr(d, name, N.View(t, o, n))
        d.ksize, d.stride, d.act = kw.get('ksize', 1), kw.get('stride', 1), kw.get('act', N.ACT_NONE)
        d.cin, d.cout, d.level_idx = kw.get('cin', 0), kw.get('cout', 0), kw.get('level_idx', 0)
        d.n_src = kw.get('n_src', 1)
        d.w_off, d.b_off = kw.get('w_off', -1), kw.get('b_off', -1)
        aux = kw.get('aux', [])
        for i in range(6):
            d.aux_off[i] = aux[i] if i < len(aux) else -1
        fp = kw.get('fparam', [])
        for i in range(8):
            d.fparam[i] = fp[i] if i < len(fp) else 0.0
        self.ops.append(d)
        return len(self.ops) - 1

    def conv(self, layer, w, b, src, dst, stride=1, act=True, dst2=None, res=None, bvec=None, pool=False):
        cout, cin, k, _ = w.shape
        if isinstance(src, list):            # 1x1 conv over the channel concatenation of several views (no copy)
            assert k == 1 and res is None and bvec is None and 1 < len(src) <= 3 and sum(v[2] for v in src) == cin
            views = src + [None] * (3 - len(src))
            return self.op(N.OP_CONV, layer, in0=views[0], in1=views[1], in2=views[2], out0=dst, out1=dst2, ksize=1, stride=1,
                           act=N.ACT_SILU if act else N.ACT_NONE, cin=cin, cout=cout, n_src=len(src), w_off=self.blob.add(w),
                           b_off=self.blob.add(b), level_idx=1 if pool else 0)
        assert src[2] == cin, (layer, src, w.shape)
        assert (dst[2] + (dst2[2] if dst2 else 0)) == cout, (layer, dst, dst2, w.shape)
        return self.op(N.OP_CONV, layer, in0=src, in1=res, in2=bvec, out0=dst, out1=dst2, ksize=k, stride=stride,
                       act=N.ACT_SILU if act else N.ACT_NONE, cin=cin, cout=cout, w_off=self.blob.add(w),
                       b_off=self.blob.add(b), level_idx=1 if pool else 0)


    def chain(self, layer, src, main, posts):
        """3x3 conv -> 1x1 [-> 1x1] fused (OP_CONV_CHAIN).  main = (w, b, dst|None); posts = [(w, b, act, dst|None), ...]"""
        w0, b0, dst0 = main
        cout, cin, k, _ = w0.shape
        assert k == 3 and src[2] == cin and 1 <= len(posts) <= 2
        d = self.ops[self.op(N.OP_CONV_CHAIN, layer, in0=src, out0=dst0, out1=posts[0][3], out2=posts[1][3] if len(posts) > 1 else None,
                             ksize=3, stride=1, act=N.ACT_SILU, cin=cin, cout=cout, w_off=self.blob.add(w0), b_off=self.blob.add(b0),
                             aux=[o for (w, b, _, _) in posts for o in (self.blob.add(w), self.blob.add(b))])]
        d.n_post = len(posts)
        prev = cout
        for i, (w, b, act, dst) in enumerate(posts):
            assert w.shape[1] == prev and w.shape[2] == 1 and (dst is None or dst[2] == w.shape[0])
            d.post_cout[i], d.post_act[i] = w.shape[0], (N.ACT_SILU if act else N.ACT_NONE)
            prev = w.shape[0]


    def conv_pair(self, layer, wb1, wb2, src, dst1, dst2, act=True):
        """Two 1x1 convs of the same input (CCVA / VoVGSCSP / SPPCSPC cv1 + cv2) as ONE conv with a split store (the two halves may
        go to different tensors): the input is read once, one launch instead of two."""
        (w1, b1), (w2, b2) = wb1, wb2
        if not merge_pairs or w1.shape[0] + w2.shape[0] > 256 or w1.shape[0] % 16:
            self.conv(layer, w1, b1, src, dst1, act=act)
            return self.conv(layer, w2, b2, src, dst2, act=act)
        return self.conv(layer, torch.cat([w1, w2], 0), torch.cat([b1, b2], 0), src, dst1, act=act, dst2=dst2)


merge_pairs = True


def _gs_perm(c):
    return list(range(0, c, 2)) + list(range(1, c, 2))


def lower(layers, fz, nc=1, fuse_chains=None):
    """layers: arch.parse() output; fz: fold.fold_state_dict() output.  Returns a Plan."""
    import os
    if fuse_chains is None:
        fuse_chains = os.environ.get('RY_FUSE_CHAINS', '1') != '0'
    fuse_pool = os.environ.get('RY_FUSE_POOL', '1') != '0'
    P = Plan()
    img = P.tensor(3, 0, N.RY_F32, N.T_EXTERNAL, N.X_IMAGE)
    pred = P.tensor(5 + nc, 0, N.RY_F32, N.T_EXTERNAL, N.X_PRED)
    raws = [P.tensor(5 + nc, 0, N.RY_F32, N.T_EXTERNAL, N.X_RAW0 + j) for j in range(3)]
    levels = {}                                 # reference layer -> pyramid level of its output

    # concat elimination: a layer consumed by a Concat writes into the Concat's tensor
    home = {}
    for L in layers:
        if L.kind == 'Concat':
            srcs = L.sources()
            lvl = None
            total = sum(layers[s].c2 for s in srcs)
            t = P.tensor(total, 0)              # level patched when the first source is lowered
            off = 0
            for s in srcs:
                assert s not in home, 'a layer feeding two Concats would need a copy op'
                home[s] = (t, off, layers[s].c2)
                off += layers[s].c2
            P.layer_out[L.i] = (t, 0, total)

    def out_view(L, lvl):
        if L.i in home:
            v = home[L.i]
            P.tensors[v[0]].level = lvl
            return v
        return P.full(P.tensor(L.c2, lvl))

    def W(key):
        return fz[key + '.weight'], fz[key + '.bias']

    def gsconv(layer, p, src, dst, k, s, act, lvl_out):
        """GSConv (common.py:3807-3825): dst quarter layout [x1 even | dw even | x1 odd | dw odd]."""
        w1, b1 = W(f'{p}.cv1.conv')
        w2, b2 = W(f'{p}.cv2.conv')
        c_ = w1.shape[0]
        h = c_ // 2
        perm = _gs_perm(c_)
        t, o, n = dst
        assert n == 2 * c_ and h % 16 == 0
        P.conv(layer, w1[perm], b1[perm], src, (t, o, h), s, act, dst2=(t, o + c_, h))
        P.op(N.OP_DW5, layer, in0=(t, o, h), in1=(t, o + c_, h), out0=(t, o + h, h), out1=(t, o + c_ + h, h), ksize=5,
             act=N.ACT_SILU if act else N.ACT_NONE, cin=c_, cout=c_, w_off=P.blob.add(w2[perm].reshape(c_, 25)),
             b_off=P.blob.add(b2[perm]))

    def attention(layer, p, kind, src, dst, q, k):
        wq, bq = W(f'{p}.query_conv.conv')
        wk, bk = W(f'{p}.key_conv.conv')
        wv, bv = W(f'{p}.value_conv.conv')
        s, t = bn_affine(fz, f'{p}.bn')
        s1, t1 = bn_affine(fz, f'{p}.bn1')
        C, Cq = wv.shape[0], wq.shape[0]
        assert wq.shape[1] == 8 and wv.shape[1] == 1
        qk_pack = torch.cat([wq.reshape(-1), bq.reshape(-1), wk.reshape(-1), bk.reshape(-1), s.reshape(-1), t.reshape(-1)])
        P.op(kind, layer, in0=src, out0=dst, cin=C, cout=C, w_off=P.blob.add(wv.reshape(C)),
             b_off=P.blob.add(bv), aux=[P.blob.add(s1), P.blob.add(t1), P.blob.add(qk_pack)], fparam=[float(fz[f'{p}.gamma'].item())])

    skip = set()
    for L in layers:
        if L.i in skip or L.kind == 'Concat':
            if L.kind == 'Concat':
                levels[L.i] = P.tensors[P.layer_out[L.i][0]].level
            continue
        p, a = f'model.{L.i}', L.args
        srcs = L.sources()
        g = Group([L.i], len(P.ops))
        if L.i > 0 and L.kind != 'IDetect':
            g.inputs = [(s, P.layer_out[s]) for s in srcs]
        x = P.layer_out[srcs[0]] if L.i > 0 else None
        lvl_in = levels[srcs[0]] if L.i > 0 else 0

        if L.kind == 'RepS_Block':
            assert L.i == 0 and a[0] == 3 and a[2] == 3 and a[3] == 2, 'only the stem RepS_Block(3->c, k3, s2) is lowered'
            w, b = W(f'{p}.reparam_conv')
            dst = out_view(L, 1)
            g.inputs = [(-1, (img, 0, 3))]
            P.op(N.OP_STEM, L.i, in0=(img, 0, 3), out0=dst, ksize=3, stride=2, act=N.ACT_SILU, cin=3, cout=w.shape[0],
                 w_off=P.blob.add(w), b_off=P.blob.add(b))
            lvl = 1
        elif L.kind == 'DER_Block':                                  # common.py:3644-3654
            c1, lvl = a[0], lvl_in
            # x1 / x4_1 / x4_3 stay separate dense tensors: cv1 reads its three inputs through three TMA maps, so neither the
            # concat nor channel-slice (96-byte runs at a 288-byte pitch) traffic exists
            x1, x41, x43 = (P.full(P.tensor(c1, lvl)) for _ in range(3))
            x2, x3, x42 = (P.full(P.tensor(c1, lvl)) for _ in range(3))
            h1, h2 = P.full(P.tensor(c1 // 2, lvl)), P.full(P.tensor(c1 // 2, lvl))
            P.conv(L.i, *W(f'{p}.stage1.0.reparam_conv'), x, x1)
            P.conv(L.i, *W(f'{p}.stage2.0.reparam_conv'), x1, x2)
            if c1 <= 64 and fuse_chains:
                # small-channel stages: 3x3 -> 1x1 [-> 1x1] chains in one kernel; x3, x4_2 and the 3x3 outputs stay on chip
                R, C = (lambda q: W(f'{p}.{q}.0.reparam_conv')), (lambda q: W(f'{p}.{q}.conv'))
                P.chain(L.i, x2, (*R('stage3'), None), [(*C('cv0_1'), True, h1)])
                P.chain(L.i, h1, (*R('stage4'), None), [(*C('cv0_2'), True, x41), (*C('cv1_1'), True, h2)])
                P.chain(L.i, h2, (*R('stage5'), None), [(*C('cv1_2'), True, None), (*C('cv2_1'), True, h1)])
                P.chain(L.i, h1, (*R('stage6'), None), [(*C('cv2_2'), True, x43)])
            else:
                P.conv(L.i, *W(f'{p}.stage3.0.reparam_conv'), x2, x3)
                for j, (src, stage, dstv) in enumerate(((x3, 4, x41), (x41, 5, x42), (x42, 6, x43))):
                    P.conv(L.i, *W(f'{p}.cv{j}_1.conv'), src, h1)
                    P.conv(L.i, *W(f'{p}.stage{stage}.0.reparam_conv'), h1, h2)
                    P.conv(L.i, *W(f'{p}.cv{j}_2.conv'), h2, dstv)
            nxt = layers[L.i + 1] if L.i + 1 < len(layers) else None
            users = [M.i for M in layers if M.i > L.i and L.i in M.sources()]
            if fuse_pool and nxt is not None and nxt.kind == 'MP' and users == [nxt.i]:
                # MP (common.py:32-38) is the only consumer: its 2x2 max is fused into cv1's epilogue, the full-size map is never stored
                skip.add(nxt.i)
                g.layers.append(nxt.i)
                g.out_layer = nxt.i
                lvl += 1
                dst = out_view(nxt, lvl)
                P.conv(L.i, *W(f'{p}.cv1.conv'), [x1, x41, x43], dst, pool=True)
                P.layer_out[nxt.i], levels[nxt.i] = dst, lvl
            else:
                dst = out_view(L, lvl)
                P.conv(L.i, *W(f'{p}.cv1.conv'), [x1, x41, x43], dst)
        elif L.kind == 'MP':
            lvl = lvl_in + 1
            dst = out_view(L, lvl)
            P.op(N.OP_MAXPOOL2, L.i, in0=x, out0=dst, cin=x[2], cout=x[2])
        elif L.kind == 'SPPCSPC':                                    # common.py:284-290
            c_, lvl = a[1], lvl_in
            ta, tb, tc = (P.full(P.tensor(c_, lvl)) for _ in range(3))
            cat4, cat2 = P.tensor(4 * c_, lvl), P.tensor(2 * c_, lvl)
            P.conv_pair(L.i, W(f'{p}.cv1.conv'), W(f'{p}.cv2.conv'), x, ta, (cat2, c_, c_))
            P.conv(L.i, *W(f'{p}.cv3.conv'), ta, tb)
            P.conv(L.i, *W(f'{p}.cv4.conv'), tb, (cat4, 0, c_))
            P.op(N.OP_SPP, L.i, in0=(cat4, 0, c_), out0=(cat4, c_, c_), out1=(cat4, 2 * c_, c_), out2=(cat4, 3 * c_, c_),
                 cin=c_, cout=c_)
            P.conv(L.i, *W(f'{p}.cv5.conv'), P.full(cat4), tc)
            P.conv(L.i, *W(f'{p}.cv6.conv'), tc, (cat2, 0, c_))
            dst = out_view(L, lvl)
            P.conv(L.i, *W(f'{p}.cv7.conv'), P.full(cat2), dst)
        elif L.kind == 'GSConv':
            k, s = a[2], a[3]
            lvl = lvl_in + (1 if s == 2 else 0)
            dst = out_view(L, lvl)
            gsconv(L.i, p, x, dst, k, s, True, lvl)
        elif L.kind == 'Upsample':
            lvl = lvl_in - 1
            dst = out_view(L, lvl)
            P.op(N.OP_UPSAMPLE2, L.i, in0=x, out0=dst, cin=x[2], cout=x[2])
        elif L.kind == 'VoVGSCSP':                                   # common.py:3858-3861, 3837-3838
            c_, lvl = a[1] // 2, lvl_in
            cat = P.tensor(2 * c_, lvl)
            t, g0, g1 = (P.full(P.tensor(c_, lvl)) for _ in range(3))
            P.conv_pair(L.i, W(f'{p}.cv1.conv'), W(f'{p}.cv2.conv'), x, t, (cat, 0, c_))
            gsconv(L.i, f'{p}.gsb.0.conv_lighting.0', t, g0, 1, 1, True, lvl)
            gsconv(L.i, f'{p}.gsb.0.conv_lighting.1', g0, g1, 3, 1, False, lvl)
            P.conv(L.i, *W(f'{p}.gsb.0.shortcut.conv'), t, (cat, c_, c_), act=False, res=g1)
            dst = out_view(L, lvl)
            P.conv(L.i, *W(f'{p}.cv3.conv'), P.full(cat), dst)
        elif L.kind == 'Conv':
            lvl = lvl_in + (1 if a[3] == 2 else 0)
            dst = out_view(L, lvl)
            P.conv(L.i, *W(f'{p}.conv'), x, dst, stride=a[3])
        elif L.kind == 'CA':                                         # common.py:3797-3802 -> [B, C] vector
            lvl = lvl_in
            C = a[0]
            dst = P.full(P.tensor(C, 0, N.RY_F32, N.T_VEC))
            P.op(N.OP_CA, L.i, in0=x, out0=dst, cin=C, cout=C, w_off=P.blob.add(fz[f'{p}.f1.weight'].reshape(C // 16, C)),
                 aux=[P.blob.add(fz[f'{p}.f2.weight'].reshape(C, C // 16))])
        elif L.kind == 'CCVA':                                       # common.py:2654-2655, 3781-3786
            c_, lvl = a[1] // 2, lvl_in
            nxt = layers[L.i + 1] if L.i + 1 < len(layers) else None
            bvec = None
            if nxt is not None and nxt.kind == 'ADD' and nxt.sources() == [L.i, L.i - 1] and layers[L.i - 1].kind == 'CA':
                bvec = P.layer_out[L.i - 1]                          # ADD (common.py:3345-3349) folded into cv3's epilogue
                skip.add(nxt.i)
                g.layers.append(nxt.i)
                g.inputs.append((L.i - 1, bvec))
            cat = P.tensor(2 * c_, lvl)
            ta, tb = P.full(P.tensor(c_, lvl)), P.full(P.tensor(c_, lvl))
            q = k = None                                 # q/k projections are fused into the attention kernels
            P.conv_pair(L.i, W(f'{p}.cv1.conv'), W(f'{p}.cv2.conv'), x, ta, (cat, c_, c_))
            attention(L.i, f'{p}.m', N.OP_CRISSCROSS, ta, tb, q, k)
            attention(L.i, f'{p}.m1', N.OP_VERTICAL, tb, (cat, 0, c_), q, k)
            tgt = nxt if bvec is not None else L
            dst = out_view(tgt, lvl)
            P.conv(L.i, *W(f'{p}.cv3.conv'), P.full(cat), dst, bvec=bvec)
            if bvec is not None:
                P.layer_out[nxt.i], levels[nxt.i] = dst, lvl
        elif L.kind == 'ADD':
            raise NotImplementedError('stand-alone ADD (not preceded by CA/CCVA) is not part of Rep-YOLO')
        elif L.kind == 'RepConv':
            lvl = lvl_in
            dst = out_view(L, lvl)
            P.conv(L.i, *W(f'{p}.rbr_reparam'), x, dst, stride=a[3])
        elif L.kind == 'IDetect':                                    # yolo.py:135-168
            nc_, anchors, chs = a
            g.inputs = [(s, P.layer_out[s]) for s in srcs]
            ag = fz[f'{p}.anchor_grid'].reshape(len(chs), -1)
            for j, s in enumerate(srcs):
                w, b = W(f'{p}.m.{j}')
                stride = float(2 ** levels[s])
                P.op(N.OP_DETECT, L.i, in0=P.layer_out[s], out0=(pred, 0, 5 + nc_), out1=(raws[j], 0, 5 + nc_), ksize=1, stride=1,
                     cin=w.shape[1], cout=w.shape[0], level_idx=j, w_off=P.blob.add(w), b_off=P.blob.add(b),
                     fparam=[stride] + [float(v) for v in ag[j]])
            dst, lvl = None, 0
        else:
            raise NotImplementedError(L.kind)
        if dst is not None and not (L.kind == 'CCVA' and bvec is not None):
            P.layer_out[L.i] = dst
        if L.kind == 'CCVA' and bvec is not None:
            g.out_layer = L.i + 1
        levels[L.i] = lvl
        g.output = dst
        g.last_op = len(P.ops)
        P.groups.append(g)
    P.levels = levels
    return P
