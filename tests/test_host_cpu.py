"""-m "not gpu": host logic of the product package + the C-ABI library loads and exports every declared symbol."""
import ctypes
import importlib
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def test_library_builds_and_exports_every_declared_symbol():
    b = importlib.import_module('rep-yolo_b200._build')
    so = b.build()                                     # nvcc cross-compiles sm_100a without a GPU
    header = open(os.path.join(ROOT, 'include', 'repyolo_b200.h')).read()
    declared = sorted(set(re.findall(r'^(?:int|void|const char \*)\s*(ry_[a-z_0-9]+)\s*\(', header, re.M)))
    assert len(declared) >= 14
    lib = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ry_abi_version() == 5
    N = importlib.import_module('rep-yolo_b200._lib')
    assert sorted(N.EXPORTS) == declared               # the ctypes binding covers the whole header
    assert lib.ry_abi_sizeof(1) == ctypes.sizeof(N.OpDesc) and lib.ry_abi_sizeof(0) == ctypes.sizeof(N.TensorDesc)
    N.lib()                                            # full binding incl. its layout self-check


def test_model_state_dict_matches_reference_inventory():
    import repyolo_b200 as R
    keys = json.load(open(os.path.join(GOLDEN, 'state_keys.json')))
    m = R.Model()
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == {k: v for k, v in keys['unfused']}
    assert m.save == keys['save'] and m.stride.tolist() == keys['stride']
    det = m.model[-1]
    assert (det.nl, det.na, det.no, det.nc) == (3, 3, 6, 1) and m.names == ['0']
    assert R.rep_yolo_cfg()['anchors'][2] == [44, 114, 48, 172, 80, 112]


def test_product_fold_matches_reference_fuse(oracle_model):
    """Fold KAT for the PRODUCT's fold pass (rep-yolo_b200/fold.py) against the reference Model.fuse() digest."""
    import repyolo_b200 as R
    _, _, sd, _ = oracle_model
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    assert m.fuse() is m and m.fuse() is m             # returns self, idempotent
    ref = np.load(os.path.join(GOLDEN, 'fold_digest.npz'))
    n = 0
    for k in ref.files:
        if k not in m._fused:
            assert '.ia.' in k or '.im.' in k, k
            continue
        f = m._fused[k].double().flatten()
        idx = torch.linspace(0, f.numel() - 1, 8).long()
        d = np.concatenate([[f.sum().item(), f.abs().sum().item(), (f * f).sum().item()], f[idx].numpy()])
        assert abs(d[0] - ref[k][0]) <= 2e-6 * ref[k][1] + 1e-7, k           # signed sum: tolerance relative to sum |w|
        np.testing.assert_allclose(d[1:], ref[k][1:], rtol=2e-6, atol=2e-7, err_msg=k)
        n += 1
    assert n >= 392


def test_lowering_invariants(oracle_model):
    import repyolo_b200 as R
    planner = importlib.import_module('rep-yolo_b200.planner')
    N = importlib.import_module('rep-yolo_b200._lib')
    _, _, sd, _ = oracle_model
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    P = planner.lower(m._layers, m._fused, 1)
    kinds = [o.kind for o in P.ops]
    n_chain = sum(1 + o.n_post for o in P.ops if o.kind == N.OP_CONV_CHAIN)                        # fused 3x3 -> 1x1 [-> 1x1]
    n_pair = sum(1 for o in P.ops if o.kind == N.OP_CONV and o.out1.tensor >= 0 and o.out1.tensor != o.out0.tensor)   # cv1 + cv2 merged
    assert kinds.count(N.OP_CONV) + kinds.count(N.OP_DETECT) + kinds.count(N.OP_STEM) + n_chain + n_pair == 130   # dense convs
    n_pool = sum(1 for o in P.ops if o.kind == N.OP_CONV and o.level_idx == 1)                      # MP fused into DER_Block.cv1
    assert kinds.count(N.OP_DW5) == 18 and kinds.count(N.OP_MAXPOOL2) + n_pool == 6 and kinds.count(N.OP_UPSAMPLE2) == 2
    assert kinds.count(N.OP_CRISSCROSS) == 6 and kinds.count(N.OP_VERTICAL) == 6 and kinds.count(N.OP_CA) == 6
    written = {}
    for o in P.ops:                                    # channel ranges written by different ops never partially overlap
        for v in (o.out0, o.out1, o.out2):
            if v.tensor < 0 or P.tensors[v.tensor].kind == N.T_EXTERNAL:
                continue
            if o.kind in (N.OP_ATTN_QK,):
                continue                               # q/k scratch is reused by the two attention modules of a CCVA
            rng = written.setdefault(v.tensor, [])
            if (v.c_off, v.c_off + v.c_len) in rng:
                continue                               # whole-range rewrite = scratch reuse (DER h1/h2)
            for lo, hi in rng:
                assert v.c_off >= hi or v.c_off + v.c_len <= lo, ('partially overlapping writes', v.tensor)
            rng.append((v.c_off, v.c_off + v.c_len))
    for t, rng in written.items():                     # and every channel of every tensor is produced
        assert sum(hi - lo for lo, hi in rng) == P.tensors[t].channels, t
    flops = 0
    for o in P.ops:
        if o.kind in (N.OP_CONV, N.OP_DETECT, N.OP_STEM):
            lvl = P.tensors[o.out0.tensor].level if o.kind != N.OP_DETECT else P.tensors[o.in0.tensor].level
            if o.kind == N.OP_CONV and o.level_idx == 1:
                lvl -= 1                               # fused MaxPool2d: the conv runs on the un-pooled grid
            flops += 2 * o.cin * o.cout * o.ksize ** 2 * (640 >> lvl) ** 2
        elif o.kind == N.OP_CONV_CHAIN:
            px, prev = (640 >> P.tensors[o.in0.tensor].level) ** 2, o.cout
            flops += 2 * o.cin * o.cout * 9 * px
            for i in range(o.n_post):
                flops += 2 * prev * o.post_cout[i] * px
                prev = o.post_cout[i]
    assert abs(flops / 1e9 - 68.733) < 0.01            # dense-conv GFLOP / image @640


def test_no_cpu_fallback():
    import repyolo_b200 as R
    m = R.Model().fuse()
    with pytest.raises(R.NativeError):
        m(torch.rand(1, 3, 64, 64))
    with pytest.raises(R.NativeError):
        R.non_max_suppression(torch.rand(1, 10, 6))
    with pytest.raises(RuntimeError):
        R.Model()(torch.rand(1, 3, 64, 64))            # unfused forward is not part of the deployed path


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'rep-yolo_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src, os.path.join(dirpath, f)


def test_compat_install_patches_reference_modules():
    """compat.install() rebinds attempt_load / TracedModel / non_max_suppression inside (stand-in) reference modules,
    including the copies `from x import y` left in an already imported caller."""
    import sys
    import types
    import repyolo_b200 as R
    compat = importlib.import_module('rep-yolo_b200.compat')
    fake = {}
    for name in ('models', 'models.experimental', 'utils', 'utils.torch_utils', 'utils.general', 'detect'):
        fake[name] = types.ModuleType(name)
    def ref_fn(*a, **k):
        raise AssertionError('reference implementation called')
    for mod, attr in (('models.experimental', 'attempt_load'), ('utils.torch_utils', 'TracedModel'),
                      ('utils.general', 'non_max_suppression'), ('detect', 'attempt_load'), ('detect', 'non_max_suppression')):
        f = types.FunctionType(ref_fn.__code__, globals(), attr)
        f.__module__ = 'models.experimental' if attr == 'attempt_load' else ('utils.torch_utils' if attr == 'TracedModel' else 'utils.general')
        setattr(fake[mod], attr, f)
    saved = {k: sys.modules.get(k) for k in fake}
    sys.modules.update(fake)
    try:
        patched = compat.install()
        assert ('detect', 'attempt_load') in patched and ('utils.general', 'non_max_suppression') in patched
        assert fake['detect'].non_max_suppression is R.non_max_suppression
        assert fake['utils.torch_utils'].TracedModel is compat.TracedModel
        assert fake['models.experimental'].attempt_load is compat.attempt_load
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_from_reference_accepts_reference_state_dict(oracle_model):
    """from_reference() needs only .yaml / .state_dict() / .names of the reference Model (same parameter names)."""
    import types
    import repyolo_b200 as R
    arch = importlib.import_module('rep-yolo_b200.arch')
    _, _, sd, _ = oracle_model
    ref = types.SimpleNamespace(yaml=arch.rep_yolo_cfg(), state_dict=lambda: sd, names=['person'])
    m = R.from_reference(ref)
    assert m.names == ['person'] and m.fuse() is m
    t = R.TracedModel(m, 'cpu', 640)
    assert t.model is m and t.detect_layer is m.model[-1]
