"""-m gpu: the tcgen05 implicit-GEMM conv kernel (through ry_plan_* / ry_run_ops) vs torch fp32 conv2d on the same
bf16-rounded operands.  Tolerance: bf16 output rounding (2^-8 relative) + fp32 accumulation-order noise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # cin, cout, k, s, H, W, B, act, kwargs
    (64, 64, 1, 1, 32, 32, 2, True, {}),
    (128, 256, 1, 1, 32, 32, 2, True, {}),
    (64, 64, 3, 1, 32, 32, 2, True, {}),
    (128, 128, 3, 1, 64, 64, 1, False, {}),
    (48, 48, 3, 1, 32, 32, 2, True, {}),                 # 16-channel K blocks (32 B swizzle)
    (24, 24, 3, 1, 32, 32, 2, True, {}),                 # K block padded by TMA zero fill
    (48, 24, 1, 1, 32, 32, 2, True, {}),
    (144, 48, 1, 1, 64, 64, 1, True, {}),
    (32, 64, 1, 1, 32, 32, 1, True, {}),                 # 32-channel K blocks (64 B swizzle)
    (96, 32, 3, 1, 32, 32, 1, True, {}),
    (512, 512, 3, 1, 32, 32, 2, True, {}),               # two N tiles, many K blocks
    (256, 1024, 1, 1, 32, 32, 1, True, {}),              # four N tiles
    (128, 64, 3, 2, 64, 64, 2, True, {}),                # stride 2 through the parity-phase tensor maps
    (64, 32, 3, 2, 32, 64, 1, True, {}),
    (128, 64, 3, 1, 32, 32, 2, True, dict(c_pad_in=64, c_pad_out=32)),      # channel-offset views (concat slots)
    (128, 128, 1, 1, 32, 32, 2, False, dict(with_res=True)),                # residual epilogue (GSBottleneck shortcut)
    (128, 128, 1, 1, 32, 32, 2, True, dict(with_bvec=True)),                # broadcast add (CA + CCVA)
    (128, 64, 1, 1, 32, 32, 2, True, dict(split=True)),                     # split store (GSConv shuffle)
    (256, 128, 3, 1, 32, 96, 1, True, {}),                                  # non-square, partial tiles
    (512, 256, 1, 1, 640 // 32 * 32, 640 // 32 * 32, 1, True, {}),           # large M (persistent loop, TMEM double buffer)
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(str(v) for v in c[:7]))
def test_conv_vs_torch(case):
    from gpu_util import single_conv_engine, nchw_to_arena, arena_to_nchw
    cin, cout, k, s, H, W, B, act, kw = case
    g = torch.Generator().manual_seed(cin * 131 + cout * 7 + k + s)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g) * 0.5
    x = torch.randn(B, cin, H, W, generator=g)
    eng, vin, dst, dst2, res, bvec = single_conv_engine(w, b, B, H, W, s, act, **kw)
    for t in range(len(eng.plan_ir.tensors)):
        eng.tensor(t).fill_(7.0)                      # poison: untouched channels must stay untouched
    nchw_to_arena(eng, vin, x)
    xb, wb = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv2d(xb, wb, b, s, k // 2)
    if act:
        ref = F.silu(ref)
    if res is not None:
        r = torch.randn(ref.shape, generator=g)
        nchw_to_arena(eng, res, r)
        ref = ref + r.bfloat16().float()
    if bvec is not None:
        v = torch.randn(B, cout, 1, 1, generator=g)
        nchw_to_arena(eng, bvec, v)
        ref = ref + v
    eng.run_ops(0, 1)
    torch.cuda.synchronize()
    if dst2 is None:
        got = arena_to_nchw(eng, dst)
    else:
        got = torch.cat([arena_to_nchw(eng, dst), arena_to_nchw(eng, dst2)], 1)
    err = (got - ref).abs()
    tol = 2e-2 + 1e-2 * ref.abs()
    assert bool((err <= tol).all()), f'max err {err.max().item():.4f} at ref {ref.flatten()[err.argmax()].item():.4f}'
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 4e-3, rel
    # channels outside the destination view are untouched
    full = eng.tensor(dst[0]).float().cpu()
    mask = torch.ones(full.shape[-1], dtype=torch.bool)
    mask[dst[1]:dst[1] + dst[2]] = False
    if dst2 is not None:
        mask[dst2[1]:dst2[1] + dst2[2]] = False
    if mask.any():
        assert bool((full[..., mask] == 7.0).all())


def test_multi_source_1x1_conv():
    """1x1 conv over the channel concatenation of three separate tensors (DER_Block.cv1 without the torch.cat)."""
    import repyolo_b200 as R
    from gpu_util import N, planner, nchw_to_arena, arena_to_nchw
    g = torch.Generator().manual_seed(77)
    B, H, W, cs, cout = 2, 32, 32, (48, 24, 48), 48
    cin = sum(cs)
    w = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g) * 0.5
    xs = [torch.randn(B, c, H, W, generator=g) for c in cs]
    P = planner.Plan()
    tins = [P.full(P.tensor(c, 0)) for c in cs]
    dst = P.full(P.tensor(cout, 0))
    P.conv(0, w, b, tins, dst)
    eng = R.NativeEngine(P, 1, 'cuda:0')
    eng.bind(B, H, W)
    for v, x in zip(tins, xs):
        nchw_to_arena(eng, v, x)
    eng.run_ops(0, 1)
    torch.cuda.synchronize()
    got = arena_to_nchw(eng, dst)
    ref = F.silu(F.conv2d(torch.cat(xs, 1).bfloat16().float(), w.bfloat16().float(), b))
    assert float((got - ref).norm() / ref.norm()) < 4e-3


@pytest.mark.parametrize('case', [
    # sources, cout, image H, image W, pyramid level of the conv's map, B
    ((48, 24, 48), 48, 128, 128, 1, 2),       # DER_Block L1 cv1 + MP: 64x64 map, tile teams
    ((48, 24, 48), 128, 128, 128, 2, 2),      # L3: 32x32 map
    ((128, 64, 128), 256, 320, 320, 3, 2),    # L5: 40x40 map, one N tile, tiles spanning two images
    ((256, 128, 256), 512, 320, 320, 4, 3),   # L7: 20x20 map, two N tiles, odd batch
    ((64,), 64, 96, 160, 2, 1),               # single source, non-square 24x40 map
])
def test_1x1_conv_with_fused_maxpool(case):
    """DER_Block.cv1 with the following MP (common.py:32-38: MaxPool2d(2, 2)) fused in the epilogue: only the pooled map is stored."""
    import repyolo_b200 as R
    from gpu_util import planner, nchw_to_arena, arena_to_nchw
    cs, cout, IH, IW, lvl, B = case
    H, W = IH >> lvl, IW >> lvl
    g = torch.Generator().manual_seed(sum(cs) + cout + H)
    cin = sum(cs)
    w = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g) * 0.5
    xs = [torch.randn(B, c, H, W, generator=g) for c in cs]
    P = planner.Plan()
    tins = [P.full(P.tensor(c, lvl)) for c in cs]
    tout = P.tensor(cout + 16, lvl + 1)
    dst = (tout, 8, cout)
    P.conv(0, w, b, tins if len(tins) > 1 else tins[0], dst, pool=True)
    eng = R.NativeEngine(P, 1, 'cuda:0')
    eng.bind(B, IH, IW)
    eng.tensor(tout).fill_(7.0)
    for v, x in zip(tins, xs):
        nchw_to_arena(eng, v, x)
    eng.run_ops(0, 1)
    torch.cuda.synchronize()
    got = arena_to_nchw(eng, dst)
    ref = F.max_pool2d(F.silu(F.conv2d(torch.cat(xs, 1).bfloat16().float(), w.bfloat16().float(), b)), 2, 2)
    assert got.shape == ref.shape
    err = (got - ref).abs()
    assert bool((err <= 2e-2 + 1e-2 * ref.abs()).all()), float(err.max())
    assert float((got - ref).norm() / ref.norm()) < 4e-3
    full = eng.tensor(tout).float().cpu()
    assert bool((full[..., :8] == 7.0).all()) and bool((full[..., 8 + cout:] == 7.0).all())
