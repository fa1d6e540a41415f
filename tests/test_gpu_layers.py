"""-m gpu: teacher-forced per-module parity and end-to-end sanity (8d(3)).

Every lowered group (one reference layer, or CA/CCVA/ADD fused) gets the ORACLE's fp32 input(s) cast to bf16, runs
through ry_run_ops, and is compared with the oracle's fp32 output of that layer.  Stated tolerances (relative L2),
= at most 2x the reference's own bf16-vs-fp32 error measured in the survey (Appendix D.3):
    single conv / RepConv / GSConv <= 8e-3 | composite blocks (DER, SPPCSPC, VoVGSCSP, CCVA+ADD) <= 3e-2
    data movement (MP, Upsample) <= 2e-3 | CA <= 6e-3
Decode (teacher-forced from fp32 head inputs): |d xy| <= 0.25 px, |d wh| <= 3e-2*wh + 0.1, |d obj|,|d cls| <= 6e-3
(wh = (2*sigmoid)^2 * anchor amplifies the bf16 operand rounding of the logit by up to 8*anchor; the same-operand check
against the oracle is 2e-3 absolute).
"""
import pytest
import torch

from oracle import repyolo_oracle as O

pytestmark = pytest.mark.gpu

TOL = {'RepS_Block': 8e-3, 'DER_Block': 3e-2, 'MP': 2e-3, 'SPPCSPC': 3e-2, 'GSConv': 8e-3, 'Upsample': 2e-3,
       'VoVGSCSP': 3e-2, 'Conv': 8e-3, 'CA': 6e-3, 'CCVA': 3e-2, 'RepConv': 8e-3}


@pytest.fixture(scope='module', params=[(2, 128), (1, 640), (1, (96, 160))], ids=['b2x128', 'b1x640', 'b1x96x160'])
def bound(request, oracle_model):
    import repyolo_b200 as R
    B, size = request.param
    layers, save, sd, fz = oracle_model
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    H, W = (size, size) if isinstance(size, int) else size          # non-square: letterboxed detect.py inputs, VerticalAttention's view chain
    x0 = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(100 + H))
    outs, pred, raws = O.forward_fused(fz, layers, save, x0)
    eng = m.engine('cuda:0')
    eng.bind(B, H, W)
    return dict(m=m, eng=eng, x0=x0, outs=outs, layers=layers, fz=fz, B=B, size=size)


def test_groups_teacher_forced(bound):
    from gpu_util import nchw_to_arena, arena_to_nchw, rel_l2
    eng, outs, layers = bound['eng'], bound['outs'], bound['layers']
    report, bad = [], []
    for g in eng.plan_ir.groups:
        kind = layers[g.layers[0]]['kind']
        if kind == 'IDetect':
            continue
        image = None
        for src, view in g.inputs:
            if src == -1:
                image = bound['x0'].cuda()
            else:
                nchw_to_arena(eng, view, outs[src])
        eng.run_ops(g.first_op, g.last_op, image=image)
        torch.cuda.synchronize()
        got = arena_to_nchw(eng, g.output)
        ref = outs[g.out_layer]
        assert got.shape == ref.shape, (g.layers, got.shape, ref.shape)
        e = rel_l2(got, ref)
        report.append((g.layers, kind, round(e, 5)))
        if not (e <= TOL[kind]):
            bad.append((g.layers, kind, e, TOL[kind]))
    print('\nteacher-forced rel-L2 per group:', report)
    assert not bad, bad


def test_detect_decode_teacher_forced(bound):
    """Head 1x1 conv + decode fused in the GEMM epilogue.  Inputs: the oracle's fp32 feature maps normalised to rms 1
    (random-weight Rep-YOLO can drive logits to +-1e3 where any bf16 operand rounding saturates the sigmoid).
    (a) vs the oracle run on the same bf16-rounded operands: only fp32 accumulation order differs -> tight;
    (b) vs the pure fp32 oracle: the stated decode tolerance of SURVEY.md 8d(4)."""
    import torch.nn.functional as F
    from gpu_util import nchw_to_arena
    eng, outs, layers, fz, B, size = (bound[k] for k in ('eng', 'outs', 'layers', 'fz', 'B', 'size'))
    g = eng.plan_ir.groups[-1]
    feats = [o / o.pow(2).mean().sqrt() for o in (outs[62], outs[63], outs[64])]
    for (src, view), f in zip(g.inputs, feats):
        nchw_to_arena(eng, view, f)
    H, W = (size, size) if isinstance(size, int) else size
    pred, raws = eng._outputs(B, H, W)
    eng.run_ops(g.first_op, g.last_op, pred=pred, raws=raws)
    torch.cuda.synchronize()
    p = pred.cpu()
    ag = fz['model.65.anchor_grid']
    # (a) same operands
    heads_b = [F.conv2d(f.bfloat16().float(), fz[f'model.65.m.{j}.weight'].bfloat16().float(), fz[f'model.65.m.{j}.bias'])
               for j, f in enumerate(feats)]
    pred_b, raws_b = O.decode_heads(heads_b, ag)
    for a, b in zip(raws, raws_b):
        assert a.shape == b.shape
        assert bool(((a.cpu() - b).abs() <= 2e-4 * (1 + b.abs())).all()), float((a.cpu() - b).abs().max())
    assert bool(((p - pred_b).abs() <= 2e-3 + 1e-4 * pred_b.abs()).all()), float((p - pred_b).abs().max())
    # (b) fp32 oracle, stated tolerance
    pred_ref, _ = O.decode_heads(O.run_fused_layer(fz, layers[-1], feats), ag)
    assert p.shape == pred_ref.shape
    assert float((p[..., :2] - pred_ref[..., :2]).abs().max()) <= 0.25
    dwh = (p[..., 2:4] - pred_ref[..., 2:4]).abs()
    assert bool((dwh <= 3e-2 * pred_ref[..., 2:4] + 0.1).all()), float(dwh.max())
    assert float((p[..., 4:] - pred_ref[..., 4:]).abs().max()) <= 6e-3


def test_idetect_fuseforward_signature(bound):
    """IDetect.fuseforward(list) mutates the list in place and returns (pred, list) like models/yolo.py:135-168."""
    m, outs = bound['m'], bound['outs']
    xs = [outs[62].cuda(), outs[63].cuda(), outs[64].cuda()]
    keep = xs
    pred, lst = m.model[-1].fuseforward(xs)
    assert lst is keep and lst[0].dim() == 5 and lst[0].shape[-1] == 6
    assert pred.shape[1] == sum(3 * o.shape[2] * o.shape[3] for o in (outs[62], outs[63], outs[64]))


def test_end_to_end_default_init():
    """in the default-init (collapsed, near-linear) regime the whole chain must stay within 2x the
    reference's own bf16 drift: |d xywh| <= 0.08 px, |d obj|,|d cls| <= 3.2e-3."""
    import repyolo_b200 as R
    layers, save, sd, fz = O.make_model(seed=0, mode='default')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(2, 3, 320, 320, generator=torch.Generator().manual_seed(123))
    pred, raws = m(x.cuda())
    _, pred_ref, raws_ref = O.forward_fused(fz, layers, save, x)
    d = (pred.cpu() - pred_ref).abs()
    assert float(d[..., :4].max()) <= 0.08, float(d[..., :4].max())
    assert float(d[..., 4:].max()) <= 3.2e-3, float(d[..., 4:].max())
    assert len(raws) == 3 and raws[0].shape == raws_ref[0].shape


def test_forward_then_nms_matches_oracle_nms():
    """Forward on the GPU, then NMS on GPU vs oracle NMS on the SAME candidates: bit-exact detections."""
    import repyolo_b200 as R
    from oracle import nms_oracle
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(2, 3, 640, 640, generator=torch.Generator().manual_seed(7))
    pred, _ = m(x.cuda())
    assert pred.shape == (2, 25200, 6) and bool(torch.isfinite(pred).all())
    for conf, iou in ((0.25, 0.45), (0.001, 0.65)):
        got = R.non_max_suppression(pred, conf, iou)
        ref = nms_oracle.non_max_suppression(pred.cpu(), conf, iou)
        for a, b in zip(got, ref):
            assert a.shape == b.shape and a.cpu().numpy().tobytes() == b.numpy().tobytes()


def test_config4_1280_runs_and_nms_matches():
    """BASELINE config 4 shape (1280x1280: 160/80/40 maps, 100800 candidates): the whole path runs, output is finite,
    and NMS on the GPU candidates is bit-exact against the oracle NMS."""
    import repyolo_b200 as R
    from oracle import nms_oracle
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(1, 3, 1280, 1280, generator=torch.Generator().manual_seed(11))
    pred, raws = m(x.cuda())
    assert pred.shape == (1, 100800, 6) and bool(torch.isfinite(pred).all())
    assert raws[0].shape == (1, 3, 160, 160, 6)
    got = R.non_max_suppression(pred, 0.25, 0.45)
    ref = nms_oracle.non_max_suppression(pred.cpu(), 0.25, 0.45)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and a.cpu().numpy().tobytes() == b.numpy().tobytes()


def test_default_init_teacher_forced_1280_der_block():
    """DER_Block L1 at 1280x1280 (640^2 maps, halo tiles + multi-source cv1) against the oracle, teacher-forced."""
    import repyolo_b200 as R
    from gpu_util import nchw_to_arena, arena_to_nchw, rel_l2
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    eng = m.engine('cuda:0')
    eng.bind(1, 1280, 1280)
    g = [gr for gr in eng.plan_ir.groups if gr.layers[0] == 1][0]
    x_in = torch.randn(1, 48, 640, 640, generator=torch.Generator().manual_seed(5)).abs()
    nchw_to_arena(eng, g.inputs[0][1], x_in)
    eng.run_ops(g.first_op, g.last_op)
    torch.cuda.synchronize()
    got = arena_to_nchw(eng, g.output)
    ref = O.run_fused_layer(fz, layers[1], x_in.bfloat16().float())
    if g.out_layer == 2:                              # the MP that follows (L2) is fused into cv1's epilogue
        ref = O.run_fused_layer(fz, layers[2], ref)
    assert rel_l2(got, ref) <= 3e-2


def test_uint8_image_input_matches_float_input():
    """uint8 NCHW input (what detect.py:73 ships to the device) with the /255 fused in the stem == float input / 255."""
    import repyolo_b200 as R
    layers, save, sd, fz = O.make_model(seed=0, mode='default')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    xb = torch.randint(0, 256, (2, 3, 128, 128), dtype=torch.uint8, generator=torch.Generator().manual_seed(4)).cuda()
    p8, _ = m(xb)
    pf, _ = m(xb.float() / 255.0)
    p8b, _ = m(xb)                                   # switching back and forth keeps working
    assert torch.equal(p8, p8b)
    d = (p8 - pf).abs()
    assert float(d[..., :4].max()) <= 0.05 and float(d[..., 4:].max()) <= 2e-3, (float(d[..., :4].max()), float(d[..., 4:].max()))


def test_tta_augment_matches_reference_recipe():
    """Model.forward(augment=True) (models/yolo.py:570-585): three scales, middle one flipped, de-scaled and concatenated.
    Checked against the same recipe driven through the CPU oracle (default init: near-linear regime)."""
    import math
    import torch.nn.functional as F
    import repyolo_b200 as R
    layers, save, sd, fz = O.make_model(seed=0, mode='default')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(1, 3, 192, 192, generator=torch.Generator().manual_seed(9))
    pred, none = m(x.cuda(), augment=True)
    assert none is None
    ys = []
    for si, fi in zip((1, 0.83, 0.67), (None, 3, None)):
        xi = x.flip(fi) if fi else x
        if si != 1:
            s = (int(192 * si), int(192 * si))
            xi = F.interpolate(xi, size=s, mode='bilinear', align_corners=False)
            hw = math.ceil(192 * si / 32) * 32
            xi = F.pad(xi, [0, hw - s[1], 0, hw - s[0]], value=0.447)
        yi = O.forward_fused(fz, layers, save, xi)[1]
        yi[..., :4] /= si
        if fi == 3:
            yi[..., 0] = 192 - yi[..., 0]
        ys.append(yi)
    ref = torch.cat(ys, 1)
    assert pred.shape == ref.shape
    d = (pred.cpu() - ref).abs()
    assert float(d[..., :4].max()) <= 0.15 and float(d[..., 4:].max()) <= 3.2e-3, (float(d[..., :4].max()), float(d[..., 4:].max()))
    p2, _ = m(x.cuda())                               # plain forward still works after the shape changes
    assert torch.allclose(p2, pred[:, :p2.shape[1]], atol=0, rtol=0)


def test_idetect_output_contracts():
    """end2end / include_nms / export branches of IDetect.fuseforward (models/yolo.py:158-166, convert() :189-199) against
    the fixture minted from the REFERENCE's own head (tests/golden/make_golden.py step 5), teacher-forced from the
    reference's fp32 L62-L64 outputs: decode tolerance of SURVEY 8d(4) on the boxes, 6e-3 on the scores."""
    import os
    import numpy as np
    import repyolo_b200 as R
    from conftest import GOLDEN
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    g = np.load(os.path.join(GOLDEN, 'convert_64.npz'))
    lay = np.load(os.path.join(GOLDEN, 'layers_64.npz'))
    feats = lambda: [torch.from_numpy(lay[f'layer{i}']).cuda() for i in (62, 63, 64)]
    det = m.model[-1]
    pred, raws = det.fuseforward(feats())
    ref_pred = torch.from_numpy(g['pred'])
    tol_xy, wh_ref = 0.25, ref_pred[..., 2:4]
    assert float((pred.cpu()[..., :2] - ref_pred[..., :2]).abs().max()) <= tol_xy
    assert bool(((pred.cpu()[..., 2:4] - wh_ref).abs() <= 3e-2 * wh_ref + 0.1).all())
    det.end2end = True
    e2e = det.fuseforward(feats())
    assert torch.equal(e2e, pred)                                        # yolo.py:160-161: the concatenated decode only
    assert float((e2e.cpu()[..., 4:] - torch.from_numpy(g['end2end'])[..., 4:]).abs().max()) <= 6e-3
    det.end2end, det.include_nms = False, True
    (box, score), = det.fuseforward(feats())
    rbox, rscore = torch.from_numpy(g['box']), torch.from_numpy(g['score'])
    assert box.shape == rbox.shape and score.shape == rscore.shape
    # xyxy = cxcy -/+ wh/2: |d| <= |d xy| + |d wh| / 2
    bound = tol_xy + 0.5 * (3e-2 * wh_ref.repeat(1, 1, 2) + 0.1)
    assert bool(((box.cpu() - rbox).abs() <= bound).all()), float((box.cpu() - rbox).abs().max())
    assert float((score.cpu() - rscore).abs().max()) <= 6e-3
    # and exactly the reference's arithmetic on the native decode (same matrix product, same score = cls * obj)
    cbox, cscore = O.convert(pred.cpu())
    assert torch.equal(box.cpu(), cbox) and torch.equal(score.cpu(), cscore)
    det.include_nms, det.export = False, True
    out = det.fuseforward(feats())
    assert isinstance(out, list) and len(out) == 3 and torch.equal(out[0], raws[0])
    det.export = False


def test_tta_augment_half_input_is_cast_not_reinterpreted():
    """test.py:104 sends img.half(); the augmented path must normalise the dtype like the plain forward."""
    import repyolo_b200 as R
    layers, save, sd, fz = O.make_model(seed=0, mode='default')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(12)).cuda()
    ref, _ = m(x.half().float(), augment=True)
    got, _ = m(x.half(), augment=True)
    assert torch.equal(got, ref)
    with pytest.raises(R.NativeError):
        m.engine(x.device).forward(x.half())          # the engine itself never reinterprets a foreign dtype
    with pytest.raises(R.NativeError):
        m.engine(x.device).forward(x.permute(0, 1, 3, 2))


def test_load_state_dict_after_fuse_invalidates():
    import repyolo_b200 as R
    layers, save, sd, fz = O.make_model(seed=0, mode='default')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    x = torch.rand(1, 3, 64, 64).cuda()
    m(x)
    m.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):
        m(x)
    m.fuse()
    m(x)


def test_detect_decode_nc2_generic_head():
    """Detect epilogue with na*(nc+5) = 21 columns (nc = 2): the generic column split (columns >= 18 used to be
    left unstaged).  Teacher-forced from normalised random features, same-operand check + stated decode tolerance."""
    import torch.nn.functional as F
    import repyolo_b200 as R
    from gpu_util import nchw_to_arena
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated', nc=2)
    m = R.Model(nc=2)
    m.load_state_dict(sd, strict=True)
    m.fuse()
    B, H, W = 2, 96, 160
    eng = m.engine('cuda:0')
    eng.bind(B, H, W)
    g = eng.plan_ir.groups[-1]
    gen = torch.Generator().manual_seed(21)
    feats = [torch.randn(B, c, H >> l, W >> l, generator=gen) for c, l in ((256, 3), (512, 4), (1024, 5))]
    for (src, view), f in zip(g.inputs, feats):
        nchw_to_arena(eng, view, f)
    pred, raws = eng._outputs(B, H, W)
    assert pred.shape[-1] == 7
    eng.run_ops(g.first_op, g.last_op, pred=pred, raws=raws)
    torch.cuda.synchronize()
    ag = fz['model.65.anchor_grid']
    heads_b = [F.conv2d(f.bfloat16().float(), fz[f'model.65.m.{j}.weight'].bfloat16().float(), fz[f'model.65.m.{j}.bias'])
               for j, f in enumerate(feats)]
    pred_b, raws_b = O.decode_heads(heads_b, ag)
    for a, b in zip(raws, raws_b):
        assert a.shape == b.shape
        assert bool(((a.cpu() - b).abs() <= 2e-4 * (1 + b.abs())).all()), float((a.cpu() - b).abs().max())
    p = pred.cpu()
    assert bool(((p - pred_b).abs() <= 2e-3 + 1e-4 * pred_b.abs()).all()), float((p - pred_b).abs().max())


# ---- BASELINE batch sizes (the B=64 code paths -- tile ranges spanning images, the two-image per-image-vector
# staging, magic-division image indices at 6.5 M pixels, thousands of tiles per persistent CTA -- ran unchecked) ----
def _arena_put_tiled(eng, view, x4, B):
    """oracle activations of a few images, tiled along the batch to B images, written into a plan view (GPU-side permute)."""
    from gpu_util import nchw_to_arena
    rep = B // x4.shape[0]
    nchw_to_arena(eng, view, x4.cuda().repeat(rep, *([1] * (x4.dim() - 1))))


def _arena_get(eng, view):
    t, off, n = view
    src = eng.tensor(t)
    if src.dim() == 2:
        return src[:, off:off + n].float().reshape(src.shape[0], n, 1, 1)
    return src[..., off:off + n].float().permute(0, 3, 1, 2)


def _run_groups_tiled(m, layers, outs, x0, B, H, W, n_ref, want_first_layers=None):
    """Teacher-forced groups at batch B with the oracle's n_ref images repeated B / n_ref times.  Checks (1) images
    0..n_ref-1 and the LAST n_ref images against the oracle (stated tolerances), (2) every image equals its period-n_ref
    twin bit for bit (a tile, image index or staging slip at large B shows up as a difference between twins)."""
    from gpu_util import rel_l2
    eng = m.engine('cuda:0')
    eng.bind(B, H, W)
    report, bad = [], []
    for g in eng.plan_ir.groups:
        kind = layers[g.layers[0]]['kind']
        if kind == 'IDetect' or (want_first_layers is not None and g.layers[0] not in want_first_layers):
            continue
        image = None
        for src, view in g.inputs:
            if src == -1:
                image = x0.cuda().repeat(B // n_ref, 1, 1, 1).contiguous()
            else:
                _arena_put_tiled(eng, view, outs[src], B)
        eng.run_ops(g.first_op, g.last_op, image=image)
        torch.cuda.synchronize()
        got = _arena_get(eng, g.output)
        ref = outs[g.out_layer]
        assert got.shape[1:] == ref.shape[1:] and got.shape[0] == B
        twins = got.reshape(B // n_ref, n_ref, *got.shape[1:])
        if not bool((twins == twins[:1]).all()):
            bad.append((g.layers, kind, 'twin images differ', float((twins - twins[:1]).abs().max())))
        e_first, e_last = rel_l2(got[:n_ref].cpu(), ref), rel_l2(got[B - n_ref:].cpu(), ref)
        report.append((g.layers, kind, round(e_first, 5), round(e_last, 5)))
        if not (e_first <= TOL[kind] and e_last <= TOL[kind]):
            bad.append((g.layers, kind, e_first, e_last, TOL[kind]))
        del got, twins
    print('\nteacher-forced rel-L2 per group at B =', B, report)
    assert not bad, bad
    return eng


def test_groups_teacher_forced_batch64_640(oracle_model):
    """BASELINE configs[1]: every lowered group at B = 64 @ 640x640 (oracle on 4 images: the GPU batch is those 4 repeated
    16 times, checked on images 0-3 and 60-63 + twin equality over all 64), then Detect + decode at B = 64."""
    import torch.nn.functional as F
    import repyolo_b200 as R
    layers, save, sd, fz = oracle_model
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    B, H, W, n_ref = 64, 640, 640, 4
    x0 = torch.rand(n_ref, 3, H, W, generator=torch.Generator().manual_seed(64))
    outs, pred_ref, raws_ref = O.forward_fused(fz, layers, save, x0)
    eng = _run_groups_tiled(m, layers, outs, x0, B, H, W, n_ref)
    # Detect head + decode (normalised features, see test_detect_decode_teacher_forced)
    g = eng.plan_ir.groups[-1]
    feats = [o / o.pow(2).mean().sqrt() for o in (outs[62], outs[63], outs[64])]
    for (src, view), f in zip(g.inputs, feats):
        _arena_put_tiled(eng, view, f, B)
    pred, raws = eng._outputs(B, H, W)
    eng.run_ops(g.first_op, g.last_op, pred=pred, raws=raws)
    torch.cuda.synchronize()
    tw = pred.reshape(B // n_ref, n_ref, *pred.shape[1:])
    assert bool((tw == tw[:1]).all())
    heads_b = [F.conv2d(f.bfloat16().float(), fz[f'model.65.m.{j}.weight'].bfloat16().float(), fz[f'model.65.m.{j}.bias'])
               for j, f in enumerate(feats)]
    pred_b, raws_b = O.decode_heads(heads_b, fz['model.65.anchor_grid'])
    for sl in (slice(0, n_ref), slice(B - n_ref, B)):
        p = pred[sl].cpu()
        assert bool(((p - pred_b).abs() <= 2e-3 + 1e-4 * pred_b.abs()).all()), float((p - pred_b).abs().max())
        for a, b in zip(raws, raws_b):
            assert bool(((a[sl].cpu() - b).abs() <= 2e-4 * (1 + b.abs())).all())


def test_der_block_teacher_forced_batch16_1280(oracle_model):
    """BASELINE configs[3]: DER_Block L1 (+ fused MP) and the stem at B = 16 @ 1280x1280 (oracle on 2 images, tiled 8x)."""
    import repyolo_b200 as R
    layers, save, sd, fz = oracle_model
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    B, H, W, n_ref = 16, 1280, 1280, 2
    x0 = torch.rand(n_ref, 3, H, W, generator=torch.Generator().manual_seed(1280))
    outs = {}
    with torch.no_grad():
        y = O.run_fused_layer(fz, layers[0], x0)
        outs[0] = y
        y = O.run_fused_layer(fz, layers[1], y)
        outs[1] = y
        outs[2] = O.run_fused_layer(fz, layers[2], y)
    _run_groups_tiled(m, layers, outs, x0, B, H, W, n_ref, want_first_layers=(0, 1))


def test_cuda_graph_replay_matches_eager():
    """Model.cuda_graph: the forward pass replayed as one CUDA graph (static output slots, captured the second time an input
    address is seen) returns exactly what the eager launches return -- alternating input buffers, with and without the fused
    decode filter, across a shape change."""
    import repyolo_b200 as R
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    gen = torch.Generator().manual_seed(77)
    xs = [torch.rand(2, 3, 128, 160, generator=gen).cuda() for _ in range(2)]
    want = [m(x)[0].clone() for x in xs]
    m.cuda_graph = True
    eng = m.engine('cuda:0')
    for it in range(8):                                      # calls 0-1 eager (first sight), 2-3 capture, 4+ replay
        pred, raws = m(xs[it % 2])
        assert torch.equal(pred, want[it % 2]), it
        assert len(raws) == 3 and raws[0].shape == (2, 3, 16, 20, 6)
    assert len(eng._graphs) == 2
    xs[0].copy_(xs[1])                                       # same address, new content: the replay reads the new content
    assert torch.equal(m(xs[0])[0], want[1])
    m.decode_filter = 0.25
    for it in range(6):
        pred, _ = m(xs[1])
        assert torch.equal(pred, want[1]) and hasattr(pred, '_ry_cand')
        a = R.non_max_suppression(pred, 0.25, 0.45)
        b = R.non_max_suppression(want[1], 0.25, 0.45)
        assert all(torch.equal(p, q) for p, q in zip(a, b))
    m.decode_filter = None
    y = torch.rand(1, 3, 64, 64, generator=gen).cuda()       # another shape: new engine / binding, graphs of the old one untouched
    m.cuda_graph = False
    ref = m(y)[0].clone()
    m.cuda_graph = True
    for it in range(4):
        assert torch.equal(m(y)[0], ref)
    for it in range(3):
        assert torch.equal(m(xs[1])[0], want[1])
