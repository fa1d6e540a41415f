"""Pins the CPU oracle (oracle/) to fixtures produced by the real reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import nms_oracle
from oracle import repyolo_oracle as O


def rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def test_state_dict_inventory_matches_reference():
    keys = json.load(open(os.path.join(GOLDEN, 'state_keys.json')))
    layers, save = O.build_graph()
    mine = {k: list(v) for k, v in O.param_shapes(layers).items()}
    ref = {k: v for k, v in keys['unfused']}
    assert mine == ref
    assert save == keys['save']
    assert list(O.STRIDES) == keys['stride']


def test_fold_matches_reference_fuse(oracle_model):
    """Fold KAT: own fold of the unfused synthetic weights vs the reference's Model.fuse()."""
    _, _, _, fz = oracle_model
    ref = np.load(os.path.join(GOLDEN, 'fold_digest.npz'))
    n = 0
    for k in ref.files:
        if k not in fz:
            assert '.ia.' in k or '.im.' in k, k          # implicit layers stay in the tree, unused (yolo.py:170-182)
            continue
        d = O_digest(fz[k])
        np.testing.assert_allclose(d, ref[k], rtol=2e-6, atol=1e-7, err_msg=k)
        n += 1
    assert n >= 196 * 2


def O_digest(t):
    f = t.detach().double().flatten()
    idx = torch.linspace(0, f.numel() - 1, 8).long()
    return np.concatenate([[f.sum().item(), f.abs().sum().item(), (f * f).sum().item()], f[idx].numpy()])


@pytest.mark.parametrize('tag', ['64', '64x96'])
def test_teacher_forced_layers_match_reference(oracle_model, tag):
    layers, save, _, fz = oracle_model
    g = np.load(os.path.join(GOLDEN, f'layers_{tag}.npz'))
    x0 = torch.from_numpy(g['x'])
    if tag == '64':
        outs = [torch.from_numpy(g[f'layer{i}']) for i in range(len(layers) - 1)]
        for i, L in enumerate(layers[:-1]):
            y = O.run_fused_layer(fz, L, O.layer_inputs(layers, outs, x0, i))
            assert y.shape == outs[i].shape
            assert rel_l2(y, outs[i]) < 1e-5, (i, L['kind'])
        heads = O.run_fused_layer(fz, layers[-1], [outs[62], outs[63], outs[64]])
        pred, raws = O.decode_heads(heads, fz['model.65.anchor_grid'])
        np.testing.assert_allclose(pred.numpy(), g['pred'], rtol=1e-5, atol=1e-4)
        for j in range(3):
            np.testing.assert_allclose(raws[j].numpy(), g[f'raw{j}'], rtol=1e-5, atol=1e-5)
    else:   # rectangular input: end-to-end (short chain at 64x96 keeps fp32 drift tiny) + the H != W attention views
        outs, pred, raws = O.forward_fused(fz, layers, save, x0)
        for i, L in enumerate(layers[:-1]):
            if L['kind'] in ('CCVA', 'RepConv'):
                assert rel_l2(outs[i], torch.from_numpy(g[f'layer{i}'])) < 1e-4, i
        assert rel_l2(pred, torch.from_numpy(g['pred'])) < 1e-4


def test_unfused_equals_fused(oracle_model):
    """The reference's own consistency check: fused vs unfused forward."""
    layers, save, sd, fz = oracle_model
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(11))
    a = O.forward_unfused(sd, layers, save, x)
    b, _, _ = O.forward_fused(fz, layers, save, x)
    for i in (0, 1, 9, 14, 21):
        assert rel_l2(a[i], b[i]) < 1e-4, i


def _cases():
    g = np.load(os.path.join(GOLDEN, 'nms_cases.npz'))
    meta = json.loads(bytes(g['meta']).decode())
    return g, meta


def test_nms_oracle_bit_exact_vs_reference_outputs():
    g, meta = _cases()
    for name, kw in meta.items():
        pred = torch.from_numpy(g[f'{name}.pred'])
        outs = nms_oracle.non_max_suppression(pred, **kw)
        counts = g[f'{name}.counts']
        assert [o.shape[0] for o in outs] == counts.tolist(), name
        got = torch.cat(outs, 0).numpy()
        assert got.tobytes() == g[f'{name}.out'].tobytes(), name


def test_greedy_nms_bit_exact_vs_torchvision_cpu():
    tv = pytest.importorskip('torchvision')
    gen = torch.Generator().manual_seed(7)
    for n, thr, quant in ((2000, 0.45, None), (3000, 0.65, None), (2000, 0.5, 8.0), (1, 0.5, None), (500, 0.6, 2.0)):
        xy = torch.rand(n, 2, generator=gen) * 300
        wh = torch.rand(n, 2, generator=gen) * 80 + 2
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=gen)
        if quant:
            boxes = (boxes / quant).round() * quant
            scores = (scores * 20).round() / 20
        boxes = boxes + (torch.randint(0, 3, (n, 1), generator=gen).float() * 4096)
        ref = tv.ops.nms(boxes, scores, thr).numpy()
        got = nms_oracle.greedy_nms(boxes.numpy(), scores.numpy(), thr)
        assert np.array_equal(ref, got), (n, thr, quant)
        assert np.array_equal(ref[:300], nms_oracle.greedy_nms(boxes.numpy(), scores.numpy(), thr, 300))


def test_convert_and_end2end_contracts_match_reference(oracle_model):
    """include_nms -> IDetect.convert() (models/yolo.py:189-199) and end2end -> cat(z) (yolo.py:160-161): the oracle's
    restatement on the reference's own head inputs reproduces the reference-minted fixture bit for bit."""
    layers, save, sd, fz = oracle_model
    g = np.load(os.path.join(GOLDEN, 'convert_64.npz'))
    lay = np.load(os.path.join(GOLDEN, 'layers_64.npz'))
    feats = [torch.from_numpy(lay[f'layer{i}']) for i in (62, 63, 64)]
    pred, _ = O.decode_heads(O.run_fused_layer(fz, layers[-1], feats), fz['model.65.anchor_grid'])
    assert rel_l2(pred, torch.from_numpy(g['pred'])) <= 1e-6
    box, score = O.convert(torch.from_numpy(g['pred']))
    assert np.array_equal(box.numpy(), g['box']) and np.array_equal(score.numpy(), g['score'])
    assert np.array_equal(g['end2end'], g['pred'])


def test_nms_apriori_labels_match_reference():
    """general.py:981-987 (test.py --save-hybrid): label rows appended behind the filtered candidates; fixture minted from the
    reference's own non_max_suppression(labels=[...]) by tests/golden/make_golden.py step 6."""
    g = np.load(os.path.join(GOLDEN, 'nms_labels.npz'))
    pred = torch.from_numpy(g['pred'])
    labels = [torch.from_numpy(g[f'labels{b}']) for b in range(pred.shape[0])]
    for c in range(2):
        kw = json.loads(bytes(g[f'c{c}.kw']).decode())
        outs = nms_oracle.non_max_suppression(pred, labels=labels, **kw)
        assert [o.shape[0] for o in outs] == g[f'c{c}.counts'].tolist()
        assert torch.cat(outs, 0).numpy().tobytes() == g[f'c{c}.out'].tobytes()
