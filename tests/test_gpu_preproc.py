"""-m gpu: letterbox / preprocess / scale_coords kernels (through the C ABI) against the CPU oracle and the fixtures minted
from the real reference -- bit-exact (uint8 pixels incl. cv2's fixed-point bilinear; fp32 boxes)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import letterbox_oracle as LO

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location('make_golden_letterbox', os.path.join(HERE, 'golden', 'make_golden_letterbox.py'))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(HERE, 'golden', 'letterbox_cases.npz'))


@pytest.mark.parametrize('i', range(len(G.CASES)))
def test_letterbox_golden(golden, i):
    import repyolo_b200 as R
    shape, new_shape, auto, fill, up = G.CASES[i]
    img0 = G.image(i, shape)
    img, ratio, pad = R.letterbox(img0, new_shape, auto=auto, scaleFill=fill, scaleup=up, stride=32)
    chw = img.cpu().numpy()[:, :, ::-1].transpose(2, 0, 1)
    assert chw.shape == golden[f'img_{i}'].shape and np.array_equal(chw, golden[f'img_{i}'])
    np.testing.assert_array_equal(np.array([ratio[0], ratio[1], pad[0], pad[1]], np.float64), golden[f'ratio_pad_{i}'])
    if not fill and up:                                   # the LoadImages defaults: fused RGB / CHW packing
        out, _, _ = R.preprocess(torch.from_numpy(img0).cuda(), new_shape, stride=32, auto=auto)
        assert np.array_equal(out.cpu().numpy(), golden[f'img_{i}'])


@pytest.mark.parametrize('shape,size', [((1080, 1920), 640), ((480, 640), 640), ((720, 1280), 1280), ((333, 500), 640), ((2, 3), 64)])
def test_preprocess_vs_oracle(shape, size):
    import repyolo_b200 as R
    img0 = np.random.default_rng(shape[0]).integers(0, 256, (shape[0], shape[1], 3), dtype=np.uint8)
    got, ratio, pad = R.preprocess(img0, size, stride=32)
    ref = LO.preprocess(img0, size, stride=32)
    assert got.shape == ref.shape and np.array_equal(got.cpu().numpy(), ref)


def test_preprocess_feeds_the_model():
    """raw BGR image -> preprocess (GPU) -> uint8 batch -> Model.forward: the whole device-side path of detect.py:66-90"""
    import repyolo_b200 as R
    from oracle import repyolo_oracle as O
    layers, save, sd, fz = O.make_model(seed=0, mode='default')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    img0 = np.random.default_rng(11).integers(0, 256, (90, 150, 3), dtype=np.uint8)
    x, ratio, pad = R.preprocess(img0, 128, stride=32)
    assert x.shape[1] % 32 == 0 and x.shape[2] % 32 == 0
    pred, _ = m(x.unsqueeze(0))
    ref_x = torch.from_numpy(LO.preprocess(img0, 128, stride=32)).cuda().float() / 255.0
    pred2, _ = m(ref_x.unsqueeze(0))
    d = (pred - pred2).abs()
    assert float(d[..., :4].max()) <= 0.05 and float(d[..., 4:].max()) <= 2e-3


@pytest.mark.parametrize('i', range(len(G.CASES)))
def test_scale_coords_golden(golden, i):
    import repyolo_b200 as R
    shape = G.CASES[i][0]
    img1 = golden[f'img_{i}'].shape[1:]
    det = torch.zeros(40, 6)
    det[:, :4] = G.boxes(i)
    det = det.cuda()
    ret = R.scale_coords(img1, det[:, :4], (shape[0], shape[1], 3))          # strided view, in place, like detect.py:114
    assert ret.data_ptr() == det.data_ptr()
    assert det[:, :4].cpu().numpy().tobytes() == golden[f'scaled_{i}'].tobytes()
    assert det[:, :4].round().cpu().numpy().tobytes() == golden[f'scaled_round_{i}'].tobytes()
    rp = golden[f'ratio_pad_{i}']
    b = G.boxes(i).cuda()
    R.scale_coords(img1, b, (shape[0], shape[1], 3), ratio_pad=((rp[0], rp[1]), (rp[2], rp[3])))
    assert b.cpu().numpy().tobytes() == golden[f'scaled_rp_{i}'].tobytes()


def test_scale_detections_uses_device_counts():
    import repyolo_b200 as R
    g = torch.Generator().manual_seed(5)
    out = (torch.rand(3, 300, 6, generator=g) * 600.0).cuda()
    counts = torch.tensor([300, 0, 17], dtype=torch.int32).cuda()
    shapes = [(480, 640, 3), (333, 500, 3), (1080, 1920, 3)]
    ref = out.clone().cpu().numpy()
    for i, n in enumerate([300, 0, 17]):
        if n:
            ref[i, :n, :4] = np.rint(LO.scale_coords((384, 640), ref[i, :n, :4].copy(), shapes[i]))
    got = R.scale_detections(out, counts, (384, 640), shapes)
    assert got.cpu().numpy().tobytes() == ref.tobytes()                      # rows past the count are untouched


def test_no_cpu_fallback():
    import repyolo_b200 as R
    with pytest.raises(R.NativeError):
        R.scale_coords((64, 64), torch.zeros(4, 4), (50, 60, 3))
