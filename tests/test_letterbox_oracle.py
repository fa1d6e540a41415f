"""-m "not gpu": the pre-/post-processing oracle (oracle/letterbox_oracle.py) against fixtures minted from the real reference
(tests/golden/make_golden_letterbox.py: utils.datasets.letterbox incl. cv2.resize, utils.general.scale_coords)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import letterbox_oracle as LO

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location('make_golden_letterbox', os.path.join(HERE, 'golden', 'make_golden_letterbox.py'))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(HERE, 'golden', 'letterbox_cases.npz'))


@pytest.mark.parametrize('i', range(len(G.CASES)))
def test_letterbox_matches_reference(golden, i):
    shape, new_shape, auto, fill, up = G.CASES[i]
    img0 = G.image(i, shape)
    img, ratio, pad = LO.letterbox(img0, new_shape, auto=auto, scaleFill=fill, scaleup=up, stride=32)
    chw = np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))
    ref = golden[f'img_{i}']
    assert chw.shape == ref.shape and chw.tobytes() == ref.tobytes()          # bit-exact incl. cv2's fixed-point bilinear
    np.testing.assert_array_equal(np.array([ratio[0], ratio[1], pad[0], pad[1]], np.float64), golden[f'ratio_pad_{i}'])


@pytest.mark.parametrize('i', range(len(G.CASES)))
def test_scale_coords_matches_reference(golden, i):
    shape = G.CASES[i][0]
    ref_img = golden[f'img_{i}']
    b = G.boxes(i).numpy().copy()
    got = LO.scale_coords(ref_img.shape[1:], b, (shape[0], shape[1], 3))
    assert got.tobytes() == golden[f'scaled_{i}'].tobytes()
    assert np.rint(got).tobytes() == golden[f'scaled_round_{i}'].tobytes()
    rp = golden[f'ratio_pad_{i}']
    got2 = LO.scale_coords(ref_img.shape[1:], G.boxes(i).numpy().copy(), (shape[0], shape[1], 3), ratio_pad=((rp[0], rp[1]), (rp[2], rp[3])))
    assert got2.tobytes() == golden[f'scaled_rp_{i}'].tobytes()


def test_resize_matches_cv2_when_available():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(7)
    for _ in range(60):
        H, W = (int(v) for v in rng.integers(2, 200, 2))
        nh, nw = (int(v) for v in rng.integers(1, 260, 2))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR), LO.resize_linear_u8(img, nw, nh)), (H, W, nh, nw)
