"""TEST INFRASTRUCTURE (CPU oracle) -- numpy restatement of the reference's image pre-/post-processing either side of the hot
path.  Imported only by tests/, never by the product.

  letterbox      utils/datasets.py:984-1014   (resize + pad to a stride multiple, value 114)
  preprocess     utils/datasets.py:191-195    (letterbox -> BGR to RGB -> HWC to CHW -> contiguous uint8)
  scale_coords   utils/general.py:319-332     (+ clip_coords 335-340), fp32 like the reference's torch ops on CPU

cv2.resize(INTER_LINEAR) on uint8 is third-party arithmetic (opencv-python, un-vendored; installed pin 4.x): restated from the
published OpenCV algorithm (modules/imgproc/src/resize.cpp: 11-bit fixed-point coefficients, HResizeLinear / VResizeLinear
<uchar, int, short>):
    fx = float((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx; border columns clamp sx and zero fx
    alpha = (round_half_even((1 - fx) * 2048), round_half_even(fx * 2048)) as int16; rows: the same with CLAMPED row indices
    H[x] = S[sx] * a0 + S[sx + 1] * a1                                  (int32)
    dst  = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
Pinned: tests/golden/letterbox_*.npz hold outputs of the REAL reference (cv2 included) minted by tests/golden/make_golden_letterbox.py.
"""
import math

import numpy as np


def _coeffs(dst_n, src_n):
    """per destination index: source index (clamped) and the int16 coefficient pair of OpenCV's linear resize"""
    scale = 1.0 / (float(dst_n) / float(src_n))                           # double, as resize.cpp computes it
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _round_short(x):
    return np.rint(x.astype(np.float32)).astype(np.int32)                  # cvRound: half to even; values are within int16


def resize_linear_u8(img, new_w, new_h):
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC"""
    H, W = img.shape[:2]
    sx, fx = _coeffs(new_w, W)
    lo = sx < 0
    fx[lo] = 0.0; sx[lo] = 0
    hi = sx >= W - 1
    fx[hi] = 0.0; sx[hi] = W - 1
    a0 = _round_short((np.float32(1.0) - fx) * np.float32(2048.0))
    a1 = _round_short(fx * np.float32(2048.0))
    sx1 = np.minimum(sx + 1, W - 1)                                         # a1 == 0 wherever sx + 1 would be outside
    sy, fy = _coeffs(new_h, H)
    b0 = _round_short((np.float32(1.0) - fy) * np.float32(2048.0))
    b1 = _round_short(fy * np.float32(2048.0))
    y0 = np.clip(sy, 0, H - 1)
    y1 = np.clip(sy + 1, 0, H - 1)
    src = img.astype(np.int32)
    hrow = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]     # [H, new_w, C] int32
    h0, h1 = hrow[y0], hrow[y1]
    out = (((b0[:, None, None] * (h0 >> 4)) >> 16) + ((b1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_geometry(shape, new_shape=(640, 640), auto=True, scaleFill=False, scaleup=True, stride=32):
    """the shape arithmetic of letterbox (datasets.py:986-1012): (new_unpad (w, h), ratio, (dw, dh), top, bottom, left, right)"""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    elif scaleFill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = new_shape[1] / shape[1], new_shape[0] / shape[0]
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, ratio, (dw, dh), top, bottom, left, right


def letterbox(img, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32):
    shape = img.shape[:2]
    new_unpad, ratio, (dw, dh), top, bottom, left, right = letterbox_geometry(shape, new_shape, auto, scaleFill, scaleup, stride)
    if shape[::-1] != new_unpad:
        img = resize_linear_u8(img, new_unpad[0], new_unpad[1])
    out = np.empty((img.shape[0] + top + bottom, img.shape[1] + left + right, img.shape[2]), np.uint8)
    out[...] = np.asarray(color, np.uint8)
    out[top:top + img.shape[0], left:left + img.shape[1]] = img
    return out, ratio, (dw, dh)


def preprocess(img0, img_size=640, stride=32, auto=True):
    """LoadImages.__next__ (datasets.py:191-195): BGR HWC uint8 -> letterboxed RGB CHW uint8"""
    img = letterbox(img0, img_size, stride=stride, auto=auto)[0]
    return np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    """general.py:319-332 on an fp32 [n, >=4] array (in place, returns it); clip_coords included"""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    c = coords
    c[:, [0, 2]] -= np.float32(pad[0])
    c[:, [1, 3]] -= np.float32(pad[1])
    c[:, :4] /= np.float32(gain)
    c[:, 0] = np.clip(c[:, 0], 0, img0_shape[1]); c[:, 1] = np.clip(c[:, 1], 0, img0_shape[0])
    c[:, 2] = np.clip(c[:, 2], 0, img0_shape[1]); c[:, 3] = np.clip(c[:, 3], 0, img0_shape[0])
    return c
