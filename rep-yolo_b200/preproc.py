"""The callers either side of the hot path, with the reference's signatures, executed by the CUDA library:

    letterbox(img, new_shape, color, auto, scaleFill, scaleup, stride)     utils/datasets.py:984-1014
    preprocess(img0, img_size, stride, auto)                               LoadImages.__next__, utils/datasets.py:191-195
    scale_coords(img1_shape, coords, img0_shape, ratio_pad)                utils/general.py:319-340 (clip_coords included)
    scale_detections(out, counts, img1_shape, img0_shapes)                 the per-image loop of detect.py:109-114, no host sync

The shape arithmetic (a dozen scalar operations per image) stays on the host and is the reference's own, line for line in
meaning; every pixel / box is computed on the GPU, bit-exact with cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder and with the
reference's torch fp32 chain.  No CPU fallback: host images are uploaded (uint8, the un-resized original), CPU box tensors are rejected.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as N


def _geometry(shape, new_shape, auto, scaleFill, scaleup, stride):
    # datasets.py:986-1012
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:                                        # only scale down (better test mAP)
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:                                               # minimum rectangle
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    elif scaleFill:                                        # stretch
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = new_shape[1] / shape[1], new_shape[0] / shape[0]
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, ratio, (dw, dh), top, bottom, left, right


def _as_device_image(img, device):
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(np.ascontiguousarray(img))
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise ValueError('expected a uint8 HWC image with 3 channels')
    if not img.is_cuda:
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise N.NativeError('letterbox: no CUDA device (there is no CPU fallback on this path)')
        img = img.to(device, non_blocking=True)
    if img.stride(2) != 1 or img.stride(1) != 3:
        img = img.contiguous()
    return img


def _run(img, new_shape, color, auto, scaleFill, scaleup, stride, planar_rgb, out, device):
    img = _as_device_image(img, device)
    H0, W0 = int(img.shape[0]), int(img.shape[1])
    new_unpad, ratio, pad, top, bottom, left, right = _geometry((H0, W0), new_shape, auto, scaleFill, scaleup, stride)
    H1, W1 = new_unpad[1] + top + bottom, new_unpad[0] + left + right
    shape = (3, H1, W1) if planar_rgb else (H1, W1, 3)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=img.device)
    elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != img.device:
        raise ValueError(f'out must be a contiguous uint8 tensor of shape {shape} on {img.device}')
    col = (C.c_int * 3)(*[int(c) for c in (color if not np.isscalar(color) else (color,) * 3)])
    with torch.cuda.device(img.device):
        st = torch.cuda.current_stream(img.device).cuda_stream
        N.check(N.lib().ry_letterbox_u8(img.data_ptr(), H0, W0, int(img.stride(0)), out.data_ptr(), H1, W1, new_unpad[0], new_unpad[1],
                                        left, top, col, int(planar_rgb), C.c_void_p(st)), 'ry_letterbox_u8')
    return out, ratio, pad


def letterbox(img, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleFill=False, scaleup=True, stride=32, out=None,
              device=None):
    """Resize and pad an HWC uint8 image to a stride multiple (reference signature; returns (HWC uint8 CUDA tensor, ratio, (dw, dh)))."""
    return _run(img, new_shape, color, auto, scaleFill, scaleup, stride, False, out, device)


def preprocess(img0, img_size=640, stride=32, auto=True, out=None, device=None):
    """BGR HWC uint8 image (numpy or tensor) -> letterboxed RGB CHW uint8 CUDA tensor, ready for ``Model.forward`` (which takes
    uint8 and fuses the /255 of detect.py:76).  Returns (img [3, H1, W1], ratio, (dw, dh))."""
    return _run(img0, img_size, (114, 114, 114), auto, False, True, stride, True, out, device)


def _gain_pad(img1_shape, img0_shape, ratio_pad):
    # general.py:321-326
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    return float(gain), (float(pad[0]), float(pad[1]))


def _scale(coords, count, n_max, img1_shape, img0_shape, ratio_pad, round_result):
    gain, pad = _gain_pad(img1_shape, img0_shape, ratio_pad)
    with torch.cuda.device(coords.device):
        st = torch.cuda.current_stream(coords.device).cuda_stream
        N.check(N.lib().ry_scale_coords(coords.data_ptr(), count, n_max, int(coords.stride(0)), C.c_float(float(np.float32(pad[0]))),
                                        C.c_float(float(np.float32(pad[1]))), C.c_float(float(np.float32(gain))), int(img0_shape[1]),
                                        int(img0_shape[0]), int(round_result), C.c_void_p(st)), 'ry_scale_coords')


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    """Rescale xyxy boxes from img1_shape to img0_shape in place and clip them (reference signature; returns ``coords``).
    ``coords``: fp32 CUDA tensor [n, >= 4] whose columns are contiguous (``det[:, :4]`` of an NMS result is fine)."""
    if not coords.is_cuda:
        raise N.NativeError('scale_coords: coords must be a CUDA tensor (no CPU fallback on this path)')
    if coords.dtype != torch.float32 or coords.dim() != 2 or coords.shape[1] < 4 or (coords.shape[0] > 1 and coords.stride(1) != 1):
        raise ValueError('scale_coords: expected an fp32 [n, >= 4] tensor with unit column stride')
    if coords.shape[0]:
        _scale(coords, None, int(coords.shape[0]), img1_shape, img0_shape, ratio_pad, False)
    return coords


def scale_detections(out, counts, img1_shape, img0_shapes, ratio_pads=None, round_result=True):
    """detect.py:109-114 for a padded NMS result (``nms_padded``): boxes of image i are rescaled to img0_shapes[i], clipped and
    rounded in place, using the device-side counts (no host sync).  Returns ``out``."""
    if not out.is_cuda or out.dtype != torch.float32 or out.dim() != 3 or not out.is_contiguous():
        raise N.NativeError('scale_detections: expected the contiguous fp32 CUDA tensor [B, max_det, 6] of nms_padded')
    for i in range(out.shape[0]):
        rp = None if ratio_pads is None else ratio_pads[i]
        _scale(out[i], counts.data_ptr() + 4 * i, int(out.shape[1]), img1_shape, img0_shapes[i], rp, round_result)
    return out
