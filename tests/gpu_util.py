"""Helpers shared by the -m gpu tests (CUDA path through the C ABI vs the CPU oracle)."""
import importlib

import torch

import repyolo_b200 as R

N = importlib.import_module('rep-yolo_b200._lib')
planner = importlib.import_module('rep-yolo_b200.planner')


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-12))


def nchw_to_arena(eng, view, x):
    """write an fp32 NCHW tensor (or [B,C,1,1] vector) into a plan view as bf16 NHWC / fp32 [B,C]"""
    t, off, n = view
    dst = eng.tensor(t)
    if dst.dim() == 2:
        dst[:, off:off + n] = x.reshape(x.shape[0], -1).to(dst.device, dst.dtype)
    else:
        dst[..., off:off + n] = x.permute(0, 2, 3, 1).to(dst.device, dst.dtype)


def arena_to_nchw(eng, view):
    t, off, n = view
    src = eng.tensor(t)
    if src.dim() == 2:
        return src[:, off:off + n].float().cpu().reshape(src.shape[0], n, 1, 1)
    return src[..., off:off + n].float().permute(0, 3, 1, 2).contiguous().cpu()


def single_conv_engine(w, b, B, H, W, stride=1, act=True, c_pad_in=0, c_pad_out=0, with_res=False, with_bvec=False,
                       split=False):
    """One CONV op reading channels [c_pad_in, c_pad_in+cin) of a wider tensor and writing at channel offset c_pad_out."""
    cout, cin, k, _ = w.shape
    P = planner.Plan()
    lvl_out = 1 if stride == 2 else 0
    tin = P.tensor(cin + 2 * c_pad_in, 0)
    extra = 16 if split else 0
    tout = P.tensor(cout + 2 * c_pad_out + extra, lvl_out)
    res = P.full(P.tensor(cout, lvl_out)) if with_res else None
    bvec = P.full(P.tensor(cout, 0, N.RY_F32, N.T_VEC)) if with_bvec else None
    if split:
        h = cout // 2
        dst, dst2 = (tout, c_pad_out, h), (tout, c_pad_out + h + extra, h)
    else:
        dst, dst2 = (tout, c_pad_out, cout), None
    P.conv(0, w, b, (tin, c_pad_in, cin), dst, stride, act, dst2=dst2, res=res, bvec=bvec)
    eng = R.NativeEngine(P, 1, 'cuda:0')
    eng.bind(B, H, W)
    return eng, (tin, c_pad_in, cin), dst, dst2, res, bvec
