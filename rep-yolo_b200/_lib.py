"""ctypes binding of include/repyolo_b200.h.  There is no CPU or PyTorch fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, 'csrc', 'librepyolo_b200.so')

RY_BF16, RY_F32, RY_U8 = 0, 1, 2
T_MAP, T_VEC, T_EXTERNAL = 0, 1, 2
X_IMAGE, X_PRED, X_RAW0 = 0, 1, 2
(OP_STEM, OP_CONV, OP_DW5, OP_MAXPOOL2, OP_SPP, OP_UPSAMPLE2, OP_CA, OP_ATTN_QK, OP_CRISSCROSS, OP_VERTICAL,
 OP_DETECT, OP_CONV_CHAIN) = range(1, 13)
ACT_NONE, ACT_SILU = 0, 1


class TensorDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('dtype', C.c_int32), ('channels', C.c_int32), ('level', C.c_int32),
                ('slot', C.c_int32), ('pad_', C.c_int32)]


class View(C.Structure):
    _fields_ = [('tensor', C.c_int32), ('c_off', C.c_int32), ('c_len', C.c_int32)]


class OpDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('layer', C.c_int32),
                ('in0', View), ('in1', View), ('in2', View), ('out0', View), ('out1', View), ('out2', View),
                ('ksize', C.c_int32), ('stride', C.c_int32), ('act', C.c_int32), ('cin', C.c_int32), ('cout', C.c_int32),
                ('level_idx', C.c_int32), ('n_src', C.c_int32), ('pad_', C.c_int32), ('n_post', C.c_int32), ('post_cout', C.c_int32 * 2),
                ('post_act', C.c_int32 * 2), ('pad2_', C.c_int32), ('w_off', C.c_int64), ('b_off', C.c_int64), ('aux_off', C.c_int64 * 6),
                ('fparam', C.c_float * 8)]


EXPORTS = ['ry_abi_version', 'ry_abi_sizeof', 'ry_last_error', 'ry_plan_create', 'ry_plan_destroy', 'ry_plan_workspace_bytes',
           'ry_plan_bind', 'ry_plan_tensor_info', 'ry_plan_num_candidates', 'ry_plan_launch_count', 'ry_forward',
           'ry_run_ops', 'ry_plan_set_image_dtype', 'ry_plan_set_profiling', 'ry_plan_op_times', 'ry_nms_workspace_bytes', 'ry_nms_launch_count', 'ry_nms', 'ry_decode_filter', 'ry_nms_filtered', 'ry_letterbox_u8', 'ry_scale_coords', 'ry_nchw_to_nhwc_bf16']

_lib = None
ABI_VERSION = 5


class NativeError(RuntimeError):
    pass


def lib():
    """Loads csrc/librepyolo_b200.so (built by __graft_entry__.build() / rep-yolo_b200/_build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NativeError(f'{SO_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(there is no CPU / PyTorch fallback for this path)')
    L = C.CDLL(SO_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    L.ry_abi_version.restype = i32
    L.ry_last_error.restype = C.c_char_p
    L.ry_plan_create.argtypes = [C.POINTER(TensorDesc), i32, C.POINTER(OpDesc), i32, vp, sz, i32, i32, C.POINTER(vp)]
    L.ry_plan_destroy.argtypes = [vp]
    L.ry_plan_destroy.restype = None
    L.ry_plan_workspace_bytes.argtypes = [vp, i32, i32, i32, C.POINTER(sz)]
    L.ry_plan_bind.argtypes = [vp, i32, i32, i32, vp, sz]
    L.ry_plan_tensor_info.argtypes = [vp, i32, C.POINTER(sz), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.ry_plan_num_candidates.argtypes = [vp, C.POINTER(i32)]
    L.ry_plan_launch_count.argtypes = [vp, C.POINTER(i32)]
    L.ry_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.ry_run_ops.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ry_plan_set_image_dtype.argtypes = [vp, i32]
    L.ry_plan_set_profiling.argtypes = [vp, i32]
    L.ry_plan_op_times.argtypes = [vp, C.POINTER(C.c_float), i32]
    L.ry_nms_workspace_bytes.argtypes = [i32, i32, i32, i32, C.POINTER(sz)]
    L.ry_nms_launch_count.argtypes = [i32, i32, i32, i32, C.POINTER(i32)]
    L.ry_nms.argtypes = [vp, i32, i32, i32, C.c_float, C.c_double, vp, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    L.ry_decode_filter.argtypes = [vp, vp, C.c_float, vp, vp, vp, vp, vp, vp]
    L.ry_nms_filtered.argtypes = [vp, vp, i32, i32, i32, C.c_float, C.c_double, vp, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    L.ry_letterbox_u8.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, C.POINTER(i32), i32, vp]
    L.ry_scale_coords.argtypes = [vp, vp, i32, i32, C.c_float, C.c_float, C.c_float, i32, i32, i32, vp]
    L.ry_nchw_to_nhwc_bf16.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, vp]
    for name in EXPORTS:
        if name not in ('ry_last_error', 'ry_plan_destroy', 'ry_abi_version', 'ry_abi_sizeof'):
            getattr(L, name).restype = i32
    if L.ry_abi_version() != ABI_VERSION:
        raise NativeError('ABI version mismatch between _lib.py and librepyolo_b200.so')
    L.ry_abi_sizeof.argtypes = [i32]
    L.ry_abi_sizeof.restype = i32
    if L.ry_abi_sizeof(0) != C.sizeof(TensorDesc) or L.ry_abi_sizeof(1) != C.sizeof(OpDesc):
        raise NativeError('struct layout mismatch between _lib.py and include/repyolo_b200.h')
    _lib = L
    return L


def check(rc: int, what: str = ''):
    if rc != 0:
        raise NativeError(f'{what}: {lib().ry_last_error().decode()}')
