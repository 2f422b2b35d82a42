"""ctypes binding of libsimvg_b200.so (the C ABI declared in include/simvg_b200.h).

There is deliberately no fallback: if the library is missing or the device is not sm_100, the
product path raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIMVGB_LIB") or os.path.join(_HERE, "libsimvg_b200.so")   # SIMVGB_LIB: A/B builds when tuning
_lib = None

c_int, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

EPI_BF16, EPI_GELU, EPI_RESID, EPI_F32, EPI_ATOMIC = range(5)


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("a_mn_major", c_int), ("b_mn_major", c_int),
        ("lda", c_i64), ("ldb", c_i64),
        ("A", c_vp), ("B", c_vp),
        ("epilogue", c_int), ("k_splits", c_int),
        ("bias", c_vp), ("out_bf16", c_vp), ("out2_bf16", c_vp), ("out_f32", c_vp), ("res_f32", c_vp),
        ("ldo", c_i64),
        ("scale", c_f32), ("scale_cols", c_int),
        ("row_scale", c_vp), ("rows_per_scale", c_int),
        ("accumulate", c_int),
    ]


def lib():
    """Loads the shared library (building it is `python -m simvg_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "simvg_b200: %s is missing — run `python -m simvg_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.simvgb_last_error.restype = ctypes.c_char_p
        L.simvgb_version.restype = c_int
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("simvg_b200 %s failed (%d): %s" % (what, rc, lib().simvgb_last_error().decode()))


def stream_ptr():
    return c_vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_vp(0) if t is None else c_vp(t.data_ptr())


def require_device(t):
    if not t.is_cuda:
        raise RuntimeError("simvg_b200 kernels run on sm_100a CUDA tensors only (got a %s tensor; no CPU fallback)" % t.device)
