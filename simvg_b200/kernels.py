"""Tensor-level wrappers over the C ABI of libsimvg_b200.so (include/simvg_b200.h).

PyTorch only owns memory and streams here; every function below launches hand-written sm_100a kernels on the
current stream and raises if the library or a CUDA device is missing (no CPU / eager fallback).
"""
import ctypes
import os

import torch

from . import _lib as L
from ._lib import EPI_ATOMIC, EPI_BF16, EPI_F32, EPI_GELU, EPI_RESID  # noqa: F401

bf16 = torch.bfloat16
f32 = torch.float32

# Launch counter: number of libsimvg_b200 kernel launches since the last reset (bench.py reports it).
_launches = [0]


# Optional per-family device timing (bench.py's roofline leg): family -> list of (start event, end event, flops).
_prof = [None]
_prof_detail = [False]   # per-shape GEMM families (tools/gemm_shapes.py)


def profile_start():
    _prof[0] = {}


def profile_stop():
    """-> {family: (launches, total_ms, total_flops)}; synchronises."""
    data, _prof[0] = _prof[0], None
    torch.cuda.synchronize()
    out = {}
    for fam, evs in (data or {}).items():
        out[fam] = (len(evs), sum(a.elapsed_time(b) for a, b, _ in evs), sum(f for _, _, f in evs))
    return out


# NVTX ranges per fused op / per layer (SIMVGB_NVTX=1): visible in nsys / ncu --nvtx timelines; off by default (two host calls
# per range).  The ranges follow the reference's module names so a timeline reads like the reference's forward.
_nvtx_on = [os.environ.get("SIMVGB_NVTX", "0") not in ("", "0")]


class nvtx:
    """with K.nvtx("encoder.layer3.fwd"): ...   — no-op unless SIMVGB_NVTX=1 (or kernels.enable_nvtx())."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _nvtx_on[0]:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _nvtx_on[0]:
            torch.cuda.nvtx.range_pop()
        return False


def enable_nvtx(on=True):
    _nvtx_on[0] = bool(on)


def nvtx_push(name):
    if _nvtx_on[0]:
        torch.cuda.nvtx.range_push(name)


def nvtx_pop():
    if _nvtx_on[0]:
        torch.cuda.nvtx.range_pop()


class _timed:
    def __init__(self, family, flops):
        self.family, self.flops = family, flops

    def __enter__(self):
        if _nvtx_on[0]:
            torch.cuda.nvtx.range_push(self.family)
        if _prof[0] is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _prof[0] is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof[0].setdefault(self.family, []).append((self.e0, e1, self.flops))
        if _nvtx_on[0]:
            torch.cuda.nvtx.range_pop()
        return False


def launch_count():
    return _launches[0]


def reset_launch_count():
    _launches[0] = 0


class AttnArgs(ctypes.Structure):
    _fields_ = [
        ("B", L.c_int), ("H", L.c_int), ("Lv", L.c_int), ("Lt", L.c_int), ("head_dim", L.c_int),
        ("qkv_v", L.c_vp), ("qkv_t", L.c_vp), ("text_pad", L.c_vp),
        ("out_v", L.c_vp), ("out_t", L.c_vp), ("lse", L.c_vp),
        ("dout_v", L.c_vp), ("dout_t", L.c_vp), ("dqkv_v", L.c_vp), ("dqkv_t", L.c_vp),
        ("delta", L.c_vp), ("dq_acc_v", L.c_vp), ("dq_acc_t", L.c_vp), ("q_scale", L.c_f32), ("delta_ready", L.c_int),
    ]


class LnBwdArgs(ctypes.Structure):
    _fields_ = [
        ("mode", L.c_int), ("C", L.c_int), ("rows", L.c_i64),
        ("x", L.c_vp), ("dy", L.c_vp), ("dy_is_f32", L.c_int),
        ("gamma", L.c_vp), ("mean", L.c_vp), ("rstd", L.c_vp),
        ("dgamma", L.c_vp), ("dbeta", L.c_vp),
        ("dres_in", L.c_vp), ("dres_out", L.c_vp), ("dyb", L.c_vp),
        ("row_scale", L.c_vp), ("rows_per_scale", L.c_int),
        ("dbias_prev", L.c_vp), ("dx", L.c_vp), ("u", L.c_vp),
        ("delta", L.c_vp), ("delta_L", L.c_int), ("delta_H", L.c_int), ("delta_stride", L.c_int), ("delta_vbase", L.c_int),
    ]


def _p(t):
    return None if t is None else t.data_ptr()


def _lib_setup():
    lib = L.lib()
    if not getattr(lib, "_simvgb_typed", False):
        lib.simvgb_attn_lse_stride.restype = L.c_int
        lib.simvgb_attn_text_offset.restype = L.c_int
        lib.simvgb_ln_fwd.argtypes = [L.c_vp, L.c_int, L.c_vp, L.c_int, L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_i64, L.c_int,
                                      L.c_f32, L.c_int, L.c_vp]
        lib.simvgb_colsum.argtypes = [L.c_vp, L.c_int, L.c_vp, L.c_vp, L.c_vp, L.c_int, L.c_i64, L.c_int, L.c_i64, L.c_vp]
        lib.simvgb_cast_bf16.argtypes = [L.c_vp, L.c_vp, L.c_i64, L.c_vp]
        lib.simvgb_im2col_patch.argtypes = [L.c_vp, L.c_vp, L.c_int, L.c_int, L.c_int, L.c_vp]
        lib.simvgb_embed_vision.argtypes = [L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_int, L.c_int, L.c_int, L.c_vp]
        lib.simvgb_embed_text.argtypes = [L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_int, L.c_int, L.c_int, L.c_vp]
        lib.simvgb_sumsq.argtypes = [L.c_vp, L.c_i64, L.c_vp, L.c_vp]
        lib.simvgb_adam_amsgrad.argtypes = [L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_i64, L.c_f32, L.c_f32, L.c_f32,
                                            L.c_f32, L.c_f32, L.c_int, L.c_vp, L.c_f32, L.c_vp, L.c_f32, L.c_vp]
        lib.simvgb_adam_amsgrad_dev.argtypes = [L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_vp, L.c_i64, L.c_vp, L.c_f32, L.c_f32,
                                                L.c_f32, L.c_f32, L.c_vp, L.c_f32, L.c_vp, L.c_vp]
        lib.simvgb_head_xattn_ws_floats.argtypes = [L.c_int, L.c_int, L.c_int, L.c_int]
        lib.simvgb_head_xattn_ws_floats.restype = ctypes.c_longlong
        lib._simvgb_typed = True
    return lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------- GEMM
def _gemm_args(A, B, M, N, K, *, a_mn=False, b_mn=False, epilogue=EPI_BF16, bias=None, out=None, out2=None, res=None,
               scale=1.0, scale_cols=0, row_scale=None, rows_per_scale=1, k_splits=1, accumulate=False, ldo=None):
    L.require_device(A)
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
    a.lda, a.ldb = A.stride(0), B.stride(0)
    a.A, a.B = A.data_ptr(), B.data_ptr()
    a.epilogue, a.k_splits = epilogue, k_splits
    a.bias = _p(bias)
    if out is None:
        if epilogue in (EPI_BF16, EPI_GELU):
            out = torch.empty(M, N, device=A.device, dtype=bf16)
        elif epilogue == EPI_ATOMIC:
            out = torch.zeros(M, N, device=A.device, dtype=f32)
        else:
            out = torch.empty(M, N, device=A.device, dtype=f32)
    if epilogue in (EPI_BF16, EPI_GELU):
        a.out_bf16 = out.data_ptr()
        if epilogue == EPI_GELU:
            a.out2_bf16 = out2.data_ptr()
    else:
        a.out_f32 = out.data_ptr()
    a.res_f32 = _p(res)
    a.ldo = ldo if ldo is not None else out.stride(0)
    a.scale, a.scale_cols = scale, scale_cols
    a.row_scale = _p(row_scale)
    a.rows_per_scale = rows_per_scale
    a.accumulate = int(accumulate)
    fam = "gemm" if not _prof_detail[0] else "gemm M=%d N=%d K=%d a_mn=%d b_mn=%d epi=%d ks=%d" % (M, N, K, a_mn, b_mn, epilogue, k_splits)
    return a, out, fam, 2.0 * M * N * K


def gemm(A, B, M, N, K, **kw):
    """C[M,N] = epilogue(A(M,K) @ B(N,K)^T).  A: [M,K] (or [K,M] if a_mn), B: [N,K] (or [K,N] if b_mn), bf16.

    Returns `out` (allocated when None: bf16 for the BF16/GELU epilogues, fp32 otherwise)."""
    lib = _lib_setup()
    a, out, fam, flops = _gemm_args(A, B, M, N, K, **kw)
    with _timed(fam, flops):
        L.check(lib.simvgb_gemm(ctypes.byref(a), L.c_vp(_stream())), "gemm")
    _launches[0] += 1
    return out


def gemm_pair(first, second):
    """Two independent GEMMs in one persistent launch (simvgb_gemm_pair): `first` / `second` are (A, B, M, N, K, kwargs)
    tuples with the arguments of gemm().  Used for the vision-expert / text-expert problems of a multiway layer."""
    lib = _lib_setup()
    a0, out0, fam, fl0 = _gemm_args(*first[:5], **first[5])
    a1, out1, _, fl1 = _gemm_args(*second[:5], **second[5])
    with _timed(fam, fl0 + fl1):
        L.check(lib.simvgb_gemm_pair(ctypes.byref(a0), ctypes.byref(a1), L.c_vp(_stream())), "gemm_pair")
    _launches[0] += 1
    return out0, out1


def wgrad_splits(M, N, K, clusters=74, epi_kb=8):
    """Split-K factor for a weight-gradient GEMM (few output tiles, very long K).

    2-CTA kernel (N > 128, M >= 256; 256 x 256 tiles on `clusters` = SMs / 2 clusters): minimise
    waves x (k-blocks per split + epilogue), waves = ceil(tiles * splits / clusters) — the persistent grid runs whole waves, so
    e.g. 36 tiles x 5 splits = 180 items cost 3 waves where 36 x 2 = 72 items cost one.  Every split keeps >= 32 k-blocks so
    its fp32 atomic epilogue stays amortised.  Measured on B200 (tools/wgrad_ks_ab.py, K = 102464): 768x3072 ks 5 -> 2:
    0.437 -> 0.358 ms; 3072x768 0.414 -> 0.363; 2304x768 ks 6 -> 8: 0.336 -> 0.295; 768x768 ks 17 -> 8: 0.119 -> 0.093.
    1-CTA kernel (narrow N / short M): the earlier rule (about two items per SM)."""
    kb = (K + 63) // 64
    if not (N > 128 and M >= 256):
        tiles = ((M + 127) // 128) * ((N + 255) // 256 if N > 128 else 1)
        return max(1, min((2 * 148 + tiles - 1) // tiles, kb // 32))
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    best_cost, best_ks = None, 1
    for ks in range(1, max(1, kb // 32) + 1):
        per = (kb + ks - 1) // ks
        ks_eff = (kb + per - 1) // per                      # no empty splits (the library applies the same rounding)
        waves = (tiles * ks_eff + clusters - 1) // clusters
        cost = waves * (per + epi_kb) + (0.5 * epi_kb if ks_eff > 1 else 0.0)   # atomics cost a little more than plain accumulate
        if best_cost is None or cost < best_cost - 1e-9:
            best_cost, best_ks = cost, ks_eff
    return best_ks


def _wgrad_spec(dY, X, Dout, Din, R, out):
    ks = wgrad_splits(Dout, Din, R)
    if ks == 1:
        return (dY, X, Dout, Din, R, dict(a_mn=True, b_mn=True, epilogue=EPI_F32, out=out, accumulate=True))
    return (dY, X, Dout, Din, R, dict(a_mn=True, b_mn=True, epilogue=EPI_ATOMIC, out=out, k_splits=ks))


def wgrad_pair(first, second):
    """Two weight-gradient GEMMs (dY, X, Dout, Din, R, out) in one launch."""
    return gemm_pair(_wgrad_spec(*first), _wgrad_spec(*second))


def wgrad(dY, X, Dout, Din, R, out=None):
    """dW[Dout, Din] (+)= dY[R, Dout]^T @ X[R, Din]; both operands read MN-major, fp32 atomics over split-K."""
    if out is None:
        out = torch.zeros(Dout, Din, device=dY.device, dtype=f32)
    ks = wgrad_splits(Dout, Din, R)
    if ks == 1:   # no split: plain read-modify-write accumulate, no atomics
        return gemm(dY, X, Dout, Din, R, a_mn=True, b_mn=True, epilogue=EPI_F32, out=out, accumulate=True)
    return gemm(dY, X, Dout, Din, R, a_mn=True, b_mn=True, epilogue=EPI_ATOMIC, out=out, k_splits=ks)


# ---------------------------------------------------------------------------------------------- attention
def attn_lse_stride(Lv, Lt):
    return _lib_setup().simvgb_attn_lse_stride(Lv, Lt)


def _attn_args(B, H, Lv, Lt, qkv_v, qkv_t, pad, out_v, out_t, lse):
    a = AttnArgs()
    a.B, a.H, a.Lv, a.Lt, a.head_dim = B, H, Lv, Lt, 64
    a.qkv_v, a.qkv_t, a.text_pad = _p(qkv_v), _p(qkv_t), _p(pad)
    a.out_v, a.out_t, a.lse = _p(out_v), _p(out_t), _p(lse)
    a.q_scale = 0.125
    return a


def attn_fwd(qkv_v, qkv_t, pad, B, H, Lv, Lt):
    L.require_device(qkv_v)
    lib = _lib_setup()
    D = H * 64
    dev = qkv_v.device
    out_v = torch.empty(B * Lv, D, device=dev, dtype=bf16)
    out_t = torch.empty(B * Lt, D, device=dev, dtype=bf16)
    lse = torch.empty(B, H, attn_lse_stride(Lv, Lt), device=dev, dtype=f32)
    a = _attn_args(B, H, Lv, Lt, qkv_v, qkv_t, pad, out_v, out_t, lse)
    with _timed("attn_fwd", 4.0 * B * H * (Lv + Lt) ** 2 * 64):
        L.check(lib.simvgb_attn_fwd(ctypes.byref(a), L.c_vp(_stream())), "attn_fwd")
    _launches[0] += 1
    return out_v, out_t, lse


def attn_workspace(ws, B, H, Lv, Lt, dev):
    """Backward workspaces (delta, fp32 dQ accumulators), cached per geometry in the dict `ws`.  delta is allocated ZEROED: its
    non-token slots of the virtual sequence axis must read 0 and are never written."""
    key = (B, H, Lv, Lt)
    if ws.get("key") != key:
        D = H * 64
        ws["key"] = key
        ws["delta"] = torch.zeros(B, H, attn_lse_stride(Lv, Lt), device=dev, dtype=f32)
        ws["dq_v"] = torch.empty(B * Lv, D, device=dev, dtype=f32)
        ws["dq_t"] = torch.empty(B * Lt, D, device=dev, dtype=f32)
    return ws


def attn_delta_spec(ws, B, H, Lv, Lt, which):
    """`delta=` argument of ln_bwd(mode 1) for the vision (0) / text (1) rows: the inner-attention-LN backward then also emits
    rowsum(O o dO) in the attention backward's layout, and attn_bwd(..., delta_ready=True) skips its own delta pass."""
    stride = attn_lse_stride(Lv, Lt)
    if which == 0:
        return (ws["delta"], Lv, H, stride, 0)
    return (ws["delta"], Lt, H, stride, _lib_setup().simvgb_attn_text_offset(Lv, Lt))


def attn_bwd(qkv_v, qkv_t, pad, out_v, out_t, lse, dout_v, dout_t, B, H, Lv, Lt, ws=None, delta_ready=False):
    lib = _lib_setup()
    D = H * 64
    dev = qkv_v.device
    dqkv_v = torch.empty(B * Lv, 3 * D, device=dev, dtype=bf16)
    dqkv_t = torch.empty(B * Lt, 3 * D, device=dev, dtype=bf16)
    if ws is None:
        ws = {}
    attn_workspace(ws, B, H, Lv, Lt, dev)
    a = _attn_args(B, H, Lv, Lt, qkv_v, qkv_t, pad, out_v, out_t, lse)
    a.delta_ready = int(delta_ready)
    a.dout_v, a.dout_t = _p(dout_v), _p(dout_t)
    a.dqkv_v, a.dqkv_t = _p(dqkv_v), _p(dqkv_t)
    a.delta, a.dq_acc_v, a.dq_acc_t = _p(ws["delta"]), _p(ws["dq_v"]), _p(ws["dq_t"])
    with _timed("attn_bwd", 10.0 * B * H * (Lv + Lt) ** 2 * 64):
        L.check(lib.simvgb_attn_bwd(ctypes.byref(a), L.c_vp(_stream())), "attn_bwd")
    _launches[0] += 3 if delta_ready else 5
    return dqkv_v, dqkv_t


# ---------------------------------------------------------------------------------------------- row kernels
def ln_fwd(x, gamma, beta, eps, out_dtype=bf16, gelu=False):
    L.require_device(x)
    lib = _lib_setup()
    R, C = x.shape
    y = torch.empty(R, C, device=x.device, dtype=out_dtype)
    mean = torch.empty(R, device=x.device, dtype=f32)
    rstd = torch.empty(R, device=x.device, dtype=f32)
    L.check(lib.simvgb_ln_fwd(x.data_ptr(), int(x.dtype == bf16), y.data_ptr(), int(out_dtype == bf16), gamma.data_ptr(),
                              beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), R, C, eps, int(gelu), _stream()), "ln_fwd")
    _launches[0] += 1
    return y, mean, rstd


def ln_bwd(mode, x, dy, gamma, mean, rstd, dgamma, dbeta, *, dres_in=None, dres_out=None, dyb=None, row_scale=None,
           rows_per_scale=1, dbias_prev=None, dx=None, u=None, delta=None):
    lib = _lib_setup()
    R, C = dy.shape
    a = LnBwdArgs()
    a.mode, a.C, a.rows = mode, C, R
    a.x, a.dy, a.dy_is_f32 = _p(x), dy.data_ptr(), int(dy.dtype == f32)
    a.gamma, a.mean, a.rstd = gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.dgamma, a.dbeta = dgamma.data_ptr(), dbeta.data_ptr()
    a.dres_in, a.dres_out, a.dyb = _p(dres_in), _p(dres_out), _p(dyb)
    a.row_scale, a.rows_per_scale = _p(row_scale), rows_per_scale
    a.dbias_prev, a.dx, a.u = _p(dbias_prev), _p(dx), _p(u)
    if delta is not None:
        a.delta, a.delta_L, a.delta_H, a.delta_stride, a.delta_vbase = delta[0].data_ptr(), delta[1], delta[2], delta[3], delta[4]
    L.check(lib.simvgb_ln_bwd(ctypes.byref(a), L.c_vp(_stream())), "ln_bwd")
    _launches[0] += 1


def colsum(x, out=None, out_bf16=None, row_scale=None, rows_per_scale=1, C=None, ld=None):
    lib = _lib_setup()
    R = x.shape[0]
    C = C if C is not None else x.shape[1]
    ld = ld if ld is not None else x.stride(0)
    L.check(lib.simvgb_colsum(x.data_ptr(), int(x.dtype == bf16), _p(out), _p(out_bf16), _p(row_scale), rows_per_scale,
                              R, C, ld, _stream()), "colsum")
    _launches[0] += 1
    return out


def cast_bf16(src, dst):
    """dst (bf16, same numel, contiguous) = src (fp32, contiguous)."""
    lib = _lib_setup()
    L.check(lib.simvgb_cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "cast_bf16")
    _launches[0] += 1
    return dst


def im2col_patch(img, P):
    L.require_device(img)
    lib = _lib_setup()
    B, _, S, _ = img.shape
    cols = torch.empty(B * (S // P) ** 2, 3 * P * P, device=img.device, dtype=bf16)
    L.check(lib.simvgb_im2col_patch(img.data_ptr(), cols.data_ptr(), B, S, P, _stream()), "im2col_patch")
    _launches[0] += 1
    return cols


def im2col_patch_u8(img_u8, P, mean, std, to_rgb=True):
    """uint8 [B,S,S,3] (HWC) image -> normalised bf16 patch matrix [B*N, 3*P*P] (Normalize + transpose + im2col fused)."""
    L.require_device(img_u8)
    lib = _lib_setup()
    B, S, S2, C = img_u8.shape
    assert img_u8.dtype == torch.uint8 and C == 3 and S == S2 and img_u8.is_contiguous()
    cols = torch.empty(B * (S // P) ** 2, 3 * P * P, device=img_u8.device, dtype=bf16)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    sd = (ctypes.c_float * 3)(*[float(v) for v in std])
    L.check(lib.simvgb_im2col_patch_u8(L.c_vp(img_u8.data_ptr()), L.c_vp(cols.data_ptr()), B, S, P, m, sd, int(bool(to_rgb)),
                                       L.c_vp(_stream())), "im2col_patch_u8")
    _launches[0] += 1
    return cols


def embed_vision(patch, cls, posA, B, N, D):
    lib = _lib_setup()
    xv = torch.empty(B * (N + 1), D, device=patch.device, dtype=f32)
    L.check(lib.simvgb_embed_vision(patch.data_ptr(), cls.data_ptr(), posA.data_ptr(), xv.data_ptr(), B, N, D, _stream()),
            "embed_vision")
    _launches[0] += 1
    return xv


def embed_text(table, ids, pad, posB, B, Lt, D):
    lib = _lib_setup()
    xt = torch.empty(B * Lt, D, device=table.device, dtype=f32)
    L.check(lib.simvgb_embed_text(table.data_ptr(), ids.data_ptr(), _p(pad), posB.data_ptr(), xt.data_ptr(), B, Lt, D,
                                  _stream()), "embed_text")
    _launches[0] += 1
    return xt


def hungarian(cost, sizes):
    """Per-sample optimal assignment on the device.  cost: fp32 [B, nq, sum(sizes)] (detrex HungarianMatcher's batched cost
    matrix), sizes: host list of target counts.  -> [(query_idx, target_idx)] int64 device tensors per sample (scipy's
    linear_sum_assignment order), with NO device->host synchronisation: the result lengths min(nq, n_b) are host-known."""
    L.require_device(cost)
    lib = _lib_setup()
    B, nq, ttot = cost.shape
    if max(sizes, default=0) > 32 or nq > 32:
        raise RuntimeError("simvgb_hungarian handles at most 32 queries / 32 targets per sample (got nq=%d, max targets %d)" % (nq, max(sizes)))
    off = [0]
    for n in sizes:
        off.append(off[-1] + int(n))
    offsets = torch.tensor(off, dtype=torch.int32).to(cost.device, non_blocking=True)
    kmax = max(1, min(nq, max(sizes, default=1)))
    out_q = torch.empty(B, kmax, dtype=torch.int64, device=cost.device)
    out_t = torch.empty(B, kmax, dtype=torch.int64, device=cost.device)
    c = cost.contiguous().float()
    L.check(lib.simvgb_hungarian(L.c_vp(c.data_ptr()), B, nq, ttot, L.c_vp(offsets.data_ptr()), L.c_vp(out_q.data_ptr()),
                                 L.c_vp(out_t.data_ptr()), kmax, L.c_vp(_stream())), "hungarian")
    _launches[0] += 1
    return [(out_q[b, :min(nq, int(n))], out_t[b, :min(nq, int(n))]) for b, n in enumerate(sizes)]


def sumsq(g, out):
    lib = _lib_setup()
    L.check(lib.simvgb_sumsq(g.data_ptr(), g.numel(), out.data_ptr(), _stream()), "sumsq")
    _launches[0] += 1


def adam_amsgrad(p, g, m, v, vmax, lr, beta1, beta2, eps, weight_decay, step, grad_sumsq=None, max_norm=0.0, ema=None,
                 ema_decay=0.0):
    """Clip (global norm from `grad_sumsq`) + Adam(amsgrad) on flat fp32 buffers; `ema`: shadow weights updated in the same pass."""
    lib = _lib_setup()
    L.check(lib.simvgb_adam_amsgrad(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), vmax.data_ptr(), p.numel(),
                                    lr, beta1, beta2, eps, weight_decay, step, _p(grad_sumsq), max_norm, _p(ema), ema_decay,
                                    _stream()), "adam_amsgrad")
    _launches[0] += 1


def adam_amsgrad_dev(p, g, m, v, vmax, hyper, beta1, beta2, eps, weight_decay, grad_sumsq=None, max_norm=0.0, ema=None):
    """Adam(amsgrad) with {lr, 1-beta1^t, sqrt(1-beta2^t), ema_decay} read from the device tensor `hyper` (graph-capturable)."""
    lib = _lib_setup()
    L.check(lib.simvgb_adam_amsgrad_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), vmax.data_ptr(), p.numel(),
                                        hyper.data_ptr(), beta1, beta2, eps, weight_decay, _p(grad_sumsq), max_norm, _p(ema),
                                        _stream()), "adam_amsgrad_dev")
    _launches[0] += 1


# ---------------------------------------------------------------------------------------------- object-token head (fp32)
class HeadLinArgs(ctypes.Structure):
    _fields_ = [("x", L.c_vp), ("x2", L.c_vp), ("W", L.c_vp), ("b", L.c_vp), ("drop_u", L.c_vp), ("y", L.c_vp), ("dy", L.c_vp),
                ("dx", L.c_vp), ("dx2", L.c_vp), ("dW", L.c_vp), ("db", L.c_vp),
                ("R", L.c_int), ("N", L.c_int), ("K", L.c_int), ("n_split", L.c_int), ("relu", L.c_int), ("k_splits", L.c_int),
                ("drop_p", L.c_f32)]


class HeadLnArgs(ctypes.Structure):
    _fields_ = [("a", L.c_vp), ("b", L.c_vp), ("drop_u", L.c_vp), ("gamma", L.c_vp), ("beta", L.c_vp), ("y", L.c_vp),
                ("mean", L.c_vp), ("rstd", L.c_vp), ("dy", L.c_vp), ("da", L.c_vp), ("db", L.c_vp), ("dgamma", L.c_vp),
                ("dbeta", L.c_vp), ("R", L.c_int), ("C", L.c_int), ("drop_p", L.c_f32), ("eps", L.c_f32)]


class HeadAttnArgs(ctypes.Structure):
    _fields_ = [("q", L.c_vp), ("k", L.c_vp), ("v", L.c_vp), ("kpm", L.c_vp), ("drop_u", L.c_vp), ("ctx", L.c_vp), ("P", L.c_vp),
                ("dctx", L.c_vp), ("dq", L.c_vp), ("dk", L.c_vp), ("dv", L.c_vp),
                ("B", L.c_int), ("nq", L.c_int), ("nk", L.c_int), ("H", L.c_int), ("ldq", L.c_int), ("ldk", L.c_int), ("ldc", L.c_int),
                ("scale", L.c_f32), ("drop_p", L.c_f32)]


class HeadXAttnArgs(ctypes.Structure):
    _fields_ = [("q", L.c_vp), ("kin", L.c_vp), ("val", L.c_vp), ("Wk", L.c_vp), ("bk", L.c_vp), ("Wv", L.c_vp), ("bv", L.c_vp),
                ("kpm", L.c_vp), ("drop_u", L.c_vp), ("ctx", L.c_vp), ("P", L.c_vp), ("z", L.c_vp), ("psum", L.c_vp),
                ("dctx", L.c_vp), ("dq", L.c_vp), ("dkin", L.c_vp), ("dval", L.c_vp), ("dWk", L.c_vp), ("dbk", L.c_vp),
                ("dWv", L.c_vp), ("dbv", L.c_vp),
                ("B", L.c_int), ("nq", L.c_int), ("N", L.c_int), ("E", L.c_int), ("H", L.c_int), ("drop_p", L.c_f32), ("scale", L.c_f32),
                ("ws", L.c_vp), ("ws_floats", ctypes.c_longlong)]


def _f32c(t):
    """fp32, contiguous (the head kernels take plain row-major fp32 buffers)."""
    if t is None:
        return None
    assert t.dtype == f32 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t


def head_lin_fwd(x, W, b=None, x2=None, n_split=0, relu=False, drop_u=None, drop_p=0.0, k_splits=1, out=None):
    """y = dropout(relu?((x + x2 [first n_split outputs only]) W^T + b)); x [R, K], W [N, K] -> [R, N]."""
    L.require_device(x)
    lib = _lib_setup()
    R, Kd = x.shape
    N = W.shape[0]
    if out is None:
        out = (torch.zeros if k_splits > 1 else torch.empty)(R, N, device=x.device, dtype=f32)
    a = HeadLinArgs()
    a.x, a.x2, a.W, a.b, a.drop_u, a.y = _p(_f32c(x)), _p(_f32c(x2)), _p(_f32c(W)), _p(b), _p(drop_u), out.data_ptr()
    a.R, a.N, a.K, a.n_split, a.relu, a.k_splits, a.drop_p = R, N, Kd, (N if x2 is not None and n_split <= 0 else n_split), int(relu), k_splits, drop_p
    L.check(lib.simvgb_head_lin_fwd(ctypes.byref(a), L.c_vp(_stream())), "head_lin_fwd")
    _launches[0] += 1
    return out


def head_lin_bwd(dy, x, W, *, y=None, x2=None, n_split=0, relu=False, drop_u=None, drop_p=0.0, dx=None, dx2=None, dW=None, db=None):
    """Accumulates dx (+ dx2) += dY_eff W, dW += dY_eff^T (x [+ x2]), db += colsum(dY_eff)."""
    lib = _lib_setup()
    R, Kd = x.shape
    N = W.shape[0]
    a = HeadLinArgs()
    a.x, a.x2, a.W, a.drop_u, a.y, a.dy = _p(_f32c(x)), _p(_f32c(x2)), _p(_f32c(W)), _p(drop_u), _p(y), _p(_f32c(dy))
    a.dx, a.dx2, a.dW, a.db = _p(dx), _p(dx2), _p(dW), _p(db)
    a.R, a.N, a.K, a.n_split, a.relu, a.k_splits, a.drop_p = R, N, Kd, (N if x2 is not None and n_split <= 0 else n_split), int(relu), 1, drop_p
    L.check(lib.simvgb_head_lin_bwd(ctypes.byref(a), L.c_vp(_stream())), "head_lin_bwd")
    _launches[0] += (1 if dW is not None else 0) + (0 if dx is None and dx2 is None else (2 if (x2 is not None and a.n_split < N) else 1))


def head_lnres_fwd(a_in, b_in, gamma, beta, eps=1e-5, drop_u=None, drop_p=0.0):
    """y = LN(a + dropout(b)) -> (y, mean, rstd)."""
    L.require_device(a_in)
    lib = _lib_setup()
    R, C = a_in.shape
    y = torch.empty_like(a_in)
    mean = torch.empty(R, device=a_in.device, dtype=f32)
    rstd = torch.empty(R, device=a_in.device, dtype=f32)
    a = HeadLnArgs()
    a.a, a.b, a.drop_u, a.gamma, a.beta, a.y, a.mean, a.rstd = _p(_f32c(a_in)), _p(_f32c(b_in)), _p(drop_u), _p(gamma), _p(beta), y.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.R, a.C, a.drop_p, a.eps = R, C, drop_p, eps
    L.check(lib.simvgb_head_lnres(ctypes.byref(a), 0, L.c_vp(_stream())), "head_lnres_fwd")
    _launches[0] += 1
    return y, mean, rstd


def head_lnres_bwd(dy, a_in, b_in, gamma, mean, rstd, dgamma, dbeta, da=None, db=None, drop_u=None, drop_p=0.0, eps=1e-5):
    lib = _lib_setup()
    R, C = a_in.shape
    a = HeadLnArgs()
    a.a, a.b, a.drop_u, a.gamma, a.mean, a.rstd, a.dy = _p(_f32c(a_in)), _p(_f32c(b_in)), _p(drop_u), _p(gamma), mean.data_ptr(), rstd.data_ptr(), _p(_f32c(dy))
    a.da, a.db, a.dgamma, a.dbeta = _p(da), _p(db), dgamma.data_ptr(), dbeta.data_ptr()
    a.R, a.C, a.drop_p, a.eps = R, C, drop_p, eps
    L.check(lib.simvgb_head_lnres(ctypes.byref(a), 1, L.c_vp(_stream())), "head_lnres_bwd")
    _launches[0] += 1


def _attn_small_args(q, k, v, B, nq, nk, H, ldq, ldk, ldc, scale, kpm, drop_u, drop_p):
    a = HeadAttnArgs()
    a.q, a.k, a.v, a.kpm, a.drop_u = q.data_ptr(), k.data_ptr(), v.data_ptr(), _p(kpm), _p(drop_u)
    a.B, a.nq, a.nk, a.H, a.ldq, a.ldk, a.ldc, a.scale, a.drop_p = B, nq, nk, H, ldq, ldk, ldc, scale, drop_p
    return a


def head_attn_small_fwd(q, k, v, B, nq, nk, H, scale, kpm=None, drop_u=None, drop_p=0.0):
    """q [B*nq, *], k / v [B*nk, *] (2-D views with unit column stride; row strides may differ: packed projections) -> ctx [B*nq, H*32], P."""
    L.require_device(q)
    lib = _lib_setup()
    ctx = torch.empty(B * nq, H * 32, device=q.device, dtype=f32)
    P = torch.empty(B, H, nq, nk, device=q.device, dtype=f32)
    assert q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1 and k.stride(0) == v.stride(0)
    a = _attn_small_args(q, k, v, B, nq, nk, H, q.stride(0), k.stride(0), ctx.stride(0), scale, kpm, drop_u, drop_p)
    a.ctx, a.P = ctx.data_ptr(), P.data_ptr()
    L.check(lib.simvgb_head_attn_small(ctypes.byref(a), 0, L.c_vp(_stream())), "head_attn_small_fwd")
    _launches[0] += 1
    return ctx, P


def head_attn_small_bwd(dctx, q, k, v, P, dq, dk, dv, B, nq, nk, H, scale, drop_u=None, drop_p=0.0):
    """Accumulates into dq / dk / dv (views with the same strides as q / k / v)."""
    lib = _lib_setup()
    assert dq.stride(0) == q.stride(0) and dk.stride(0) == k.stride(0) and dv.stride(0) == k.stride(0) and dctx.is_contiguous()
    a = _attn_small_args(q, k, v, B, nq, nk, H, q.stride(0), k.stride(0), dctx.stride(0), scale, None, drop_u, drop_p)
    a.P, a.dctx, a.dq, a.dk, a.dv = P.data_ptr(), dctx.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    L.check(lib.simvgb_head_attn_small(ctypes.byref(a), 1, L.c_vp(_stream())), "head_attn_small_bwd")
    _launches[0] += 1


def _xattn_args(q, kin, val, Wk, bk, Wv, bv, B, nq, N, kpm, drop_u, drop_p):
    a = HeadXAttnArgs()
    a.q, a.kin, a.val = _p(_f32c(q)), _p(_f32c(kin)), _p(_f32c(val))
    a.Wk, a.bk, a.Wv, a.bv = _p(_f32c(Wk)), _p(_f32c(bk)), _p(_f32c(Wv)), _p(_f32c(bv))
    a.kpm, a.drop_u, a.B, a.nq, a.N, a.E, a.H, a.drop_p, a.scale = _p(kpm), _p(drop_u), B, nq, N, 256, 8, drop_p, 32 ** -0.5
    return a


def _xattn_ws(lib, a, B, nq, N, backward, dev):
    """Scratch of one absorbed cross-attention call (stream-ordered: freed back to the caching allocator right after the launch)."""
    n = lib.simvgb_head_xattn_ws_floats(B, nq, N, backward)
    ws = torch.empty(n, device=dev, dtype=f32)
    a.ws, a.ws_floats = ws.data_ptr(), n
    return ws


def head_xattn_fwd(q, kin, val, Wk, bk, Wv, bv, B, nq, N, kpm=None, drop_u=None, drop_p=0.0):
    """Absorbed-projection cross-attention over the image memory -> (ctx [B*nq, 256], P [B*nq, 8, N], z [B*nq, 8, 256], psum [B*nq, 8])."""
    L.require_device(q)
    lib = _lib_setup()
    R = B * nq
    ctx = torch.empty(R, 256, device=q.device, dtype=f32)
    P = torch.empty(R, 8, N, device=q.device, dtype=f32)
    z = torch.empty(R, 8, 256, device=q.device, dtype=f32)
    psum = torch.empty(R, 8, device=q.device, dtype=f32)
    a = _xattn_args(q, kin, val, Wk, bk, Wv, bv, B, nq, N, kpm, drop_u, drop_p)
    a.ctx, a.P, a.z, a.psum = ctx.data_ptr(), P.data_ptr(), z.data_ptr(), psum.data_ptr()
    ws = _xattn_ws(lib, a, B, nq, N, 0, q.device)
    L.check(lib.simvgb_head_xattn(ctypes.byref(a), 0, L.c_vp(_stream())), "head_xattn_fwd")
    del ws
    _launches[0] += 5
    return ctx, P, z, psum


def head_xattn_bwd(dctx, q, kin, val, Wk, bk, Wv, bv, P, z, psum, B, nq, N, dq, dkin, dval, dWk, dbk, dWv, dbv, kpm=None, drop_u=None,
                   drop_p=0.0):
    lib = _lib_setup()
    a = _xattn_args(q, kin, val, Wk, bk, Wv, bv, B, nq, N, kpm, drop_u, drop_p)
    a.P, a.z, a.psum, a.dctx = P.data_ptr(), z.data_ptr(), psum.data_ptr(), _p(_f32c(dctx))
    a.dq, a.dkin, a.dval, a.dWk, a.dbk, a.dWv, a.dbv = (t.data_ptr() for t in (dq, dkin, dval, dWk, dbk, dWv, dbv))
    ws = _xattn_ws(lib, a, B, nq, N, 1, q.device)
    L.check(lib.simvgb_head_xattn(ctypes.byref(a), 1, L.c_vp(_stream())), "head_xattn_bwd")
    del ws
    _launches[0] += 9
