"""GPU bring-up check for the tcgen05 GEMM (run under gpurun). Prints max errors vs torch fp32 matmul
of the same bf16 inputs, for every operand-major / epilogue combination, then a quick timing."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simvg_b200 import _lib as L  # noqa: E402


def gemm(A, B, M, N, K, a_mn=0, b_mn=0, epi=L.EPI_F32, bias=None, res=None, k_splits=1, scale=1.0, scale_cols=0,
         row_scale=None, rows_per_scale=1, accumulate=0, out_f32=None):
    dev = A.device
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.a_mn_major, a.b_mn_major = a_mn, b_mn
    a.lda, a.ldb = A.stride(0), B.stride(0)
    a.A, a.B = A.data_ptr(), B.data_ptr()
    a.epilogue, a.k_splits = epi, k_splits
    a.bias = bias.data_ptr() if bias is not None else None
    outs = {}
    if epi in (L.EPI_BF16, L.EPI_GELU):
        outs["bf16"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        a.out_bf16 = outs["bf16"].data_ptr()
        if epi == L.EPI_GELU:
            outs["bf16_2"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            a.out2_bf16 = outs["bf16_2"].data_ptr()
    else:
        outs["f32"] = out_f32 if out_f32 is not None else torch.zeros(M, N, device=dev, dtype=torch.float32)
        a.out_f32 = outs["f32"].data_ptr()
    if res is not None:
        a.res_f32 = res.data_ptr()
    a.ldo = N
    a.scale, a.scale_cols = scale, scale_cols
    a.row_scale = row_scale.data_ptr() if row_scale is not None else None
    a.rows_per_scale = rows_per_scale
    a.accumulate = accumulate
    L.check(L.lib().simvgb_gemm(ctypes.byref(a), L.stream_ptr()), "gemm")
    return outs


def relerr(x, ref):
    return ((x.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item()


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    L.check(L.lib().simvgb_device_check(0), "device_check")
    results = {}
    ok = True

    def report(name, err, tol):
        nonlocal ok
        good = err == err and err < tol
        ok = ok and good
        results[name] = err
        print("%-46s err %.3e  %s" % (name, err, "OK" if good else "FAIL"), flush=True)

    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 768), (1000, 768, 768), (333, 128, 192), (2048, 2304, 768)]:
        X = torch.randn(M, K, device=dev).bfloat16()
        W = torch.randn(N, K, device=dev).bfloat16()
        ref = X.float() @ W.float().t()
        # NT (both K-major)
        o = gemm(X, W, M, N, K)["f32"]
        torch.cuda.synchronize()
        report("NT f32 %dx%dx%d" % (M, N, K), relerr(o, ref), 1e-5)
        # A MN-major: stored [K, M]
        Xt = X.t().contiguous()
        if M % 8 == 0:
            o = gemm(Xt, W, M, N, K, a_mn=1)["f32"]
            torch.cuda.synchronize()
            report("A-mn f32 %dx%dx%d" % (M, N, K), relerr(o, ref), 1e-5)
        Wt = W.t().contiguous()
        o = gemm(X, Wt, M, N, K, b_mn=1)["f32"]
        torch.cuda.synchronize()
        report("B-mn f32 %dx%dx%d" % (M, N, K), relerr(o, ref), 1e-5)
        if M % 8 == 0:
            o = gemm(Xt, Wt, M, N, K, a_mn=1, b_mn=1)["f32"]
            torch.cuda.synchronize()
            report("AB-mn f32 %dx%dx%d" % (M, N, K), relerr(o, ref), 1e-5)
            o = gemm(Xt, Wt, M, N, K, a_mn=1, b_mn=1, epi=L.EPI_ATOMIC, k_splits=3)["f32"]
            torch.cuda.synchronize()
            report("AB-mn atomic splitK3 %dx%dx%d" % (M, N, K), relerr(o, ref), 1e-5)

    # epilogues
    M, N, K = 1000, 768, 768
    X = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = X.float() @ W.float().t() + bias
    o = gemm(X, W, M, N, K, epi=L.EPI_BF16, bias=bias, scale=0.125, scale_cols=256)["bf16"]
    ref_s = ref.clone()
    ref_s[:, :256] *= 0.125
    report("epi BF16+bias+scale", relerr(o, ref_s), 1e-2)
    outs = gemm(X, W, M, N, K, epi=L.EPI_GELU, bias=bias)
    report("epi GELU u", relerr(outs["bf16"], ref), 1e-2)
    report("epi GELU g", relerr(outs["bf16_2"], torch.nn.functional.gelu(ref)), 1e-2)
    res = torch.randn(M, N, device=dev)
    rsc = torch.rand(10, device=dev)
    o = gemm(X, W, M, N, K, epi=L.EPI_RESID, bias=bias, res=res, row_scale=rsc, rows_per_scale=100)["f32"]
    ref_r = res + rsc.repeat_interleave(100)[:, None] * ref
    report("epi RESID+rowscale", relerr(o, ref_r), 1e-5)
    prev = torch.randn(M, N, device=dev)
    o = gemm(X, W, M, N, K, epi=L.EPI_F32, accumulate=1, out_f32=prev.clone())["f32"]
    report("epi F32 accumulate", relerr(o, prev + ref - bias), 1e-5)
    # small-N (BN=128 path) with ragged N
    M, N, K = 500, 72, 256
    X = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    o = gemm(X, W, M, N, K)["f32"]
    report("BN128 ragged N=72", relerr(o, X.float() @ W.float().t()), 1e-5)

    # timing
    def bench(M, N, K, a_mn=0, b_mn=0, epi=L.EPI_BF16, k_splits=1, iters=20):
        A = (torch.randn(K, M, device=dev) if a_mn else torch.randn(M, K, device=dev)).bfloat16()
        B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).bfloat16()
        out = torch.zeros(M, N, device=dev) if epi in (L.EPI_F32, L.EPI_ATOMIC) else None
        for _ in range(3):
            gemm(A, B, M, N, K, a_mn, b_mn, epi, k_splits=k_splits, out_f32=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            gemm(A, B, M, N, K, a_mn, b_mn, epi, k_splits=k_splits, out_f32=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = 2.0 * M * N * K / ms / 1e9
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        Af = A.t() if a_mn else A
        Bf = B if b_mn else B.t()
        for _ in range(3):
            torch.matmul(Af, Bf)
        t0.record()
        for _ in range(iters):
            torch.matmul(Af, Bf)
        t1.record()
        torch.cuda.synchronize()
        tfc = 2.0 * M * N * K / (t0.elapsed_time(t1) / iters) / 1e9
        print("time M=%d N=%d K=%d a_mn=%d b_mn=%d epi=%d ks=%d: %.3f ms  %.0f TFLOP/s (cuBLAS %.0f)" % (M, N, K, a_mn, b_mn, epi, k_splits, ms, tf, tfc), flush=True)
        results["tflops_%d_%d_%d_%d%d_%d" % (M, N, K, a_mn, b_mn, epi)] = tf

    if ok:
        bench(102464, 2304, 768)
        bench(102464, 768, 768)
        bench(102464, 3072, 768, epi=L.EPI_GELU)
        bench(102464, 768, 3072)
        bench(8192, 8192, 8192)
        bench(2304, 768, 102464, a_mn=1, b_mn=1, epi=L.EPI_ATOMIC, k_splits=8)
        bench(3072, 768, 102464, a_mn=1, b_mn=1, epi=L.EPI_ATOMIC, k_splits=6)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_check.json", "w") as f:
        json.dump({"ok": ok, "results": results}, f, indent=1)
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
