"""Steady-state cost of the head kernels' smallest launches (one GPU): chained launches inside a CUDA graph, per-launch time.
Tooling for csrc/headops.cu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simvg_b200 import kernels as K  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def timed(name, fn, n=50, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    print("%-52s %7.2f us / launch" % (name, best), flush=True)


def main():
    for R, Kd, N in ((64, 256, 256), (64, 256, 2048), (64, 2048, 256), (640, 256, 256), (1344, 768, 256), (64, 256, 4)):
        x = torch.randn(R, Kd, device=dev)
        W = torch.randn(N, Kd, device=dev) * 0.05
        b = torch.randn(N, device=dev)
        y = torch.empty(R, N, device=dev)
        timed("lin_fwd R=%d K=%d N=%d" % (R, Kd, N), lambda: K.head_lin_fwd(x, W, b, out=y))
        timed("  torch addmm", lambda: torch.addmm(b, x, W.t(), out=y))
        dy = torch.randn(R, N, device=dev)
        dx = torch.zeros(R, Kd, device=dev)
        dW = torch.zeros(N, Kd, device=dev)
        db = torch.zeros(N, device=dev)
        timed("  lin_bwd (dx + dW + db)", lambda: K.head_lin_bwd(dy, x, W, dx=dx, dW=dW, db=db))
    R, C = 64, 256
    a, bb = torch.randn(R, C, device=dev), torch.randn(R, C, device=dev)
    gm, bt = torch.randn(C, device=dev), torch.randn(C, device=dev)
    timed("lnres_fwd R=64", lambda: K.head_lnres_fwd(a, bb, gm, bt))
    yv, mu, rs = K.head_lnres_fwd(a, bb, gm, bt)
    dy = torch.randn(R, C, device=dev)
    da, dbb, dg, dbt = torch.zeros(R, C, device=dev), torch.zeros(R, C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    timed("lnres_bwd R=64", lambda: K.head_lnres_bwd(dy, a, bb, gm, mu, rs, dg, dbt, da=da, db=dbb))
    timed("  torch layer_norm fwd", lambda: torch.nn.functional.layer_norm(a + bb, (C,), gm, bt))
    z = torch.zeros(1, device=dev)
    timed("aten add_ (1 element)", lambda: z.add_(1.0))
    B, nq, N = 64, 1, 1600
    q = torch.randn(B * nq, 256, device=dev)
    kin, val = torch.randn(B * N, 256, device=dev), torch.randn(B * N, 256, device=dev)
    Wk, Wv = torch.randn(256, 256, device=dev) * 0.05, torch.randn(256, 256, device=dev) * 0.05
    bk, bv = torch.randn(256, device=dev), torch.randn(256, device=dev)
    timed("xattn fwd (5 launches) B=64 nq=1 N=1600", lambda: K.head_xattn_fwd(q, kin, val, Wk, bk, Wv, bv, B, nq, N), n=10)
    ctx, P, zz, psum = K.head_xattn_fwd(q, kin, val, Wk, bk, Wv, bv, B, nq, N)
    dctx = torch.randn_like(ctx)
    outs = [torch.zeros_like(t) for t in (q, kin, val, Wk, bk, Wv, bv)]
    timed("xattn bwd (9 launches)", lambda: K.head_xattn_bwd(dctx, q, kin, val, Wk, bk, Wv, bv, P, zz, psum, B, nq, N, *outs), n=10)


if __name__ == "__main__":
    main()
