"""Small-shape launches of the round-2 kernels (head fp32 kernels, device Hungarian, uint8 patch gather, EMA in the optimiser
pass, delta fused into the inner-LN backward, final attention kernels), for compute-sanitizer.  Tooling, not product.

    for t in memcheck racecheck initcheck; do compute-sanitizer --tool $t python tools/sanitizer_cases_head.py; done
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simvg_b200 import kernels as K  # noqa: E402


def main():
    torch.manual_seed(0)
    dev = "cuda"
    r = lambda *s: torch.randn(*s, device=dev)  # noqa: E731
    # ---- linears: ragged rows / columns, position add on the first outputs, ReLU + dropout, split contraction in the backward
    R, Kd, N = 70, 200, 300
    x, x2, W, b = r(R, Kd), r(R, Kd), r(N, Kd) * 0.1, r(N)
    u = torch.rand(R, N, device=dev)
    y = K.head_lin_fwd(x, W, b, x2=x2, n_split=128, relu=True, drop_u=u, drop_p=0.1)
    dx, dx2, dW, db = torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(W), torch.zeros_like(b)
    K.head_lin_bwd(r(R, N), x, W, y=y, x2=x2, n_split=128, relu=True, drop_u=u, drop_p=0.1, dx=dx, dx2=dx2, dW=dW, db=db)
    xb, Wb = r(1100, 96), r(40, 96)          # many rows: the weight gradient splits the row contraction
    dWb, dbb = torch.zeros_like(Wb), torch.zeros(40, device=dev)
    K.head_lin_bwd(r(1100, 40), xb, Wb, dx=torch.zeros_like(xb), dW=dWb, db=dbb)
    # ---- residual + LayerNorm (C = 256 and 512)
    for C in (256, 512):
        a, bb, gm, bt = r(37, C), r(37, C), r(C), r(C)
        uu = torch.rand(37, C, device=dev)
        yy, mu, rs = K.head_lnres_fwd(a, bb, gm, bt, drop_u=uu, drop_p=0.1)
        K.head_lnres_bwd(r(37, C), a, bb, gm, mu, rs, torch.zeros(C, device=dev), torch.zeros(C, device=dev),
                         da=torch.zeros_like(a), db=torch.zeros_like(a), drop_u=uu, drop_p=0.1)
    # ---- few-keys attention with a padding mask
    B, nq, nk, H, E = 3, 4, 20, 8, 256
    q, k, v = r(B * nq, E), r(B * nk, E), r(B * nk, E)
    kpm = torch.zeros(B, nk, dtype=torch.uint8, device=dev)
    kpm[1, 7:] = 1
    ctx, P = K.head_attn_small_fwd(q, k, v, B, nq, nk, H, 32 ** -0.5, kpm=kpm)
    K.head_attn_small_bwd(r(B * nq, E), q, k, v, P, torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v), B, nq, nk, H, 32 ** -0.5)
    # ---- absorbed cross-attention: key count not a multiple of any tile, several queries, mask, dropout
    for (B, nq, N) in ((2, 3, 70), (2, 1, 333)):
        q, kin, val = r(B * nq, E), r(B * N, E), r(B * N, E)
        Wk, Wv, bk, bv = r(E, E) * 0.05, r(E, E) * 0.05, r(E), r(E)
        kpm = torch.zeros(B, N, dtype=torch.uint8, device=dev)
        kpm[0, N // 2:] = 1
        uu = torch.rand(B * nq * H * N, device=dev)
        ctx, P, z, psum = K.head_xattn_fwd(q, kin, val, Wk, bk, Wv, bv, B, nq, N, kpm=kpm, drop_u=uu, drop_p=0.1)
        outs = [torch.zeros_like(t) for t in (q, kin, val, Wk, bk, Wv, bv)]
        K.head_xattn_bwd(r(B * nq, E), q, kin, val, Wk, bk, Wv, bv, P, z, psum, B, nq, N, *outs, kpm=kpm, drop_u=uu, drop_p=0.1)
    # ---- device Hungarian, uint8 patch gather, optimiser pass with EMA
    K.hungarian(torch.rand(3, 5, 9, device=dev), [2, 3, 4])
    img = torch.randint(0, 256, (2, 64, 64, 3), dtype=torch.uint8, device=dev)
    K.im2col_patch_u8(img, 16, [123.675, 116.28, 103.53], [58.395, 57.12, 57.375], True)
    n = 1003
    p, g, m, vv, vm, ema = r(n), r(n), torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev), r(n)
    ss = torch.zeros(1, device=dev)
    K.sumsq(g, ss)
    K.adam_amsgrad(p, g, m, vv, vm, 1e-3, 0.9, 0.98, 1e-9, 0.0, 1, grad_sumsq=ss, max_norm=0.15, ema=ema, ema_decay=0.1)
    # ---- attention fwd + bwd with delta produced by the inner-attention-LN backward (two query tiles per item + a single-tile item)
    B, H, Lv, Lt = 2, 4, 325, 20
    D = H * 64
    qkv_v = (r(B * Lv, 3 * D) * 0.5).bfloat16()
    qkv_t = (r(B * Lt, 3 * D) * 0.5).bfloat16()
    pad = torch.zeros(B, Lt, dtype=torch.uint8, device=dev)
    pad[1, 13:] = 1
    o_v, o_t, lse = K.attn_fwd(qkv_v, qkv_t, pad, B, H, Lv, Lt)
    ws = K.attn_workspace({}, B, H, Lv, Lt, dev)
    gam = r(D)
    for gidx, (o, rows, L) in enumerate(((o_v, B * Lv, Lv), (o_t, B * Lt, Lt))):
        _, mi, ri = K.ln_fwd(o, gam, torch.zeros_like(gam), 1e-5)
        dO = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
        K.ln_bwd(1, o, r(rows, D).bfloat16(), gam, mi, ri, torch.zeros(D, device=dev), torch.zeros(D, device=dev), dx=dO,
                 delta=K.attn_delta_spec(ws, B, H, Lv, Lt, gidx))
        if gidx == 0:
            dOv = dO
        else:
            dOt = dO
    K.attn_bwd(qkv_v, qkv_t, pad, o_v, o_t, lse, dOv, dOt, B, H, Lv, Lt, ws=ws, delta_ready=True)
    torch.cuda.synchronize()
    print("round-2 sanitizer cases done")


if __name__ == "__main__":
    main()
