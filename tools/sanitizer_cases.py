"""Small-shape launches of every tcgen05 / mbarrier kernel, for compute-sanitizer (tools/sanitize.sh).  Tooling, not product."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simvg_b200 import kernels as K  # noqa: E402


def main():
    torch.manual_seed(0)
    dev = "cuda"
    # GEMM: 1-CTA kernel (narrow N), 2-CTA kernel (pair), MN-major operands, split-K atomics
    X = torch.randn(520, 256, device=dev).bfloat16()
    W = torch.randn(72, 256, device=dev).bfloat16()
    K.gemm(X, W, 520, 72, 256, epilogue=K.EPI_F32)
    Xv, Xt = torch.randn(1000, 256, device=dev).bfloat16(), torch.randn(264, 256, device=dev).bfloat16()
    Wv, Wt = torch.randn(512, 256, device=dev).bfloat16(), torch.randn(512, 256, device=dev).bfloat16()
    bv = torch.randn(512, device=dev)
    K.gemm_pair((Xv, Wv, 1000, 512, 256, dict(epilogue=K.EPI_BF16, bias=bv)), (Xt, Wt, 264, 512, 256, dict(epilogue=K.EPI_BF16, bias=bv)))
    dY = torch.randn(1000, 512, device=dev).bfloat16()
    g = torch.zeros(512, 256, device=dev)
    K.gemm(dY, Xv, 512, 256, 1000, a_mn=True, b_mn=True, epilogue=K.EPI_ATOMIC, k_splits=2, out=g)
    # attention fwd + bwd: vision tail tile + text tile + padding
    B, H, Lv, Lt = 2, 2, 197, 20
    D = H * 64
    qkv_v = (torch.randn(B * Lv, 3 * D, device=dev) * 0.5).bfloat16()
    qkv_t = (torch.randn(B * Lt, 3 * D, device=dev) * 0.5).bfloat16()
    pad = torch.zeros(B, Lt, dtype=torch.uint8, device=dev)
    pad[1, 13:] = 1
    o_v, o_t, lse = K.attn_fwd(qkv_v, qkv_t, pad, B, H, Lv, Lt)
    K.attn_bwd(qkv_v, qkv_t, pad, o_v, o_t, lse, torch.randn_like(o_v), torch.randn_like(o_t), B, H, Lv, Lt)
    # row kernels
    x = torch.randn(300, 768, device=dev)
    gm, bt = torch.randn(768, device=dev), torch.randn(768, device=dev)
    y, mean, rstd = K.ln_fwd(x, gm, bt, 1e-5)
    u = torch.randn(300, 3072, device=dev).bfloat16()
    g3 = torch.randn(3072, device=dev)
    yf, m2, r2 = K.ln_fwd(u, g3, torch.zeros_like(g3), 1e-5, gelu=True)
    dg, db, dbias = (torch.zeros(3072, device=dev) for _ in range(3))
    du = torch.empty(300, 3072, device=dev, dtype=torch.bfloat16)
    K.ln_bwd(2, None, torch.randn(300, 3072, device=dev).bfloat16(), g3, m2, r2, dg, db, dx=du, u=u, dbias_prev=dbias)
    torch.cuda.synchronize()
    print("sanitizer cases done")


if __name__ == "__main__":
    main()
