"""-m gpu: every CUDA kernel (called through the C ABI via simvg_b200.kernels) against fp32 PyTorch math on the same inputs.

Tolerances.  Kernels that *output fp32* from bf16 operands are compared with an fp32 matmul of the same bf16 values: only
the accumulation order differs -> 1e-5 relative (max-norm).  Kernels that *output bf16* carry one bf16 rounding (2^-9
relative per element): 4e-3 of the reference's max magnitude.  Attention additionally rounds P (and dS) to bf16 before the
second MMA: 1.5e-2 of max magnitude.  End-to-end (north_star: 1e-3 relative on model outputs) is asserted in
tests/test_gpu_model.py.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K(lib):
    from simvg_b200 import kernels
    kernels.L.check(lib.simvgb_device_check(0), "device_check")
    return kernels


def maxrel(x, ref):
    return ((x.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30)).item()


DEV = "cuda"


@pytest.mark.parametrize("M,N,K_", [(128, 256, 64), (1000, 768, 768), (333, 128, 192), (2048, 2304, 768), (500, 72, 256), (1280, 3072, 768)])
def test_gemm_all_operand_majors(K, M, N, K_):
    torch.manual_seed(0)
    X = torch.randn(M, K_, device=DEV).bfloat16()
    W = torch.randn(N, K_, device=DEV).bfloat16()
    ref = X.float() @ W.float().t()
    assert maxrel(K.gemm(X, W, M, N, K_, epilogue=K.EPI_F32), ref) < 1e-5
    Wt = W.t().contiguous()
    if N % 8 == 0:
        assert maxrel(K.gemm(X, Wt, M, N, K_, b_mn=True, epilogue=K.EPI_F32), ref) < 1e-5
    if M % 8 == 0 and N % 8 == 0:
        Xt = X.t().contiguous()
        assert maxrel(K.gemm(Xt, W, M, N, K_, a_mn=True, epilogue=K.EPI_F32), ref) < 1e-5
        assert maxrel(K.gemm(Xt, Wt, M, N, K_, a_mn=True, b_mn=True, epilogue=K.EPI_F32), ref) < 1e-5
        out = torch.zeros(M, N, device=DEV)
        K.gemm(Xt, Wt, M, N, K_, a_mn=True, b_mn=True, epilogue=K.EPI_ATOMIC, k_splits=3, out=out)
        assert maxrel(out, ref) < 1e-5


def test_gemm_pair_equals_two_gemms(K):
    """simvgb_gemm_pair: the vision-expert and text-expert problems of one layer in a single persistent launch — different
    M, operand majors, epilogues and split-K factors — must equal two separate launches bit for bit (same tile schedule per
    problem, deterministic epilogues) and fall back cleanly when one problem is not eligible for the 2-CTA kernel."""
    torch.manual_seed(4)
    Mv, Mt, N, K_ = 3000, 520, 768, 768
    Xv, Xt = torch.randn(Mv, K_, device=DEV).bfloat16(), torch.randn(Mt, K_, device=DEV).bfloat16()
    Wv, Wt = torch.randn(N, K_, device=DEV).bfloat16(), torch.randn(N, K_, device=DEV).bfloat16()
    bv, bt = torch.randn(N, device=DEV), torch.randn(N, device=DEV)
    rv, rt = torch.randn(Mv, N, device=DEV), torch.randn(Mt, N, device=DEV)
    # forward-style: bf16 output with bias / scale, and residual epilogue
    for kw_v, kw_t in ((dict(epilogue=K.EPI_BF16, bias=bv, scale=0.125, scale_cols=256), dict(epilogue=K.EPI_BF16, bias=bt, scale=0.125, scale_cols=256)),
                       (dict(epilogue=K.EPI_RESID, bias=bv, res=rv), dict(epilogue=K.EPI_RESID, bias=bt, res=rt)),
                       (dict(epilogue=K.EPI_F32), dict(epilogue=K.EPI_BF16, b_mn=True))):
        Wt_use = Wt.t().contiguous() if kw_t.get("b_mn") else Wt
        pv, pt = K.gemm_pair((Xv, Wv, Mv, N, K_, kw_v), (Xt, Wt_use, Mt, N, K_, kw_t))
        sv, st = K.gemm(Xv, Wv, Mv, N, K_, **kw_v), K.gemm(Xt, Wt_use, Mt, N, K_, **kw_t)
        assert torch.equal(pv, sv) and torch.equal(pt, st)
    # weight-gradient style: MN-major operands, split-K atomics for the long problem, plain accumulate for the short one
    dYv, dYt = torch.randn(Mv, N, device=DEV).bfloat16(), torch.randn(Mt, N, device=DEV).bfloat16()
    gv, gt = torch.zeros(N, K_, device=DEV), torch.zeros(N, K_, device=DEV)
    K.wgrad_pair((dYv, Xv, N, K_, Mv, gv), (dYt, Xt, N, K_, Mt, gt))
    assert maxrel(gv, dYv.float().t() @ Xv.float()) < 1e-5
    assert maxrel(gt, dYt.float().t() @ Xt.float()) < 1e-5
    # one problem too narrow for the 2-CTA kernel (N <= 128): two launches behind the same call
    Wn = torch.randn(72, K_, device=DEV).bfloat16()
    pv, pn = K.gemm_pair((Xv, Wv, Mv, N, K_, dict(epilogue=K.EPI_F32)), (Xt, Wn, Mt, 72, K_, dict(epilogue=K.EPI_F32)))
    assert maxrel(pv, Xv.float() @ Wv.float().t()) < 1e-5 and maxrel(pn, Xt.float() @ Wn.float().t()) < 1e-5


def test_gemm_epilogues(K):
    torch.manual_seed(1)
    M, N, K_ = 1000, 768, 768
    X = torch.randn(M, K_, device=DEV).bfloat16()
    W = (torch.randn(N, K_, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV)
    ref = X.float() @ W.float().t() + bias
    o = K.gemm(X, W, M, N, K_, epilogue=K.EPI_BF16, bias=bias, scale=0.125, scale_cols=256)
    r = ref.clone()
    r[:, :256] *= 0.125
    assert maxrel(o, r) < 4e-3
    u = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    g = torch.empty_like(u)
    K.gemm(X, W, M, N, K_, epilogue=K.EPI_GELU, bias=bias, out=u, out2=g)
    assert maxrel(u, ref) < 4e-3 and maxrel(g, torch.nn.functional.gelu(ref)) < 4e-3
    res = torch.randn(M, N, device=DEV)
    rs = torch.rand(10, device=DEV)
    o = K.gemm(X, W, M, N, K_, epilogue=K.EPI_RESID, bias=bias, res=res, row_scale=rs, rows_per_scale=100)
    assert maxrel(o, res + rs.repeat_interleave(100)[:, None] * ref) < 1e-5
    prev = torch.randn(M, N, device=DEV)
    o = K.gemm(X, W, M, N, K_, epilogue=K.EPI_F32, out=prev.clone(), accumulate=True)
    assert maxrel(o, prev + ref - bias) < 1e-5


def test_tensor_map_cache_and_nvtx_ranges(K):
    """TMA descriptors are cached on (pointer, shape, stride, box): a repeated launch on the same buffers re-encodes nothing and
    computes the same result; a recycled pointer with another shape misses instead of reusing a stale descriptor.  NVTX ranges
    (SIMVGB_NVTX / kernels.enable_nvtx) wrap launches without changing them."""
    import ctypes
    from simvg_b200 import _lib as L

    def stats():
        h, m = ctypes.c_longlong(), ctypes.c_longlong()
        L.lib().simvgb_tmap_cache_stats(ctypes.byref(h), ctypes.byref(m))
        return h.value, m.value

    torch.manual_seed(2)
    M, N, K_ = 512, 768, 768
    X = torch.randn(M, K_, device=DEV).bfloat16()
    W = (torch.randn(N, K_, device=DEV) * 0.05).bfloat16()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    K.gemm(X, W, M, N, K_, epilogue=K.EPI_BF16, out=out)
    first = out.clone()
    h0, m0 = stats()
    K.enable_nvtx(True)
    try:
        with K.nvtx("test.range"):
            K.gemm(X, W, M, N, K_, epilogue=K.EPI_BF16, out=out)
    finally:
        K.enable_nvtx(False)
    h1, m1 = stats()
    assert h1 - h0 >= 2 and m1 == m0
    assert torch.equal(out, first)
    # same base pointers, half the rows: other descriptors, correct result
    o2 = K.gemm(X[:256], W, 256, N, K_, epilogue=K.EPI_BF16, out=out[:256])
    h2, m2 = stats()
    assert m2 > m1
    assert maxrel(o2, X[:256].float() @ W.float().t()) < 4e-3


def test_gemm_full_size_identity_is_exact(K):
    """Size-independent property at the cfg2 problem size (M = 64*1601 rows): X @ I == X bit-for-bit."""
    M, D = 64 * 1601, 768
    torch.manual_seed(2)
    X = torch.randn(M, D, device=DEV).bfloat16()
    eye = torch.eye(D, device=DEV).bfloat16()
    out = K.gemm(X, eye, M, D, D, epilogue=K.EPI_F32)
    assert torch.equal(out, X.float())
    out_b = K.gemm(X, eye, M, D, D, epilogue=K.EPI_BF16)
    assert torch.equal(out_b, X)


def _attn_ref(qkv_v, qkv_t, pad, B, H, Lv, Lt, dout_v=None, dout_t=None):
    D = H * 64
    x = torch.cat([qkv_v.view(B, Lv, 3 * D), qkv_t.view(B, Lt, 3 * D)], 1).float().requires_grad_(True)
    q, k, v = x.split(D, dim=-1)
    L = Lv + Lt
    q, k, v = (t.view(B, L, H, 64).transpose(1, 2) for t in (q, k, v))
    s = q @ k.transpose(-1, -2)
    kpm = torch.cat([torch.zeros(B, Lv, dtype=torch.bool, device=x.device), pad.bool()], 1)
    s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, L, D)
    if dout_v is None:
        return o, None
    o.backward(torch.cat([dout_v.view(B, Lv, D), dout_t.view(B, Lt, D)], 1).float())
    return o, x.grad


@pytest.mark.parametrize("B,H,Lv,Lt,pads", [(2, 2, 100, 20, [0, 7]), (2, 12, 197, 20, [3, 12]), (3, 4, 401, 20, [0, 5, 15]),
                                            (2, 2, 256, 20, [1, 0]), (2, 12, 1601, 20, [2, 9]), (1, 4, 2305, 20, [4]),
                                            (2, 2, 120, 16, [15, 0]), (1, 2, 5, 20, [18])])
def test_attention_fwd_bwd(K, B, H, Lv, Lt, pads):
    torch.manual_seed(3)
    D = H * 64
    qkv_v = (torch.randn(B * Lv, 3 * D, device=DEV) * 0.7)
    qkv_t = (torch.randn(B * Lt, 3 * D, device=DEV) * 0.7)
    qkv_v[:, :D] *= 0.125
    qkv_t[:, :D] *= 0.125
    qkv_v, qkv_t = qkv_v.bfloat16(), qkv_t.bfloat16()
    pad = torch.zeros(B, Lt, dtype=torch.uint8, device=DEV)
    for b, n in enumerate(pads):
        if n:
            pad[b, Lt - n:] = 1
    dout_v = torch.randn(B * Lv, D, device=DEV).bfloat16()
    dout_t = torch.randn(B * Lt, D, device=DEV).bfloat16()
    o_v, o_t, lse = K.attn_fwd(qkv_v, qkv_t, pad, B, H, Lv, Lt)
    dqkv_v, dqkv_t = K.attn_bwd(qkv_v, qkv_t, pad, o_v, o_t, lse, dout_v, dout_t, B, H, Lv, Lt)
    o_ref, g = _attn_ref(qkv_v, qkv_t, pad, B, H, Lv, Lt, dout_v, dout_t)
    tol = 1.5e-2
    assert maxrel(o_v, o_ref[:, :Lv].reshape(B * Lv, D)) < tol
    assert maxrel(o_t, o_ref[:, Lv:].reshape(B * Lt, D)) < tol
    gv, gt = g[:, :Lv].reshape(B * Lv, 3 * D), g[:, Lv:].reshape(B * Lt, 3 * D)
    for got, want in ((dqkv_v, gv), (dqkv_t, gt)):
        assert maxrel(got[:, :D], want[:, :D] * 0.125) < tol      # kernel returns d/d(unscaled q)
        assert maxrel(got[:, D:2 * D], want[:, D:2 * D]) < tol
        assert maxrel(got[:, 2 * D:], want[:, 2 * D:]) < tol


def test_attention_full_size_properties(K):
    """cfg2 problem size (B=64, H=12, L=1621).  Softmax rows sum to one: with V == 1 the output is exactly 1 for every
    query row whatever Q/K/padding are; and with dO == 0 every gradient is exactly 0."""
    B, H, Lv, Lt = 64, 12, 1601, 20
    D = H * 64
    torch.manual_seed(4)
    qkv_v = torch.randn(B * Lv, 3 * D, device=DEV).bfloat16()
    qkv_t = torch.randn(B * Lt, 3 * D, device=DEV).bfloat16()
    qkv_v[:, 2 * D:] = 1
    qkv_t[:, 2 * D:] = 1
    pad = (torch.arange(Lt, device=DEV)[None, :] >= torch.randint(5, Lt + 1, (B, 1), device=DEV)).to(torch.uint8)
    o_v, o_t, lse = K.attn_fwd(qkv_v, qkv_t, pad, B, H, Lv, Lt)
    assert (o_v.float() - 1).abs().max().item() <= 2 ** -7 and (o_t.float() - 1).abs().max().item() <= 2 ** -7
    assert torch.isfinite(lse[:, :, :Lv]).all()
    z_v, z_t = torch.zeros_like(o_v), torch.zeros_like(o_t)
    dv, dt = K.attn_bwd(qkv_v, qkv_t, pad, o_v, o_t, lse, z_v, z_t, B, H, Lv, Lt)
    assert dv.float().abs().max().item() == 0 and dt.float().abs().max().item() == 0


@pytest.mark.parametrize("C", [256, 768, 1024, 3072, 4096])
@pytest.mark.parametrize("in_bf16", [False, True])
def test_layernorm_fwd(K, C, in_bf16):
    torch.manual_seed(5)
    R = 777
    x = torch.randn(R, C, device=DEV) * 2 + 0.5
    if in_bf16:
        x = x.bfloat16()
    g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), g, b, 1e-5)
    y, mean, rstd = K.ln_fwd(x, g, b, 1e-5, out_dtype=torch.float32)
    assert maxrel(y, ref) < 1e-5
    assert torch.allclose(mean, x.float().mean(-1), atol=1e-5)
    yb, _, _ = K.ln_fwd(x, g, b, 1e-5)
    assert maxrel(yb, ref) < 4e-3


def _ln_ref_bwd(x, dy, g):
    x = x.float().requires_grad_(True)
    gg = g.clone().requires_grad_(True)
    bb = torch.zeros_like(g).requires_grad_(True)
    y = torch.nn.functional.layer_norm(x, (x.shape[-1],), gg, bb, 1e-5)
    y.backward(dy.float())
    return x.grad, gg.grad, bb.grad


@pytest.mark.parametrize("C", [768, 1024])
def test_layernorm_bwd_residual_mode(K, C):
    torch.manual_seed(6)
    B, Lr = 6, 50
    R = B * Lr
    x = torch.randn(R, C, device=DEV)
    dy = torch.randn(R, C, device=DEV).bfloat16()
    g = torch.randn(C, device=DEV)
    dres = torch.randn(R, C, device=DEV)
    rs = torch.rand(B, device=DEV) + 0.5
    _, mean, rstd = K.ln_fwd(x, g, torch.zeros_like(g), 1e-5)
    dg, db, dbias = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    out = dres.clone()
    dyb = torch.empty(R, C, device=DEV, dtype=torch.bfloat16)
    K.ln_bwd(0, x, dy, g, mean, rstd, dg, db, dres_in=out, dres_out=out, dyb=dyb, row_scale=rs, rows_per_scale=Lr, dbias_prev=dbias)
    dx_ref, dg_ref, db_ref = _ln_ref_bwd(x, dy, g)
    want = dres + dx_ref
    assert maxrel(out, want) < 1e-5
    scaled = want * rs.repeat_interleave(Lr)[:, None]
    assert maxrel(dyb, scaled) < 4e-3
    assert maxrel(dbias, scaled.sum(0)) < 1e-4 and maxrel(dg, dg_ref) < 1e-4 and maxrel(db, db_ref) < 1e-4
    # fp32 dy, no incoming residual gradient (the final encoder LayerNorm)
    out2 = torch.empty(R, C, device=DEV)
    K.ln_bwd(0, x, dy.float(), g, mean, rstd, torch.zeros(C, device=DEV), torch.zeros(C, device=DEV), dres_in=None, dres_out=out2)
    assert maxrel(out2, dx_ref) < 1e-5


@pytest.mark.parametrize("C", [768, 3072, 4096])
def test_layernorm_bwd_inner_and_gelu_modes(K, C):
    torch.manual_seed(7)
    R = 300
    u = torch.randn(R, C, device=DEV).bfloat16()
    gl = torch.nn.functional.gelu(u.float()).bfloat16()
    dy = torch.randn(R, C, device=DEV).bfloat16()
    g = torch.randn(C, device=DEV)
    _, mean, rstd = K.ln_fwd(gl, g, torch.zeros_like(g), 1e-5)
    dg, db, dbias = (torch.zeros(C, device=DEV) for _ in range(3))
    dx = torch.empty(R, C, device=DEV, dtype=torch.bfloat16)
    K.ln_bwd(1, gl, dy, g, mean, rstd, dg, db, dx=dx)
    dx_ref, dg_ref, db_ref = _ln_ref_bwd(gl, dy, g)
    assert maxrel(dx, dx_ref) < 4e-3 and maxrel(dg, dg_ref) < 1e-4 and maxrel(db, db_ref) < 1e-4
    dg.zero_(); db.zero_()
    # mode 2: the LN input is gelu(u), recomputed in-kernel from u (never stored); fused forward = LN(gelu(u))
    yf, mean2, rstd2 = K.ln_fwd(u, g, torch.zeros_like(g), 1e-5, out_dtype=torch.float32, gelu=True)
    uf = u.float().requires_grad_(True)
    gg = g.clone().requires_grad_(True)
    y_ref = torch.nn.functional.layer_norm(torch.nn.functional.gelu(uf), (C,), gg, torch.zeros_like(g), 1e-5)
    assert maxrel(yf, y_ref) < 1e-5
    y_ref.backward(dy.float())
    du = torch.empty(R, C, device=DEV, dtype=torch.bfloat16)
    K.ln_bwd(2, None, dy, g, mean2, rstd2, dg, db, dx=du, u=u, dbias_prev=dbias)
    assert maxrel(du, uf.grad) < 4e-3 and maxrel(dbias, uf.grad.sum(0)) < 5e-3 and maxrel(dg, gg.grad) < 1e-4


def test_colsum_cast_embed_im2col(K):
    torch.manual_seed(8)
    R, C = 1234, 768
    x = torch.randn(R, C, device=DEV)
    rs = torch.rand(R // 2 + 1, device=DEV)
    out = torch.zeros(C, device=DEV)
    ob = torch.empty(R, C, device=DEV, dtype=torch.bfloat16)
    K.colsum(x, out=out, out_bf16=ob, row_scale=rs, rows_per_scale=2)
    sc = x * rs.repeat_interleave(2)[:R, None]
    assert maxrel(out, sc.sum(0)) < 1e-5 and torch.equal(ob, sc.bfloat16())
    xb = x.bfloat16()
    out.zero_()
    K.colsum(xb[:, :256], out=out[:256], C=256, ld=C)
    assert maxrel(out[:256], xb[:, :256].float().sum(0)) < 1e-5
    dst = torch.empty(R, C, device=DEV, dtype=torch.bfloat16)
    assert torch.equal(K.cast_bf16(x, dst), x.bfloat16())
    # im2col + GEMM == Conv2d(k=P, s=P)
    B, S, P, D = 2, 64, 16, 256
    img = torch.randn(B, 3, S, S, device=DEV)
    w = torch.randn(D, 3, P, P, device=DEV) * 0.05
    cols = K.im2col_patch(img, P)
    ref = torch.nn.functional.conv2d(img.bfloat16().float(), w.bfloat16().float(), stride=P).flatten(2).transpose(1, 2).reshape(-1, D)
    got = K.gemm(cols, w.view(D, -1).bfloat16().contiguous(), cols.shape[0], D, 3 * P * P, epilogue=K.EPI_F32)
    assert maxrel(got, ref) < 1e-5
    N = (S // P) ** 2
    patch, cls, posA = torch.randn(B * N, D, device=DEV), torch.randn(D, device=DEV), torch.randn(N + 3, D, device=DEV)
    xv = K.embed_vision(patch, cls, posA, B, N, D).view(B, N + 1, D)
    want = torch.cat([cls.expand(B, 1, D), patch.view(B, N, D)], 1) + posA[2:2 + N + 1]
    assert torch.equal(xv, want)
    table, posB = torch.randn(50, D, device=DEV), torch.randn(1024, D, device=DEV)
    ids = torch.randint(0, 50, (B, 20), device=DEV)
    pad = (torch.rand(B, 20, device=DEV) > 0.7).to(torch.uint8)
    xt = K.embed_text(table, ids, pad, posB, B, 20, D).view(B, 20, D)
    assert torch.equal(xt, (table[ids] + posB[2:22]) * (1 - pad.float())[..., None])


def test_fused_adam_amsgrad_matches_torch(K):
    torch.manual_seed(9)
    n = 100003
    p0 = torch.randn(n, device=DEV)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=5e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=0, amsgrad=True)
    pad = (n + 3) // 4 * 4
    p = torch.zeros(pad, device=DEV); p[:n] = p0
    m, v, vmax = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    ss = torch.zeros(1, device=DEV)
    for step in range(1, 6):
        g = torch.randn(n, device=DEV) * (10.0 if step % 2 else 0.01)
        ref_p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 0.15)
        opt.step()
        gp = torch.zeros(pad, device=DEV); gp[:n] = g
        ss.zero_()
        K.sumsq(gp, ss)
        assert abs(ss.sqrt().item() - g.norm().item()) < 1e-3 * g.norm().item()
        K.adam_amsgrad(p, gp, m, v, vmax, 5e-4, 0.9, 0.98, 1e-9, 0.0, step, grad_sumsq=ss, max_norm=0.15)
        assert (p[:n] - ref_p.detach()).abs().max().item() < 2e-6
