"""-m gpu: round-2 components — fused optimiser as a torch.optim.Optimizer (groups, frozen layers, EMA), the uint8 input path,
and the 2-GPU gradient-exchange equivalence (NCCL; skipped with fewer than two devices)."""
import copy
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _small_model(seed=0, S=128, P=32, freeze_layer=-1, drop_path_rate=0.0):
    from simvg_b200.models import build_model
    from tools.synth import model_cfg
    cfg = model_cfg("base", S, P, drop_path_rate=drop_path_rate)
    cfg["vis_enc"]["freeze_layer"] = freeze_layer
    torch.manual_seed(seed)
    return build_model(cfg)


def _step_loss(model, b):
    losses, _ = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                      text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
    return losses["loss_total"]


def test_fused_optimizer_tracks_torch_adam_on_the_real_model(lib):
    """Three fused steps (clip 0.15 + Adam-amsgrad over the reference's three lr groups + EMA) against
    clip_grad_norm_ + torch.optim.Adam(amsgrad=True) + the reference's EMA arithmetic fed the SAME gradients."""
    from simvg_b200.optim import FusedAdamAMSGrad
    from tools.synth import make_batch
    model = _small_model().cuda().eval()
    twin = copy.deepcopy(model)
    opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5, lr_lan_enc=5e-5, grad_norm_clip=0.15, ema_alpha=0.999)
    named = list(twin.named_parameters())
    ref = torch.optim.Adam([{"params": [p for n, p in named if "vis_enc" in n], "lr": 5e-5},
                            {"params": [p for n, p in named if "lan_enc" in n], "lr": 5e-5},
                            {"params": [p for n, p in named if "vis_enc" not in n and "lan_enc" not in n], "lr": 5e-4}],
                           betas=(0.9, 0.98), eps=1e-9, weight_decay=0, amsgrad=True)
    shadow = {n: p.detach().clone() for n, p in named}
    b = make_batch(4, 128, seed=3, device=DEV)
    for t in range(3):
        opt.zero_grad()
        _step_loss(model, b).backward()
        for (n, p), (_, q) in zip(model.named_parameters(), named):
            q.grad = p.grad.detach().clone()
        opt.step()
        torch.nn.utils.clip_grad_norm_([q for _, q in named], 0.15)
        ref.step()
        decay = min(0.999, (t + 1.0) / (t + 10.0))                         # models/utils.py:149
        for n, q in named:
            shadow[n].copy_(decay * shadow[n] + (1 - decay) * q.detach())
    worst = max(float((p.detach() - q.detach()).abs().max()) for (_, p), (_, q) in zip(model.named_parameters(), named))
    assert worst < 5e-6, worst
    worst_ema = max(float((opt.ema_view(p) - shadow[n]).abs().max()) for n, p in model.named_parameters())
    assert worst_ema < 5e-6, worst_ema
    # checkpoint round trip through torch's own format on the device
    sd = opt.state_dict()
    assert len(sd["param_groups"]) == 3 and sd["param_groups"][0]["lr"] == 5e-5 and len(sd["param_groups"][1]["params"]) == 0
    ref_sd = ref.state_dict()
    k = max(sd["state"])
    assert torch.allclose(sd["state"][k]["exp_avg"], ref_sd["state"][k]["exp_avg"], atol=1e-7)
    assert float(sd["state"][k]["step"]) == 3.0


def test_foreign_zero_grad_does_not_lose_gradients(lib):
    """model.zero_grad() (set_to_none) replaces p.grad by fresh tensors outside the flat buffers: step() must pick them up."""
    from simvg_b200.optim import FusedAdamAMSGrad
    from tools.synth import make_batch
    model = _small_model().cuda().eval()
    opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5)
    b = make_batch(2, 128, seed=4, device=DEV)
    model.zero_grad(set_to_none=True)
    before = model.head.class_embed_decoder.weight.detach().clone()
    _step_loss(model, b).backward()
    opt.step()
    assert float((model.head.class_embed_decoder.weight.detach() - before).abs().max()) > 0
    seg = [s for s in opt.segments if s.name == "rest"][0]
    p = model.head.class_embed_decoder.weight
    assert p.grad.data_ptr() == seg.fb.grad_of(seg.fb.index(p)).data_ptr()


def test_frozen_layers_are_not_updated(lib):
    from simvg_b200.optim import FusedAdamAMSGrad
    from tools.synth import make_batch
    model = _small_model(freeze_layer=2).cuda().train()
    opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5, grad_norm_clip=0.15, weight_decay=0.01)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    b = make_batch(2, 128, seed=5, device=DEV)
    for _ in range(2):
        opt.zero_grad()
        _step_loss(model, b).backward()
        opt.step()
    for n, p in model.named_parameters():
        moved = float((p.detach() - before[n]).abs().max()) > 0
        if not p.requires_grad:
            assert not moved, n                                         # no update, no weight decay
            assert p.grad is None or float(p.grad.abs().max()) == 0, n
        elif "mask_token" not in n:
            assert moved, n


def test_uint8_input_path_equals_host_normalisation(lib):
    """uint8 HWC image -> (normalise + transpose + im2col) fused on the device == mmcv.imnormalize on the host followed by the
    float NCHW path (pipelines/transforms.py:126-155): identical patch matrix, hence identical losses."""
    from simvg_b200 import kernels as K
    from tools.synth import make_batch
    B, S, P = 3, 128, 32
    g = torch.Generator().manual_seed(6)
    u8 = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    rgb = u8.flip(-1).float()                                             # to_rgb: BGR -> RGB
    inv = torch.tensor([float(1.0 / s) for s in std], dtype=torch.float64).float()
    host = ((rgb - torch.tensor(mean)) * inv).permute(0, 3, 1, 2).contiguous()
    cols_ref = K.im2col_patch(host.cuda(), P)
    cols = K.im2col_patch_u8(u8.cuda(), P, mean, std, True)
    assert torch.equal(cols, cols_ref)
    cols_bgr = K.im2col_patch_u8(u8.cuda(), P, mean, std, False)
    host_bgr = ((u8.float() - torch.tensor(mean)) * inv).permute(0, 3, 1, 2).contiguous()
    assert torch.equal(cols_bgr, K.im2col_patch(host_bgr.cuda(), P))
    model = _small_model().cuda().eval()
    b = make_batch(B, S, seed=7, device=DEV)
    b_f = dict(b, img=host.cuda())
    b_u = dict(b, img=u8.cuda())
    lf, lu = _step_loss(model, b_f), _step_loss(model, b_u)
    assert abs(float(lf) - float(lu)) <= 1e-6 * abs(float(lf))
    lu.backward()
    assert float(model.vis_enc.beit3.vision_embed.proj.weight.grad.abs().sum()) > 0


# ------------------------------------------------------------------------------------------------ two GPUs
def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    from datetime import timedelta
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank), timeout=timedelta(seconds=90))
    from simvg_b200.optim import FlatDDP, FusedAdamAMSGrad
    from simvg_b200.runtime import GraphedTrainStep
    from tools.synth import make_batch
    res = {}
    batches = [make_batch(3, 128, seed=50 + 10 * i + rank, device="cuda") for i in range(2)]

    from simvg_b200.models.heads.tgqs_kd_detr_head import native as nat

    def fresh(native_head=True):
        # The head kernels' backward accumulates a few gradients with fp32 atomics: two runs differ in the last bit of the gradient
        # entering the encoder, and the encoder backward's bf16 roundings turn that into ~1e-3 differences on this tiny problem
        # (tools/diag_grad_noise.py).  Legs (1)/(2) compare two separate backward passes element-wise, so they run the head on
        # its deterministic op-by-op path; leg (3) runs the product path.
        nat._linear_native[0] = native_head
        m = _small_model(seed=11).cuda().train()
        for mod in m.modules():           # stochastic layers off: runs must be comparable
            if hasattr(mod, "use_native"):
                mod.use_native = native_head
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
            if hasattr(mod, "attn_drop") and isinstance(mod.attn_drop, float):
                mod.attn_drop = 0.0
        return m, FusedAdamAMSGrad(m, lr=2e-4, lr_vis_enc=2e-5, grad_norm_clip=0.15)

    # (1) local gradients without any exchange
    m0, o0 = fresh(native_head=False)
    o0.zero_grad()
    _step_loss(m0, batches[0]).backward()
    local = [s.fb.grad.clone() for s in o0.segments]
    gathered = [[torch.empty_like(g) for _ in range(world)] for g in local]
    for g, lst in zip(local, gathered):
        dist.all_gather(lst, g)
    want = [torch.stack(lst).mean(0) for lst in gathered]
    # (2) overlap mode: ranges reduced from inside the encoder backward (+ sparse text-embedding rows)
    m1, o1 = fresh(native_head=False)
    d1 = FlatDDP(m1, o1)
    d1.broadcast_parameters()
    o1.zero_grad()
    _step_loss(m1, batches[0]).backward()
    d1.finish()
    torch.cuda.synchronize()
    res["overlap_err"] = max(float((s.fb.grad - w).abs().max() / w.abs().max()) for s, w in zip(o1.segments, want))
    chk = [torch.empty_like(o1.segments[0].fb.grad) for _ in range(world)]
    dist.all_gather(chk, o1.segments[0].fb.grad)
    res["ranks_identical"] = bool(torch.equal(chk[0], chk[1]))
    o1.step()
    # (3) deferred mode under the two-graph runtime: same parameters after two steps as the eager overlap loop
    o1.zero_grad()
    _step_loss(m1, batches[1]).backward()
    d1.finish()
    o1.step()
    nat._linear_native[0] = True
    for mod in m1.modules():
        if hasattr(mod, "use_native"):
            mod.use_native = True
    m2, o2 = fresh()
    d2 = FlatDDP(m2, o2, deferred=True)
    d2.broadcast_parameters()
    step = GraphedTrainStep(m2, o2, d2, warmup=1)
    for i in range(2):
        bb = batches[i]
        step(bb["img"], bb["ref_expr_inds"], bb["img_metas"], bb["text_attention_mask"], torch.stack(bb["gt_bbox"]))
    torch.cuda.synchronize()
    keep = [(p, q) for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters())
            if "k_proj" not in n and "in_proj_bias" not in n]       # zero-gradient key biases random-walk under Adam
    a = torch.cat([p.detach().flatten() for p, _ in keep])
    c = torch.cat([q.detach().flatten() for _, q in keep])
    res["graph_vs_eager_rel"] = float((a - c).norm() / a.norm())
    res["graphs"] = step.graph_opt is not None and len(step.plan) > 2     # backward cut into chunk graphs
    torch.save(res, os.path.join(out, "ddp%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_two_gpu_gradient_exchange_equivalence(lib, tmp_path):
    """After FlatDDP.finish() every rank's flat gradient equals the mean of the per-rank gradients computed without DDP
    (the reference's MMDistributedDataParallel semantics, tools/train.py:102-103), bit-identical across ranks; and the
    two-graph deferred runtime lands on the same parameters as the eager overlapped loop."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = torch.load(tmp_path / ("ddp%d.pt" % r))
        assert res["overlap_err"] < 1e-4, res
        assert res["ranks_identical"], res
        assert res["graphs"] and res["graph_vs_eager_rel"] < 2e-3, res


def test_device_hungarian_matches_scipy(lib):
    """simvgb_hungarian against scipy.optimize.linear_sum_assignment (the reference's solver) on random rectangular problems:
    more queries than targets, fewer, equal, empty target sets — identical assignments, ascending query order."""
    from scipy.optimize import linear_sum_assignment
    from simvg_b200 import kernels as K
    g = torch.Generator().manual_seed(8)
    for nq in (1, 3, 10, 32):
        sizes = [int(x) for x in torch.randint(0, 8, (9,), generator=g)] + [nq, min(32, nq + 5), 1]
        C = torch.randn(len(sizes), nq, sum(sizes), generator=g)
        got = K.hungarian(C.cuda(), sizes)
        off = 0
        for b, n in enumerate(sizes):
            r, c = linear_sum_assignment(C[b, :, off:off + n].double().numpy())
            off += n
            assert got[b][0].cpu().tolist() == list(r) and got[b][1].cpu().tolist() == list(c), (nq, b, n)


def test_nq10_matcher_on_device_equals_scipy_path(lib):
    """HungarianMatcher(nq = 10) on CUDA tensors (device kernel) returns what the scipy path returns on the same CPU tensors."""
    from simvg_b200.core.criterion.criterion import HungarianMatcher
    g = torch.Generator().manual_seed(9)
    B, nq = 6, 10
    logits, boxes = torch.randn(B, nq, 2, generator=g), torch.rand(B, nq, 4, generator=g) * 0.4 + 0.3
    targets = []
    for b in range(B):
        n = 1 + b % 3
        targets.append({"labels": torch.zeros(n, dtype=torch.int64), "boxes": torch.rand(n, 4, generator=g) * 0.3 + 0.3})
    m = HungarianMatcher(cost_class=1, cost_bbox=5.0, cost_giou=2.0, cost_class_type="ce_cost")
    want = m({"pred_logits": logits, "pred_boxes": boxes}, targets)
    got = m({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda()},
            [{k: v.cuda() for k, v in t.items()} for t in targets])
    for (wq, wt), (gq, gt) in zip(want, got):
        assert gq.is_cuda and wq.tolist() == gq.cpu().tolist() and wt.tolist() == gt.cpu().tolist()


# ------------------------------------------------------------------------------------------------ native head kernels
def _mr(x, ref):
    return ((x.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30)).item()


def test_head_linear_kernels_match_torch(lib):
    """simvgb_head_lin_fwd / _bwd: position add on the first n_split outputs, bias, ReLU, dropout from given uniforms, split-K."""
    from simvg_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1)
    for R, N, Kd, n_split, relu, p in ((64, 768, 256, 512, False, 0.0), (70, 2048, 256, 0, True, 0.1), (640, 256, 256, 256, False, 0.0),
                                       (5, 2, 256, 0, False, 0.0)):
        x = torch.randn(R, Kd, device=DEV, generator=g).requires_grad_(True)
        x2 = torch.randn(R, Kd, device=DEV, generator=g).requires_grad_(True) if n_split else None
        W = (torch.randn(N, Kd, device=DEV, generator=g) * 0.1).requires_grad_(True)
        b = torch.randn(N, device=DEV, generator=g).requires_grad_(True)
        u = torch.rand(R, N, device=DEV, generator=g) if p > 0 else None
        y = K.head_lin_fwd(x.detach(), W.detach(), b.detach(), x2=None if x2 is None else x2.detach(), n_split=n_split, relu=relu,
                           drop_u=u, drop_p=p)
        xin = x.unsqueeze(1) if x2 is None else None
        if x2 is None:
            ref = torch.nn.functional.linear(x, W, b)
        else:
            ref = torch.cat([torch.nn.functional.linear(x + x2, W[:n_split], b[:n_split]),
                             torch.nn.functional.linear(x, W[n_split:], b[n_split:])], 1)
        if relu:
            ref = ref.relu()
        if u is not None:
            ref = ref * (u >= p).float() / (1 - p)
        assert _mr(y, ref) < 1e-5, (R, N, Kd)
        dy = torch.randn(R, N, device=DEV, generator=g)
        ref.backward(dy)
        dx, dx2 = torch.zeros(R, Kd, device=DEV), torch.zeros(R, Kd, device=DEV)
        dW, db = torch.zeros(N, Kd, device=DEV), torch.zeros(N, device=DEV)
        K.head_lin_bwd(dy, x.detach(), W.detach(), y=y, x2=None if x2 is None else x2.detach(), n_split=n_split, relu=relu, drop_u=u,
                       drop_p=p, dx=dx, dx2=None if x2 is None else dx2, dW=dW, db=db)
        assert _mr(dx, x.grad) < 1e-4 and _mr(dW, W.grad) < 1e-4 and _mr(db, b.grad) < 1e-4, (R, N, Kd)
        if x2 is not None:
            assert _mr(dx2, x2.grad) < 1e-4
    # split-K forward (FFN second layer): zeroed output, bias from split 0
    x, W, b = torch.randn(64, 2048, device=DEV, generator=g), torch.randn(256, 2048, device=DEV, generator=g) * 0.05, torch.randn(256, device=DEV, generator=g)
    assert _mr(K.head_lin_fwd(x, W, b, k_splits=8), torch.nn.functional.linear(x, W, b)) < 1e-5


def test_head_lnres_and_small_attention_match_torch(lib):
    from simvg_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(2)
    R, C, p = 70, 256, 0.1
    a = torch.randn(R, C, device=DEV, generator=g).requires_grad_(True)
    b = torch.randn(R, C, device=DEV, generator=g).requires_grad_(True)
    gm, bt = torch.randn(C, device=DEV, generator=g).requires_grad_(True), torch.randn(C, device=DEV, generator=g).requires_grad_(True)
    u = torch.rand(R, C, device=DEV, generator=g)
    mask = (u >= p).float() / (1 - p)
    ref = torch.nn.functional.layer_norm(a + b * mask, (C,), gm, bt, 1e-5)
    y, mean, rstd = K.head_lnres_fwd(a.detach(), b.detach(), gm.detach(), bt.detach(), drop_u=u, drop_p=p)
    assert _mr(y, ref) < 1e-5
    dy = torch.randn(R, C, device=DEV, generator=g)
    ref.backward(dy)
    da, db, dg, dbt = (torch.zeros_like(t) for t in (a, b, gm, bt))
    K.head_lnres_bwd(dy, a.detach(), b.detach(), gm.detach(), mean, rstd, dg, dbt, da=da, db=db, drop_u=u, drop_p=p)
    assert _mr(da, a.grad) < 1e-4 and _mr(db, b.grad) < 1e-4 and _mr(dg, gm.grad) < 1e-4 and _mr(dbt, bt.grad) < 1e-4
    # few-keys attention: packed q|k|v rows (self-attention, nq = nk = 10) and separate projections with a key padding mask (nk = 20)
    B, H, E = 5, 8, 256
    for nq, nk, packed in ((10, 10, True), (1, 1, True), (3, 20, False)):
        if packed:
            qkv = torch.randn(B * nq, 3 * E, device=DEV, generator=g).requires_grad_(True)
            q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
            kpm = None
        else:
            q = torch.randn(B * nq, E, device=DEV, generator=g).requires_grad_(True)
            k = torch.randn(B * nk, E, device=DEV, generator=g).requires_grad_(True)
            v = torch.randn(B * nk, E, device=DEV, generator=g).requires_grad_(True)
            kpm = torch.zeros(B, nk, dtype=torch.uint8, device=DEV)
            kpm[1, 12:] = 1
            kpm[3, 5:] = 1
        u = torch.rand(B, H, nq, nk, device=DEV, generator=g)
        scale = 32 ** -0.5
        qh = q.reshape(B, nq, H, 32).transpose(1, 2)
        kh = k.reshape(B, nk, H, 32).transpose(1, 2)
        vh = v.reshape(B, nk, H, 32).transpose(1, 2)
        sc = (qh @ kh.transpose(-1, -2)) * scale
        if kpm is not None:
            sc = sc.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
        pr = sc.softmax(-1)
        ref = ((pr * (u >= p).float() / (1 - p)) @ vh).transpose(1, 2).reshape(B * nq, E)
        ctx, P = K.head_attn_small_fwd(q.detach(), k.detach(), v.detach(), B, nq, nk, H, scale, kpm=kpm, drop_u=u, drop_p=p)
        assert _mr(ctx, ref) < 1e-5 and _mr(P, pr) < 1e-5, (nq, nk)
        dctx = torch.randn(B * nq, E, device=DEV, generator=g)
        ref.backward(dctx)
        if packed:
            dqkv = torch.zeros_like(qkv)
            K.head_attn_small_bwd(dctx, q.detach(), k.detach(), v.detach(), P, dqkv[:, :E], dqkv[:, E:2 * E], dqkv[:, 2 * E:], B, nq, nk, H,
                                  scale, drop_u=u, drop_p=p)
            assert _mr(dqkv, qkv.grad) < 1e-4, (nq, nk)
        else:
            dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
            K.head_attn_small_bwd(dctx, q.detach(), k.detach(), v.detach(), P, dq, dk, dv, B, nq, nk, H, scale, drop_u=u, drop_p=p)
            assert _mr(dq, q.grad) < 1e-4 and _mr(dk, k.grad) < 1e-4 and _mr(dv, v.grad) < 1e-4, (nq, nk)


@pytest.mark.parametrize("B,nq,N,p", [(3, 1, 1600, 0.0), (2, 10, 100, 0.1), (2, 3, 333, 0.1), (1, 5, 40, 0.0)])
def test_head_absorbed_cross_attention_matches_torch(lib, B, nq, N, p):
    """simvgb_head_xattn (key-parallel passes, no atomics) against nn.MultiheadAttention arithmetic written out in torch with the
    same dropout mask: context, saved probabilities, and every gradient (query, keys, values, both projections) — gradients
    ACCUMULATE into their buffers, as the decoder stack relies on."""
    from simvg_b200 import kernels as K
    E, H = 256, 8
    g = torch.Generator(device="cuda").manual_seed(11)
    mk = lambda *s: torch.randn(*s, device=DEV, generator=g)  # noqa: E731
    R = B * nq
    q, kin, val = mk(R, E), mk(B * N, E), mk(B * N, E)
    Wk, Wv, bk, bv = mk(E, E) * 0.06, mk(E, E) * 0.06, mk(E) * 0.1, mk(E) * 0.1
    kpm = torch.zeros(B, N, dtype=torch.bool, device=DEV)
    kpm[0, N // 3:] = True
    u = torch.rand(R * H * N, device=DEV, generator=g) if p > 0 else None
    dctx = mk(R, E)
    scale = 32 ** -0.5
    leaves = [t.clone().requires_grad_(True) for t in (q, kin, val, Wk, bk, Wv, bv)]
    q_, kin_, val_, Wk_, bk_, Wv_, bv_ = leaves
    kp = (kin_ @ Wk_.t() + bk_).view(B, N, H, 32).permute(0, 2, 1, 3)
    vp = (val_ @ Wv_.t() + bv_).view(B, N, H, 32).permute(0, 2, 1, 3)
    qh = (q_ * scale).view(B, nq, H, 32).permute(0, 2, 1, 3)
    sc = (qh @ kp.transpose(-1, -2)).masked_fill(kpm[:, None, None, :], float("-inf"))
    P_ref = sc.softmax(-1)                                       # [B, H, nq, N]
    Pd = P_ref
    if u is not None:
        m = (u.view(B, nq, H, N).permute(0, 2, 1, 3) >= p).float() / (1 - p)
        Pd = P_ref * m
    ctx_ref = (Pd @ vp).permute(0, 2, 1, 3).reshape(R, E)
    (ctx_ref * dctx).sum().backward()
    ctx, P, z, psum = K.head_xattn_fwd(q, kin, val, Wk, bk, Wv, bv, B, nq, N, kpm=kpm.to(torch.uint8), drop_u=u, drop_p=p)
    assert _mr(ctx, ctx_ref.detach()) < 2e-5
    assert _mr(P.view(B, nq, H, N).permute(0, 2, 1, 3), P_ref.detach()) < 2e-5
    base = 0.5
    outs = [torch.full_like(t, base) for t in (q, kin, val, Wk, bk, Wv, bv)]
    dq, dkin, dval, dWk, dbk, dWv, dbv = outs
    K.head_xattn_bwd(dctx, q, kin, val, Wk, bk, Wv, bv, P, z, psum, B, nq, N, dq, dkin, dval, dWk, dbk, dWv, dbv,
                     kpm=kpm.to(torch.uint8), drop_u=u, drop_p=p)
    for name, got, leaf in zip(("dq", "dkin", "dval", "dWk", "dbk", "dWv", "dbv"), outs, leaves):
        if name == "dbk":      # the key bias shifts every score of a row equally: its gradient is rounding noise around zero
            assert (got - base).abs().max() < 1e-4 and leaf.grad.abs().max() < 1e-4
            continue
        assert _mr(got - base, leaf.grad) < 2e-4, (name, _mr(got - base, leaf.grad))


@pytest.mark.parametrize("nq,N,masked", [(1, 400, False), (10, 100, True), (1, 20, True), (4, 20, False)])
def test_native_decoder_stack_matches_op_by_op_path(lib, nq, N, masked):
    """DetrTransformerDecoder on the fused head kernels (one autograd node) against its own op-by-op PyTorch path (the
    implementation the CPU tests pin to the oracle): stacked outputs, input gradients and EVERY parameter gradient — object-token
    decoder geometry (absorbed cross-attention over N > 32 keys) and text-guided query generation geometry (N = 20 keys)."""
    from simvg_b200.models.heads.tgqs_kd_detr_head.transformer import DetrTransformerDecoder
    torch.manual_seed(3)
    B, E = 4, 256
    dec = DetrTransformerDecoder(embed_dim=E, num_heads=8, attn_dropout=0.1, feedforward_dim=2048 if N > 32 else 512, ffn_dropout=0.1,
                                 num_layers=3, return_intermediate=N > 32, post_norm=True).cuda().eval()
    for p_ in dec.parameters():
        if p_.dim() > 1:
            torch.nn.init.xavier_uniform_(p_)
        else:
            torch.nn.init.normal_(p_, std=0.2)
    g = torch.Generator(device="cuda").manual_seed(4)
    mk = lambda *s: torch.randn(*s, device=DEV, generator=g)  # noqa: E731
    query, qpos, key, kpos = mk(B, nq, E) * 0.0, mk(B, nq, E), mk(B, N, E), mk(B, N, E)
    kpm = None
    if masked:
        kpm = torch.zeros(B, N, dtype=torch.bool, device=DEV)
        kpm[0, N // 2:] = True
        kpm[2, 3:] = True
    res = {}
    for native in (False, True):
        dec.use_native = native
        dec.zero_grad(set_to_none=True)
        q_, qp_, k_ = query.clone().requires_grad_(True), qpos.clone().requires_grad_(True), key.clone().requires_grad_(True)
        out = dec(q_, k_, k_, query_pos=qp_, key_pos=kpos, key_padding_mask=kpm)
        w = torch.randn(out.shape, device=DEV, generator=torch.Generator(device="cuda").manual_seed(5))
        (out * w).sum().backward()
        res[native] = (out.detach(), q_.grad, qp_.grad, k_.grad, {n: p_.grad.clone() for n, p_ in dec.named_parameters()})
    (o0, dq0, dp0, dk0, g0), (o1, dq1, dp1, dk1, g1) = res[False], res[True]
    assert o0.shape == o1.shape and _mr(o1, o0) < 2e-5
    assert _mr(dq1, dq0) < 2e-4 and _mr(dp1, dp0) < 2e-4 and _mr(dk1, dk0) < 2e-4
    for n in g0:
        if g0[n].abs().max() < 1e-6:     # mathematically zero gradients (rounding noise on either path): the self-attention q / k
            assert g1[n].abs().max() < 1e-6, n   # projections when nq = 1 (softmax over one key); all of layer 0's self-attention
            continue                             # input projection here (content queries start at zero)
        assert _mr(g1[n], g0[n]) < 5e-4, (n, _mr(g1[n], g0[n]))
    # training mode (dropout on): finite outputs and gradients
    dec.use_native = True
    dec.train()
    out = dec(query.clone().requires_grad_(True), key, key, query_pos=qpos, key_pos=kpos, key_padding_mask=kpm)
    out.sum().backward()
    assert torch.isfinite(out).all() and all(torch.isfinite(p_.grad).all() for p_ in dec.parameters())
