"""-m gpu: the registry-built model (CUDA kernels behind BEIT3 / TextGuidedQuerySelectKDDETRHead / MIXDETRMB) against the
CPU oracle and the committed golden vectors, on identical seeded weights and inputs.

north_star tolerance: model outputs (losses, predicted boxes) within 1e-3 relative of the reference path.  Hidden features
carry bf16 operand rounding through 12 layers (measured ~5e-3 relative L2) and gradients ~2-4 %; those looser bounds are
asserted explicitly so a regression in either direction is visible."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _build(S, P, seed, blw=None, nq=1):
    from simvg_b200.models import build_model
    from tools.synth import model_cfg, synth_state_dict
    cfg = model_cfg("base", S, P, num_decoder_layers=3, branch_loss_weight=blw, num_queries=nq)
    model = build_model(cfg)
    sd = synth_state_dict({k: v.float() for k, v in model.state_dict().items()}, seed=seed)
    model.load_state_dict(sd)
    return cfg, model.cuda().eval(), sd


def test_cfg1_against_golden(lib, golden_dir):
    """BASELINE configs[0] (ViT-B/16, 224x224, bs=2): CUDA path vs the reference's own output (golden fixture)."""
    from tools.synth import make_batch
    fx = torch.load(os.path.join(golden_dir, "cfg1_train_step.pt"), weights_only=False)
    cfg, model, _ = _build(fx["S"], fx["P"], fx["weight_seed"])
    b = make_batch(fx["B"], fx["S"], seed=fx["batch_seed"], device="cuda")
    losses, preds = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
    for k, v in fx["losses"].items():
        assert abs(float(losses[k]) - v) <= 1e-3 * max(abs(v), 1e-6), (k, float(losses[k]), v)
    S = fx["S"]
    assert (preds[0]["pred_bboxes"].cpu() - fx["pred_dec"]).abs().max().item() < 1e-3 * S
    assert (preds[1]["pred_bboxes"].cpu() - fx["pred_tok"]).abs().max().item() < 1e-3 * S
    with torch.no_grad():
        img_f, txt_f, cls_f = model.vis_enc(b["img"], b["ref_expr_inds"], b["text_attention_mask"])
    assert rel(img_f[:, ::28, ::64], fx["img_feat_slice"]) < 2e-2
    assert rel(txt_f[:, :, ::64], fx["text_feat_slice"]) < 2e-2
    assert rel(cls_f, fx["cls_feat"]) < 2e-2
    losses["loss_total"].backward()
    got = {n: float(p.grad.norm()) for n, p in model.named_parameters() if p.grad is not None}
    bad = [(k, got.get(k, 0.0), n) for k, n in fx["grad_norms"].items() if n > 1e-4 and abs(got.get(k, 0.0) - n) > 0.1 * n]
    assert len(bad) <= 5, bad[:10]


@pytest.mark.parametrize("S,P,blw,nq", [(128, 32, None, 1), (64, 16, {"decoder": 1.0}, 1), (96, 32, None, 10)])
def test_train_step_against_oracle(lib, S, P, blw, nq):
    from oracle import simvg_oracle as O
    from tools.synth import make_batch
    cfg, model, sd = _build(S, P, seed=3, blw=blw, nq=nq)
    B = 3
    b = make_batch(B, S, seed=12, device="cuda")
    c = make_batch(B, S, seed=12, device="cpu")
    osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k) for k, v in sd.items()}
    om = O.OracleModel(osd, "base", S, P, cfg["head"])
    ol, op, _ = om.forward_train(c["img"], c["ref_expr_inds"], copy.deepcopy(c["img_metas"]), c["text_attention_mask"], c["gt_bbox"])
    ol["loss_total"].backward()
    losses, preds = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
    losses["loss_total"].backward()
    assert set(losses) == set(ol)
    for k in ol:
        assert abs(float(losses[k]) - float(ol[k])) <= 1e-3 * max(abs(float(ol[k])), 1e-6), (k, float(losses[k]), float(ol[k]))
    assert (preds[0]["pred_bboxes"].cpu() - op[0]["pred_bboxes"]).abs().max().item() < 1e-3 * S
    if op[1]["pred_bboxes"] is not None:
        assert (preds[1]["pred_bboxes"].cpu() - op[1]["pred_bboxes"]).abs().max().item() < 1e-3 * S
    # gradients: direction and magnitude
    cos, n_checked = [], 0
    for n, p in model.named_parameters():
        og = osd[n].grad
        if og is None or og.norm() < 1e-7:
            continue
        g = p.grad.detach().float().cpu()
        cos.append((torch.nn.functional.cosine_similarity(g.flatten(), og.flatten(), dim=0).item(), n, float(og.norm())))
        n_checked += 1
    cos.sort()
    med, p5 = cos[len(cos) // 2][0], cos[len(cos) // 20][0]
    # nq=10 on a 3x3-patch image averages over far fewer tokens, so bf16 operand noise shows up more in the gradients
    lim_med, lim_p5 = (0.995, 0.97) if nq == 1 else (0.99, 0.95)
    assert n_checked > 300 and med > lim_med and p5 > lim_p5, (n_checked, cos[:6], med, p5)
    # parameters unused by the graph keep exactly-zero gradients
    assert float(model.vis_enc.beit3.vision_embed.mask_token.grad.abs().sum()) == 0


def test_inference_and_accuracy_metric(lib):
    from oracle import simvg_oracle as O
    from tools.synth import make_batch
    cfg, model, sd = _build(128, 32, seed=4)
    b = make_batch(4, 128, seed=5, device="cuda")
    c = make_batch(4, 128, seed=5, device="cpu")
    preds = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=False, text_attention_mask=b["text_attention_mask"])
    om = O.OracleModel(sd, "base", 128, 32, cfg["head"])
    op, _ = om.forward_test(c["img"], c["ref_expr_inds"], copy.deepcopy(c["img_metas"]), c["text_attention_mask"])
    for got, want in zip(preds, op):
        assert set(got) == {"pred_bboxes", "pred_masks", "predict_classes"}
        assert (got["pred_bboxes"].cpu() - want["pred_bboxes"]).abs().max().item() < 1e-3 * 128
        assert torch.equal(got["predict_classes"].cpu(), want["predict_classes"])
    a1 = O.accuracy_at_05(preds[0]["pred_bboxes"].cpu(), c["gt_bbox"])
    a2 = O.accuracy_at_05(op[0]["pred_bboxes"], c["gt_bbox"])
    assert float(a1) == float(a2)


def test_training_reduces_loss_and_droppath_runs(lib):
    """A few fused-optimiser steps in train() mode (DropPath + decoder dropout active) on a fixed batch."""
    from simvg_b200.models import build_model
    from simvg_b200.optim import FusedAdamAMSGrad
    from tools.synth import make_batch, model_cfg
    torch.manual_seed(0)
    model = build_model(model_cfg("base", 128, 32)).cuda().train()
    opt = FusedAdamAMSGrad(model, lr=2e-4, lr_vis_enc=2e-5, grad_norm_clip=0.15)
    b = make_batch(8, 128, seed=1, device="cuda")
    hist = []
    for _ in range(12):
        opt.zero_grad()
        losses, _ = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
        losses["loss_total"].backward()
        opt.step()
        hist.append(float(losses["loss_total"]))
    assert all(h == h for h in hist) and min(hist[-3:]) < hist[0], hist
    sd = opt.state_dict()
    assert len(sd["param_groups"]) == 3 and sd["param_groups"][0]["lr"] == 2e-5 and sd["param_groups"][2]["lr"] == 2e-4


def test_graphed_train_step_matches_eager(lib):
    """GraphedTrainStep (whole step captured in one CUDA graph, Adam scalars read from the device) reproduces the eager
    train loop: same losses step by step and the same parameters afterwards (stochastic layers off so runs are comparable)."""
    from simvg_b200.models import build_model
    from simvg_b200.optim import FusedAdamAMSGrad
    from simvg_b200.runtime import GraphedTrainStep
    from tools.synth import make_batch, model_cfg

    def make():
        torch.manual_seed(5)
        m = build_model(model_cfg("base", 128, 32, drop_path_rate=0.0)).cuda().train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
            if hasattr(mod, "attn_drop") and isinstance(mod.attn_drop, float):
                mod.attn_drop = 0.0
        return m, FusedAdamAMSGrad(m, lr=2e-4, lr_vis_enc=2e-5, grad_norm_clip=0.15)

    batches = [make_batch(4, 128, seed=10 + i, device="cuda") for i in range(2)]
    m1, o1 = make()
    eager = []
    for it in range(5):
        b = batches[it % 2]
        o1.zero_grad()
        losses, _ = m1(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                       text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
        losses["loss_total"].backward()
        o1.step()
        eager.append(float(losses["loss_total"]))
    m2, o2 = make()
    step = GraphedTrainStep(m2, o2, warmup=1)
    graphed = []
    for it in range(5):
        b = batches[it % 2]
        losses, preds = step(b["img"], b["ref_expr_inds"], b["img_metas"], b["text_attention_mask"], torch.stack(b["gt_bbox"]))
        graphed.append(float(losses["loss_total"]))
    assert step.launches_per_step > 100
    assert o2.t == o1.t == 5
    # Same kernels, same data: the first step is identical and the second differs by accumulation order only.  Later steps drift
    # apart chaotically on this tiny problem (two EAGER runs differ by up to 2.5e-3 at step 4, tools/diag_graph_eager.py), so
    # they are only bracketed.
    for i, (a, g) in enumerate(zip(eager, graphed)):
        assert abs(a - g) <= (2e-4 if i < 2 else 2e-2) * abs(a), (eager, graphed)
    # Parameters: two runs differ by fp32 atomic ordering.  Adam turns gradients that are mathematically ZERO (every key-
    # projection bias: softmax is invariant to a constant added to all scores) into +-lr random walks of pure rounding noise,
    # so those are excluded; everything else must agree closely.
    names = [n for n, _ in m1.named_parameters()]
    keep = [i for i, n in enumerate(names) if "k_proj" not in n and "in_proj_bias" not in n]
    l1, l2 = list(m1.parameters()), list(m2.parameters())
    p1 = torch.cat([l1[i].detach().flatten() for i in keep])
    p2 = torch.cat([l2[i].detach().flatten() for i in keep])
    assert rel(p2, p1) < 2e-3, rel(p2, p1)
    assert preds[0]["pred_bboxes"].shape == (4, 4)


@pytest.mark.gpu
def test_chunked_backward_graphs_match_the_single_graph_step():
    """The multi-GPU runtime's step (backward cut into per-layer-chunk graphs + a separate optimiser graph) run on ONE GPU, where
    no collective sits between the chunks: same losses and parameters as the single-graph step."""
    from simvg_b200.models import build_model
    from simvg_b200.optim import FusedAdamAMSGrad
    from simvg_b200.runtime import GraphedTrainStep
    from tools.synth import make_batch, model_cfg

    def run(chunk_layers):
        torch.manual_seed(5)
        m = build_model(model_cfg("base", 128, 32, drop_path_rate=0.0)).cuda().train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
            if hasattr(mod, "attn_drop") and isinstance(mod.attn_drop, float):
                mod.attn_drop = 0.0
        o = FusedAdamAMSGrad(m, lr=2e-4, lr_vis_enc=2e-5, grad_norm_clip=0.15)
        step = GraphedTrainStep(m, o, warmup=1, chunk_layers=chunk_layers)
        batches = [make_batch(4, 128, seed=10 + i, device="cuda") for i in range(2)]
        ls = []
        for it in range(3):
            b = batches[it % 2]
            losses, _ = step(b["img"], b["ref_expr_inds"], b["img_metas"], b["text_attention_mask"], torch.stack(b["gt_bbox"]))
            ls.append(float(losses["loss_total"]))
        names = [n for n, _ in m.named_parameters()]
        keep = [p.detach().flatten() for n, p in m.named_parameters() if "k_proj" not in n and "in_proj_bias" not in n]
        return ls, torch.cat(keep), len(step.plan), step.graph_opt is not None, names

    l0, p0, n0, opt0, _ = run(0)
    l2, p2, n2, opt2, _ = run(2)
    assert n0 == 1 and not opt0
    assert n2 == 1 + 6 + 1 and opt2          # forward/head graph + chunks (11,10) .. (1,1) + (0,0); the optimiser graph is separate
    for i, (a, b) in enumerate(zip(l0, l2)):
        assert abs(a - b) <= (2e-4 if i < 2 else 2e-2) * abs(a), (l0, l2)
    assert float((p0 - p2).norm() / p0.norm()) < 2e-3
