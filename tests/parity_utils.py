"""Shared helpers of the -m gpu model-level parity tests: build product + oracles on identical weights, run one train step on
each, and reduce the comparison to per-tensor relative L2 errors (not cosines: a 5 % scale error must fail)."""
import copy

import torch


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build_product(vit, S, P, seed, dec_layers=3, blw=None, nq=1, enc_layers=None):
    """-> (model_cfg, CUDA model in eval(), CPU fp32 state dict).  enc_layers truncates the encoder (ViT-L at test cost)."""
    from simvg_b200.models import build_model
    from tools.synth import model_cfg, synth_state_dict
    cfg = model_cfg(vit, S, P, num_decoder_layers=dec_layers, branch_loss_weight=blw, num_queries=nq)
    model = build_model(cfg)
    if enc_layers is not None:
        enc = model.vis_enc
        enc.beit3.encoder.layers = torch.nn.ModuleList(list(enc.beit3.encoder.layers)[:enc_layers])
        enc.beit3.encoder.num_layers = enc_layers
        enc.cfg["layers"] = enc_layers
        enc.drop_path_probs = enc.drop_path_probs[:enc_layers]
    sd = synth_state_dict({k: v.float() for k, v in model.state_dict().items()}, seed=seed)
    model.load_state_dict(sd)
    return cfg, model.cuda().eval(), sd


def oracle_step(sd, cfg, vit, S, P, batch, emulate_bf16=False, enc_layers=None, backward=True):
    """One CPU oracle train step -> (losses, preds, {param name: grad})."""
    from oracle import bf16_emulation as E
    from oracle import simvg_oracle as O
    osd = {k: v.clone().requires_grad_(backward and v.dtype.is_floating_point and "empty_weight" not in k) for k, v in sd.items()}
    om = (E.OracleModelBF16 if emulate_bf16 else O.OracleModel)(osd, vit, S, P, cfg["head"])
    if enc_layers is not None:
        om.cfg["layers"] = enc_layers
    with torch.set_grad_enabled(backward):
        losses, preds, out = om.forward_train(batch["img"], batch["ref_expr_inds"], copy.deepcopy(batch["img_metas"]),
                                              batch["text_attention_mask"], batch["gt_bbox"])
        if backward:
            losses["loss_total"].backward()
    grads = {k: v.grad for k, v in osd.items() if v.requires_grad and v.grad is not None}
    return losses, preds, grads, om


def product_step(model, batch, backward=True):
    for p in model.parameters():
        p.grad = None
    with torch.set_grad_enabled(backward):
        losses, preds = model(batch["img"], batch["ref_expr_inds"], batch["img_metas"], return_loss=True,
                              text_attention_mask=batch["text_attention_mask"], gt_bbox=batch["gt_bbox"])
        if backward:
            losses["loss_total"].backward()
    torch.cuda.synchronize()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    return losses, preds, grads


def grad_errors(got, want, floor=1e-6):
    """Per-tensor relative L2 error for every parameter whose reference gradient norm exceeds `floor` times the largest
    gradient norm in the model.  Skipped as carrying no information: gradients that are MATHEMATICALLY zero and therefore pure
    rounding noise in every implementation — the encoder's key-projection biases (a constant added to all keys of a row shifts
    every score of that row equally and softmax is shift-invariant) — and parameters the graph does not use.
    -> sorted [(err, name, ref_norm)], n_skipped."""
    top = max(float(g.norm()) for g in want.values())
    out, skipped = [], 0
    for n, og in want.items():
        nrm = float(og.norm())
        if nrm <= floor * top or (".k_proj." in n and n.endswith(".bias")):
            skipped += 1
            continue
        g = got.get(n)
        if g is None:
            out.append((1.0, n, nrm))
            continue
        out.append((rel_l2(g, og), n, nrm))
    out.sort(reverse=True)
    return out, skipped


def loss_errors(got, want):
    return {k: abs(float(got[k]) - float(want[k])) / max(abs(float(want[k])), 1e-6) for k in want}


def summarize(errs):
    v = sorted(e for e, _, _ in errs)
    return {"n": len(v), "max": v[-1], "p99": v[int(0.99 * (len(v) - 1))], "median": v[len(v) // 2], "worst": [(round(e, 5), n) for e, n, _ in errs[:6]]}
