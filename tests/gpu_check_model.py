"""GPU check: full model (CUDA kernels) vs the CPU oracle on identical weights/inputs — outputs, losses, gradients."""
import copy
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import simvg_oracle as O  # noqa: E402
from simvg_b200.models import build_model  # noqa: E402
from tools.synth import make_batch, model_cfg  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    S, P, B = int(os.environ.get("S", 224)), int(os.environ.get("P", 16)), int(os.environ.get("B", 2))
    torch.manual_seed(6666)
    cfg = model_cfg("base", S, P, num_decoder_layers=3)
    model = build_model(cfg)
    # non-trivial LayerNorm / bias values so every gradient path is exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.02)
        model.vis_enc.beit3.vision_embed.cls_token.normal_(0, 0.02)
    sd_cpu = {k: v.detach().clone().float() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    batch = make_batch(B, S, device="cuda")
    cpu_batch = make_batch(B, S, device="cpu")

    # ---- oracle
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k) for k, v in sd_cpu.items()}
    om = O.OracleModel(sd, "base", S, P, cfg["head"])
    t0 = time.time()
    o_losses, o_preds, o_out = om.forward_train(cpu_batch["img"], cpu_batch["ref_expr_inds"], copy.deepcopy(cpu_batch["img_metas"]),
                                                cpu_batch["text_attention_mask"], cpu_batch["gt_bbox"])
    o_losses["loss_total"].backward()
    print("oracle fwd+bwd %.1fs" % (time.time() - t0), flush=True)

    # ---- product
    losses, preds = model(batch["img"], batch["ref_expr_inds"], batch["img_metas"], return_loss=True,
                          text_attention_mask=batch["text_attention_mask"], gt_bbox=batch["gt_bbox"])
    losses["loss_total"].backward()
    torch.cuda.synchronize()
    res = {}
    with torch.no_grad():
        f_img, f_txt, f_cls = model.vis_enc(batch["img"], batch["ref_expr_inds"], batch["text_attention_mask"])
        o_img, o_txt, o_cls = O.beit3_forward(sd, om.cfg, cpu_batch["img"], cpu_batch["ref_expr_inds"],
                                              cpu_batch["text_attention_mask"], prefix="vis_enc.")
    res["enc_img"] = rel(f_img, o_img)
    res["enc_txt"] = rel(f_txt, o_txt)
    res["enc_cls"] = rel(f_cls, o_cls)
    for k in o_losses:
        res["loss:" + k] = abs(losses[k].item() - o_losses[k].item()) / max(abs(o_losses[k].item()), 1e-12)
    res["pred_dec"] = rel(preds[0]["pred_bboxes"], o_preds[0]["pred_bboxes"])
    res["pred_tok"] = rel(preds[1]["pred_bboxes"], o_preds[1]["pred_bboxes"])
    grads = {n: p.grad for n, p in model.named_parameters()}
    worst = []
    for n, g in grads.items():
        og = sd[n].grad
        if og is None and g is None:
            continue
        if og is None:
            og = torch.zeros_like(sd[n])
        if g is None:
            g = torch.zeros_like(og)
        if og.norm() == 0 and g.float().norm() == 0:
            continue
        worst.append((rel(g, og), n, og.norm().item()))
    worst.sort(reverse=True)
    res["grad_worst"] = worst[:12]
    res["grad_median"] = sorted(w[0] for w in worst)[len(worst) // 2]
    for k, v in res.items():
        print(k, v, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/model_check.json", "w") as f:
        json.dump(res, f, indent=1, default=str)


if __name__ == "__main__":
    main()
