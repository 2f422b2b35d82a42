"""Tooling: find a synthetic-box training setup the PRODUCT learns quickly (used to size tests/acc_parity.py).
python tools/acc_search.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from acc_parity import task_batch, build_arm_model, accuracy_at_05  # noqa: E402
from simvg_b200.optim import FusedAdamAMSGrad  # noqa: E402

S, P, B = 128, 16, 32
for layers, lr, lr_enc, clip, steps in ((2, 5e-4, 2e-4, 1.0, 2000), (2, 5e-4, 5e-5, 0.15, 2000), (4, 1e-3, 5e-4, 1.0, 2000)):
    cfg, model, _ = build_arm_model(S, P, layers)
    model = model.cuda().eval()
    opt = FusedAdamAMSGrad(model, lr=lr, lr_vis_enc=lr_enc, betas=(0.9, 0.98), eps=1e-9, grad_norm_clip=clip)
    t0 = time.time()
    for it in range(steps):
        b = task_batch(B, S, seed=1000 + it, device="cuda")
        opt.zero_grad()
        losses, _ = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True, text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
        losses["loss_total"].backward()
        opt.step()
        if it % 250 == 249:
            acc = [[], []]
            with torch.no_grad():
                for j in range(4):
                    e = task_batch(B, S, seed=900000 + j, device="cuda")
                    pg = model(e["img"], e["ref_expr_inds"], e["img_metas"], return_loss=False, text_attention_mask=e["text_attention_mask"])
                    for k in range(2):
                        acc[k].append(float(accuracy_at_05(pg[k]["pred_bboxes"], e["gt_bbox"])))
            print("layers %d lr %.0e/%.0e clip %.2f step %4d loss %.3f acc dec %.1f tok %.1f (%.0fs)" % (layers, lr, lr_enc, clip, it + 1, float(losses["loss_total"]), sum(acc[0]) / 4, sum(acc[1]) / 4, time.time() - t0), flush=True)
