"""Bisect the product's convergence on the Acc@0.5 task (tests/acc_parity.py): native head kernels on / off x fused / torch Adam.
Product arm only (27 s per 1500 steps).  Tooling."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from acc_parity import CLIP, LR, LR_ENC, build_arm_model, task_batch  # noqa: E402
from simvg_b200.models.heads.tgqs_kd_detr_head import native as nat  # noqa: E402
from simvg_b200.optim import FusedAdamAMSGrad  # noqa: E402

S, P, B, LAYERS = 128, 16, 32, 2
STEPS = int(os.environ.get("STEPS", 1500))


def run(native_head, fused, seed_shift=0):
    nat._linear_native[0] = native_head
    cfg, model, _ = build_arm_model(S, P, LAYERS)
    model = model.cuda().eval()
    for mod in model.modules():
        if hasattr(mod, "use_native"):
            mod.use_native = native_head
    if fused:
        opt = FusedAdamAMSGrad(model, lr=LR, lr_vis_enc=LR_ENC, betas=(0.9, 0.98), eps=1e-9, grad_norm_clip=CLIP)
    else:
        named = list(model.named_parameters())
        opt = torch.optim.Adam([{"params": [p for n, p in named if "vis_enc" in n], "lr": LR_ENC},
                                {"params": [p for n, p in named if "vis_enc" not in n], "lr": LR}],
                               betas=(0.9, 0.98), eps=1e-9, weight_decay=0, amsgrad=True)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[int(0.7 * STEPS)], gamma=0.1)
    curve = []
    for it in range(STEPS):
        b = task_batch(B, S, seed=1000 + it + seed_shift, device="cuda")
        opt.zero_grad()
        losses, _ = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
        losses["loss_total"].backward()
        if not fused:
            torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], CLIP)
        opt.step()
        sched.step()
        curve.append(float(losses["loss_total"].detach()))
    m = int(0.7 * STEPS)
    mean = lambda v: sum(v) / len(v)  # noqa: E731
    print("native_head=%s fused_adam=%s : loss@250 %.3f  @500 %.3f  before-decay %.3f  final %.3f"
          % (native_head, fused, mean(curve[225:275]), mean(curve[475:525]), mean(curve[m - 50:m]), mean(curve[-50:])), flush=True)


if __name__ == "__main__":
    for nh, fu in ((True, True), (False, True), (True, False), (False, False)):
        run(nh, fu)
