"""Run-to-run gradient noise of the small model (one GPU): two identical fresh runs, native head on / off; worst parameters."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simvg_b200.models import build_model  # noqa: E402
from simvg_b200.optim import FusedAdamAMSGrad  # noqa: E402
from tools.synth import make_batch, model_cfg  # noqa: E402


def run(native, train=True):
    from simvg_b200.models.heads.tgqs_kd_detr_head import native as nat
    nat._linear_native[0] = native
    torch.manual_seed(11)
    m = build_model(model_cfg("base", 128, 32, drop_path_rate=0.0)).cuda()
    m = m.train() if train else m.eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
        if hasattr(mod, "attn_drop") and isinstance(mod.attn_drop, float):
            mod.attn_drop = 0.0
        if hasattr(mod, "use_native"):
            mod.use_native = native
    o = FusedAdamAMSGrad(m, lr=2e-4, lr_vis_enc=2e-5, grad_norm_clip=0.15)
    b = make_batch(3, 128, seed=50, device="cuda")
    o.zero_grad()
    losses, _ = m(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                  text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
    losses["loss_total"].backward()
    torch.cuda.synchronize()
    return float(losses["loss_total"]), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}, \
        [s.fb.grad.clone() for s in o.segments]


def cmp(tag, a, b):
    la, ga, sa = a
    lb, gb, sb = b
    print(tag, "loss", la, lb)
    for i, (x, y) in enumerate(zip(sa, sb)):
        if x.numel():
            print("  segment %d: max|d| / max|g| = %.3e   (max|g| %.3e)" % (i, float((x - y).abs().max() / y.abs().max()), float(y.abs().max())))
    worst = sorted(((float((ga[n] - gb[n]).abs().max()), float(gb[n].abs().max()), n) for n in ga), reverse=True)[:8]
    for d, mx, n in worst:
        print("    %-70s |d| %.3e  max|g| %.3e" % (n[-70:], d, mx))


def main():
    a, b = run(True), run(True)
    cmp("native vs native", a, b)
    c, d = run(False), run(False)
    cmp("torch-head vs torch-head", c, d)
    cmp("native vs torch-head", a, c)


if __name__ == "__main__":
    main()
