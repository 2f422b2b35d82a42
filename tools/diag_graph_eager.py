"""Noise floor of the small-model train loop: eager twice, single-graph, chunked graphs (N = 1).  Prints the loss sequences."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simvg_b200.models import build_model  # noqa: E402
from simvg_b200.optim import FusedAdamAMSGrad  # noqa: E402
from simvg_b200.runtime import GraphedTrainStep  # noqa: E402
from tools.synth import make_batch, model_cfg  # noqa: E402


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


def main():
    batches = [make_batch(4, 128, seed=10 + i, device="cuda") for i in range(2)]
    n = int(os.environ.get("STEPS", 5))
    out = {}
    for name in ("eager1", "eager2"):
        m, o = make()
        ls = []
        for it in range(n):
            b = batches[it % 2]
            o.zero_grad()
            losses, _ = m(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
            losses["loss_total"].backward()
            o.step()
            ls.append(float(losses["loss_total"]))
        out[name] = ls
    for name, cl in (("graph", 0), ("chunked2", 2), ("chunked5", 5)):
        m, o = make()
        step = GraphedTrainStep(m, o, warmup=1, chunk_layers=cl)
        ls = []
        for it in range(n):
            b = batches[it % 2]
            losses, _ = step(b["img"], b["ref_expr_inds"], b["img_metas"], b["text_attention_mask"], torch.stack(b["gt_bbox"]))
            ls.append(float(losses["loss_total"]))
        out[name] = ls + [len(step.plan)]
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main()
