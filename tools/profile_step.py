"""Per-kernel device-time table of one train step (torch.profiler / CUPTI), written to gpurun_out/step_profile.txt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CONFIGS  # noqa: E402
from simvg_b200.models import build_model  # noqa: E402
from simvg_b200.optim import FusedAdamAMSGrad  # noqa: E402
from tools.synth import make_batch, model_cfg  # noqa: E402


def main():
    cfg_name = os.environ.get("CFG", "cfg2")
    vit, S, P, bs, dec, blw = CONFIGS[cfg_name]
    bs = int(os.environ.get("BS", bs))
    torch.manual_seed(6666)
    model = build_model(model_cfg(vit, S, P, num_decoder_layers=dec, branch_loss_weight=blw)).cuda().train()
    opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5, grad_norm_clip=0.15)
    b = make_batch(bs, S, device="cuda")

    def step():
        opt.zero_grad()
        losses, _ = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
        losses["loss_total"].backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if os.environ.get("NOPROF"):   # under ncu: `--profile-from-start off` captures only these steps
        torch.cuda.cudart().cudaProfilerStart()
        for _ in range(int(os.environ.get("STEPS", 1))):
            step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    os.makedirs("gpurun_out", exist_ok=True)
    txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70)
    with open("gpurun_out/step_profile_%s.txt" % cfg_name, "w") as f:
        f.write(txt)
    print(txt[-6000:])


if __name__ == "__main__":
    main()
