"""Per-shape GEMM device time inside one real train step (cfg2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS
from simvg_b200 import kernels as K
from simvg_b200.models import build_model
from simvg_b200.optim import FusedAdamAMSGrad
from tools.synth import make_batch, model_cfg
vit, S, P, bs, dec, blw = CONFIGS[os.environ.get("CFG", "cfg2")]
torch.manual_seed(6666)
model = build_model(model_cfg(vit, S, P, num_decoder_layers=dec, branch_loss_weight=blw)).cuda().train()
opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5, grad_norm_clip=0.15)
b = make_batch(bs, S, device="cuda")
def step():
    opt.zero_grad()
    losses, _ = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True, text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
    losses["loss_total"].backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
K._prof_detail[0] = True
K.profile_start()
step()
prof = K.profile_stop()
rows = sorted(prof.items(), key=lambda kv: -kv[1][1])
tot = sum(v[1] for k, v in rows if k.startswith("gemm"))
print("total gemm ms %.2f" % tot)
for k, (n, ms, fl) in rows[:40]:
    print("%-62s n=%3d  total %7.3f ms  avg %7.3f ms  %6.0f TF/s" % (k, n, ms, ms / n, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0))
