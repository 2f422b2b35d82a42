"""Acc@0.5 parity on held-out synthetic boxes (north_star): train the CUDA product and the CPU oracle (the reference's
arithmetic) from IDENTICAL weights on IDENTICAL batches, then evaluate both on the same held-out batches.

Task: the box is recoverable from the image — a bright rectangle on N(0, 0.3^2) noise at the ground-truth box; text tokens are
random.  Stochastic layers are off (eval-mode forward with losses, as in smoke()) so the two runs are comparable step by step.
Writes gpurun_out/acc_parity.json.   python tests/acc_parity.py [steps] [batch] [img_size]
"""
import copy
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import simvg_oracle as O  # noqa: E402  (checker: lives under tests/ because only tests may use the oracle)
from simvg_b200.models import build_model  # noqa: E402
from simvg_b200.optim import FusedAdamAMSGrad  # noqa: E402
from tools.synth import make_batch, model_cfg  # noqa: E402


def task_batch(B, S, seed):
    b = make_batch(B, S, seed=seed)
    g = torch.Generator().manual_seed(seed + 77)
    img = 0.3 * torch.randn(B, 3, S, S, generator=g)
    for i, box in enumerate(b["gt_bbox"]):
        x0, y0, x1, y1 = [int(round(float(v))) for v in box]
        img[i, :, y0:y1 + 1, x0:x1 + 1] += 2.0
    b["img"] = img
    return b


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    P, n_eval = 16, 8
    lr, lr_enc, clip = 5e-4, 5e-5, 0.15
    torch.manual_seed(6666)
    cfg = model_cfg("base", S, P, num_decoder_layers=3, drop_path_rate=0.0)
    model = build_model(cfg)
    sd0 = {k: v.detach().clone().float() for k, v in model.state_dict().items()}

    # ---- CUDA product
    model = model.cuda().eval()
    opt = FusedAdamAMSGrad(model, lr=lr, lr_vis_enc=lr_enc, betas=(0.9, 0.98), eps=1e-9, grad_norm_clip=clip)
    # ---- CPU oracle (reference arithmetic) + torch Adam(amsgrad) with the reference's two LR groups and clip
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k) for k, v in sd0.items()}
    params = [v for v in sd.values() if v.requires_grad]
    ref_opt = torch.optim.Adam([{"params": [v for k, v in sd.items() if v.requires_grad and "vis_enc" in k], "lr": lr_enc},
                                {"params": [v for k, v in sd.items() if v.requires_grad and "vis_enc" not in k], "lr": lr}],
                               betas=(0.9, 0.98), eps=1e-9, weight_decay=0, amsgrad=True)
    om = O.OracleModel(sd, "base", S, P, cfg["head"])

    curve = []
    t_gpu = t_cpu = 0.0
    for it in range(steps):
        b = task_batch(B, S, seed=1000 + it)
        t0 = time.perf_counter()
        opt.zero_grad()
        losses, _ = model(b["img"].cuda(), b["ref_expr_inds"].cuda(), b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"].cuda(), gt_bbox=[t.cuda() for t in b["gt_bbox"]])
        losses["loss_total"].backward()
        opt.step()
        lg = float(losses["loss_total"])
        t_gpu += time.perf_counter() - t0
        t0 = time.perf_counter()
        ol, _, _ = om.forward_train(b["img"], b["ref_expr_inds"], copy.deepcopy(b["img_metas"]), b["text_attention_mask"], b["gt_bbox"])
        ref_opt.zero_grad()
        ol["loss_total"].backward()
        torch.nn.utils.clip_grad_norm_(params, clip)
        ref_opt.step()
        lc = float(ol["loss_total"])
        t_cpu += time.perf_counter() - t0
        curve.append((lg, lc))
        if it % 10 == 0 or it == steps - 1:
            print("step %3d  loss cuda %.4f  oracle %.4f  rel %.2e" % (it, lg, lc, abs(lg - lc) / max(abs(lc), 1e-9)), flush=True)

    # ---- held-out evaluation (fixed seeds never seen in training)
    res = {"cuda": {"dec": [], "tok": []}, "oracle": {"dec": [], "tok": []}}
    box_diff = 0.0
    for j in range(n_eval):
        b = task_batch(B, S, seed=900000 + j)
        with torch.no_grad():
            pg = model(b["img"].cuda(), b["ref_expr_inds"].cuda(), b["img_metas"], return_loss=False,
                       text_attention_mask=b["text_attention_mask"].cuda())
        pc, _ = om.forward_test(b["img"], b["ref_expr_inds"], copy.deepcopy(b["img_metas"]), b["text_attention_mask"])
        for name, k in (("dec", 0), ("tok", 1)):
            res["cuda"][name].append(float(O.accuracy_at_05(pg[k]["pred_bboxes"].cpu(), b["gt_bbox"])))
            res["oracle"][name].append(float(O.accuracy_at_05(pc[k]["pred_bboxes"], b["gt_bbox"])))
        box_diff = max(box_diff, float((pg[0]["pred_bboxes"].cpu() - pc[0]["pred_bboxes"]).abs().max()) / S)
    mean = lambda v: sum(v) / len(v)  # noqa: E731
    out = {"steps": steps, "batch": B, "img_size": S, "patch": P, "held_out_images": n_eval * B,
           "acc05_decoder": {"cuda": mean(res["cuda"]["dec"]), "oracle": mean(res["oracle"]["dec"])},
           "acc05_token": {"cuda": mean(res["cuda"]["tok"]), "oracle": mean(res["oracle"]["tok"])},
           "final_loss": {"cuda": curve[-1][0], "oracle": curve[-1][1]},
           "first_loss": {"cuda": curve[0][0], "oracle": curve[0][1]},
           "max_rel_loss_gap": max(abs(a - c) / max(abs(c), 1e-9) for a, c in curve),
           "max_pred_box_gap_frac_of_image": box_diff,
           "seconds": {"cuda_train": t_gpu, "oracle_train": t_cpu}, "host_threads": os.cpu_count()}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/acc_parity.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
