"""Acc@0.5 parity on held-out synthetic boxes (north_star): train the CUDA product and the oracle (the reference's arithmetic,
plain eager fp32 PyTorch) from IDENTICAL weights on IDENTICAL batches, then evaluate both on the same held-out batches.

Task: the box is recoverable from the image — a bright rectangle on N(0, 0.3^2) noise at the ground-truth box; text tokens are
random.  Both arms run on the GPU box: the product through its sm_100a kernels, the oracle as fp32 eager ops on cuda:0 with TF32
off (the same code the CPU tests pin against the reference's files; on the device only so that thousands of steps are
affordable).  Stochastic layers are off so the two runs are comparable step by step.  The encoder is the first `layers`
ViT-B/16 layers (the task needs no depth; the oracle's cost scales with it).
    python tests/acc_parity.py [steps] [batch] [img_size] [layers]        -> gpurun_out/acc_parity.json
tests/test_gpu_acc_parity.py asserts on the result (both arms >= 80 %, |dAcc| <= 2 points, final-loss gap <= 5 %).
"""
import copy
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import make_batch, model_cfg  # noqa: E402

LR, LR_ENC, CLIP = 5e-4, 2e-4, 1.0


def task_batch(B, S, seed, device="cpu"):
    b = make_batch(B, S, seed=seed)
    g = torch.Generator().manual_seed(seed + 77)
    img = 0.3 * torch.randn(B, 3, S, S, generator=g)
    for i, box in enumerate(b["gt_bbox"]):
        x0, y0, x1, y1 = [int(round(float(v))) for v in box]
        img[i, :, y0:y1 + 1, x0:x1 + 1] += 2.0
    b["img"] = img
    if device != "cpu":
        for k in ("img", "ref_expr_inds", "text_attention_mask"):
            b[k] = b[k].to(device)
        b["gt_bbox"] = [t.to(device) for t in b["gt_bbox"]]
    return b


def build_arm_model(S, P, layers, seed=6666):
    """-> (cfg, product model on CPU with its encoder truncated to `layers`, fp32 state dict of the same weights)."""
    from simvg_b200.models import build_model
    torch.manual_seed(seed)
    cfg = model_cfg("base", S, P, num_decoder_layers=3, drop_path_rate=0.0)
    model = build_model(cfg)
    enc = model.vis_enc
    enc.beit3.encoder.layers = torch.nn.ModuleList(list(enc.beit3.encoder.layers)[:layers])
    enc.beit3.encoder.num_layers = layers
    enc.cfg["layers"] = layers
    enc.drop_path_probs = enc.drop_path_probs[:layers]
    sd0 = {k: v.detach().clone().float() for k, v in model.state_dict().items()}
    return cfg, model, sd0


def accuracy_at_05(pred, gt_list):
    """apis/test.py:70-79 (mmdet aligned bbox_overlaps, eps 1e-6) batched on whatever device the boxes live on."""
    gt = torch.stack(gt_list).to(pred)
    lt, rb = torch.max(gt[:, :2], pred[:, :2]), torch.min(gt[:, 2:], pred[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    a1 = (gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])
    a2 = (pred[:, 2] - pred[:, 0]) * (pred[:, 3] - pred[:, 1])
    return ((inter / (a1 + a2 - inter).clamp(min=1e-6)) >= 0.5).float().mean() * 100.0


class _RefArm:
    """One reference arm: the oracle's arithmetic (fp32, or with the product's bf16 operand roundings emulated) trained with
    torch Adam(amsgrad) over the reference's lr groups + clip_grad_norm_ + the shared step schedule."""

    def __init__(self, cls, sd0, cfg, S, P, layers, steps, dev):
        self.sd = {k: v.clone().to(dev).requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k) for k, v in sd0.items()}
        self.params = [v for v in self.sd.values() if v.requires_grad]
        self.opt = torch.optim.Adam([{"params": [v for k, v in self.sd.items() if v.requires_grad and "vis_enc" in k], "lr": LR_ENC},
                                     {"params": [v for k, v in self.sd.items() if v.requires_grad and "vis_enc" not in k], "lr": LR}],
                                    betas=(0.9, 0.98), eps=1e-9, weight_decay=0, amsgrad=True)
        self.om = cls(self.sd, "base", S, P, cfg["head"])
        self.om.cfg["layers"] = layers
        self.sched = torch.optim.lr_scheduler.MultiStepLR(self.opt, milestones=[int(0.7 * steps)], gamma=0.1)
        self.seconds = 0.0
        self.curve = []

    def step(self, b):
        t0 = time.perf_counter()
        out = self.om.forward_train(b["img"], b["ref_expr_inds"], copy.deepcopy(b["img_metas"]), b["text_attention_mask"], b["gt_bbox"])
        self.opt.zero_grad()
        out[0]["loss_total"].backward()
        torch.nn.utils.clip_grad_norm_(self.params, CLIP)
        self.opt.step()
        self.sched.step()
        self.curve.append(float(out[0]["loss_total"].detach()))
        torch.cuda.synchronize()
        self.seconds += time.perf_counter() - t0

    def predict(self, b):
        return self.om.forward_test(b["img"], b["ref_expr_inds"], copy.deepcopy(b["img_metas"]), b["text_attention_mask"])[0]


def run(steps=1500, B=32, S=128, layers=2, P=16, n_eval=8, log_every=100, out_path="gpurun_out/acc_parity.json", emulated=True):
    """Arms: `product` (sm_100a kernels), `oracle` (fp32 reference arithmetic), and — to attribute whatever separates those two —
    `oracle_bf16`: the same oracle with the product's bf16 operand roundings emulated in fp32 (oracle/bf16_emulation.py)."""
    from oracle import simvg_oracle as O   # checker: this file lives under tests/ because only tests may use the oracle
    from oracle import bf16_emulation as OE
    from simvg_b200.optim import FusedAdamAMSGrad
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    cfg, model, sd0 = build_arm_model(S, P, layers)
    # ---- CUDA product
    model = model.cuda().eval()
    opt = FusedAdamAMSGrad(model, lr=LR, lr_vis_enc=LR_ENC, betas=(0.9, 0.98), eps=1e-9, grad_norm_clip=CLIP)
    # the reference's step schedule (core/scheduler.py: decay built ON the optimiser) drives every arm: x0.1 for the last 30 % of
    # the steps, so all end converged instead of wherever their (chaotically diverging) trajectories happen to be
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[int(0.7 * steps)], gamma=0.1)
    arms = {"oracle": _RefArm(O.OracleModel, sd0, cfg, S, P, layers, steps, dev)}
    if emulated:
        arms["oracle_bf16"] = _RefArm(OE.OracleModelBF16, sd0, cfg, S, P, layers, steps, dev)

    curve = []
    t_gpu = 0.0
    for it in range(steps):
        b = task_batch(B, S, seed=1000 + it, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad()
        losses, _ = model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=True,
                          text_attention_mask=b["text_attention_mask"], gt_bbox=b["gt_bbox"])
        losses["loss_total"].backward()
        opt.step()
        sched.step()
        curve.append(float(losses["loss_total"].detach()))
        t_gpu += time.perf_counter() - t0
        for arm in arms.values():
            arm.step(b)
        if it % log_every == 0 or it == steps - 1:
            print("step %4d  loss product %.4f  %s" % (it, curve[-1], "  ".join("%s %.4f" % (k, a.curve[-1]) for k, a in arms.items())),
                  flush=True)

    # ---- held-out evaluation (fixed seeds never seen in training)
    names = ["product"] + list(arms)
    res = {n: {"dec": [], "tok": []} for n in names}
    box_diff = 0.0
    for j in range(n_eval):
        b = task_batch(B, S, seed=900000 + j, device=dev)
        with torch.no_grad():
            preds = {"product": model(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=False,
                                      text_attention_mask=b["text_attention_mask"])}
            for k, arm in arms.items():
                preds[k] = arm.predict(b)
        for n in names:
            for name, k in (("dec", 0), ("tok", 1)):
                res[n][name].append(float(accuracy_at_05(preds[n][k]["pred_bboxes"], b["gt_bbox"])))
        box_diff = max(box_diff, float((preds["product"][0]["pred_bboxes"] - preds["oracle"][0]["pred_bboxes"]).abs().max()) / S)
    mean = lambda v: sum(v) / len(v)  # noqa: E731
    curves = {"product": curve, **{k: a.curve for k, a in arms.items()}}
    out = {"steps": steps, "batch": B, "img_size": S, "patch": P, "encoder_layers": layers, "held_out_images": n_eval * B,
           "lr": LR, "lr_vis_enc": LR_ENC, "grad_norm_clip": CLIP,
           "acc05_decoder": {n: mean(res[n]["dec"]) for n in names},
           "acc05_token": {n: mean(res[n]["tok"]) for n in names},
           "final_loss_mean50": {n: mean(curves[n][-50:]) for n in names},
           "loss_before_lr_decay_mean50": {n: mean(curves[n][int(0.7 * steps) - 50:int(0.7 * steps)]) for n in names},
           "first_loss": {n: curves[n][0] for n in names},
           "first_step_rel_loss_gap": abs(curve[0] - curves["oracle"][0]) / abs(curves["oracle"][0]),
           "max_pred_box_gap_frac_of_image": box_diff,
           "seconds": {"product_train": t_gpu, **{k + "_train": a.seconds for k, a in arms.items()}},
           "oracle_device": "cuda:0 eager fp32 (TF32 off)"}
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    run(*a[:4])
