"""Diagnostic (not a test): gradient error distributions of the ViT-L 768-geometry case over several batch seeds, and of the
two oracles against each other.  python tests/gpu_parity_diag.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_utils import build_product, grad_errors, loss_errors, oracle_step, product_step, summarize  # noqa: E402
from tools.synth import make_batch  # noqa: E402

vit, S, P, B = "large", 768, 16, 2
cfg, model, sd = build_product(vit, S, P, seed=7, dec_layers=3, enc_layers=2)
for seed in (21, 22, 23):
    gb, cb = make_batch(B, S, seed=seed, device="cuda"), make_batch(B, S, seed=seed, device="cpu")
    losses, preds, grads = product_step(model, gb)
    ol, op, og, _ = oracle_step(sd, cfg, vit, S, P, cb, emulate_bf16=False, enc_layers=2)
    el, ep, eg, _ = oracle_step(sd, cfg, vit, S, P, cb, emulate_bf16=True, enc_layers=2)
    for tag, a, b_ in (("product vs fp32", grads, og), ("product vs emu", grads, eg), ("emu vs fp32", eg, og)):
        errs, _ = grad_errors(a, b_)
        s = summarize(errs)
        print("seed %d %-16s median %.4f p99 %.4f max %.4f  worst %s" % (seed, tag, s["median"], s["p99"], s["max"], s["worst"][:2]), flush=True)
    print("   loss err vs emu", max(loss_errors(losses, el).values()), "l1 margins:",
          [float((preds[i]["pred_bboxes"].cpu() - ep[i]["pred_bboxes"]).abs().max()) for i in range(2)], flush=True)
