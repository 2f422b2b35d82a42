"""Yardsticks for the attention kernels (tooling, not product): cuDNN SDPA and flash-attn 2.8 on the same box and the same
problem sizes as simvgb_attn_fwd / simvgb_attn_bwd — B x H heads of L = Lv + Lt tokens, head_dim 64, non-causal, bf16.

    python tools/attn_yardstick.py [--out gpurun_out/attn_yardstick.json]

FLOPs: forward 4 L^2 dh per (b, h), backward 10 L^2 dh (the same accounting bench.py uses).  Library kernels see one dense
[B, H, L, 64] problem without key padding (they do strictly less masking work than ours)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/attn_yardstick.json")
    args = ap.parse_args()
    from simvg_b200 import kernels as K
    shapes = [("cfg2 ViT-B 640", 64, 12, 1601, 20), ("cfg5 ViT-L 640", 64, 16, 1601, 20), ("cfg4 ViT-L 768", 16, 16, 2305, 20)]
    res = []
    for name, B, H, Lv, Lt in shapes:
        L, D = Lv + Lt, H * 64
        f_fwd, f_bwd = 4.0 * B * H * L * L * 64, 10.0 * B * H * L * L * 64
        rec = {"shape": name, "B": B, "H": H, "L": L}
        torch.manual_seed(0)
        # ---- ours
        qkv_v = (torch.randn(B * Lv, 3 * D, device="cuda") * 0.5).bfloat16()
        qkv_t = (torch.randn(B * Lt, 3 * D, device="cuda") * 0.5).bfloat16()
        pad = torch.zeros(B, Lt, dtype=torch.uint8, device="cuda")
        pad[:, 12:] = 1
        o_v, o_t, lse = K.attn_fwd(qkv_v, qkv_t, pad, B, H, Lv, Lt)
        do_v, do_t = torch.randn_like(o_v), torch.randn_like(o_t)
        ws = {}
        ms = time_ms(lambda: K.attn_fwd(qkv_v, qkv_t, pad, B, H, Lv, Lt))
        rec["ours_fwd_ms"], rec["ours_fwd_tflops"] = ms, f_fwd / ms / 1e9
        ms = time_ms(lambda: K.attn_bwd(qkv_v, qkv_t, pad, o_v, o_t, lse, do_v, do_t, B, H, Lv, Lt, ws=ws))
        rec["ours_bwd_ms"], rec["ours_bwd_tflops"] = ms, f_bwd / ms / 1e9
        del qkv_v, qkv_t, o_v, o_t, do_v, do_t, ws
        # ---- library kernels on [B, H, L, 64] / [B, L, H, 64]
        q, k, v = (torch.randn(B, H, L, 64, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for tag, backend in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("torch_flash", SDPBackend.FLASH_ATTENTION)):
            try:
                with sdpa_kernel(backend):
                    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
                    do = torch.randn_like(o)
                    ms = time_ms(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
                    rec[tag + "_fwd_ms"], rec[tag + "_fwd_tflops"] = ms, f_fwd / ms / 1e9
                    ms = time_ms(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
                    rec[tag + "_bwd_ms"], rec[tag + "_bwd_tflops"] = ms, f_bwd / ms / 1e9
            except Exception as e:  # noqa: BLE001
                rec[tag + "_error"] = repr(e)[:200]
        try:
            from flash_attn import flash_attn_func
            q2, k2, v2 = (t.detach().transpose(1, 2).contiguous().requires_grad_(True) for t in (q, k, v))
            o = flash_attn_func(q2, k2, v2)
            do = torch.randn_like(o)
            ms = time_ms(lambda: flash_attn_func(q2, k2, v2))
            rec["fa2_fwd_ms"], rec["fa2_fwd_tflops"] = ms, f_fwd / ms / 1e9
            ms = time_ms(lambda: torch.autograd.grad(o, (q2, k2, v2), do, retain_graph=True))
            rec["fa2_bwd_ms"], rec["fa2_bwd_tflops"] = ms, f_bwd / ms / 1e9
        except Exception as e:  # noqa: BLE001
            rec["fa2_error"] = repr(e)[:200]
        print(json.dumps(rec), flush=True)
        res.append(rec)
        del q, k, v
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
