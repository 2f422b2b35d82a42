#!/usr/bin/env python
"""bench.py — images/sec of SimVG's full train step (fwd + DWBD losses + bwd + grad all-reduce + clip + Adam-amsgrad).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu] [--config cfg2|cfg3|cfg4|cfg5|ref32]

Workload (BASELINE.json configs[1], the one `metric` is quoted on): ViT-B/16 BEiT-3 encoder + 6-layer object-token decoder
(as BASELINE.json names it — the head's constructor default, tgqs_kd_detr_head.py:24-48; the reference's shipped RefCOCO
configs set 3: --dec-layers 3), 640x640, bs=64 per GPU, synthetic RefCOCO-shaped inputs, random-init weights.
One "step" = one optimiser step on one batch.

  value  : whole-job img/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e    : the same metric through the public plugin API with HOST (pinned) inputs: every step copies its inputs
           host->device (prefetched on a copy stream) and reads the loss back device->host.  Default e2e input = the image as
           the dataset pipeline holds it before `Normalize` (uint8 HWC; normalise + transpose run fused in the patch-embed
           prologue on the GPU: 1 byte / sample over PCIe); `e2e_fp32` = the reference's collated float NCHW input.
  roofline: dominant kernel family (by summed device time inside a profiled step): algorithmic FLOPs / CUDA-event time
  cpu_baseline: the CPU oracle (oracle/simvg_oracle.py, kind "port") timed on this host on a bounded sample
  --impl reference: the reference's CPU-eager path (the same oracle port; the reference itself cannot be imported on the
           GPU box — its third-party deps are absent) on all host threads, same metric/config.
  --impl reference-gpu: the same oracle (plain eager fp32 PyTorch, allow_tf32=False) on cuda:0 at the largest batch that
           fits — BASELINE.md §4's "reference on GPU" baseline, the one that says something about kernel quality.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INPUT_KINDS = {"img_u8": "uint8 HWC image (pre-Normalize); normalise + transpose fused into the patch-embed prologue on the GPU",
               "img": "float32 NCHW image, normalised on the host (the reference's collated input)"}
TRAFFIC_ATTN_CFG2 = (797.8e6 + 167.1e6 + 966.7e6 + 594.6e6) / 2   # DRAM read + write of one fwd and one bwd launch (profiles/r02_kernels_ncu.md)

CONFIGS = {
    #        vit      img  patch  bs  dec_layers branch_loss_weight
    "cfg2": ("base", 640, 16, 64, 6, {"decoder": 1.0, "balanced_distill": {"token": 2.0, "distill": 1.0}}),
    "cfg3": ("large", 640, 16, 32, 3, {"decoder": 1.0, "balanced_distill": {"token": 1.0, "distill": 0.4}}),
    "cfg4": ("large", 768, 16, 16, 3, {"decoder": 1.0, "balanced_distill": {"token": 1.0, "distill": 0.4}}),
    "cfg5": ("large", 640, 16, 64, 3, {"decoder": 1.0}),
    "ref32": ("base", 640, 32, 64, 3, {"decoder": 1.0, "balanced_distill": {"token": 2.0, "distill": 1.0}}),
}


def step_gflops_per_image(vit, S, P, dec_layers):
    """2MNK per GEMM, 4 L^2 D per attention layer, train step = 3 x forward."""
    D, L = (768, 12) if vit == "base" else (1024, 24)
    N = (S // P) ** 2
    Lx = N + 1 + 20
    enc = 2 * N * 3 * P * P * D + L * (24 * Lx * D * D + 4 * Lx * Lx * D)
    head = 2 * N * D * 256 + dec_layers * 4 * N * 256 * 256
    return 3 * (enc + head) / 1e9


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, cfg_name):
    """The reference's CPU-eager train step (oracle port) on the host cores; bounded sample per step."""
    import copy

    import torch

    from oracle import simvg_oracle as O
    from simvg_b200.models import build_model
    from tools.synth import make_batch, model_cfg
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    vit, S, P, bs, dec, blw = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_bs = args.ref_batch
    torch.manual_seed(6666)
    mcfg = model_cfg(vit, S, P, num_decoder_layers=dec, branch_loss_weight=blw)
    model = build_model(mcfg)   # parameter container only (CPU tensors); the arithmetic below is the oracle's
    sd = {k: v.detach().clone().float().requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k)
          for k, v in model.state_dict().items()}
    del model
    params = [v for v in sd.values() if v.requires_grad]
    vis = [v for k, v in sd.items() if v.requires_grad and "vis_enc" in k]
    rest = [v for k, v in sd.items() if v.requires_grad and "vis_enc" not in k]
    opt = torch.optim.Adam([{"params": vis, "lr": 5e-5}, {"params": rest, "lr": 5e-4}], betas=(0.9, 0.98), eps=1e-9,
                           weight_decay=0, amsgrad=True)
    om = O.OracleModel(sd, vit, S, P, mcfg["head"])
    batch = make_batch(sample_bs, S, seed=6666)

    def step():
        losses, _, _ = om.forward_train(batch["img"], batch["ref_expr_inds"], copy.deepcopy(batch["img_metas"]),
                                        batch["text_attention_mask"], batch["gt_bbox"])
        opt.zero_grad()
        losses["loss_total"].backward()
        torch.nn.utils.clip_grad_norm_(params, 0.15)
        opt.step()
        return float(losses["loss_total"])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ips = sample_bs * args.steps / dt
    line = {
        "impl": "reference", "metric": "images/sec (train step, %dpx, bs=%d/GPU)" % (S, bs), "value": ips, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: ViT-%s/%d BEiT-3 + %d-layer decoder, %dx%d, train step (CPU sample bs=%d)" % (cfg_name, vit, P, dec, S, S, sample_bs)},
        "cpu_baseline": {"value": ips, "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": "oracle/simvg_oracle.py train step (fwd+loss+bwd+clip+Adam-amsgrad) at bs=%d, fp32 eager, %d threads" % (sample_bs, cores)},
        "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_reference_gpu(args, cfg_name):
    """The oracle (the reference's eager fp32 op sequence) on ONE B200, TF32 off, largest batch that fits (halved on OOM)."""
    import copy

    import torch

    from oracle import simvg_oracle as O
    from simvg_b200.models import build_model
    from tools.synth import make_batch, model_cfg
    if int(os.environ.get("RANK", 0)) != 0:
        return 0
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    vit, S, P, bs, dec, blw = CONFIGS[cfg_name]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.manual_seed(6666)
    mcfg = model_cfg(vit, S, P, num_decoder_layers=dec, branch_loss_weight=blw)
    model = build_model(mcfg)
    sd = {k: v.detach().clone().float().to(dev).requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k)
          for k, v in model.state_dict().items()}
    del model
    params = [v for v in sd.values() if v.requires_grad]
    vis = [v for k, v in sd.items() if v.requires_grad and "vis_enc" in k]
    rest = [v for k, v in sd.items() if v.requires_grad and "vis_enc" not in k]
    opt = torch.optim.Adam([{"params": vis, "lr": 5e-5}, {"params": rest, "lr": 5e-4}], betas=(0.9, 0.98), eps=1e-9,
                           weight_decay=0, amsgrad=True)
    om = O.OracleModel(sd, vit, S, P, mcfg["head"])
    use_bs = args.batch or bs
    while True:
        try:
            batch = make_batch(use_bs, S, seed=6666, device=dev)

            def step():
                losses, _, _ = om.forward_train(batch["img"], batch["ref_expr_inds"], copy.deepcopy(batch["img_metas"]),
                                                batch["text_attention_mask"], batch["gt_bbox"])
                opt.zero_grad()
                losses["loss_total"].backward()
                torch.nn.utils.clip_grad_norm_(params, 0.15)
                opt.step()

            for _ in range(max(args.warmup, 1)):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            break
        except torch.cuda.OutOfMemoryError:
            opt.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
            use_bs //= 2
            if use_bs < 1:
                raise
    ms = e0.elapsed_time(e1) / args.steps
    ips = use_bs / (ms * 1e-3)
    gf = step_gflops_per_image(vit, S, P, dec)
    print(json.dumps({
        "impl": "reference-gpu", "metric": "images/sec (train step, %dpx, bs=%d/GPU)" % (S, bs), "value": ips, "unit": "img/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "dtype": "f32", "data": "synthetic", "batch_used": use_bs,
        "config": {"workload": "%s: oracle (reference op sequence, eager fp32 PyTorch, TF32 off) on one B200, bs=%d" % (cfg_name, use_bs)},
        "step_tflops": ips * gf / 1e3}))
    return 0


def cpu_baseline_sample(cfg_name, budget_bs=1, iters=1):
    """Rank-0 bounded CPU sample of the same workload via the oracle (10-30 s)."""
    import copy

    import torch

    from oracle import simvg_oracle as O
    from simvg_b200.models import build_model
    from tools.synth import make_batch, model_cfg
    vit, S, P, bs, dec, blw = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(6666)
    mcfg = model_cfg(vit, S, P, num_decoder_layers=dec, branch_loss_weight=blw)
    model = build_model(mcfg)
    sd = {k: v.detach().clone().float().requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k)
          for k, v in model.state_dict().items()}
    del model
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=5e-4, betas=(0.9, 0.98), eps=1e-9, amsgrad=True)
    om = O.OracleModel(sd, vit, S, P, mcfg["head"])
    batch = make_batch(budget_bs, S, seed=6666)

    def step():
        losses, _, _ = om.forward_train(batch["img"], batch["ref_expr_inds"], copy.deepcopy(batch["img_metas"]),
                                        batch["text_attention_mask"], batch["gt_bbox"])
        opt.zero_grad()
        losses["loss_total"].backward()
        torch.nn.utils.clip_grad_norm_(params, 0.15)
        opt.step()

    step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = time.perf_counter() - t0
    return {"value": budget_bs * iters / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": "oracle train step at bs=%d x %d iters after 1 warm-up, fp32 eager, %d threads" % (budget_bs, iters, cores)}


def run_ours(args, cfg_name):
    import torch
    import torch.distributed as dist

    from simvg_b200 import kernels as K
    from simvg_b200.models import build_model
    from simvg_b200.optim import FlatDDP, FusedAdamAMSGrad
    from tools.synth import make_batch, model_cfg

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on stdout: keep stdout to the one JSON line
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vit, S, P, bs, dec, blw = CONFIGS[cfg_name]
    if args.batch:
        bs = args.batch
    if args.dec_layers:
        dec = args.dec_layers
    K.L.check(K.L.lib().simvgb_device_check(local), "device_check")
    torch.manual_seed(6666)
    mcfg = model_cfg(vit, S, P, num_decoder_layers=dec, branch_loss_weight=blw)
    model = build_model(mcfg).to(dev)
    model.train()
    opt = FusedAdamAMSGrad(model, lr=5e-4, lr_vis_enc=5e-5, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.0, grad_norm_clip=0.15)
    use_graph = not args.no_graph
    ddp = FlatDDP(model, opt, deferred=use_graph)   # graph runtime: collectives run between the two step graphs
    ddp.broadcast_parameters()

    # host batches (pinned) — two alternating synthetic batches, data seed = 6666 + rank
    host = []
    for i in range(2):
        b = make_batch(bs, S, seed=6666 + rank + 1000 * i)
        b["img"] = b["img"].pin_memory()
        b["ref_expr_inds"] = b["ref_expr_inds"].pin_memory()
        b["text_attention_mask"] = b["text_attention_mask"].pin_memory()
        b["gt_box_t"] = torch.stack(b["gt_bbox"]).pin_memory()
        # the same batch as the dataset pipeline holds it before Normalize: uint8 HWC (cv2 channel order)
        b["img_u8"] = torch.randint(0, 256, (bs, S, S, 3), dtype=torch.uint8,
                                    generator=torch.Generator().manual_seed(6666 + rank + 1000 * i)).pin_memory()
        host.append(b)

    def batch_bytes(img_key):
        return sum(host[0][k].numel() * host[0][k].element_size() for k in (img_key, "ref_expr_inds", "text_attention_mask", "gt_box_t"))

    copy_stream = torch.cuda.Stream()

    def upload(b, img_key="img"):
        with torch.cuda.stream(copy_stream):
            d = {k: b[k].to(dev, non_blocking=True) for k in ("ref_expr_inds", "text_attention_mask", "gt_box_t")}
            d["img"] = b[img_key].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d, ev

    def train_step(d, metas):
        gt = list(d["gt_box_t"].unbind(0))
        opt.zero_grad()
        losses, _preds = model(d["img"], d["ref_expr_inds"], metas, return_loss=True,
                               text_attention_mask=d["text_attention_mask"], gt_bbox=gt, rescale=False)
        losses["loss_total"].backward()
        ddp.finish()
        opt.step()
        return losses["loss_total"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Whole-step CUDA graph (simvg_b200/runtime.py): the public train-step entry point; --no-graph times the eager loop.
    gstep = None
    # Every N replays the same captured launch sequence: one graph at N = 1; at N > 1 the backward is cut into per-layer-chunk
    # graphs and each chunk's gradient range is all-reduced under the following chunks (no collective is captured; runtime.py).
    if use_graph:
        from simvg_b200.runtime import GraphedTrainStep
        gstep = GraphedTrainStep(model, opt, ddp if world > 1 else None, warmup=2,
                                 chunk_layers=None if args.chunk_layers < 0 else args.chunk_layers)

    def graphed_step(d, metas):
        losses, _preds = gstep(d["img"], d["ref_expr_inds"], metas, d["text_attention_mask"], d["gt_box_t"])
        return losses["loss_total"]

    def timed(n, e2e, img_key="img"):
        """-> ms per step (device time via CUDA events, max over ranks)."""
        res, ev = upload(host[0], img_key)
        torch.cuda.current_stream().wait_event(ev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss_host = 0.0
        for i in range(n):
            cur = res
            if e2e:
                nxt, ev = upload(host[(i + 1) % 2], img_key)      # this step's H2D copy, overlapped on the copy stream
            # graph mode: the staged device batch is copied (device->device) into the graph's static inputs, then replayed
            loss = (graphed_step if gstep is not None else train_step)(cur, host[i % 2]["img_metas"])
            if e2e:
                loss_host = float(loss)                  # D2H read of the step's result
                torch.cuda.current_stream().wait_event(ev)
                res = nxt
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, loss_host

    # roofline leg FIRST, before any graph exists (at ViT-L bs 64 an eager step and a step graph's private pool do not fit side by
    # side in 180 GB): one profiled EAGER step with CUDA events around every GEMM / attention launch (events cannot bracket
    # kernels inside a graph replay); same kernels, same shapes as the replayed step
    res, ev = upload(host[0])
    torch.cuda.current_stream().wait_event(ev)
    for _ in range(2):
        train_step(res, host[0]["img_metas"])
    torch.cuda.synchronize()
    K.profile_start()
    train_step(res, host[0]["img_metas"])
    prof = K.profile_stop()
    barrier()
    del res
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    graphed = gstep is not None

    # Input kind of the timed runs: the batch as the dataset pipeline holds it before `Normalize` (uint8 HWC; normalise +
    # transpose fused into the patch-embed prologue).  `value` has it resident in HBM, `e2e` copies it from pinned host memory
    # every step.  At N = 1 the reference's collated float32 NCHW input is timed as well (`e2e_fp32`).
    kind = "img" if args.fp32_input else "img_u8"
    # warm-up (also warms the caching allocator; in graph mode the first call of each input kind captures)
    timed(max(args.warmup, 3), e2e=False, img_key=kind)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K.reset_launch_count()
    ms, _ = timed(args.steps, e2e=False, img_key=kind)
    launches = K.launch_count() if gstep is None else gstep.launches_per_step * args.steps
    timed(max(args.warmup, 3), e2e=True, img_key=kind)      # untimed: first use of the copy stream / pinned staging buffers
    ms_e2e, last_loss = timed(args.steps, e2e=True, img_key=kind)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e_f32 = None
    # a second input kind is a second capture, whose eager warm-up needs a step's activations next to the first graph's pool:
    # only when that fits (cfg2: 36 + 36 GB; ViT-L at bs 64: 97 + 97 GB does not)
    torch.cuda.empty_cache()
    if world == 1 and kind == "img_u8" and torch.cuda.mem_get_info()[0] > 0.95 * torch.cuda.max_memory_allocated():
        timed(max(args.warmup, 3), e2e=True, img_key="img")
        ms_e2e_f32, _ = timed(args.steps, e2e=True, img_key="img")

    # attention forward + backward are one kernel family for the "dominant family" pick (bench.py used to split them, which let
    # the GEMM family win while attention as a whole was the larger share)
    if "attn_fwd" in prof and "attn_bwd" in prof:
        a, b2 = prof["attn_fwd"], prof["attn_bwd"]
        prof["attention"] = (a[0] + b2[0], a[1] + b2[1], a[2] + b2[2])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured bf16_tflops_sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PF sustained (B200_PROFILING.md)"
    fam = max((k for k in prof if k not in ("attn_fwd", "attn_bwd")), key=lambda k: prof[k][1]) if prof else None
    roofline = None
    if fam:
        n, tot_ms, fl = prof[fam]
        ach = fl / (tot_ms * 1e-3) / 1e12
        # DRAM traffic of the family's largest launch from the committed ncu --set full capture (profiles/r01_gemm_ncu.md):
        # QKV forward GEMM pair (vision 102464x2304x768 + text 1280x2304x768): 597.5 MB read+write vs 644.5 MB algorithmic.
        traffic, note = None, None
        if cfg_name == "cfg2" and fam == "gemm":
            traffic, note = 597.5e6, "bytes/launch of the largest launch (QKV fwd GEMM, both experts), ncu capture in profiles/r01_gemm_ncu.md; algorithmic 644.5e6"
        elif cfg_name == "cfg2" and fam == "attention":
            traffic, note = TRAFFIC_ATTN_CFG2, "mean DRAM bytes/launch over the family (one forward + one backward launch per layer), ncu --set full captures in profiles/r02_kernels_ncu.md; algorithmic 637e6 (fwd) / 1274e6 (bwd)"
        roofline = {"bound": "tensor", "kernel": fam, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "traffic": traffic, "traffic_note": note,
                    "launches": n, "avg_ms": tot_ms / n, "peak_source": peak_src,
                    "families": {k: {"launches": v[0], "ms": v[1], "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None}
                                 for k, v in prof.items()}}
    gf = step_gflops_per_image(vit, S, P, dec)
    ips = world * bs / (ms * 1e-3)
    ips_e2e = world * bs / (ms_e2e * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_sample(cfg_name, budget_bs=args.ref_batch, iters=1)
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "img/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    line = {
        "metric": "images/sec (train step, %dpx, bs=%d/GPU)" % (S, bs), "value": ips, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "%s: ViT-%s/%d BEiT-3 multiway encoder + %d-layer object-token decoder + DWBD losses, %dx%d, bs=%d/GPU, "
                               "full train step (fwd+loss+bwd+allreduce+clip+Adam-amsgrad)" % (cfg_name, vit, P, dec, S, S, bs),
                   "global_batch": world * bs, "seq_len": (S // P) ** 2 + 21, "parallelism": "dp%d" % world,
                   "input": INPUT_KINDS[kind],
                   "l2": "inputs+activations per step (>40 GB) far exceed the 126 MB L2; no explicit flush",
                   "operands": "bf16 GEMM/attention operands, fp32 accumulate, fp32 residual stream / LN / softmax / optimiser",
                   "launch": ("eager launches" if not graphed else
                              "whole step replayed as one CUDA graph (simvg_b200.runtime.GraphedTrainStep)" if world == 1 else
                              "CUDA graphs per step: fwd+head-bwd | encoder backward in %d-layer chunks | clip+Adam; each chunk's "
                              "gradient range is all-reduced (NCCL, asynchronous) under the following chunks"
                              % (2 if args.chunk_layers < 0 else args.chunk_layers) if (args.chunk_layers != 0) else
                              "two CUDA graphs per step (fwd+bwd | clip+Adam) with the NCCL gradient exchange between them")},
        "clocks": clocks,
        "e2e": {"value": ips_e2e, "unit": "img/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": batch_bytes(kind),
                "d2h_bytes_per_step": 4, "last_loss": last_loss, "input": INPUT_KINDS[kind]},
        "e2e_fp32": None if ms_e2e_f32 is None else {
            "value": world * bs / (ms_e2e_f32 * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e_f32,
            "h2d_bytes_per_step": batch_bytes("img"), "d2h_bytes_per_step": 4, "input": INPUT_KINDS["img"]},
        "gpu_launches": launches,
        "step_gflops_per_image": gf,
        "step_tflops": ips * gf / 1e3,   # whole job
        "step_frac_of_peak": ips * gf / 1e3 / peak_tf / world,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dec-layers", type=int, default=0)
    ap.add_argument("--ref-batch", type=int, default=2, help="images per CPU-reference step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fp32-input", action="store_true", help="time value / e2e on the reference's float32 NCHW input instead of uint8 HWC")
    ap.add_argument("--chunk-layers", type=int, default=-1,
                    help="N > 1 graph runtime: encoder layers per backward chunk graph (0 = one backward graph, exchange after it)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager launch loop instead of the whole-step CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, args.config)
    if args.impl == "reference-gpu":
        return run_reference_gpu(args, args.config)
    return run_ours(args, args.config)


if __name__ == "__main__":
    sys.exit(main())
