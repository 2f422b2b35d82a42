"""Whole-train-step CUDA graph (B200-first runtime piece; the reference has no counterpart — its step is ~4000 eager
launches driven from Python, /root/reference/simvg/apis/train.py:75-83).

One SimVG train step here is ~5000 kernel launches, ~2000 of them tiny fp32 ops of the DETR head / DWBD losses whose
launch overhead the CPU cannot hide; capturing fwd + losses + bwd + gradient all-reduce + clip + Adam once and replaying
the graph removes that host cost.  Everything inside is step-invariant: inputs are copied into static buffers, the
optimiser reads its step-dependent scalars from device memory (FusedAdamAMSGrad.advance), DropPath draws from the
graph-registered CUDA generator, and the model code performs no host synchronisation on the REC path.
"""
import torch

from simvg_b200 import kernels as K


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, ddp=None); losses, preds = step(img, ids, img_metas, mask, gt_boxes)

    img [B,3,S,S] fp32, ids / mask [B,Lt] int64, gt_boxes [B,4] (xyxy pixels) — host (ideally pinned) or device tensors.
    The first call runs `warmup` eager steps' worth of allocator warm-up on a side stream and captures; later calls with
    the same shapes and image sizes replay.  Returned tensors are the graph's static outputs (overwritten by the next call).
    """

    def __init__(self, model, optimizer, ddp=None, warmup=2):
        self.model, self.opt, self.ddp, self.warmup = model, optimizer, ddp, warmup
        self.graph = None
        self.key = None
        self.static = None
        self.out = None
        self.launches_per_step = 0

    def _eager(self, d, metas):
        self.opt.zero_grad()
        losses, preds = self.model(d["img"], d["ids"], metas, return_loss=True, text_attention_mask=d["mask"],
                                   gt_bbox=list(d["gt"].unbind(0)), rescale=False)
        losses["loss_total"].backward()
        if self.ddp is not None:
            self.ddp.finish()
        self.opt.step()
        return losses, preds

    def _capture(self, img, ids, metas, mask, gt):
        dev = next(self.model.parameters()).device
        self.static = {"img": torch.empty(img.shape, dtype=torch.float32, device=dev),
                       "ids": torch.empty(ids.shape, dtype=torch.int64, device=dev),
                       "mask": torch.empty(mask.shape, dtype=torch.int64, device=dev),
                       "gt": torch.empty(gt.shape, dtype=gt.dtype, device=dev)}
        self._upload(img, ids, mask, gt)
        self.opt.enable_graph_mode()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):   # real optimiser steps: allocator / workspace / NCCL warm-up before capture
                self.opt.advance()
                self._eager(self.static, metas)
            if self.warmup == 0:
                # one forward + backward without an optimiser step: fills the model's host-built caches (position
                # encodings, image-size tensors, attention workspaces, kernel attributes) without touching the parameters
                self.opt.zero_grad()
                losses, _ = self.model(self.static["img"], self.static["ids"], metas, return_loss=True,
                                       text_attention_mask=self.static["mask"], gt_bbox=list(self.static["gt"].unbind(0)))
                losses["loss_total"].backward()
                if self.ddp is not None:
                    self.ddp.finish()
                del losses
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        K.reset_launch_count()
        # With a process group, NCCL's watchdog thread polls CUDA events while we capture: only this thread's calls may be
        # policed ("thread_local"), otherwise its cudaEventQuery invalidates the capture.  (Round 1: NCCL capture is still
        # experimental — bench.py keeps N > 1 on eager launches unless --graph-ddp is given.)
        mode = "thread_local" if self.ddp is not None else "global"
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode=mode):   # same stream as the warm-up: autograd's AccumulateGrad nodes stay on it
            losses, preds = self._eager(self.static, metas)
        self.launches_per_step = K.launch_count()
        self.out = (losses, preds)

    def _upload(self, img, ids, mask, gt):
        self.static["img"].copy_(img, non_blocking=True)
        self.static["ids"].copy_(ids, non_blocking=True)
        self.static["mask"].copy_(mask, non_blocking=True)
        self.static["gt"].copy_(gt, non_blocking=True)

    def __call__(self, img, ref_expr_inds, img_metas, text_attention_mask, gt_boxes):
        if not self.model.training:
            raise RuntimeError("GraphedTrainStep captures a training step: call model.train() first")
        key = (tuple(img.shape), tuple(ref_expr_inds.shape), tuple(tuple(m["img_shape"][:2]) for m in img_metas))
        if self.graph is None or key != self.key:
            self.key = key
            self._capture(img, ref_expr_inds, img_metas, text_attention_mask, gt_boxes)
            # the capture pass itself does not execute: fall through and replay once for this batch
        else:
            self._upload(img, ref_expr_inds, text_attention_mask, gt_boxes)
        self.opt.advance()
        self.graph.replay()
        return self.out
