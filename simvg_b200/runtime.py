"""Whole-train-step CUDA graphs (B200-first runtime piece; the reference has no counterpart — its step is ~4000 eager
launches driven from Python, /root/reference/simvg/apis/train.py:75-86).

One SimVG train step here is thousands of kernel launches, many of them tiny fp32 ops of the DETR head / DWBD losses whose
launch overhead the CPU cannot hide; capturing them once and replaying removes that host cost.  Everything inside is
step-invariant: inputs are copied into static buffers, the optimiser reads its step-dependent scalars from device memory
(FusedAdamAMSGrad.advance), DropPath draws from the graph-registered CUDA generator, and the model code performs no host
synchronisation on the REC path.

  1 GPU : ONE graph  = zero_grad + forward + losses + backward + clip + Adam(+EMA).
  N GPUs: the backward is cut into CHUNKS of encoder layers, each its own graph:
          [zero_grad + forward + losses + head backward] -> [encoder layers 11..10] -> [9..8] -> ... -> [1] ->
          [0 + embeddings] -> [clip + Adam(+EMA)].  After each chunk's replay the flat gradient range that chunk finished is handed to an
          asynchronous NCCL all-reduce (FlatDDP.reduce_async), which runs under the following chunks' kernels; only the last
          chunk's range (layer 0 + embeddings, ~60 MB for ViT-B) and the all-gather of the text-embedding rows are exposed.
          No collective is captured (capturing them hung in round 1) and every rank replays exactly the launch sequence the
          1-GPU number is quoted on.  `chunk_layers=0` keeps the backward in one graph and exchanges everything after it.
"""
import torch

from simvg_b200 import kernels as K


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, ddp=None); losses, preds = step(img, ids, img_metas, mask, gt_boxes)

    img [B,3,S,S] fp32 (or uint8 [B,S,S,3] when the encoder was given `input_norm`), ids / mask [B,Lt] int64, gt_boxes [B,4]
    (xyxy pixels) — host (ideally pinned) or device tensors.  The first call runs `warmup` forward+backward passes WITHOUT
    an optimiser step (allocator / workspace / NCCL warm-up: parameters, Adam moments and the step count are untouched, so
    the training trajectory equals the eager loop's) and captures; later calls with the same shapes and image sizes replay.
    A new batch shape / image dtype captures another graph (again without touching the optimiser state); graphs are kept per
    shape and share one memory pool, so alternating between shapes does not re-capture.  Returned tensors are the graph's
    static outputs (overwritten by the next call with the same shape)."""

    def __init__(self, model, optimizer, ddp=None, warmup=1, chunk_layers=None):
        self.model, self.opt, self.warmup = model, optimizer, max(1, int(warmup))
        self.ddp = ddp if (ddp is not None and ddp.world > 1) else None
        if self.ddp is not None and not self.ddp.deferred:
            raise ValueError("GraphedTrainStep needs FlatDDP(..., deferred=True): collectives run between the step's graphs")
        enc = getattr(model, "vis_enc", None)
        can_chunk = enc is not None and hasattr(enc, "_defer_backward")
        if chunk_layers is None:
            chunk_layers = 2 if (self.ddp is not None and can_chunk) else 0
        self.chunk_layers = int(chunk_layers) if can_chunk else 0
        self.plan = None         # [(graph, exchange)]: exchange = None | "all" | (lo, hi, heads) handed to the process group after the replay
        self.graph_opt = None    # optimiser graph (when the step is more than one graph)
        self.key = None
        self.static = None
        self.out = None
        self.launches_per_step = 0
        self._cache = {}         # key -> (plan, graph_opt, static, out, launches, text-gradient mailbox, deferred backward)
        self._pool = None
        self._mail = None        # (ids, rows) tensors this graph's backward writes for the sparse text-embedding exchange
        self._eb = None          # the captured EncoderBackward (its buffers live across the chunk graphs)

    @property
    def graph(self):
        return None if self.plan is None else self.plan[0][0]

    def _fwd_bwd(self, d, metas):
        self.opt.zero_grad()
        losses, preds = self.model(d["img"], d["ids"], metas, return_loss=True, text_attention_mask=d["mask"],
                                   gt_bbox=list(d["gt"].unbind(0)), rescale=False)
        losses["loss_total"].backward()
        return losses, preds

    def _chunks(self, nl):
        """[(hi, lo)] top-down; layer 0 (+ embeddings) always last and alone: its exchange is the exposed one."""
        cl = self.chunk_layers
        out = [(hi, max(hi - cl + 1, 1)) for hi in range(nl - 1, 0, -cl)]
        return out + [(0, 0)]

    def _capture(self, img, ids, metas, mask, gt):
        dev = next(self.model.parameters()).device
        self.static = {"img": torch.empty(img.shape, dtype=img.dtype if img.dtype == torch.uint8 else torch.float32, device=dev),
                       "ids": torch.empty(ids.shape, dtype=torch.int64, device=dev),
                       "mask": torch.empty(mask.shape, dtype=torch.int64, device=dev),
                       "gt": torch.empty(gt.shape, dtype=gt.dtype, device=dev)}
        self._upload(img, ids, mask, gt)
        self.opt.enable_graph_mode()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                # forward + backward only: fills the model's host-built caches (position encodings, image-size tensors,
                # attention workspaces, kernel attributes) and warms NCCL without touching parameters or optimiser state
                losses, _ = self._fwd_bwd(self.static, metas)
                if self.ddp is not None:
                    self.ddp.exchange()
                del losses
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # the warm-up's activations sit in the caching allocator's default pool; the capture allocates the same amount again in its
        # private pool (ViT-L at bs 64: 2 x 97 GB > 180 GB) — hand the cached blocks back first
        torch.cuda.empty_cache()
        K.reset_launch_count()
        # a process group's watchdog thread may touch CUDA while we capture: police this thread's calls only
        mode = "thread_local" if self.ddp is not None else "global"
        multi = self.ddp is not None or self.chunk_layers > 0
        # single-graph steps of different input shapes share one memory pool (nothing else runs between replays); a
        # multi-graph step has eager collectives between its graphs, so every shape keeps its own pool (shared by its graphs)
        pool = None if multi else self._pool
        enc = getattr(self.model, "vis_enc", None)
        self.plan, self._eb = [], None

        def capture(body):
            nonlocal pool
            g = torch.cuda.CUDAGraph()
            kw = {} if pool is None else {"pool": pool}
            # same stream as the warm-up: autograd's AccumulateGrad nodes stay on it
            with torch.cuda.graph(g, stream=side, capture_error_mode=mode, **kw):
                res = body()
            if pool is None:
                pool = g.pool()
            return g, res

        def first():
            if self.chunk_layers > 0:
                enc._defer_backward = True
            try:
                res = self._fwd_bwd(self.static, metas)
            finally:
                if self.chunk_layers > 0:
                    enc._defer_backward = False
            if self.chunk_layers > 0:
                self._eb, enc._deferred = enc._deferred, None
                if self._eb is None:
                    raise RuntimeError("the encoder backward did not run inside loss.backward(): nothing to chunk")
                self._eb.start()
            elif not multi:
                self.opt.step()
            return res

        if self.chunk_layers > 0:
            g, (losses, preds) = capture(first)
            self.plan.append((g, (None, None, True)))       # head gradients are final once autograd reached the encoder node
            eb = self._eb
            for hi, lo in self._chunks(eb.nl):
                def body(hi=hi, lo=lo):
                    with torch.no_grad():
                        eb.layers(hi, lo)
                        if lo == 0:
                            eb.finish()
                g, _ = capture(body)
                rng = (eb.layer_range(lo)[0], eb.layer_range(hi)[1]) if lo > 0 else (0, eb.layer_range(0)[1])
                self.plan.append((g, (rng[0], rng[1], False)))
        else:
            g, (losses, preds) = capture(first)
            self.plan.append((g, "all" if self.ddp is not None else None))
        if not multi:
            self._pool = pool
        self.graph_opt = None
        if multi:
            self.graph_opt, _ = capture(self.opt.step)
        self.launches_per_step = K.launch_count()
        self.out = (losses, preds)
        st = getattr(getattr(self.model, "vis_enc", None), "sparse_text_grad", None)
        self._mail = None if st is None else (st.get("ids"), st.get("rows"))

    def _upload(self, img, ids, mask, gt):
        self.static["img"].copy_(img, non_blocking=True)
        self.static["ids"].copy_(ids, non_blocking=True)
        self.static["mask"].copy_(mask, non_blocking=True)
        self.static["gt"].copy_(gt, non_blocking=True)

    def __call__(self, img, ref_expr_inds, img_metas, text_attention_mask, gt_boxes):
        if not self.model.training:
            raise RuntimeError("GraphedTrainStep captures a training step: call model.train() first")
        key = (tuple(img.shape), img.dtype, tuple(ref_expr_inds.shape), tuple(tuple(m["img_shape"][:2]) for m in img_metas))
        if key != self.key:
            if self.key is not None:
                self._cache[self.key] = (self.plan, self.graph_opt, self.static, self.out, self.launches_per_step, self._mail, self._eb)
            self.key = key
            if key in self._cache:
                self.plan, self.graph_opt, self.static, self.out, self.launches_per_step, self._mail, self._eb = self._cache[key]
                st = getattr(getattr(self.model, "vis_enc", None), "sparse_text_grad", None)
                if st is not None and self._mail is not None:
                    st["ids"], st["rows"] = self._mail      # the exchange reads the buffers THIS graph's backward fills
                self._upload(img, ref_expr_inds, text_attention_mask, gt_boxes)
            else:
                self._capture(img, ref_expr_inds, img_metas, text_attention_mask, gt_boxes)
                # the capture pass itself does not execute: fall through and replay once for this batch
        else:
            self._upload(img, ref_expr_inds, text_attention_mask, gt_boxes)
        self.opt.advance()
        for g, ex in self.plan:
            g.replay()
            if self.ddp is not None and ex is not None:
                if ex == "all":
                    self.ddp.exchange()
                else:
                    self.ddp.reduce_async(ex[0], ex[1], heads=ex[2])
        if self.ddp is not None and self.chunk_layers > 0:
            self.ddp.wait()
        if self.graph_opt is not None:
            self.graph_opt.replay()
        return self.out
