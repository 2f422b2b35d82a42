"""Fused global-norm clip + Adam(amsgrad) (+ EMA) over flat parameter segments, and the data-parallel gradient exchange.

Replaces, inside the timed train step (/root/reference/simvg/apis/train.py:75-86):
  * torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.grad_norm_clip)   (:81-82)
  * torch.optim.Adam(amsgrad=True) built with three learning-rate groups by parameter-name substring
    (/root/reference/tools/train.py:78-93; hyper-parameters configs/.../refcoco_onestage.py:107-123)
  * ExponentialMovingAverage.update_params (/root/reference/simvg/models/utils.py:148-173, called at apis/train.py:85-86)
  * MMDistributedDataParallel's bucketed gradient all-reduce (tools/train.py:102-103)
by: one sum-of-squares pass + one update pass per trainable range (simvgb_sumsq / simvgb_adam_amsgrad, 36 B per parameter,
44 B with the EMA shadow), and NCCL collectives over contiguous ranges of the flat gradient buffer.

`FusedAdamAMSGrad` IS a torch.optim.Optimizer: `param_groups` is the reference's own layout — three groups
[vis_enc, lan_enc (empty in every SimVG config), everything else], parameters in `named_parameters()` order, frozen
parameters left out — so the reference's schedulers (LambdaLR / CosineAnnealingLR built on the optimizer,
/root/reference/simvg/core/scheduler.py) and its checkpoint code (utils/checkpoint.py:117-148: optimizer.state_dict())
work unchanged, and a checkpoint written by the reference's torch.optim.Adam loads into it (and vice versa).
"""
import torch
import torch.distributed as dist

from simvg_b200 import kernels as K
from simvg_b200.flat import FlatBuffer


class _Segment:
    """One flat fp32 buffer of parameters that share a learning-rate group."""

    def __init__(self, name, fb, group):
        self.name, self.fb, self.group = name, fb, group
        self.m = self.v = self.vmax = self.ema = None
        self._ranges = None

    def ensure_state(self):
        d = self.fb.data
        if self.m is None or self.m.device != d.device:
            self.m, self.v, self.vmax = torch.zeros_like(d), torch.zeros_like(d), torch.zeros_like(d)

    def ranges(self):
        """Contiguous [lo, hi) element ranges covering the trainable parameters (frozen ones — BEIT3(freeze_layer=k),
        beit3.py:78-90 — are neither clipped nor updated, exactly as if they were absent from the reference's groups)."""
        key = tuple(p.requires_grad for p in self.fb.params)
        if self._ranges is None or self._ranges[0] != key:
            out, fb = [], self.fb
            for i, p in enumerate(fb.params):
                if not p.requires_grad:
                    continue
                lo = fb.offsets[i]
                hi = fb.offsets[i + 1] if i + 1 < len(fb.offsets) else fb.numel
                if out and out[-1][1] == lo:
                    out[-1][1] = hi
                else:
                    out.append([lo, hi])
            self._ranges = (key, [tuple(r) for r in out])
        return self._ranges[1]


class FusedAdamAMSGrad(torch.optim.Optimizer):
    def __init__(self, model, lr, lr_vis_enc=None, lr_lan_enc=None, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.0,
                 amsgrad=True, grad_norm_clip=None, ema_alpha=None):
        assert amsgrad, "SimVG trains with Adam(amsgrad=True)"
        named = list(model.named_parameters())
        lr_v = lr if lr_vis_enc is None else lr_vis_enc
        lr_l = lr if lr_lan_enc is None else lr_lan_enc
        # the reference's three groups (tools/train.py:78-93), same order, same membership tests
        groups = [
            {"params": [p for n, p in named if "vis_enc" in n and p.requires_grad], "lr": lr_v},
            {"params": [p for n, p in named if "lan_enc" in n and p.requires_grad], "lr": lr_l},
            {"params": [p for n, p in named if "lan_enc" not in n and "vis_enc" not in n and p.requires_grad], "lr": lr},
        ]
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=True))
        self.betas, self.eps, self.weight_decay, self.grad_norm_clip = betas, eps, weight_decay, grad_norm_clip
        self.t = 0
        self.ema_alpha, self.ema_t = ema_alpha, 0
        self.segments = []
        enc = getattr(model, "vis_enc", None)
        enc_ids = set()
        if enc is not None and hasattr(enc, "flat"):
            fb = enc.flat()
            enc_ids = {id(p) for p in fb.params}
            self.segments.append(_Segment("vis_enc", fb, 0))
        rest = [(n, p) for n, p in named if id(p) not in enc_ids]
        vis = [(n, p) for n, p in rest if "vis_enc" in n]
        lan = [(n, p) for n, p in rest if "lan_enc" in n and "vis_enc" not in n]
        other = [(n, p) for n, p in rest if "lan_enc" not in n and "vis_enc" not in n]
        for name, members, gi in (("vis_enc_rest", vis, 0), ("lan_enc", lan, 1), ("rest", other, 2)):
            if members:
                self.segments.append(_Segment(name, FlatBuffer(members), gi))
        self._where = {}
        for s in self.segments:
            s.fb.ensure()
            s.fb.attach_grads()
            for i, p in enumerate(s.fb.params):
                self._where[id(p)] = (s, i)
        self._sumsq = None
        self._hyper = None        # device [n_segments, 4]: lr, 1 - beta1^t, sqrt(1 - beta2^t), ema decay  (graph mode)
        self.graph_mode = False
        if ema_alpha is not None:
            self.enable_ema(ema_alpha)

    # ---- learning rates: the scheduler writes param_groups[i]["lr"]; every launch reads it from there
    def _lr(self, seg):
        return float(self.param_groups[seg.group]["lr"])

    # ---- EMA of the weights, updated inside the Adam pass
    def enable_ema(self, alpha):
        """Shadow weights start as a copy of the current parameters (ExponentialMovingAverage.__init__, models/utils.py:136-145)."""
        self.ema_alpha = float(alpha)
        for s in self.segments:
            s.fb.ensure()
            s.ema = s.fb.data.clone()

    def ema_decay(self, t=None):
        t = self.ema_t if t is None else t
        return min(self.ema_alpha, (t + 1.0) / (t + 10.0))       # models/utils.py:149

    def swap_ema(self):
        """Exchange the live parameters with their averages in place (evaluation with the EMA weights: the reference's
        `model_ema.apply_shadow()` ... `model_ema.restore()`, tools/train.py:133-138); call again to swap back.  The bf16 operand
        shadows of the encoder follow the parameters on its next forward (it re-casts from the flat buffer every call)."""
        if self.ema_alpha is None:
            raise RuntimeError("EMA is not enabled (ema_alpha=None)")
        with torch.no_grad():
            for s in self.segments:
                tmp = s.fb.data.clone()
                s.fb.data.copy_(s.ema)
                s.ema.copy_(tmp)

    def ema_view(self, p):
        """The EMA shadow of parameter `p` (a view of the flat shadow buffer)."""
        s, i = self._where[id(p)]
        return s.fb.view(i, s.ema)

    # ---- CUDA-graph support: the step-dependent scalars live in device memory and are advanced BY THE GRAPH ITSELF
    # (beta^t is a running product updated by a tiny in-graph op), so replays need no per-step host->device traffic and the
    # host may run any number of steps ahead of the device.
    def enable_graph_mode(self):
        """Idempotent: captured graphs bake the ADDRESSES of these scalars in, so they are created once and never replaced."""
        if self.graph_mode and self._hyper is not None:
            return
        dev = self.segments[0].fb.data.device
        n = len(self.segments)
        self._hyper = torch.zeros(n, 4, device=dev, dtype=torch.float32)
        self._betas_dev = torch.tensor([self.betas[0], self.betas[1]], device=dev, dtype=torch.float64)
        self._pow = torch.tensor([self.betas[0] ** self.t, self.betas[1] ** self.t], device=dev, dtype=torch.float64)
        self._lr_host = [self._lr(s) for s in self.segments]
        self._lr_dev = torch.tensor(self._lr_host, device=dev, dtype=torch.float32)
        self._ema_t_dev = torch.tensor([float(self.ema_t)], device=dev, dtype=torch.float32)
        self.graph_mode = True
        for s in self.segments:
            s.ensure_state()
        if self._sumsq is None or self._sumsq.device != dev:
            self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)

    def advance(self):
        """Graph mode, once per step before the replay: host-side step count; learning rates are re-uploaded only when a
        scheduler changed them (a blocking 8-byte copy, outside the graph)."""
        self.t += 1
        if self.ema_alpha is not None:
            self.ema_t += 1
        lrs = [self._lr(s) for s in self.segments]
        if lrs != self._lr_host:
            self._lr_host = lrs
            self._lr_dev.copy_(torch.tensor(lrs, dtype=torch.float32))

    def zero_grad(self, set_to_none=False):
        """Always keeps the flat gradient buffers (one memset per segment); `set_to_none` is accepted for API compatibility."""
        for s in self.segments:
            s.fb.ensure()
            if s.fb.grad is None:
                s.fb.attach_grads()
            s.fb.zero_grad()

    def _prepare(self):
        """Re-establishes p.grad -> flat-buffer aliasing if someone replaced the gradients (model.zero_grad(), or any
        zero_grad(set_to_none=True) not routed through this optimiser): foreign p.grad tensors are copied in, so the
        step never silently works on stale zeros."""
        for s in self.segments:
            s.fb.ensure()
            s.fb.attach_grads()
            s.ensure_state()
        dev = self.segments[0].fb.data.device
        if self._sumsq is None or self._sumsq.device != dev:
            self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        with K.nvtx("optimizer.clip+adam_amsgrad"):
            return self._step(loss)

    def _step(self, loss):
        if self.graph_mode:
            self._step_graph()
            return loss
        self._prepare()
        self.t += 1
        clip = float(self.grad_norm_clip) if self.grad_norm_clip else 0.0
        if clip > 0:
            self._sumsq.zero_()
            for s in self.segments:
                for lo, hi in s.ranges():
                    K.sumsq(s.fb.grad[lo:hi], self._sumsq)
        decay = self.ema_decay() if self.ema_alpha is not None else 0.0
        for s in self.segments:
            for lo, hi in s.ranges():
                K.adam_amsgrad(s.fb.data[lo:hi], s.fb.grad[lo:hi], s.m[lo:hi], s.v[lo:hi], s.vmax[lo:hi], self._lr(s),
                               self.betas[0], self.betas[1], self.eps, self.weight_decay, self.t,
                               grad_sumsq=self._sumsq if clip > 0 else None, max_norm=clip,
                               ema=None if s.ema is None else s.ema[lo:hi], ema_decay=decay)
        if self.ema_alpha is not None:
            self.ema_t += 1
        return loss

    def _step_graph(self):
        """Same update with step-invariant launch arguments (advance() must have been called for this step)."""
        clip = float(self.grad_norm_clip) if self.grad_norm_clip else 0.0
        self._pow.mul_(self._betas_dev)                                   # beta^t (fp64 running product, on the device)
        self._hyper[:, 0] = self._lr_dev
        self._hyper[:, 1] = (1.0 - self._pow[0]).float()
        self._hyper[:, 2] = (1.0 - self._pow[1]).sqrt().float()
        if self.ema_alpha is not None:
            self._hyper[:, 3] = torch.clamp((self._ema_t_dev + 1.0) / (self._ema_t_dev + 10.0), max=self.ema_alpha)
            self._ema_t_dev.add_(1.0)
        if clip > 0:
            self._sumsq.zero_()
            for s in self.segments:
                for lo, hi in s.ranges():
                    K.sumsq(s.fb.grad[lo:hi], self._sumsq)
        for i, s in enumerate(self.segments):
            for lo, hi in s.ranges():
                K.adam_amsgrad_dev(s.fb.data[lo:hi], s.fb.grad[lo:hi], s.m[lo:hi], s.v[lo:hi], s.vmax[lo:hi], self._hyper[i],
                                   self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                   grad_sumsq=self._sumsq if clip > 0 else None, max_norm=clip,
                                   ema=None if s.ema is None else s.ema[lo:hi])

    def grad_norm(self):
        """Global gradient norm of the last step() (device scalar; reading it synchronises)."""
        return self._sumsq.sqrt()

    # ---- checkpoints: torch.optim.Adam's own format (utils/checkpoint.py:117-148 stores optimizer.state_dict()).
    # self.state[p] holds VIEWS of the flat moment buffers, so torch's packing code emits the reference layout:
    # state index = position in the concatenated groups (named_parameters order), keys step / exp_avg / exp_avg_sq /
    # max_exp_avg_sq.
    def _sync_state_views(self):
        for g in self.param_groups:
            for p in g["params"]:
                s, i = self._where[id(p)]
                s.ensure_state()
                self.state[p] = {"step": torch.tensor(float(self.t)), "exp_avg": s.fb.view(i, s.m),
                                 "exp_avg_sq": s.fb.view(i, s.v), "max_exp_avg_sq": s.fb.view(i, s.vmax)}

    def state_dict(self):
        self._sync_state_views()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("optimizer checkpoint has %d parameter groups, this model needs %d ([vis_enc, lan_enc, rest], "
                             "tools/train.py:78-93)" % (len(groups), len(self.param_groups)))
        for i, (g, mine) in enumerate(zip(groups, self.param_groups)):
            if len(g["params"]) != len(mine["params"]):
                raise ValueError("optimizer checkpoint group %d has %d parameters, the model has %d" %
                                 (i, len(g["params"]), len(mine["params"])))
        self._sync_state_views()
        views = {id(p): dict(self.state[p]) for g in self.param_groups for p in g["params"]}
        super().load_state_dict(state_dict)     # validates shapes group by group, casts, fills self.state with copies
        step = None
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state.get(p)
                if not st:
                    continue
                for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                    if st[k].shape != views[id(p)][k].shape:
                        raise ValueError("optimizer checkpoint state shape %s does not match parameter shape %s" %
                                         (tuple(st[k].shape), tuple(views[id(p)][k].shape)))
                    views[id(p)][k].copy_(st[k])
                step = int(float(st["step"]))
        if step is not None:
            self.t = step
        self._sync_state_views()
        if self.graph_mode:
            self._pow.copy_(torch.tensor([self.betas[0] ** self.t, self.betas[1] ** self.t], dtype=torch.float64))


class FlatDDP:
    """Data-parallel gradient averaging over the optimiser's flat gradient segments (one process per GPU, NCCL).

    Pure data parallelism is all the reference has; the forward/backward contain no collective.

    Two modes:
      overlap (eager launches): contiguous gradient ranges are all-reduced asynchronously in reverse-layer order from inside
          the encoder backward as they become final; finish() waits.
      deferred (`deferred=True`, used by the graphed runtime): nothing is issued during backward — the step is
          [fwd + bwd graph] -> exchange() -> [clip + Adam graph], so no collective has to be captured.
    The text-embedding table's gradient (64010 x D: 197 MB for ViT-B, a quarter of all gradient bytes) has at most
    B x Lt non-zero rows per rank: when the encoder provides them in compact form (`enc.sparse_text_grad`), ranks
    all-gather (ids, rows) and scatter-add locally instead of all-reducing the dense table."""

    def __init__(self, model, optimizer, process_group=None, deferred=False):
        self.model, self.opt, self.pg, self.deferred = model, optimizer, process_group, deferred
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._work = []
        self._use_avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        enc = getattr(model, "vis_enc", None)
        self.enc = enc
        self._enc_seg = None
        self._skip = None     # (lo, hi) of the dense text-embedding gradient inside the encoder segment, exchanged sparsely
        if enc is not None and hasattr(enc, "flat"):
            for s in optimizer.segments:
                if s.fb is enc._flat:
                    self._enc_seg = s
            enc._ddp = self if (self.world > 1 and not deferred) else None
            if self.world > 1 and getattr(enc, "sparse_text_grad", None) is not None and self._enc_seg is not None:
                fb = self._enc_seg.fb
                i = fb.names.index("text_embed")
                self._skip = (fb.offsets[i], fb.offsets[i + 1] if i + 1 < len(fb.offsets) else fb.numel)
                enc.sparse_text_grad["defer"] = True    # backward keeps (ids, rows) instead of the dense index_add_

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            for s in self.opt.segments:
                dist.broadcast(s.fb.data, src, group=self.pg)
                if s.ema is not None:
                    s.ema.copy_(s.fb.data)

    def _reduce(self, buf, async_op=True):
        if self.world <= 1 or buf.numel() == 0:
            return
        if self._use_avg:
            w = dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg, async_op=async_op)
            if async_op:
                self._work.append((w, None))
        else:
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=async_op)
            if async_op:
                self._work.append((w, buf))
            else:
                buf.div_(self.world)

    def _reduce_range(self, lo, hi, async_op=True):
        """All-reduces encoder-segment elements [lo, hi) minus the sparsely exchanged text-embedding rows."""
        g = self._enc_seg.fb.grad
        if self._skip is None or hi <= self._skip[0] or lo >= self._skip[1]:
            self._reduce(g[lo:hi], async_op)
            return
        if lo < self._skip[0]:
            self._reduce(g[lo:self._skip[0]], async_op)
        if hi > self._skip[1]:
            self._reduce(g[self._skip[1]:hi], async_op)

    def _exchange_text_rows(self):
        """all-gather (ids, rows) of every rank's text-embedding gradient and scatter-add the mean into the dense table."""
        st = self.enc.sparse_text_grad
        ids, rows = st.get("ids"), st.get("rows")
        if ids is None:
            return
        n = ids.numel()
        all_ids = torch.empty(self.world * n, dtype=ids.dtype, device=ids.device)
        all_rows = torch.empty(self.world * n, rows.shape[1], dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(all_ids, ids.reshape(-1).contiguous(), group=self.pg)
        dist.all_gather_into_tensor(all_rows, rows.contiguous(), group=self.pg)
        fb = self._enc_seg.fb
        table = fb.grad_of(fb.names.index("text_embed"))
        # index_put_(accumulate=True) sorts the indices: duplicates are summed in a fixed order, so every rank (all hold the
        # same gathered arrays) ends with bit-identical table gradients — index_add_'s atomics would let the replicas drift
        table.index_put_((all_ids,), all_rows * (1.0 / self.world), accumulate=True)

    # ---- hooks called by the encoder backward (simvg_b200/models/vis_encs/beit/beit3.py), overlap mode only
    def _attach_foreign(self):
        """If the trainer cleared gradients with model.zero_grad() / zero_grad(set_to_none=True), autograd allocated fresh p.grad
        tensors for the non-encoder parameters: copy them into the flat buffers before those are exchanged (the encoder writes
        its flat buffer directly)."""
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            return
        for s in self.opt.segments:
            if s is not self._enc_seg:
                s.fb.ensure()
                s.fb.attach_grads()

    def on_encoder_backward_start(self):
        """Head gradients are final once autograd reaches the encoder node."""
        self._attach_foreign()
        for s in self.opt.segments:
            if s is not self._enc_seg:
                self._reduce(s.fb.grad)

    def on_encoder_range_done(self, lo, hi):
        self._reduce_range(lo, hi)

    def exchange(self):
        """Deferred mode: every gradient exchange of the step, issued after backward (between the two step graphs)."""
        if self.world <= 1:
            return
        self._attach_foreign()
        for s in self.opt.segments:
            if s is self._enc_seg:
                self._reduce_range(0, s.fb.numel, async_op=False)
            else:
                self._reduce(s.fb.grad, async_op=False)
        if self._skip is not None:
            self._exchange_text_rows()

    # ---- chunked graph runtime (simvg_b200/runtime.py): collectives are issued BETWEEN the replays of the backward chunks
    def reduce_async(self, lo=None, hi=None, heads=False):
        """Asynchronous all-reduce of encoder-segment elements [lo, hi) and / or (heads=True) of every other segment.  NCCL
        orders it after the work already enqueued on the current stream; later launches on that stream overlap it."""
        if self.world <= 1:
            return
        if heads:
            for s in self.opt.segments:
                if s is not self._enc_seg:
                    self._reduce(s.fb.grad)
        if lo is not None:
            self._reduce_range(lo, hi)

    def wait(self):
        """The current stream waits for every outstanding all-reduce; then the sparse text-embedding rows are exchanged."""
        for w, buf in self._work:
            w.wait()
            if buf is not None:
                buf.div_(self.world)
        self._work = []
        if self.world > 1 and self._skip is not None:
            self._exchange_text_rows()

    def finish(self):
        """Call after backward(): waits for the outstanding all-reduces (and reduces anything not yet sent)."""
        if self.world <= 1:
            return
        if self.deferred:
            return self.exchange()
        if self._enc_seg is None or not self._work:
            self._attach_foreign()
            for s in self.opt.segments:
                if s is self._enc_seg:
                    self._reduce_range(0, s.fb.numel)
                else:
                    self._reduce(s.fb.grad)
        self.wait()
