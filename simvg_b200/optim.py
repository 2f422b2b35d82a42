"""Fused global-norm clip + Adam(amsgrad) over flat parameter segments, and the data-parallel gradient exchange.

Replaces, inside the timed train step (/root/reference/simvg/apis/train.py:75-83):
  * torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.grad_norm_clip)   (:81-82)
  * torch.optim.Adam(amsgrad=True) built with three learning-rate groups by parameter-name substring
    (/root/reference/tools/train.py:78-93; hyper-parameters configs/.../refcoco_onestage.py:107-123)
  * MMDistributedDataParallel's bucketed gradient all-reduce (tools/train.py:102-103)
by: one sum-of-squares pass + one update pass per segment (simvgb_sumsq / simvgb_adam_amsgrad, 36 B per parameter), and
one NCCL all-reduce per contiguous gradient range, issued from inside the encoder's backward as each layer's gradients
become final so the exchange overlaps the remaining backward kernels.  Unused parameters (vision_embed.mask_token, the
token branch in decoder-only configs — SURVEY Appendix C.13) simply keep zero gradients: no graph walk is needed.
"""
import torch
import torch.distributed as dist

from simvg_b200 import kernels as K
from simvg_b200.flat import FlatBuffer


class _Segment:
    def __init__(self, name, fb, lr):
        self.name, self.fb, self.lr = name, fb, lr
        self.m = self.v = self.vmax = None

    def ensure_state(self):
        d = self.fb.data
        if self.m is None or self.m.device != d.device:
            self.m, self.v, self.vmax = torch.zeros_like(d), torch.zeros_like(d), torch.zeros_like(d)


class FusedAdamAMSGrad:
    def __init__(self, model, lr, lr_vis_enc=None, lr_lan_enc=None, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.0,
                 amsgrad=True, grad_norm_clip=None):
        assert amsgrad, "SimVG trains with Adam(amsgrad=True)"
        self.betas, self.eps, self.weight_decay, self.grad_norm_clip = betas, eps, weight_decay, grad_norm_clip
        self.t = 0
        self.segments = []
        enc = getattr(model, "vis_enc", None)
        enc_ids = set()
        if enc is not None and hasattr(enc, "flat"):
            fb = enc.flat()
            enc_ids = {id(p) for p in fb.params}
            self.segments.append(_Segment("vis_enc", fb, lr if lr_vis_enc is None else lr_vis_enc))
        rest = [(n, p) for n, p in model.named_parameters() if id(p) not in enc_ids]
        vis = [(n, p) for n, p in rest if "vis_enc" in n]
        lan = [(n, p) for n, p in rest if "lan_enc" in n and "vis_enc" not in n]
        other = [(n, p) for n, p in rest if "lan_enc" not in n and "vis_enc" not in n]
        for name, group, glr in (("vis_enc_rest", vis, lr_vis_enc), ("lan_enc", lan, lr_lan_enc), ("rest", other, lr)):
            if group:
                self.segments.append(_Segment(name, FlatBuffer(group), lr if glr is None else glr))
        for s in self.segments:
            s.fb.ensure()
            s.fb.attach_grads()
        self._sumsq = None
        self._hyper = None        # device [n_segments, 4]: lr, 1 - beta1^t, sqrt(1 - beta2^t), pad  (graph mode)
        self.graph_mode = False

    # ---- CUDA-graph support: the step-dependent scalars live in device memory and are advanced BY THE GRAPH ITSELF
    # (beta^t is a running product updated by a tiny in-graph op), so replays need no per-step host->device traffic and the
    # host may run any number of steps ahead of the device.
    def enable_graph_mode(self):
        dev = self.segments[0].fb.data.device
        n = len(self.segments)
        self._hyper = torch.zeros(n, 4, device=dev, dtype=torch.float32)       # lr, 1 - beta1^t, sqrt(1 - beta2^t), pad
        self._betas_dev = torch.tensor([self.betas[0], self.betas[1]], device=dev, dtype=torch.float64)
        self._pow = torch.tensor([self.betas[0] ** self.t, self.betas[1] ** self.t], device=dev, dtype=torch.float64)
        self._lr_host = [s.lr for s in self.segments]
        self._lr_dev = torch.tensor(self._lr_host, device=dev, dtype=torch.float32)
        self.graph_mode = True
        for s in self.segments:
            s.ensure_state()
        if self._sumsq is None or self._sumsq.device != dev:
            self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)

    def advance(self):
        """Graph mode, once per step before the replay: host-side step count; learning rates are re-uploaded only when a
        scheduler changed them (a blocking 8-byte copy, outside the graph)."""
        self.t += 1
        lrs = [s.lr for s in self.segments]
        if lrs != self._lr_host:
            self._lr_host = lrs
            self._lr_dev.copy_(torch.tensor(lrs, dtype=torch.float32))

    @property
    def param_groups(self):  # scheduler-facing view (core/scheduler.py multiplies group["lr"])
        return [_LRProxy(s) for s in self.segments]

    def zero_grad(self, set_to_none=False):
        for s in self.segments:
            s.fb.ensure()
            if s.fb.grad is None:
                s.fb.attach_grads()
            s.fb.zero_grad()

    def step(self):
        if self.graph_mode:
            return self._step_graph()
        self.t += 1
        dev = self.segments[0].fb.data.device
        if self._sumsq is None or self._sumsq.device != dev:
            self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        clip = float(self.grad_norm_clip) if self.grad_norm_clip else 0.0
        if clip > 0:
            self._sumsq.zero_()
            for s in self.segments:
                K.sumsq(s.fb.grad, self._sumsq)
        for s in self.segments:
            s.ensure_state()
            K.adam_amsgrad(s.fb.data, s.fb.grad, s.m, s.v, s.vmax, s.lr, self.betas[0], self.betas[1], self.eps,
                           self.weight_decay, self.t, grad_sumsq=self._sumsq if clip > 0 else None, max_norm=clip)

    def _step_graph(self):
        """Same update with step-invariant launch arguments (advance() must have been called for this step)."""
        clip = float(self.grad_norm_clip) if self.grad_norm_clip else 0.0
        self._pow.mul_(self._betas_dev)                                   # beta^t (fp64 running product, on the device)
        self._hyper[:, 0] = self._lr_dev
        self._hyper[:, 1] = (1.0 - self._pow[0]).float()
        self._hyper[:, 2] = (1.0 - self._pow[1]).sqrt().float()
        if clip > 0:
            self._sumsq.zero_()
            for s in self.segments:
                K.sumsq(s.fb.grad, self._sumsq)
        for i, s in enumerate(self.segments):
            K.adam_amsgrad_dev(s.fb.data, s.fb.grad, s.m, s.v, s.vmax, self._hyper[i], self.betas[0], self.betas[1], self.eps,
                               self.weight_decay, grad_sumsq=self._sumsq if clip > 0 else None, max_norm=clip)

    def grad_norm(self):
        """Global gradient norm of the last step() (device scalar; reading it synchronises)."""
        return self._sumsq.sqrt()

    # torch.optim-compatible checkpoint format (utils/checkpoint.py:117-148 stores optimizer.state_dict())
    def state_dict(self):
        state, groups, idx = {}, [], 0
        for s in self.segments:
            s.ensure_state()
            ids = []
            for i, p in enumerate(s.fb.params):
                if not p.requires_grad:
                    continue
                state[idx] = {"step": torch.tensor(float(self.t)), "exp_avg": s.fb.view(i, s.m).clone(),
                              "exp_avg_sq": s.fb.view(i, s.v).clone(), "max_exp_avg_sq": s.fb.view(i, s.vmax).clone()}
                ids.append(idx)
                idx += 1
            groups.append({"lr": s.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                           "amsgrad": True, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        for s, g in zip(self.segments, sd["param_groups"]):
            s.ensure_state()
            s.lr = g["lr"]
            for i, p in enumerate(s.fb.params):
                if not p.requires_grad:
                    continue
                st = sd["state"].get(idx)
                if st is not None:
                    s.fb.view(i, s.m).copy_(st["exp_avg"])
                    s.fb.view(i, s.v).copy_(st["exp_avg_sq"])
                    s.fb.view(i, s.vmax).copy_(st["max_exp_avg_sq"])
                    self.t = int(st["step"])
                idx += 1


class _LRProxy(dict):
    def __init__(self, seg):
        super().__init__(lr=seg.lr, initial_lr=seg.lr)
        self._seg = seg

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        if k == "lr":
            self._seg.lr = float(v)


class FlatDDP:
    """Data-parallel gradient averaging over the optimiser's flat gradient segments (one process per GPU, NCCL).

    Pure data parallelism is all the reference has; the forward/backward contain no collective.  Gradient
    ranges are all-reduced asynchronously in reverse-layer order from inside the encoder backward."""

    def __init__(self, model, optimizer, process_group=None):
        self.model, self.opt, self.pg = model, optimizer, process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._work = []
        self._use_avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        enc = getattr(model, "vis_enc", None)
        self._enc_seg = None
        if enc is not None and hasattr(enc, "flat"):
            for s in optimizer.segments:
                if s.fb is enc._flat:
                    self._enc_seg = s
            enc._ddp = self if self.world > 1 else None

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            for s in self.opt.segments:
                dist.broadcast(s.fb.data, src, group=self.pg)

    def _reduce(self, buf):
        if self.world <= 1 or buf.numel() == 0:
            return
        if self._use_avg:
            self._work.append((dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg, async_op=True), None))
        else:
            self._work.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), buf))

    # ---- hooks called by the encoder backward (simvg_b200/models/vis_encs/beit/beit3.py)
    def on_encoder_backward_start(self):
        """Head gradients are final once autograd reaches the encoder node."""
        for s in self.opt.segments:
            if s is not self._enc_seg:
                self._reduce(s.fb.grad)

    def on_encoder_range_done(self, lo, hi):
        self._reduce(self._enc_seg.fb.grad[lo:hi])

    def finish(self):
        """Call after backward(): waits for the outstanding all-reduces (and reduces anything not yet sent)."""
        if self.world <= 1:
            return
        if self._enc_seg is None or not self._work:
            for s in self.opt.segments:
                self._reduce(s.fb.grad)
        for w, buf in self._work:
            w.wait()
            if buf is not None:
                buf.div_(self.world)
        self._work = []
