"""`ExponentialMovingAverage` with the reference's interface (/root/reference/simvg/models/utils.py:130-180; constructed at
tools/train.py:104, `update_params()` per iteration at simvg/apis/train.py:85-86, `apply_shadow()` / `restore()` around the
evaluation at tools/train.py:133-138, `shadow` saved with the checkpoint).

With a `FusedAdamAMSGrad` the averaged weights are one more flat fp32 stream updated INSIDE the optimiser pass
(`simvgb_adam_amsgrad(..., ema, decay)`, decay = min(alpha, (1 + t) / (10 + t))): `update_params()` then only advances the
bookkeeping the optimiser has not already done, and `apply_shadow()` / `restore()` swap two flat buffers instead of cloning and
re-loading whole state dicts.  Without one (plain `torch.optim.Adam`) the class does the reference's arithmetic with foreach ops.
Buffers: the benchmarked models register none that change during training; any that exist are averaged (or copied, when
`buffer_ema=False`) exactly as the reference does.
"""
import torch


class ExponentialMovingAverage(object):
    def __init__(self, model, alpha, buffer_ema=True, optimizer=None):
        self.model, self.alpha, self.buffer_ema = model, float(alpha), buffer_ema
        self.opt = optimizer if (optimizer is not None and hasattr(optimizer, "enable_ema")) else None
        self._step = 0
        self.param_keys = [k for k, _ in model.named_parameters()]
        self.buffer_keys = [k for k, _ in model.named_buffers()]
        self._applied = False
        if self.opt is not None:
            if self.opt.ema_alpha is None:
                self.opt.enable_ema(self.alpha)
            self._params = None
        else:
            self._params = {k: p.detach().clone() for k, p in model.named_parameters()}
        self._buffers = {k: b.detach().clone() for k, b in model.named_buffers()}
        self._backup_buffers = None

    # the reference exposes `step` (number of updates so far) and `shadow` (name -> averaged tensor)
    @property
    def step(self):
        return self.opt.ema_t if self.opt is not None else self._step

    @property
    def shadow(self):
        out = {}
        for k, p in self.model.named_parameters():
            out[k] = (self.opt.ema_view(p) if self.opt is not None else self._params[k]).detach().clone()
        out.update({k: v.clone() for k, v in self._buffers.items()})
        return out

    def load_shadow(self, shadow):
        """Restore averaged weights saved by the reference (`checkpoint["ema_state_dict"]`, simvg/utils/checkpoint.py)."""
        with torch.no_grad():
            for k, p in self.model.named_parameters():
                if k in shadow:
                    (self.opt.ema_view(p) if self.opt is not None else self._params[k]).copy_(shadow[k])
            for k in self._buffers:
                if k in shadow:
                    self._buffers[k].copy_(shadow[k])

    def update_params(self):
        """simvg/apis/train.py:85-86.  With the fused optimiser the parameter average was already updated by `optimizer.step()`."""
        decay = min(self.alpha, (self.step + 1) / (self.step + 10)) if self.opt is None else self.opt.ema_decay(max(self.opt.ema_t - 1, 0))
        with torch.no_grad():
            if self.opt is None:
                cur = [p.detach() for _, p in self.model.named_parameters()]
                sh = [self._params[k] for k in self.param_keys]
                torch._foreach_mul_(sh, decay)
                torch._foreach_add_(sh, cur, alpha=1.0 - decay)
                self._step += 1
            for k, b in self.model.named_buffers():
                if self.buffer_ema and b.is_floating_point():
                    self._buffers[k].mul_(decay).add_(b.detach(), alpha=1.0 - decay)
                else:
                    self._buffers[k].copy_(b.detach())

    def _swap(self):
        with torch.no_grad():
            if self.opt is not None:
                self.opt.swap_ema()
            else:
                for k, p in self.model.named_parameters():
                    tmp = p.detach().clone()
                    p.copy_(self._params[k])
                    self._params[k].copy_(tmp)
            for k, b in self.model.named_buffers():
                tmp = b.detach().clone()
                b.copy_(self._buffers[k])
                self._buffers[k].copy_(tmp)

    def apply_shadow(self):
        if self._applied:
            raise RuntimeError("apply_shadow() called twice without restore()")
        self._swap()
        self._applied = True

    def restore(self):
        if not self._applied:
            raise RuntimeError("restore() without apply_shadow()")
        self._swap()
        self._applied = False

    def get_model_state(self):
        return {k: v.clone().detach() for k, v in self.model.state_dict().items()}
