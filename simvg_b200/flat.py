"""Flat fp32 storage for a set of nn.Parameters.

Every parameter stays an ordinary nn.Parameter under its reference state-dict key, but its
storage is a view into ONE contiguous buffer, and so is its .grad.  That turns the optimiser step, the gradient-norm
clip, the bf16 weight refresh and the data-parallel gradient all-reduce into single streaming passes, and lets
adjacent q/k/v projection weights be read as one [3D, D] matrix without a concat.
"""
import torch

ALIGN = 64  # elements (256 B): keeps every view 16-byte aligned for TMA / vector loads


class FlatBuffer:
    def __init__(self, named_params):
        """named_params: ordered list of (name, nn.Parameter); the order defines the layout."""
        self.names = [n for n, _ in named_params]
        self.params = [p for _, p in named_params]
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.data = None
        self.grad = None
        self.shadow = None  # bf16 copy of data
        self._build(self.params[0].device if self.params else torch.device("cpu"))

    def _build(self, device):
        data = torch.zeros(self.numel, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                v = data[off:off + p.numel()].view(p.shape)
                v.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = v
        old_grad = self.grad
        self.data = data
        self.grad = None
        self.shadow = None
        if old_grad is not None:
            self.attach_grads()

    def is_valid(self):
        if self.data is None or not self.params:
            return self.data is not None
        base = self.data.data_ptr()
        for i in (0, len(self.params) // 2, len(self.params) - 1):
            p = self.params[i]
            if p.device != self.data.device or p.data_ptr() != base + 4 * self.offsets[i]:
                return False
        return True

    def ensure(self):
        """Re-flattens after the parameters were moved / re-allocated individually (e.g. model.cuda())."""
        if not self.is_valid():
            self._build(self.params[0].device)
        return self

    def view(self, p_index, buf):
        p = self.params[p_index]
        off = self.offsets[p_index]
        return buf[off:off + p.numel()].view(p.shape)

    def attach_grads(self):
        """Makes every p.grad a view of one flat (zero-initialised on creation) gradient buffer."""
        if self.grad is None or self.grad.device != self.data.device:
            self.grad = torch.zeros_like(self.data)
        if all(p.grad is None for p in self.params):   # e.g. after zero_grad(set_to_none=True): one memset
            self.grad.zero_()
            for i, p in enumerate(self.params):
                p.grad = self.view(i, self.grad)
            return self.grad
        for i, p in enumerate(self.params):
            want = self.grad.data_ptr() + 4 * self.offsets[i]
            if p.grad is None or p.grad.data_ptr() != want:
                g = self.view(i, self.grad)
                if p.grad is not None:
                    g.copy_(p.grad)
                else:
                    g.zero_()
                p.grad = g
        return self.grad

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()
            for i, p in enumerate(self.params):
                if p.grad is None:
                    p.grad = self.view(i, self.grad)

    def grad_of(self, p_index):
        return self.view(p_index, self.grad)

    def index(self, p):
        for i, q in enumerate(self.params):
            if q is p:
                return i
        raise KeyError("parameter is not part of this FlatBuffer")
