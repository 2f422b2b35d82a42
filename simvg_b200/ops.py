"""Autograd glue between PyTorch tensors and the sm_100a kernels for the head's large projections."""
import torch

from simvg_b200 import kernels as K

bf16, f32 = torch.bfloat16, torch.float32


def _to_bf16(x):
    x = x.contiguous()
    if x.dtype == bf16:
        return x
    if x.numel() % 8 == 0 and x.dtype == f32:
        return K.cast_bf16(x, torch.empty(x.shape, device=x.device, dtype=bf16))
    return x.to(bf16)


class LinearFn(torch.autograd.Function):
    """y[R, Dout] = x[R, Din] @ W[Dout, Din]^T + b on the tcgen05 GEMM (bf16 operands, fp32 accumulate / output).

    Replaces nn.Linear / 1x1 Conv2d / nn.MultiheadAttention's in-projection of the memory in the head
    (/root/reference/simvg/models/heads/tgqs_kd_detr_head/tgqs_kd_detr_head.py:74-76,377-379 and the cross-attention
    K/V projections inside detrex MultiheadAttention, transformer.py:107-112) and their backward."""

    @staticmethod
    def forward(ctx, x, W, b):
        K.L.require_device(x)
        R, Din = x.shape
        Dout = W.shape[0]
        xb, Wb = _to_bf16(x), _to_bf16(W)
        y = K.gemm(xb, Wb, R, Dout, Din, epilogue=K.EPI_F32, bias=None if b is None else b.contiguous().float())
        ctx.save_for_backward(xb, Wb)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, Wb = ctx.saved_tensors
        R, Din = xb.shape
        Dout = Wb.shape[0]
        dy = dy.contiguous().float()
        dyb = torch.empty(R, Dout, device=dy.device, dtype=bf16)
        db = torch.zeros(Dout, device=dy.device, dtype=f32) if ctx.has_bias else None
        K.colsum(dy, out=db, out_bf16=dyb)
        dx = dW = None
        if ctx.needs_input_grad[0]:
            dx = K.gemm(dyb, Wb, R, Din, Dout, b_mn=True, epilogue=K.EPI_F32)
        if ctx.needs_input_grad[1]:
            dW = K.wgrad(dyb, xb, Dout, Din, R)
        return dx, dW, db


def linear(x, W, b=None):
    """x: [..., Din] -> [..., Dout]; rows go through the tcgen05 GEMM when the shape allows (Din, Dout multiples of 8)."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.shape[1] % 8 != 0 or W.shape[0] % 8 != 0:
        raise RuntimeError("simvg_b200.ops.linear needs Din and Dout to be multiples of 8 (got %d -> %d)" % (x2.shape[1], W.shape[0]))
    y = LinearFn.apply(x2, W, b)
    return y.view(*lead, W.shape[0])
