"""CPU oracle, bf16-operand variant — TEST INFRASTRUCTURE ONLY (never imported by `simvg_b200/`).

The CUDA product feeds its tensor-core GEMMs and its attention kernels bf16 operands (fp32 accumulation, fp32 residual
stream / LayerNorm / softmax / GELU).  Against the plain fp32 oracle (oracle/simvg_oracle.py) that shows up as ~5e-3 on hidden
features and ~1-3 % on parameter gradients, which is too loose a bound to notice a small bug.  This module restates the
SAME reference algorithm (same file:line citations as simvg_oracle.py) but rounds to bf16 exactly where the product stores
or consumes bf16 — forward values AND the gradients that the backward kernels hand to a GEMM — so that what is left between
the two is accumulation order and transcendental approximations only.  tests/ compare the CUDA path against BOTH:
  fp32 oracle      -> the north_star tolerance (1e-3 on outputs) and the rounding budget on gradients;
  bf16 emulation   -> tight per-tensor bounds that separate rounding from bugs.

Rounding points (simvg_b200/models/vis_encs/beit/beit3.py::encoder_forward / encoder_backward):
  forward : im2col(image), every weight matrix, h=LN1(x), qkv, P (softmax numerators, before normalisation), O, a=LN_in(O),
            h2=LN2(x), u=fc1(h2), f=LN_F(gelu(u)); the head's input_proj operands.
  backward: the gradient w.r.t. each GEMM output (dY operands), dS inside attention, P^T for dV, and each dgrad GEMM output.
"""
import torch
import torch.nn.functional as F

from oracle import simvg_oracle as O


class _Round(torch.autograd.Function):
    """Identity up to bf16 rounding of the value (fwd) and/or of the incoming gradient (bwd)."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.bfloat16().to(x.dtype) if fwd else x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return (g.bfloat16().to(g.dtype) if ctx.bwd else g), None, None


def rb(x):        # stored as bf16, gradient arrives as bf16
    return _Round.apply(x, True, True)


def rf(x):        # operand rounded on the fly (weights, image); its gradient stays fp32
    return _Round.apply(x, True, False)


def rg(x):        # fp32 value whose gradient is handed to a GEMM as bf16
    return _Round.apply(x, False, True)


def _r(t):
    return t.bfloat16().to(t.dtype)


class _AttnCore(torch.autograd.Function):
    """softmax(QK^T + mask) V as the kernels compute it (csrc/attn_fwd.cu, attn_bwd.cu); q, k, v hold bf16 values.

    forward : e = exp(S - rowmax) in fp32, O = (bf16(e) @ V) / sum(e), O stored bf16.
    backward: P = e / sum(e) recomputed in fp32; dV = bf16(P)^T dO; dP = dO V^T; delta = rowsum(dO * O);
              dS = bf16(bf16(P) * (dP - delta)); dQ = dS K; dK = dS^T Q.   (torchscale MultiheadAttention)"""

    @staticmethod
    def forward(ctx, q, k, v, kpm):
        s = q @ k.transpose(-1, -2)
        if kpm is not None:
            s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
        e = torch.exp(s - s.amax(-1, keepdim=True))
        l = e.sum(-1, keepdim=True)
        o = _r((_r(e) @ v) / l)
        ctx.save_for_backward(q, k, v, e / l, o)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, p, o = ctx.saved_tensors
        do = _r(do)
        dv = _r(p).transpose(-1, -2) @ do
        dp = do @ v.transpose(-1, -2)
        delta = (do * o).sum(-1, keepdim=True)
        ds = _r(_r(p) * (dp - delta))   # the kernel keeps only the packed bf16 P^T live across the dP^T wait
        return ds @ k, ds.transpose(-1, -2) @ q, dv, None


def _lin(x, w, b):
    return F.linear(x, rf(w), b)


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def encoder_forward(sd, cfg, image, ids, pad_mask, prefix=""):
    """BEiT3.forward + Encoder.forward (beit3_base.py:441-488, 336-407) with the product's bf16 rounding points; eval
    semantics (no DropPath).  Expert A = vision tokens, expert B = text tokens (torchscale MultiwayNetwork, A.3)."""
    p = prefix + "beit3."
    D, H, P, eps = cfg["D"], cfg["H"], cfg["patch_size"], cfg["eps"]
    dh = D // H
    B = image.shape[0]
    x1 = rg(F.conv2d(rf(image), rf(sd[p + "vision_embed.proj.weight"]), sd[p + "vision_embed.proj.bias"], stride=P))
    x1 = x1.flatten(2).transpose(1, 2)
    x1 = torch.cat([sd[p + "vision_embed.cls_token"].expand(B, -1, -1), x1], dim=1)
    split = x1.shape[1]
    x2 = F.embedding(ids, sd[p + "text_embed.weight"])
    Lt = x2.shape[1]
    kpm = None
    if pad_mask is not None:
        kpm = torch.cat([torch.zeros(B, split, dtype=torch.bool, device=image.device), pad_mask.bool()], dim=1)
    xa = x1 + sd[p + "encoder.embed_positions.A.weight"][2:2 + split]
    xb = x2 + sd[p + "encoder.embed_positions.B.weight"][2:2 + Lt]
    if pad_mask is not None:
        xb = xb * (1 - pad_mask.unsqueeze(-1).type_as(xb))
    xs = [xa, xb]
    for li in range(cfg["layers"]):
        lp = "%sencoder.layers.%d." % (p, li)
        qkv = []
        for which, x in zip("AB", xs):
            h = rb(_ln(x, sd[lp + "self_attn_layer_norm.%s.weight" % which], sd[lp + "self_attn_layer_norm.%s.bias" % which], eps))
            sa = lp + "self_attn."
            q = rb(_lin(h, sd[sa + "q_proj.%s.weight" % which], sd[sa + "q_proj.%s.bias" % which]) * (dh ** -0.5))
            k = rb(_lin(h, sd[sa + "k_proj.%s.weight" % which], sd[sa + "k_proj.%s.bias" % which]))
            v = rb(_lin(h, sd[sa + "v_proj.%s.weight" % which], sd[sa + "v_proj.%s.bias" % which]))
            qkv.append((q, k, v))
        L = split + Lt
        q, k, v = (torch.cat([qkv[0][i], qkv[1][i]], dim=1).view(B, L, H, dh).transpose(1, 2) for i in range(3))
        o = _AttnCore.apply(q, k, v, kpm).transpose(1, 2).reshape(B, L, D)
        os_ = [o[:, :split], o[:, split:]]
        for g, which in enumerate("AB"):
            sa = lp + "self_attn."
            a = rb(_ln(os_[g], sd[sa + "inner_attn_ln.%s.weight" % which], sd[sa + "inner_attn_ln.%s.bias" % which], eps))
            xs[g] = xs[g] + rg(_lin(a, sd[sa + "out_proj.%s.weight" % which], sd[sa + "out_proj.%s.bias" % which]))
            fp = "%sffn.%s." % (lp, which)
            h2 = rb(_ln(xs[g], sd[lp + "final_layer_norm.%s.weight" % which], sd[lp + "final_layer_norm.%s.bias" % which], eps))
            u = rb(_lin(h2, sd[fp + "fc1.weight"], sd[fp + "fc1.bias"]))
            f = rb(_ln(F.gelu(u), sd[fp + "ffn_layernorm.weight"], sd[fp + "ffn_layernorm.bias"], eps))
            xs[g] = xs[g] + rg(_lin(f, sd[fp + "fc2.weight"], sd[fp + "fc2.bias"]))
    outs = [_ln(xs[g], sd[p + "encoder.layer_norm.%s.weight" % w], sd[p + "encoder.layer_norm.%s.bias" % w], eps)
            for g, w in enumerate("AB")]
    return torch.cat(outs, dim=1)


class OracleModelBF16(O.OracleModel):
    """OracleModel whose encoder and head input projection round operands to bf16 where the CUDA product does."""

    def _features(self, img, ids, text_mask):
        B, _, Hh, Ww = img.shape
        x = encoder_forward(self.sd, self.cfg, img, ids, text_mask, prefix="vis_enc.")
        Lt = ids.shape[-1]
        img_feat, text_feat, cls_feat = x[:, 1:-Lt], x[:, -Lt:], x[:, 0]
        P = self.cfg["patch_size"]
        return img_feat.transpose(-1, -2).reshape(B, -1, Hh // P, Ww // P), text_feat, cls_feat

    def _head_sd(self):
        sd = dict(self.sd)
        sd["head.input_proj.weight"] = rf(self.sd["head.input_proj.weight"])
        return sd

    def forward_train(self, img, ids, img_metas, text_attention_mask, gt_bbox, world_size=1):
        for m in img_metas:
            m["batch_input_shape"] = tuple(img.shape[-2:])
        x_mm, text_feat, cls_feat = self._features(img, ids, text_attention_mask)
        # input_proj runs on the tcgen05 GEMM (simvg_b200/ops.py::LinearFn): bf16 operands, fp32 output, bf16 dY for wgrad/dgrad
        x_mm = rf(x_mm)
        losses, out = O.head_forward_train(self._head_sd(), self.hc, x_mm, img_metas, cls_feat, text_feat, gt_bbox,
                                           text_attention_mask, world_size=world_size, proj_hook=rg)
        with torch.no_grad():
            preds = [O.get_predictions(out["decoder_branch_output"], img_metas),
                     O.get_predictions(out["token_branch_output"], img_metas)]
        return losses, preds, out
