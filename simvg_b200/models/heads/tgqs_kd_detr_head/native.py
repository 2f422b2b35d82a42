"""The DETR decoder stack of SimVG's head on the native fp32 head kernels (csrc/headops.cu).

`decoder_stack(dec, query, kin, val, query_pos, key_padding_mask)` computes what `DetrTransformerDecoder.forward` computes
(/root/reference/simvg/models/heads/tgqs_kd_detr_head/transformer.py:134-186 over detrex's post-norm BaseTransformerLayer
(self_attn, norm, cross_attn, norm, ffn, norm)) — for the object-token decoder (cross-attention over the
[B, N, E] image memory, key / value projections absorbed into the query / output side) and for the text-guided query generation
(cross-attention over the <= 32 text tokens) — as ONE autograd node that sequences ~12 fused launches per layer forward and ~20
backward, instead of the ~35 + ~70 tiny eager launches per layer of the op-by-op path (which remains the CPU implementation the
oracle tests pin, and the reference these kernels are tested against on the GPU).

Parameter gradients are accumulated straight into `p.grad` (for the fused optimiser: views of its flat gradient buffer) by the
backward kernels; the `anchor` input only keeps the node in the graph — the same arrangement as the encoder's `_EncoderFn`.
Dropout (attention probabilities, after the FFN activation, after the FFN output: p = 0.1 in training) uses ONE torch.rand per
stack call; the kernels threshold the uniforms themselves.
"""
import torch

from simvg_b200 import kernels as K

f32 = torch.float32


def _grad(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class _Arena:
    """Carves zero-initialised fp32 buffers out of one allocation (one memset launch for all of them)."""

    def __init__(self):
        self.req = []

    def want(self, *shape):
        n = 1
        for s in shape:
            n *= s
        n = (n + 63) // 64 * 64
        self.req.append((shape, n))
        return len(self.req) - 1

    def build(self, device):
        total = sum(n for _, n in self.req)
        buf = torch.zeros(max(total, 1), device=device, dtype=f32)
        out, off = [], 0
        for shape, n in self.req:
            cnt = 1
            for s in shape:
                cnt *= s
            out.append(buf[off:off + cnt].view(*shape))
            off += n
        return out


def _layer_params(layer):
    sa, ca = layer.attentions[0].attn, layer.attentions[1].attn
    ffn = layer.ffns[0].layers
    return dict(sa_w=sa.in_proj_weight, sa_b=sa.in_proj_bias, sa_ow=sa.out_proj.weight, sa_ob=sa.out_proj.bias,
                ca_w=ca.in_proj_weight, ca_b=ca.in_proj_bias, ca_ow=ca.out_proj.weight, ca_ob=ca.out_proj.bias,
                w1=ffn[0][0].weight, b1=ffn[0][0].bias, w2=ffn[1].weight, b2=ffn[1].bias,
                n0w=layer.norms[0].weight, n0b=layer.norms[0].bias, n1w=layer.norms[1].weight, n1b=layer.norms[1].bias,
                n2w=layer.norms[2].weight, n2b=layer.norms[2].bias)


def _ffn_splits(F):
    """Split-K factor of the second FFN linear (K = F).  1: the forward stays deterministic (no atomics); at R <= 640 rows the
    unsplit product is ~6 us on 16 CTAs, which the step graph hides behind nothing but is still ~0.1 % of the step."""
    return 1


def stack_forward(dec, query, qpos, kin, val, kpm, training):
    """query / qpos [B, nq, E]; kin (keys + key positions) / val [B, N, E]; kpm [B, N] bool or None.
    -> (stacked outputs [Lout, B, nq, E], saved context for stack_backward)."""
    B, nq, E = query.shape
    N = kin.shape[1]
    H, R = 8, B * nq
    dev = query.device
    absorbed = N > 32
    scale = (E // H) ** -0.5
    nl = len(dec.layers)
    p_attn = dec.layers[0].attentions[0].attn_drop if training else 0.0
    p_ffn = dec.layers[0].ffns[0].layers[0][2].p if training else 0.0
    F = dec.layers[0].ffns[0].layers[0][0].weight.shape[0]
    x = query.reshape(R, E).contiguous().float()
    qp = qpos.reshape(R, E).contiguous().float()
    kin2 = kin.reshape(B * N, E).contiguous().float()
    val2 = val.reshape(B * N, E).contiguous().float()
    kpm_u8 = None if kpm is None else kpm.to(torch.uint8).contiguous()
    # dropout uniforms for the whole stack: per layer [self-attn P | cross-attn P | ffn hidden | ffn out]
    sizes = [B * H * nq * nq, R * H * N, R * F, R * E]
    U = None
    if p_attn > 0 or p_ffn > 0:
        U = torch.rand(nl * sum(sizes), device=dev, dtype=f32)

    def u_of(li, which):
        if U is None or (which < 2 and p_attn <= 0) or (which >= 2 and p_ffn <= 0):
            return None
        off = li * sum(sizes) + sum(sizes[:which])
        return U[off:off + sizes[which]]

    f_bufs = None
    if _ffn_splits(F) > 1:                               # split-K outputs must start at zero
        ar = _Arena()
        f_idx = [ar.want(R, E) for _ in range(nl)]
        f_bufs = ar.build(dev)
    saved, outs = [], []
    post = dec.post_norm_layer
    for li, layer in enumerate(dec.layers):
        P = _layer_params(layer)
        s = dict(x=x)
        # ---- self-attention over the nq queries (q, k from x + qpos; v from x), residual, norm
        s["qkv"] = K.head_lin_fwd(x, P["sa_w"], P["sa_b"], x2=qp, n_split=2 * E)
        qkv = s["qkv"]
        s["ctx"], s["P0"] = K.head_attn_small_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B, nq, nq, H, scale,
                                                  drop_u=u_of(li, 0), drop_p=p_attn)
        s["a"] = K.head_lin_fwd(s["ctx"], P["sa_ow"], P["sa_ob"])
        s["x1"], s["m0"], s["r0"] = K.head_lnres_fwd(x, s["a"], P["n0w"], P["n0b"])
        # ---- cross-attention
        Wq, Wk, Wv = P["ca_w"][:E], P["ca_w"][E:2 * E], P["ca_w"][2 * E:]
        bq, bk, bv = P["ca_b"][:E], P["ca_b"][E:2 * E], P["ca_b"][2 * E:]
        s["q2"] = K.head_lin_fwd(s["x1"], Wq, bq, x2=qp, n_split=E)
        if absorbed:
            s["ctx2"], s["P1"], s["z"], s["psum"] = K.head_xattn_fwd(s["q2"], kin2, val2, Wk, bk, Wv, bv, B, nq, N, kpm=kpm_u8,
                                                                       drop_u=u_of(li, 1), drop_p=p_attn)
        else:
            s["kp"] = K.head_lin_fwd(kin2, Wk, bk)
            s["vp"] = K.head_lin_fwd(val2, Wv, bv)
            s["ctx2"], s["P1"] = K.head_attn_small_fwd(s["q2"], s["kp"], s["vp"], B, nq, N, H, scale, kpm=kpm_u8,
                                                       drop_u=u_of(li, 1), drop_p=p_attn)
        s["a2"] = K.head_lin_fwd(s["ctx2"], P["ca_ow"], P["ca_ob"])
        s["x2"], s["m1"], s["r1"] = K.head_lnres_fwd(s["x1"], s["a2"], P["n1w"], P["n1b"])
        # ---- FFN
        s["h"] = K.head_lin_fwd(s["x2"], P["w1"], P["b1"], relu=True, drop_u=u_of(li, 2), drop_p=p_ffn)
        s["f"] = K.head_lin_fwd(s["h"], P["w2"], P["b2"], k_splits=_ffn_splits(F), out=None if f_bufs is None else f_bufs[f_idx[li]])
        s["x3"], s["m2"], s["r2"] = K.head_lnres_fwd(s["x2"], s["f"], P["n2w"], P["n2b"], drop_u=u_of(li, 3), drop_p=p_ffn)
        x = s["x3"]
        if dec.return_intermediate or li == nl - 1:
            if post is not None:
                y, s["mp"], s["rp"] = K.head_lnres_fwd(x, None, post.weight, post.bias)
            else:
                y = x
            outs.append(y.view(B, nq, E))
        saved.append(s)
    ctx = dict(saved=saved, qp=qp, kin=kin2, val=val2, kpm=kpm_u8, shape=(B, nq, N, E, H, F), p=(p_attn, p_ffn), u_of=u_of,
               absorbed=absorbed, scale=scale, n_out=len(outs))
    return torch.stack(outs), ctx


def stack_backward(dec, ctx, dout):
    """dout [Lout, B, nq, E] -> (dquery, dqpos, dkin, dval); parameter gradients accumulated into p.grad."""
    B, nq, N, E, H, F = ctx["shape"]
    R = B * nq
    p_attn, p_ffn = ctx["p"]
    u_of, absorbed, scale = ctx["u_of"], ctx["absorbed"], ctx["scale"]
    qp, kin2, val2, kpm = ctx["qp"], ctx["kin"], ctx["val"], ctx["kpm"]
    dev = qp.device
    nl = len(dec.layers)
    dout = dout.reshape(-1, R, E).contiguous().float()
    post = dec.post_norm_layer
    ar = _Arena()
    i_qp, i_kin, i_val = ar.want(R, E), ar.want(B * N, E), ar.want(B * N, E)
    per = []
    for _ in range(nl):
        d = dict(dx3=ar.want(R, E), dx2=ar.want(R, E), df=ar.want(R, E), dh=ar.want(R, F), dx1=ar.want(R, E), da2=ar.want(R, E),
                 dctx2=ar.want(R, E), dq2=ar.want(R, E), da=ar.want(R, E), dctx=ar.want(R, E), dqkv=ar.want(R, 3 * E))
        if not absorbed:
            d["dkp"], d["dvp"] = ar.want(B * N, E), ar.want(B * N, E)
        per.append(d)
    i_dx0 = ar.want(R, E)
    bufs = ar.build(dev)
    dqp, dkin, dval = bufs[i_qp], bufs[i_kin], bufs[i_val]
    oi = ctx["n_out"] - 1
    for li in range(nl - 1, -1, -1):
        layer, s = dec.layers[li], ctx["saved"][li]
        P = _layer_params(layer)
        g = {k: bufs[v] for k, v in per[li].items()}
        dx3 = g["dx3"]      # holds the next layer's dx (zero for the last layer)
        if dec.return_intermediate or li == nl - 1:
            if post is not None:
                K.head_lnres_bwd(dout[oi], s["x3"], None, post.weight, s["mp"], s["rp"], _grad(post.weight), _grad(post.bias), da=dx3)
            else:
                dx3.add_(dout[oi])
            oi -= 1
        # ---- FFN
        K.head_lnres_bwd(dx3, s["x2"], s["f"], P["n2w"], s["m2"], s["r2"], _grad(P["n2w"]), _grad(P["n2b"]), da=g["dx2"], db=g["df"],
                         drop_u=u_of(li, 3), drop_p=p_ffn)
        K.head_lin_bwd(g["df"], s["h"], P["w2"], dx=g["dh"], dW=_grad(P["w2"]), db=_grad(P["b2"]))
        K.head_lin_bwd(g["dh"], s["x2"], P["w1"], y=s["h"], relu=True, drop_u=u_of(li, 2), drop_p=p_ffn, dx=g["dx2"],
                       dW=_grad(P["w1"]), db=_grad(P["b1"]))
        # ---- cross-attention
        K.head_lnres_bwd(g["dx2"], s["x1"], s["a2"], P["n1w"], s["m1"], s["r1"], _grad(P["n1w"]), _grad(P["n1b"]), da=g["dx1"], db=g["da2"])
        K.head_lin_bwd(g["da2"], s["ctx2"], P["ca_ow"], dx=g["dctx2"], dW=_grad(P["ca_ow"]), db=_grad(P["ca_ob"]))
        gw, gb = _grad(P["ca_w"]), _grad(P["ca_b"])
        Wq, Wk, Wv = P["ca_w"][:E], P["ca_w"][E:2 * E], P["ca_w"][2 * E:]
        bk, bv = P["ca_b"][E:2 * E], P["ca_b"][2 * E:]
        if absorbed:
            K.head_xattn_bwd(g["dctx2"], s["q2"], kin2, val2, Wk, bk, Wv, bv, s["P1"], s["z"], s["psum"], B, nq, N, g["dq2"], dkin, dval,
                             gw[E:2 * E], gb[E:2 * E], gw[2 * E:], gb[2 * E:], kpm=kpm, drop_u=u_of(li, 1), drop_p=p_attn)
        else:
            K.head_attn_small_bwd(g["dctx2"], s["q2"], s["kp"], s["vp"], s["P1"], g["dq2"], g["dkp"], g["dvp"], B, nq, N, H, scale,
                                  drop_u=u_of(li, 1), drop_p=p_attn)
            K.head_lin_bwd(g["dkp"], kin2, Wk, dx=dkin, dW=gw[E:2 * E], db=gb[E:2 * E])
            K.head_lin_bwd(g["dvp"], val2, Wv, dx=dval, dW=gw[2 * E:], db=gb[2 * E:])
        K.head_lin_bwd(g["dq2"], s["x1"], Wq, x2=qp, n_split=E, dx=g["dx1"], dx2=dqp, dW=gw[:E], db=gb[:E])
        # ---- self-attention
        K.head_lnres_bwd(g["dx1"], s["x"], s["a"], P["n0w"], s["m0"], s["r0"], _grad(P["n0w"]), _grad(P["n0b"]),
                         da=(bufs[per[li - 1]["dx3"]] if li > 0 else bufs[i_dx0]), db=g["da"])
        dx_prev = bufs[per[li - 1]["dx3"]] if li > 0 else bufs[i_dx0]
        K.head_lin_bwd(g["da"], s["ctx"], P["sa_ow"], dx=g["dctx"], dW=_grad(P["sa_ow"]), db=_grad(P["sa_ob"]))
        qkv, dqkv = s["qkv"], g["dqkv"]
        K.head_attn_small_bwd(g["dctx"], qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], s["P0"], dqkv[:, :E], dqkv[:, E:2 * E],
                              dqkv[:, 2 * E:], B, nq, nq, H, scale, drop_u=u_of(li, 0), drop_p=p_attn)
        K.head_lin_bwd(dqkv, s["x"], P["sa_w"], x2=qp, n_split=2 * E, dx=dx_prev, dx2=dqp, dW=_grad(P["sa_w"]), db=_grad(P["sa_b"]))
    return bufs[i_dx0].view(B, nq, E), dqp.view(B, nq, E), dkin.view(B, N, E), dval.view(B, N, E)


class _DecoderStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, query, qpos, kin, val, kpm, anchor):
        with K.nvtx("head.decoder.fwd"):
            out, saved = stack_forward(dec, query, qpos, kin, val, kpm, dec.training)
        ctx.dec, ctx.saved = dec, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.saved is None:
            raise RuntimeError("decoder stack backward called twice")
        with torch.no_grad(), K.nvtx("head.decoder.bwd"):
            dq, dqp, dkin, dval = stack_backward(ctx.dec, ctx.saved, dout)
        ctx.saved = None
        return None, dq, dqp, dkin, dval, None, torch.zeros(1, device=dout.device)


def decoder_stack(dec, query, kin, val, query_pos, key_padding_mask):
    """Native forward (+ autograd) of a DetrTransformerDecoder: -> [num_layers | 1, B, nq, E]."""
    K.L.require_device(query)
    need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in dec.parameters()) or query.requires_grad or
                                             query_pos.requires_grad or kin.requires_grad or val.requires_grad)
    if need_grad:
        anchor = torch.zeros(1, device=query.device, requires_grad=True)
        return _DecoderStackFn.apply(dec, query, query_pos, kin, val, key_padding_mask, anchor)
    with torch.no_grad():
        out, _ = stack_forward(dec, query, query_pos, kin, val, key_padding_mask, dec.training)
    return out


def supported(dec, query, kin):
    """Shapes the native kernels cover (everything SimVG configures): E = 256, 8 heads, nq <= 32, N > 32 or N <= 32."""
    E = query.shape[-1]
    layer = dec.layers[0]
    return (E == 256 and layer.attentions[0].num_heads == 8 and query.shape[1] <= 32 and
            layer.ffns[0].layers[0][0].weight.shape[0] % 32 == 0)


# ------------------------------------------------------------------------------------------------ nn.Linear / MLP on few rows
class _LinearFn(torch.autograd.Function):
    """y = relu?(x W^T + b) on simvgb_head_lin_fwd / _bwd (fp32).  dW / db are accumulated into the parameters' .grad in place."""

    @staticmethod
    def forward(ctx, x2d, lin, relu, anchor):
        y = K.head_lin_fwd(x2d, lin.weight, lin.bias, relu=relu)
        ctx.lin, ctx.relu = lin, relu
        ctx.save_for_backward(x2d, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, y = ctx.saved_tensors
        lin = ctx.lin
        dx = torch.zeros_like(x2d) if ctx.needs_input_grad[0] else None
        with torch.no_grad():
            K.head_lin_bwd(dy.contiguous().float(), x2d, lin.weight, y=y, relu=ctx.relu, dx=dx,
                           dW=_grad(lin.weight) if lin.weight.requires_grad else None,
                           db=_grad(lin.bias) if (lin.bias is not None and lin.bias.requires_grad) else None)
        return dx, None, None, torch.zeros(1, device=dy.device)


_linear_native = [True]    # tests / diagnostics: False routes head linears through torch's op-by-op path (the kernels' reference)


def linear(lin, x, relu=False):
    """nn.Linear `lin` (+ optional ReLU) applied to x [..., K] on the native head kernels (CUDA fp32 tensors)."""
    if not _linear_native[0]:
        y = lin(x)
        return torch.relu(y) if relu else y
    lead = x.shape[:-1]
    x2d = x.reshape(-1, x.shape[-1]).contiguous().float()
    if torch.is_grad_enabled() and (x2d.requires_grad or lin.weight.requires_grad):
        y = _LinearFn.apply(x2d, lin, relu, torch.zeros(1, device=x.device, requires_grad=True))
    else:
        y = K.head_lin_fwd(x2d, lin.weight, lin.bias, relu=relu)
    return y.view(*lead, lin.weight.shape[0])
