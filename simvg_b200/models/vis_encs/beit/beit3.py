"""BEIT3 — the registry-facing BEiT-3 multiway image-text encoder, backed by the sm_100a kernels.

Drop-in for the reference's `VIS_ENCODERS` entry `BEIT3` (/root/reference/simvg/models/vis_encs/beit/beit3.py:29-185):
same constructor kwargs, same `forward(image, question, padding_mask) -> (img_feat, text_feat, cls_feat)`, same
`.hidden_size / .beit3 / .get_num_layers() / .no_weight_decay()` surface (modeling_utils.py:73-109) and the same
state-dict keys.  The arithmetic of the reference's Encoder / EncoderLayer
(beit3_base.py:35-172,174-407) and of the torchscale leaf ops it imports is executed by
`encoder_forward` / `encoder_backward` below, which only sequence launches of libsimvg_b200 kernels.

Data layout (not a translation of the reference): vision tokens [B*Lv, D] and text tokens [B*Lt, D] live in two
separate token-major buffers for the whole encoder — the "multiway" experts A/B become two plain GEMM problems per op;
only the attention kernel sees both.  Residual stream fp32, GEMM operands bf16, accumulation fp32.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from simvg_b200 import kernels as K
from simvg_b200.flat import FlatBuffer
from simvg_b200.models.builder import VIS_ENCODERS

bf16, f32 = torch.bfloat16, torch.float32


# ------------------------------------------------------------------------------------------------ parameter holders
def _trunc_normal_(w, std):
    nn.init.trunc_normal_(w, mean=0.0, std=std, a=-std, b=std)  # modeling_utils.py:17-18


class _AB(nn.Module):
    """Multiway pair (torchscale MultiwayNetwork, A.3): expert A = vision tokens, expert B = text tokens."""

    def __init__(self, make):
        super().__init__()
        self.A = make()
        self.B = make()


class _SelfAttention(nn.Module):
    def __init__(self, D, eps):
        super().__init__()
        self.k_proj = _AB(lambda: nn.Linear(D, D))
        self.v_proj = _AB(lambda: nn.Linear(D, D))
        self.q_proj = _AB(lambda: nn.Linear(D, D))
        self.out_proj = _AB(lambda: nn.Linear(D, D))
        self.inner_attn_ln = _AB(lambda: nn.LayerNorm(D, eps=eps))


class _FFN(nn.Module):
    def __init__(self, D, F, eps):
        super().__init__()
        self.fc1 = nn.Linear(D, F)
        self.fc2 = nn.Linear(F, D)
        self.ffn_layernorm = nn.LayerNorm(F, eps=eps)


class _Layer(nn.Module):
    def __init__(self, D, F, eps):
        super().__init__()
        self.self_attn = _SelfAttention(D, eps)
        self.self_attn_layer_norm = _AB(lambda: nn.LayerNorm(D, eps=eps))
        self.ffn = _AB(lambda: _FFN(D, F, eps))
        self.final_layer_norm = _AB(lambda: nn.LayerNorm(D, eps=eps))


class _VisionEmbedding(nn.Module):
    def __init__(self, img_size, patch_size, D):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(3, D, kernel_size=patch_size, stride=patch_size)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, D))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))

    def num_position_embeddings(self):
        return self.num_patches + 1


class _Positions(nn.Module):
    def __init__(self, n_vis, n_text, D):
        super().__init__()
        self.A = nn.Embedding(n_vis, D)
        self.B = nn.Embedding(n_text, D)


class _Encoder(nn.Module):
    def __init__(self, n_layers, D, F, eps, n_pos_vis, n_pos_text):
        super().__init__()
        self.embed_positions = _Positions(n_pos_vis, n_pos_text, D)
        self.layers = nn.ModuleList([_Layer(D, F, eps) for _ in range(n_layers)])
        self.layer_norm = _AB(lambda: nn.LayerNorm(D, eps=eps))
        self.num_layers = n_layers


class _BEiT3(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D = cfg["embed_dim"]
        self.text_embed = nn.Embedding(cfg["vocab_size"], D)
        nn.init.normal_(self.text_embed.weight, mean=0, std=D ** -0.5)
        self.vision_embed = _VisionEmbedding(cfg["img_size"], cfg["patch_size"], D)
        self.encoder = _Encoder(cfg["layers"], D, cfg["ffn_dim"], cfg["eps"],
                                self.vision_embed.num_position_embeddings() + 2, cfg["max_source_positions"])


_GROUP_FIELDS = ("ln1_w", "ln1_b", "q_w", "k_w", "v_w", "q_b", "k_b", "v_b", "in_w", "in_b", "o_w", "o_b",
                 "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fl_w", "fl_b", "fc2_w", "fc2_b")


def _group_params(layer, which):
    sa = layer.self_attn
    g = lambda m: getattr(m, which)  # noqa: E731
    ffn = g(layer.ffn)
    return [g(layer.self_attn_layer_norm).weight, g(layer.self_attn_layer_norm).bias,
            g(sa.q_proj).weight, g(sa.k_proj).weight, g(sa.v_proj).weight,
            g(sa.q_proj).bias, g(sa.k_proj).bias, g(sa.v_proj).bias,
            g(sa.inner_attn_ln).weight, g(sa.inner_attn_ln).bias,
            g(sa.out_proj).weight, g(sa.out_proj).bias,
            g(layer.final_layer_norm).weight, g(layer.final_layer_norm).bias,
            ffn.fc1.weight, ffn.fc1.bias, ffn.ffn_layernorm.weight, ffn.ffn_layernorm.bias,
            ffn.fc2.weight, ffn.fc2.bias]


class _G:
    """fp32 / bf16 / grad views of one expert's parameters in one layer."""
    __slots__ = ("w", "wb", "g", "Wqkv", "bqkv", "gWqkv", "gbqkv")


@VIS_ENCODERS.register_module()
class BEIT3(nn.Module):
    def __init__(self, img_size=384, patch_size=32, vit_type="base", drop_path_rate=0.1, vocab_size=64010,
                 norm_layer=nn.LayerNorm, freeze_layer=-1, vision_embed_proj_interpolate=False, pretrain=None):
        super().__init__()
        if vit_type == "base":      # modeling_utils.py:21-44
            D, H, L = 768, 12, 12
            dpr = drop_path_rate
        elif vit_type == "large":   # modeling_utils.py:47-70; beit3.py:54 passes `rop_path_rate` -> DropPath stays 0
            D, H, L = 1024, 16, 24
            dpr = 0.0
        else:
            raise TypeError("please select the <vit_type> from ['base','large']")
        self.cfg = dict(img_size=img_size, patch_size=patch_size, vocab_size=vocab_size, embed_dim=D, heads=H, layers=L,
                        ffn_dim=4 * D, eps=1e-5, max_source_positions=1024, drop_path_rate=dpr)
        self.beit3 = _BEiT3(self.cfg)
        self.apply(self._init_weights)
        self.hidden_size = D
        self.vision_embed_proj_interpolate = vision_embed_proj_interpolate
        self.drop_path_probs = [float(p) for p in np.linspace(0, dpr, L)] if dpr > 0 else [0.0] * L
        self._flat = None
        self._attn_ws = {}
        self._defer_backward = False   # simvg_b200.runtime.GraphedTrainStep (N > 1): backward is stashed in _deferred, run in chunks
        self._deferred = None
        self._ddp = None   # set by simvg_b200.optim.FlatDDP: gradient ranges are all-reduced as they become final
        # Data-parallel runs exchange the text-embedding gradient as (ids, rows) instead of the dense [vocab, D] table:
        # FlatDDP sets "defer", the backward then leaves the compact form here (simvg_b200/optim.py::_exchange_text_rows).
        self.sparse_text_grad = {"defer": False, "ids": None, "rows": None}
        # GPU input path: a uint8 [B,S,S,3] batch (the dataset pipeline's image before `Normalize`) is normalised inside the
        # patch-embedding prologue with these constants — the reference's img_norm_cfg
        # (/root/reference/configs/_base_/datasets/detection/refcoco-unc.py:5-6; pipelines/transforms.py:126-155).
        self.input_norm = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
        if isinstance(pretrain, str):
            self.load_model_and_may_interpolate(pretrain)
        if freeze_layer >= 0:
            self.frozen_stages = min(freeze_layer, L)
            self._freeze_stages()

    # ---- reference surface -------------------------------------------------------------------
    def _init_weights(self, m):  # modeling_utils.py:102-109
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_num_layers(self):
        return self.beit3.encoder.num_layers

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "beit3.encoder.embed_positions.A.weight", "beit3.vision_embed.cls_token",
                "logit_scale"}

    def _freeze_stages(self):  # beit3.py:78-90
        for i in range(1, self.frozen_stages + 1):
            m = self.beit3.encoder.layers[i - 1]
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def load_model_and_may_interpolate(self, ckpt_path, model_key="model|module", model_prefix=""):
        """Checkpoint load with bicubic pos-embed / patch-proj interpolation (beit3.py:92-174)."""
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)   # BEiT-3 releases are plain pickled dicts
        sd = None
        for key in model_key.split("|"):
            if key in ckpt:
                sd = ckpt[key]
                break
        if sd is None:
            sd = ckpt
        pk = "beit3.encoder.embed_positions.A.weight"
        if pk in sd:
            pe = sd[pk]
            n_patches = self.beit3.vision_embed.num_patches
            extra = self.beit3.vision_embed.num_position_embeddings() + 2 - n_patches
            orig = int((pe.shape[-2] - extra) ** 0.5)
            new = int(n_patches ** 0.5)
            if orig != new:
                tok = pe[extra:].reshape(-1, orig, orig, pe.shape[-1]).permute(0, 3, 1, 2).float()
                tok = torch.nn.functional.interpolate(tok, size=(new, new), mode="bicubic", align_corners=False)
                tok = tok.permute(0, 2, 3, 1).flatten(1, 2)
                sd[pk] = torch.cat((pe[:extra].unsqueeze(0), tok), dim=1).squeeze(0)
        wk = "beit3.vision_embed.proj.weight"
        if wk in sd and sd[wk].shape != self.beit3.vision_embed.proj.weight.shape and self.vision_embed_proj_interpolate:
            sd[wk] = torch.nn.functional.interpolate(sd[wk].float(), size=self.beit3.vision_embed.proj.weight.shape[-2:],
                                                     mode="bicubic", align_corners=False)
        missing, unexpected = self.load_state_dict(sd, strict=False)
        return missing, unexpected

    # ---- flat storage ------------------------------------------------------------------------
    def flat(self):
        """Flat fp32 storage of all encoder parameters (q/k/v weights and biases of an expert adjacent)."""
        if self._flat is None:
            b = self.beit3
            order = [("text_embed", b.text_embed.weight), ("proj_w", b.vision_embed.proj.weight),
                     ("proj_b", b.vision_embed.proj.bias), ("mask_token", b.vision_embed.mask_token),
                     ("cls_token", b.vision_embed.cls_token), ("posA", b.encoder.embed_positions.A.weight),
                     ("posB", b.encoder.embed_positions.B.weight),
                     ("fin_A_w", b.encoder.layer_norm.A.weight), ("fin_A_b", b.encoder.layer_norm.A.bias),
                     ("fin_B_w", b.encoder.layer_norm.B.weight), ("fin_B_b", b.encoder.layer_norm.B.bias)]
            self._n_global = len(order)
            for li, layer in enumerate(b.encoder.layers):
                for which in ("A", "B"):
                    for fld, p in zip(_GROUP_FIELDS, _group_params(layer, which)):
                        order.append(("l%d.%s.%s" % (li, which, fld), p))
            self._flat = FlatBuffer(order)
        return self._flat.ensure()

    def _views(self, need_grad):
        """Per-layer, per-expert parameter views (fp32 data, bf16 shadow, fp32 grad)."""
        fb = self.flat()
        if fb.shadow is None or fb.shadow.device != fb.data.device:
            fb.shadow = torch.empty(fb.numel, device=fb.data.device, dtype=bf16)
        K.cast_bf16(fb.data, fb.shadow)
        if need_grad:
            fb.attach_grads()
        D = self.cfg["embed_dim"]
        nf = len(_GROUP_FIELDS)
        glob = {}
        for i in range(self._n_global):
            glob[fb.names[i]] = (fb.params[i].data, fb.view(i, fb.shadow), fb.grad_of(i) if need_grad else None)
        layers = []
        idx = self._n_global
        for _ in range(self.cfg["layers"]):
            pair = []
            for _g in range(2):
                G = _G()
                G.w = {f: fb.params[idx + j].data for j, f in enumerate(_GROUP_FIELDS)}
                G.wb = {f: fb.view(idx + j, fb.shadow) for j, f in enumerate(_GROUP_FIELDS)}
                G.g = {f: fb.grad_of(idx + j) for j, f in enumerate(_GROUP_FIELDS)} if need_grad else None
                qo = fb.offsets[idx + 2]
                bo = fb.offsets[idx + 5]
                G.Wqkv = fb.shadow[qo:qo + 3 * D * D].view(3 * D, D)
                G.bqkv = fb.data[bo:bo + 3 * D]
                if need_grad:
                    G.gWqkv = fb.grad[qo:qo + 3 * D * D].view(3 * D, D)
                    G.gbqkv = fb.grad[bo:bo + 3 * D]
                pair.append(G)
                idx += nf
            layers.append(pair)
        return glob, layers

    # ---- forward -----------------------------------------------------------------------------
    def forward(self, image, question, padding_mask, **kwargs):
        """-> (img_feat [B,N,D], text_feat [B,Lt,D], cls_feat [B,D])   (beit3.py:176-185)."""
        K.L.require_device(image)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad:
            anchor = torch.zeros(1, device=image.device, requires_grad=True)
            xv, xt = _EncoderFn.apply(self, image, question, padding_mask, anchor)
        else:
            xv, xt, _ = encoder_forward(self, image, question, padding_mask, save=False)
        cls_feat, img_feat, text_feat = xv[:, 0], xv[:, 1:], xt
        return img_feat, text_feat, cls_feat


# ------------------------------------------------------------------------------------------------ kernel sequencing
def _drop_path_scales(mod, B, device):
    """Per layer: (attn_scale[B] | None, ffn_scale[B] | None) = Bernoulli(keep)/keep  (timm drop_path, A.8).
    All layers are drawn with one bernoulli launch over a [layers, 2, B] keep-probability tensor."""
    probs = mod.drop_path_probs
    if not mod.training or not any(p > 0.0 for p in probs):
        return [(None, None)] * len(probs)
    keep = getattr(mod, "_dp_keep", None)
    if keep is None or keep.device != device or keep.shape[2] != B:
        keep = torch.tensor([1.0 - p for p in probs], dtype=f32).view(-1, 1, 1).expand(len(probs), 2, B).contiguous().to(device)
        mod._dp_keep = keep
    s = torch.bernoulli(keep).div_(keep)
    return [((s[i, 0], s[i, 1]) if p > 0.0 else (None, None)) for i, p in enumerate(probs)]


def encoder_forward(mod, image, ids, pad_mask, save):
    cfg = mod.cfg
    D, H, F, P, eps = cfg["embed_dim"], cfg["heads"], cfg["ffn_dim"], cfg["patch_size"], cfg["eps"]
    raw_u8 = image.dtype == torch.uint8
    if raw_u8:
        B, S, S2, _ = image.shape
    else:
        B, _, S, S2 = image.shape
    assert S == cfg["img_size"] and S2 == cfg["img_size"], \
        "Input image size (%d*%d) doesn't match model (%d*%d)." % (S, S2, cfg["img_size"], cfg["img_size"])  # A.6
    N = (S // P) ** 2
    Lv, Lt = N + 1, ids.shape[1]
    if Lt == 0:
        raise NotImplementedError("BEIT3 (SimVG) always attends over image + text tokens")
    Rs = (B * Lv, B * Lt)
    Ls = (Lv, Lt)
    glob, layers = mod._views(need_grad=save)
    pad_u8 = None if pad_mask is None else (pad_mask != 0).to(torch.uint8).contiguous()
    ids = ids.contiguous().long()
    if raw_u8:
        nrm = mod.input_norm
        cols = K.im2col_patch_u8(image.contiguous(), P, nrm["mean"], nrm["std"], nrm.get("to_rgb", True))
    else:
        cols = K.im2col_patch(image.contiguous().float(), P)
    patch = K.gemm(cols, glob["proj_w"][1].view(D, 3 * P * P), B * N, D, 3 * P * P, epilogue=K.EPI_F32, bias=glob["proj_b"][0])
    x = [K.embed_vision(patch, glob["cls_token"][0], glob["posA"][0], B, N, D),
         K.embed_text(glob["text_embed"][0], ids, pad_u8, glob["posB"][0], B, Lt, D)]
    dps = _drop_path_scales(mod, B, image.device)
    saved = []
    scale_q = (D // H) ** -0.5
    for li, pair in enumerate(layers):
        K.nvtx_push("encoder.layers.%d.fwd" % li)
        dp1, dp2 = dps[li]
        dpv = (dp1, dp2)
        sv = [dict(), dict()]
        # Every projection is one launch for BOTH experts (K.gemm_pair): the text problem (B*Lt rows) rides in the tail of
        # the vision problem's persistent grid instead of paying its own launch + ramp.
        h, st1 = [None, None], [None, None]
        for g, G in enumerate(pair):
            h[g], m1, r1 = K.ln_fwd(x[g], G.w["ln1_w"], G.w["ln1_b"], eps)
            st1[g] = (m1, r1)
        qkv = K.gemm_pair(*[(h[g], G.Wqkv, Rs[g], 3 * D, D, dict(epilogue=K.EPI_BF16, bias=G.bqkv, scale=scale_q, scale_cols=D))
                            for g, G in enumerate(pair)])
        for g in range(2):
            sv[g].update(x_in=x[g], h=h[g], m1=st1[g][0], r1=st1[g][1], qkv=qkv[g])
        o_v, o_t, lse = K.attn_fwd(qkv[0], qkv[1], pad_u8, B, H, Lv, Lt)
        o = (o_v, o_t)
        a, sti = [None, None], [None, None]
        for g, G in enumerate(pair):
            a[g], mi, ri = K.ln_fwd(o[g], G.w["in_w"], G.w["in_b"], eps)
            sti[g] = (mi, ri)
        xmid = K.gemm_pair(*[(a[g], G.wb["o_w"], Rs[g], D, D, dict(epilogue=K.EPI_RESID, bias=G.w["o_b"], res=x[g], row_scale=dpv[0],
                                                               rows_per_scale=Ls[g])) for g, G in enumerate(pair)])
        h2, st2 = [None, None], [None, None]
        for g, G in enumerate(pair):
            h2[g], m2, r2 = K.ln_fwd(xmid[g], G.w["ln2_w"], G.w["ln2_b"], eps)
            st2[g] = (m2, r2)
        # fc1 + bias; the exact-erf GELU is applied inside the FFN LayerNorm kernel (LN_F(gelu(u))) and recomputed in the
        # backward, so the activation itself is never written to HBM.
        u = K.gemm_pair(*[(h2[g], G.wb["fc1_w"], Rs[g], F, D, dict(epilogue=K.EPI_BF16, bias=G.w["fc1_b"])) for g, G in enumerate(pair)])
        f, stf = [None, None], [None, None]
        for g, G in enumerate(pair):
            f[g], mf, rf = K.ln_fwd(u[g], G.w["fl_w"], G.w["fl_b"], eps, gelu=True)
            stf[g] = (mf, rf)
        xn = K.gemm_pair(*[(f[g], G.wb["fc2_w"], Rs[g], D, F, dict(epilogue=K.EPI_RESID, bias=G.w["fc2_b"], res=xmid[g], row_scale=dpv[1],
                                                             rows_per_scale=Ls[g])) for g, G in enumerate(pair)])
        for g in range(2):
            if save:
                sv[g].update(o=o[g], a=a[g], mi=sti[g][0], ri=sti[g][1], xmid=xmid[g], h2=h2[g], m2=st2[g][0], r2=st2[g][1],
                             u=u[g], f=f[g], mf=stf[g][0], rf=stf[g][1])
            x[g] = xn[g]
        if save:
            saved.append(dict(g=sv, lse=lse, dp=(dp1, dp2)))
        K.nvtx_pop()
    outs, fin = [], []
    for g, which in enumerate(("A", "B")):
        y, mF, rF = K.ln_fwd(x[g], glob["fin_%s_w" % which][0], glob["fin_%s_b" % which][0], eps, out_dtype=f32)
        outs.append(y)
        fin.append((x[g], mF, rF))
    ctx = None
    if save:
        ctx = dict(layers=saved, fin=fin, cols=cols, ids=ids, pad=pad_u8, shape=(B, N, Lv, Lt), views=(glob, layers))
    return outs[0].view(B, Lv, D), outs[1].view(B, Lt, D), ctx


def layer_flat_range(mod, li):
    """Flat-buffer element range [lo, hi) holding every parameter of encoder layer li (both experts).  The buffer is laid out
    [embeddings + final LayerNorms | layer 0 | layer 1 | ...], so range(0)[0] is also the end of the global parameters."""
    nf = len(_GROUP_FIELDS)
    fb = mod.flat()
    i0 = mod._n_global + li * 2 * nf
    i1 = i0 + 2 * nf
    return fb.offsets[i0], (fb.offsets[i1] if i1 < len(fb.offsets) else fb.numel)


class EncoderBackward:
    """The encoder backward as a resumable sequence: start() (final LayerNorms), layers(hi, lo) (any contiguous block of layers,
    top down), finish() (embeddings).  `encoder_backward` runs it in one go; the multi-GPU graph runtime
    (simvg_b200/runtime.py) captures it in chunks so that the gradient exchange of one chunk overlaps the next chunk's kernels.
    Writes (accumulates) every encoder parameter gradient into the flat gradient buffer."""

    def __init__(self, mod, ctx, dxv, dxt):
        self.mod, self.ctx = mod, ctx
        cfg = mod.cfg
        self.D, self.H, self.F, self.P = cfg["embed_dim"], cfg["heads"], cfg["ffn_dim"], cfg["patch_size"]
        self.B, self.N, self.Lv, self.Lt = ctx["shape"]
        self.Rs, self.Ls = (self.B * self.Lv, self.B * self.Lt), (self.Lv, self.Lt)
        self.fb = mod.flat()
        self.glob, self.layer_views = ctx["views"]
        self.dev = dxv.device
        self.ddp = getattr(mod, "_ddp", None)
        self.nl = len(self.layer_views)
        self.dys = (dxv.reshape(self.Rs[0], self.D).contiguous().float(), dxt.reshape(self.Rs[1], self.D).contiguous().float())
        self.dres = [None, None]
        self.dyb = None

    def layer_range(self, li):
        return layer_flat_range(self.mod, li)

    def start(self):
        ctx, glob, layers, nl, Rs, Ls, D, dev = self.ctx, self.glob, self.layer_views, self.nl, self.Rs, self.Ls, self.D, self.dev
        if self.ddp is not None:
            self.ddp.on_encoder_backward_start()
        self.dyb = [torch.empty(Rs[g], D, device=dev, dtype=bf16) for g in range(2)]
        for g, which in enumerate(("A", "B")):
            xf, mF, rF = ctx["fin"][g]
            Gl = layers[nl - 1][g]
            self.dres[g] = torch.empty(Rs[g], D, device=dev, dtype=f32)
            K.ln_bwd(0, xf, self.dys[g], glob["fin_%s_w" % which][0], mF, rF, glob["fin_%s_w" % which][2], glob["fin_%s_b" % which][2],
                     dres_in=None, dres_out=self.dres[g], dyb=self.dyb[g], row_scale=ctx["layers"][nl - 1]["dp"][1],
                     rows_per_scale=Ls[g], dbias_prev=Gl.g["fc2_b"])

    def layers(self, hi, lo):
        mod, ctx, layers, Rs, Ls, dev = self.mod, self.ctx, self.layer_views, self.Rs, self.Ls, self.dev
        D, H, F, B, Lv, Lt = self.D, self.H, self.F, self.B, self.Lv, self.Lt
        dres, dyb, fb, ddp = self.dres, self.dyb, self.fb, self.ddp
        nf = len(_GROUP_FIELDS)
        for li in range(hi, lo - 1, -1):
            K.nvtx_push("encoder.layers.%d.bwd" % li)
            sl = ctx["layers"][li]
            dp1, _dp2 = sl["dp"]
            Gs = layers[li]
            svs = sl["g"]
            # ---- FFN:  x = xmid + dp2 * fc2(LN_F(gelu(fc1(LN2(xmid)))))      (each GEMM: both experts in one launch)
            K.wgrad_pair(*[(dyb[g], svs[g]["f"], D, F, Rs[g], Gs[g].g["fc2_w"]) for g in range(2)])
            df = K.gemm_pair(*[(dyb[g], Gs[g].wb["fc2_w"], Rs[g], F, D, dict(b_mn=True, epilogue=K.EPI_BF16)) for g in range(2)])
            du = [torch.empty(Rs[g], F, device=dev, dtype=bf16) for g in range(2)]
            for g, G in enumerate(Gs):
                K.ln_bwd(2, None, df[g], G.w["fl_w"], svs[g]["mf"], svs[g]["rf"], G.g["fl_w"], G.g["fl_b"], dx=du[g], u=svs[g]["u"],
                         dbias_prev=G.g["fc1_b"])
            del df
            K.wgrad_pair(*[(du[g], svs[g]["h2"], F, D, Rs[g], Gs[g].g["fc1_w"]) for g in range(2)])
            dh2 = K.gemm_pair(*[(du[g], Gs[g].wb["fc1_w"], Rs[g], D, F, dict(b_mn=True, epilogue=K.EPI_BF16)) for g in range(2)])
            del du
            for g, G in enumerate(Gs):
                K.ln_bwd(0, svs[g]["xmid"], dh2[g], G.w["ln2_w"], svs[g]["m2"], svs[g]["r2"], G.g["ln2_w"], G.g["ln2_b"], dres_in=dres[g],
                         dres_out=dres[g], dyb=dyb[g], row_scale=dp1, rows_per_scale=Ls[g], dbias_prev=G.g["o_b"])
            del dh2
            # ---- attention output:  xmid = x_in + dp1 * out_proj(LN_inner(O))
            K.wgrad_pair(*[(dyb[g], svs[g]["a"], D, D, Rs[g], Gs[g].g["o_w"]) for g in range(2)])
            da = K.gemm_pair(*[(dyb[g], Gs[g].wb["o_w"], Rs[g], D, D, dict(b_mn=True, epilogue=K.EPI_BF16)) for g in range(2)])
            dO = [torch.empty(Rs[g], D, device=dev, dtype=bf16) for g in range(2)]
            ws = K.attn_workspace(mod._attn_ws, B, H, Lv, Lt, dev)
            for g, G in enumerate(Gs):
                # the inner-attention-LN backward also emits delta = rowsum(O o dO) per head for the attention backward
                K.ln_bwd(1, svs[g]["o"], da[g], G.w["in_w"], svs[g]["mi"], svs[g]["ri"], G.g["in_w"], G.g["in_b"], dx=dO[g],
                         delta=K.attn_delta_spec(ws, B, H, Lv, Lt, g))
            del da
            svv, svt = svs
            dqkv = K.attn_bwd(svv["qkv"], svt["qkv"], ctx["pad"], svv["o"], svt["o"], sl["lse"], dO[0], dO[1], B, H, Lv, Lt,
                              ws=ws, delta_ready=True)
            for g, G in enumerate(Gs):
                K.colsum(dqkv[g], out=G.gbqkv)
            K.wgrad_pair(*[(dqkv[g], svs[g]["h"], 3 * D, D, Rs[g], Gs[g].gWqkv) for g in range(2)])
            dh = K.gemm_pair(*[(dqkv[g], Gs[g].Wqkv, Rs[g], D, 3 * D, dict(b_mn=True, epilogue=K.EPI_BF16)) for g in range(2)])
            for g, G in enumerate(Gs):
                sv = svs[g]
                if li > 0:
                    Gp = layers[li - 1][g]
                    K.ln_bwd(0, sv["x_in"], dh[g], G.w["ln1_w"], sv["m1"], sv["r1"], G.g["ln1_w"], G.g["ln1_b"], dres_in=dres[g],
                             dres_out=dres[g], dyb=dyb[g], row_scale=ctx["layers"][li - 1]["dp"][1], rows_per_scale=Ls[g],
                             dbias_prev=Gp.g["fc2_b"])
                else:
                    K.ln_bwd(0, sv["x_in"], dh[g], G.w["ln1_w"], sv["m1"], sv["r1"], G.g["ln1_w"], G.g["ln1_b"], dres_in=dres[g],
                             dres_out=dres[g])
            del dh
            ctx["layers"][li] = None  # release this layer's activations
            if li > 0:
                # every gradient of layer li is final now (its fc2 bias was written by layer li+1's LN1 backward)
                i0 = mod._n_global + li * 2 * nf
                i1 = i0 + 2 * nf
                _zero_frozen(fb, i0, i1)   # BEFORE the range is handed to the (asynchronous, in-place) all-reduce
            if ddp is not None and li > 0:
                ddp.on_encoder_range_done(fb.offsets[i0], fb.offsets[i1] if i1 < len(fb.offsets) else fb.numel)
            K.nvtx_pop()

    def finish(self):
        mod, ctx, glob, dres, fb, ddp, dev = self.mod, self.ctx, self.glob, self.dres, self.fb, self.ddp, self.dev
        B, N, Lv, Lt, D, P = self.B, self.N, self.Lv, self.Lt, self.D, self.P
        nf = len(_GROUP_FIELDS)
        # ---- embeddings (Encoder.forward_embedding, beit3_base.py:317-334 ; VisionEmbedding / TextEmbedding A.6-A.7)
        dv = dres[0].view(B, Lv, D)
        glob["posA"][2][2:2 + Lv].add_(dv.sum(0))
        glob["cls_token"][2].view(D).add_(dv[:, 0].sum(0))
        dpatch = torch.empty(B * N, D, device=dev, dtype=bf16)
        K.colsum(dv[:, 1:].reshape(B * N, D), out=glob["proj_b"][2], out_bf16=dpatch)
        K.wgrad(dpatch, ctx["cols"], D, 3 * P * P, B * N, out=glob["proj_w"][2].view(D, 3 * P * P))
        dt = dres[1].view(B, Lt, D)
        if ctx["pad"] is not None:
            dt = dt * (1.0 - ctx["pad"].view(B, Lt, 1).float())
        glob["posB"][2][2:2 + Lt].add_(dt.sum(0))
        st = mod.sparse_text_grad
        if st["defer"]:   # data parallel: ranks all-gather these <= B*Lt rows instead of all-reducing the dense table
            st["ids"], st["rows"] = ctx["ids"].reshape(-1), dt.reshape(B * Lt, D).contiguous()
        else:
            glob["text_embed"][2].index_add_(0, ctx["ids"].reshape(-1), dt.reshape(B * Lt, D))
        _zero_frozen(fb, 0, mod._n_global + 2 * nf)
        if ddp is not None:   # layer 0 + the embedding / final-LN parameters at the front of the buffer
            i1 = mod._n_global + 2 * nf
            ddp.on_encoder_range_done(0, fb.offsets[i1] if i1 < len(fb.offsets) else fb.numel)


def encoder_backward(mod, ctx, dxv, dxt):
    """Writes (accumulates) every encoder parameter gradient into the flat gradient buffer."""
    eb = EncoderBackward(mod, ctx, dxv, dxt)
    eb.start()
    eb.layers(eb.nl - 1, 0)
    eb.finish()


def _zero_frozen(fb, i0, i1):
    """Frozen parameters (BEIT3(freeze_layer=k), beit3.py:78-90) keep zero gradients; the optimiser skips them entirely."""
    for i in range(i0, min(i1, len(fb.params))):
        if not fb.params[i].requires_grad:
            fb.grad_of(i).zero_()


class _EncoderFn(torch.autograd.Function):
    """Autograd node for the whole encoder.  Parameter gradients are accumulated straight into the flat gradient buffer
    (each p.grad is a view of it) by the backward kernels; `anchor` only keeps the node in the graph."""

    @staticmethod
    def forward(ctx, mod, image, ids, pad_mask, anchor):
        xv, xt, saved = encoder_forward(mod, image, ids, pad_mask, save=True)
        ctx.mod = mod
        ctx.saved = saved
        return xv, xt

    @staticmethod
    def backward(ctx, dxv, dxt):
        mod, saved = ctx.mod, ctx.saved
        if saved is None:
            raise RuntimeError("BEIT3 backward called twice (activations are released after the first pass)")
        B, N, Lv, Lt = saved["shape"]
        D = mod.cfg["embed_dim"]
        dev = saved["cols"].device
        if dxv is None:
            dxv = torch.zeros(B, Lv, D, device=dev)
        if dxt is None:
            dxt = torch.zeros(B, Lt, D, device=dev)
        with torch.no_grad():
            if getattr(mod, "_defer_backward", False):
                # the multi-GPU graph runtime runs the encoder backward itself, in separately captured chunks
                mod._deferred = EncoderBackward(mod, saved, dxv, dxt)
            else:
                encoder_backward(mod, saved, dxv, dxt)
        ctx.saved = None
        return None, None, None, None, torch.zeros(1, device=dev)
