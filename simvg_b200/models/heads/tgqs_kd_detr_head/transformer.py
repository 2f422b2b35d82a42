"""Post-norm DETR decoder used by the object-token branch and by text-guided query generation.

Mirrors the module tree (and therefore the state-dict keys) of
/root/reference/simvg/models/heads/tgqs_kd_detr_head/transformer.py:93-235 built from detrex's BaseTransformerLayer /
MultiheadAttention / FFN / TransformerLayerSequence:
    layers.J.attentions.{0,1}.attn.{in_proj_weight,in_proj_bias,out_proj.*}, layers.J.ffns.0.layers.{0.0,1}.*,
    layers.J.norms.{0,1,2}.*, post_norm_layer.*
Layout is batch-first internally ([B, n, E]); the large key/value projections of the image memory ([B*N, E] rows) run on
the tcgen05 GEMM (simvg_b200.ops.linear), the per-query math (nq is 1..10) is latency-bound and stays in small fp32 ops.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from simvg_b200 import ops

_BIG_ROWS = 4096  # projections with at least this many rows go to the tensor-core GEMM


_ABSORB_MIN_KEYS = 256   # cross-attention against at least this many keys uses the absorbed-projection form


def _proj(x, W, b):
    if x.is_cuda and x.numel() // x.shape[-1] >= _BIG_ROWS:
        return ops.linear(x, W, b)
    return F.linear(x, W, b)


class _Attention(nn.Module):
    """detrex MultiheadAttention wrapper (A.9): positions are added to q/k only, output = identity + attn(...)."""

    def __init__(self, embed_dim, num_heads, attn_drop):
        super().__init__()
        self.embed_dim, self.num_heads, self.attn_drop = embed_dim, num_heads, attn_drop
        self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=attn_drop)  # parameter holder (+ its init)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, key_padding_mask=None,
                k_in=None):
        # batch-first: query [B, nq, E], key/value [B, nk, E], key_padding_mask [B, nk] (True = ignore)
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        q_in = query if query_pos is None else query + query_pos
        if k_in is None:   # the decoder passes key + key_pos precomputed once for all of its layers
            k_in = key if key_pos is None else key + key_pos
        E, H = self.embed_dim, self.num_heads
        dh = E // H
        # q/k/v parts of the packed in_proj parameters as ONE unbind each: three separate slices cost three zero-filled
        # full-size gradient buffers + copies + accumulations per attention in the backward (~300 launches per step)
        Wq, Wk, Wv = self.attn.in_proj_weight.view(3, E, E).unbind(0)
        bq, bk, bv = self.attn.in_proj_bias.view(3, E).unbind(0)
        B, nq, nk = q_in.shape[0], q_in.shape[1], k_in.shape[1]
        if nk == 1 and key_padding_mask is None:
            # One key (the REC decoder's self-attention: a single object query attends to itself): softmax over one score is
            # exactly 1 whatever q and k are, so out = out_proj(dropout(1) * v_proj(value)) and the q / k projections get exactly
            # zero gradient — as they do in the reference, which computes all of it (A.9).  Saves ~35 tiny launches per layer.
            v = F.linear(value, Wv, bv).view(B, 1, H, dh)
            if self.training and self.attn_drop > 0:
                v = v * F.dropout(torch.ones(B, 1, H, 1, device=v.device, dtype=v.dtype), p=self.attn_drop)
            o = v.expand(B, nq, H, dh).reshape(B, nq, E)
            o = F.linear(o, self.attn.out_proj.weight, self.attn.out_proj.bias)
            return identity + o
        q = F.linear(q_in, Wq, bq).view(B, nq, H, dh) * (dh ** -0.5)
        if nk >= _ABSORB_MIN_KEYS and nq * H <= 128:
            o = self._absorbed(q, k_in, value, Wk, bk, Wv, bv, key_padding_mask)
        else:
            k = _proj(k_in, Wk, bk)
            v = _proj(value, Wv, bv)
            q = q.transpose(1, 2)
            k = k.view(B, nk, H, dh).transpose(1, 2)
            v = v.view(B, nk, H, dh).transpose(1, 2)
            s = q @ k.transpose(-1, -2)
            if key_padding_mask is not None:
                s = s.masked_fill(key_padding_mask.view(B, 1, 1, nk), float("-inf"))
            p = F.softmax(s, dim=-1)
            if self.training and self.attn_drop > 0:
                p = F.dropout(p, p=self.attn_drop)
            o = (p @ v).transpose(1, 2).reshape(B, nq, E)
        o = F.linear(o, self.attn.out_proj.weight, self.attn.out_proj.bias)
        return identity + o  # proj_drop = 0

    def _absorbed(self, q, k_in, value, Wk, bk, Wv, bv, key_padding_mask):
        """Few queries against a long memory (the object-token decoder: nq = 1..10 queries, N = 1600 image tokens): the key and
        value projections are absorbed into the query / output side, so the [B*N, E] projected keys and values are never formed.
            scores[b,i,h,k] = q[b,i,h] . (Wk_h x_k[b,k] + bk_h) = (Wk_h^T q[b,i,h]) . x_k[b,k] + q[b,i,h] . bk_h
            out[b,i,h]      = sum_k p[b,i,h,k] (Wv_h x_v[b,k] + bv_h) = Wv_h (sum_k p x_v[b,k]) + bv_h sum_k p
        Same arithmetic as nn.MultiheadAttention (A.9), reassociated: per layer it reads the memory twice instead of running two
        [B*N, E] x [E, E] projections plus their transposes, casts and gradient accumulations (1.9 ms -> ~0.2 ms per layer at cfg2)."""
        B, nq, H, dh = q.shape
        E, nk = self.embed_dim, k_in.shape[1]
        Wk, bk = Wk.view(H, dh, E), bk.view(H, dh)
        Wv, bv = Wv.view(H, dh, E), bv.view(H, dh)
        u = torch.einsum("bihd,hde->bihe", q, Wk).reshape(B, nq * H, E)          # Wk_h^T q
        c = torch.einsum("bihd,hd->bih", q, bk).reshape(B, nq * H, 1)
        s = torch.baddbmm(c, u, k_in.transpose(1, 2))                              # [B, nq*H, nk]
        if key_padding_mask is not None:
            s = s.masked_fill(key_padding_mask.view(B, 1, nk), float("-inf"))
        p = F.softmax(s, dim=-1)
        if self.training and self.attn_drop > 0:
            p = F.dropout(p, p=self.attn_drop)
        z = torch.bmm(p, value).view(B, nq, H, E)                                  # sum_k p x_v
        psum = p.sum(-1).view(B, nq, H, 1)
        o = torch.einsum("bihe,hde->bihd", z, Wv) + psum * bv
        return o.reshape(B, nq, E)


class _FFN(nn.Module):
    """detrex FFN (A.9): Linear -> ReLU -> Dropout -> Linear -> Dropout, + identity."""

    def __init__(self, embed_dim, feedforward_dim, ffn_drop):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dim, feedforward_dim), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_dim, embed_dim), nn.Dropout(ffn_drop))

    def forward(self, x):
        return x + self.layers(x)


class _DecoderLayer(nn.Module):
    """operation_order = (self_attn, norm, cross_attn, norm, ffn, norm)  (transformer.py:122; A.10)."""

    def __init__(self, embed_dim, num_heads, attn_drop, feedforward_dim, ffn_drop):
        super().__init__()
        self.attentions = nn.ModuleList([_Attention(embed_dim, num_heads, attn_drop) for _ in range(2)])
        self.ffns = nn.ModuleList([_FFN(embed_dim, feedforward_dim, ffn_drop)])
        self.norms = nn.ModuleList([nn.LayerNorm(embed_dim) for _ in range(3)])
        self.embed_dim = embed_dim

    def forward(self, query, key, value, query_pos, key_pos, key_padding_mask, cross_k_in=None):
        query = self.norms[0](self.attentions[0](query, query, query, query_pos=query_pos, key_pos=query_pos))
        query = self.norms[1](self.attentions[1](query, key, value, query_pos=query_pos, key_pos=key_pos,
                                                 key_padding_mask=key_padding_mask, k_in=cross_k_in))
        return self.norms[2](self.ffns[0](query))


class DetrTransformerDecoder(nn.Module):
    def __init__(self, embed_dim=256, num_heads=8, attn_dropout=0.1, feedforward_dim=2048, ffn_dropout=0.1, num_layers=6,
                 post_norm=True, return_intermediate=True, batch_first=False):
        super().__init__()
        self.layers = nn.ModuleList([_DecoderLayer(embed_dim, num_heads, attn_dropout, feedforward_dim, ffn_dropout)
                                     for _ in range(num_layers)])
        self.num_layers = num_layers
        self.return_intermediate = return_intermediate
        self.embed_dim = embed_dim
        self.post_norm_layer = nn.LayerNorm(embed_dim) if post_norm else None
        self.use_native = True   # CUDA tensors run on the fused head kernels (native.py); False = op-by-op path (tests' reference)

    def forward(self, query, key, value, query_pos=None, key_pos=None, key_padding_mask=None):
        """Batch-first tensors.  Returns [num_layers | 1, B, nq, E]  (transformer.py:134-186: the shared post-norm is
        applied to every returned intermediate while the un-normed query feeds the next layer)."""
        inter = []
        # key + key_pos is the same for every layer (the memory is not updated by a decoder): form it once
        cross_k_in = key if key_pos is None else key + key_pos
        if query.is_cuda and self.use_native and query_pos is not None:
            from . import native
            if native.supported(self, query, cross_k_in):
                return native.decoder_stack(self, query, cross_k_in, value, query_pos, key_padding_mask)
        for layer in self.layers:
            query = layer(query, key, value, query_pos, key_pos, key_padding_mask, cross_k_in=cross_k_in)
            if self.return_intermediate:
                inter.append(self.post_norm_layer(query) if self.post_norm_layer is not None else query)
        if not self.return_intermediate:
            if self.post_norm_layer is not None:
                query = self.post_norm_layer(query)
            return query[None]
        return torch.stack(inter)


class DetrTransformerEncoder(nn.Module):
    """Constructed by the reference head and dropped when only_decoder=True (transformer.py:192-194); SimVG never runs
    it, so only the constructor signature is kept."""

    def __init__(self, embed_dim=256, num_heads=8, attn_dropout=0.1, feedforward_dim=2048, ffn_dropout=0.1, num_layers=6,
                 post_norm=False, batch_first=False):
        super().__init__()
        self.embed_dim, self.num_layers = embed_dim, num_layers

    def forward(self, *args, **kwargs):
        raise NotImplementedError("DetrTransformerEncoder is never executed by SimVG (every config builds the transformer with "
                                  "only_decoder=True, transformer.py:192-194); only its constructor is part of the surface")


class DetrTransformer(nn.Module):
    def __init__(self, encoder=None, decoder=None, only_decoder=False):
        super().__init__()
        if not only_decoder:
            raise NotImplementedError("SimVG configs use only_decoder=True (the BEiT-3 encoder replaces the DETR encoder)")
        self.decoder = decoder
        self.embed_dim = decoder.embed_dim
        self.only_decoder = only_decoder
        for p in self.parameters():  # transformer.py:200-203
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, memory, mask, query_embed, pos_embed):
        """memory / pos_embed: [B, N, E] token-major (the reference's [B,E,h,w] flattened), mask [B, N] bool,
        query_embed [B, nq, E].  Returns hidden states [layers, B, nq, E]."""
        target = torch.zeros_like(query_embed)
        return self.decoder(target, memory, memory, query_pos=query_embed, key_pos=pos_embed, key_padding_mask=mask)
