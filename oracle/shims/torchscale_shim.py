"""torchscale (unpinned, requirements.txt:13; API of the 0.2.0 era) leaf ops restated for the shim-import of the
reference — SURVEY Appendix A.1-A.8.  TEST INFRASTRUCTURE ONLY."""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F


class EncoderConfig:
    """A.1 — only the fields the reference reads; unknown kwargs are ignored (which is why `rop_path_rate=` at
    beit3.py:54 silently leaves ViT-L without DropPath)."""

    def __init__(self, **kw):
        g = kw.pop
        self.encoder_embed_dim = g("encoder_embed_dim", 768)
        self.encoder_attention_heads = g("encoder_attention_heads", 12)
        self.encoder_ffn_embed_dim = g("encoder_ffn_embed_dim", 3072)
        self.encoder_layers = g("encoder_layers", 12)
        self.encoder_normalize_before = g("encoder_normalize_before", True)
        self.normalize_output = g("normalize_output", True)
        self.activation_fn = g("activation_fn", "gelu")
        self.dropout = g("dropout", 0.0)
        self.drop_path_rate = g("drop_path_rate", 0.0)
        self.attention_dropout = g("attention_dropout", 0.0)
        self.activation_dropout = g("activation_dropout", 0.0)
        self.no_scale_embedding = g("no_scale_embedding", True)
        self.layernorm_embedding = g("layernorm_embedding", False)
        self.moe_freq = g("moe_freq", 0)
        self.moe_top1_expert = g("moe_top1_expert", False)
        self.moe_expert_count = g("moe_expert_count", 0)
        self.rel_pos_buckets = g("rel_pos_buckets", 0)
        self.max_rel_pos = g("max_rel_pos", 0)
        self.deepnorm = g("deepnorm", False)
        self.subln = g("subln", True)
        self.bert_init = g("bert_init", False)
        self.multiway = g("multiway", False)
        self.share_encoder_input_output_embed = g("share_encoder_input_output_embed", False)
        self.max_source_positions = g("max_source_positions", 1024)
        self.no_output_layer = g("no_output_layer", False)
        self.layernorm_eps = g("layernorm_eps", 1e-5)
        self.vocab_size = g("vocab_size", -1)
        self.img_size = g("img_size", 224)
        self.patch_size = g("patch_size", 16)
        self.in_chans = g("in_chans", 3)
        self.checkpoint_activations = g("checkpoint_activations", False)
        self.fsdp = g("fsdp", False)
        self.ddp_rank = g("ddp_rank", 0)
        self.xpos_rel_pos = g("xpos_rel_pos", False)
        self.xpos_scale_base = g("xpos_scale_base", 512)
        if self.deepnorm:
            self.encoder_normalize_before = False
            self.subln = False
        if self.subln:
            self.encoder_normalize_before = True
            self.deepnorm = False


def init_bert_params(module):
    pass  # bert_init=False on this path


class Unused(nn.Module):
    def __init__(self, *a, **k):
        raise RuntimeError("this torchscale component is not on SimVG's hot path (moe_freq=0, rel_pos_buckets=0)")


# ---- A.3 multiway
class MultiwayNetwork(nn.Module):
    def __init__(self, module, dim=1):
        super().__init__()
        self.dim = dim
        self.A = module
        self.B = copy.deepcopy(module)
        self.B.reset_parameters()
        self.split_position = -1

    def forward(self, x, **kwargs):
        if self.split_position == -1:
            return self.A(x, **kwargs)
        if self.split_position == 0:
            return self.B(x, **kwargs)
        x1, x2 = torch.split(x, [self.split_position, x.size(self.dim) - self.split_position], dim=self.dim)
        return torch.cat([self.A(x1, **kwargs), self.B(x2, **kwargs)], dim=self.dim)


class MutliwayEmbedding(MultiwayNetwork):
    def __init__(self, modules, dim=1):
        nn.Module.__init__(self)
        self.dim = dim
        assert len(modules) == 2
        self.A, self.B = modules
        self.split_position = -1


def MultiwayWrapper(args, module, dim=1):
    return MultiwayNetwork(module, dim=dim) if args.multiway else module


def set_split_position(position):
    def apply_fn(module):
        if hasattr(module, "split_position"):
            module.split_position = position
    return apply_fn


# ---- A.6 / A.7 embeddings
class VisionEmbedding(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, contain_mask_token=False,
                 prepend_cls_token=False):
        super().__init__()
        img_size, patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.patch_shape = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.img_size, self.patch_size = img_size, patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if contain_mask_token else None
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if prepend_cls_token else None

    def num_position_embeddings(self):
        return self.num_patches if self.cls_token is None else self.num_patches + 1

    def forward(self, x, masked_position=None, **kwargs):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x = self.proj(x).flatten(2).transpose(1, 2)
        batch_size, seq_len, _ = x.size()
        if masked_position is not None:
            assert self.mask_token is not None
            mask_token = self.mask_token.expand(batch_size, seq_len, -1)
            w = masked_position.unsqueeze(-1).type_as(mask_token)
            x = x * (1 - w) + mask_token * w
        if self.cls_token is not None:
            x = torch.cat((self.cls_token.expand(batch_size, -1, -1), x), dim=1)
        return x


class TextEmbedding(nn.Embedding):
    def reset_parameters(self):
        nn.init.normal_(self.weight, mean=0, std=self.embedding_dim ** -0.5)
        self._fill_padding_idx_with_zero()


class PositionalEmbedding(nn.Embedding):
    def forward(self, x, positions=None, **kwargs):
        if positions is None:
            # being consistent with Fairseq, which starts from 2.
            positions = torch.arange(2, x.size(1) + 2, device=x.device).long().unsqueeze(0)
        return F.embedding(positions, self.weight, self.padding_idx, self.max_norm, self.norm_type,
                           self.scale_grad_by_freq, self.sparse)


# ---- A.8 DropPath (timm drop_path)
class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


# ---- A.5 FFN
class FeedForwardNetwork(nn.Module):
    def __init__(self, embed_dim, ffn_dim, activation_fn, dropout, activation_dropout, layernorm_eps, subln=False):
        super().__init__()
        assert activation_fn == "gelu"
        self.embed_dim = embed_dim
        self.activation_dropout_module = nn.Dropout(activation_dropout)
        self.dropout_module = nn.Dropout(dropout)
        self.fc1 = nn.Linear(self.embed_dim, ffn_dim)
        self.fc2 = nn.Linear(ffn_dim, self.embed_dim)
        self.ffn_layernorm = nn.LayerNorm(ffn_dim, eps=layernorm_eps) if subln else None

    def reset_parameters(self):
        self.fc1.reset_parameters()
        self.fc2.reset_parameters()
        if self.ffn_layernorm is not None:
            self.ffn_layernorm.reset_parameters()

    def forward(self, x):
        x_shape = x.shape
        x = x.reshape(-1, x.size(-1))
        x = self.fc1(x)
        x = F.gelu(x.float()).type_as(x)
        x = self.activation_dropout_module(x)
        if self.ffn_layernorm is not None:
            x = self.ffn_layernorm(x)
        x = self.fc2(x)
        x = x.view(x_shape)
        return self.dropout_module(x)


def make_experts(*a, **k):
    raise RuntimeError("MoE is not on SimVG's hot path")


# ---- A.4 attention
class MultiheadAttention(nn.Module):
    def __init__(self, args, embed_dim, num_heads, dropout=0.0, self_attention=False, encoder_decoder_attention=False,
                 subln=False):
        super().__init__()
        self.args = args
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.scaling = self.head_dim ** -0.5
        self.self_attention, self.encoder_decoder_attention = self_attention, encoder_decoder_attention
        assert self.self_attention ^ self.encoder_decoder_attention
        self.k_proj = MultiwayWrapper(args, nn.Linear(embed_dim, embed_dim, bias=True))
        self.v_proj = MultiwayWrapper(args, nn.Linear(embed_dim, embed_dim, bias=True))
        self.q_proj = MultiwayWrapper(args, nn.Linear(embed_dim, embed_dim, bias=True))
        self.out_proj = MultiwayWrapper(args, nn.Linear(embed_dim, embed_dim, bias=True))
        self.inner_attn_ln = MultiwayWrapper(args, nn.LayerNorm(self.embed_dim, eps=args.layernorm_eps)) \
            if subln and self.self_attention else None
        self.dropout_module = nn.Dropout(dropout)
        self.xpos = None

    def forward(self, query, key, value, incremental_state=None, key_padding_mask=None, attn_mask=None, rel_pos=None):
        bsz, tgt_len, embed_dim = query.size()
        src_len = key.size(1)
        q = self.q_proj(query)
        k = self.k_proj(key)
        v = self.v_proj(value)
        q *= self.scaling
        q = q.view(bsz, tgt_len, self.num_heads, self.head_dim).transpose(1, 2).reshape(bsz * self.num_heads, tgt_len, self.head_dim)
        k = k.view(bsz, src_len, self.num_heads, self.head_dim).transpose(1, 2).reshape(bsz * self.num_heads, src_len, self.head_dim)
        v = v.view(bsz, src_len, self.num_heads, self.head_dim).transpose(1, 2).reshape(bsz * self.num_heads, src_len, self.head_dim)
        attn_weights = torch.bmm(q, k.transpose(1, 2))
        if attn_mask is not None:
            attn_weights = torch.nan_to_num(attn_weights)
            attn_weights += attn_mask.unsqueeze(0)
        if key_padding_mask is not None:
            attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len)
            attn_weights = attn_weights.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2).to(torch.bool), float("-inf"))
            attn_weights = attn_weights.view(bsz * self.num_heads, tgt_len, src_len)
        if rel_pos is not None:
            attn_weights = attn_weights + rel_pos.view(attn_weights.size())
        attn_weights = F.softmax(attn_weights, dim=-1, dtype=torch.float32).type_as(attn_weights)
        attn_probs = self.dropout_module(attn_weights)
        attn = torch.bmm(attn_probs, v)
        attn = attn.transpose(0, 1).reshape(tgt_len, bsz, embed_dim).transpose(0, 1)
        if self.inner_attn_ln is not None:
            attn = self.inner_attn_ln(attn)
        attn = self.out_proj(attn)
        attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len).transpose(1, 0)
        return attn, attn_weights
