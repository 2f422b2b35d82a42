"""CPU oracle — TEST INFRASTRUCTURE ONLY (never imported by the product path `simvg_b200/`).

A plain-PyTorch (fp32 or fp64, eager, CPU-friendly) restatement of SimVG's vision-language fusion train step:
BEiT-3 multiway encoder -> text-guided query generation -> token branch / DETR decoder branch -> heads -> DWBD losses
-> predictions.  It is written functionally over a *reference-keyed* state dict, so the same
weights drive (i) the reference's own files imported over shims in the build container (oracle/make_golden.py),
(ii) this oracle and (iii) the CUDA product.  Each function cites the reference file:line it follows; leaf ops that live
in un-vendored third-party packages (torchscale, detrex, detectron2 — unpinned in /root/reference/requirements.txt:13
and README.md:59-64) restate the published algorithm.

Parity status: the reference ships NO tests, golden vectors or fixtures -> upstream parity is UNPINNED.
What pins this oracle instead: tests/golden/*.pt, generated in the build container by executing the reference's own
beit3_base.py / beit3.py / tgqs_kd_detr_head.py / transformer.py / heads/utils.py / criterion.py / mix_detr_mb.py
verbatim over shimmed leaf ops (oracle/shims, oracle/make_golden.py); tests/test_oracle_golden.py checks the oracle
against them.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math

import torch
import torch.nn.functional as F

try:
    from scipy.optimize import linear_sum_assignment
except Exception:  # pragma: no cover
    linear_sum_assignment = None


# =============================================================================================== encoder (BEiT-3)
def encoder_config(vit_type, img_size, patch_size, drop_path_rate=0.0):
    """modeling_utils.py:21-70 (+ the `rop_path_rate` typo at beit3.py:54: ViT-L never gets DropPath)."""
    if vit_type == "base":
        return dict(D=768, H=12, F=3072, layers=12, img_size=img_size, patch_size=patch_size, eps=1e-5,
                    drop_path_rate=drop_path_rate)
    if vit_type == "large":
        return dict(D=1024, H=16, F=4096, layers=24, img_size=img_size, patch_size=patch_size, eps=1e-5, drop_path_rate=0.0)
    raise TypeError("please select the <vit_type> from ['base','large']")


def _mw(sd, prefix, leaf, x, split, fn):
    """torchscale MultiwayNetwork.forward (A.3): split at `split` along dim 1, expert A | expert B, concat."""
    xa, xb = x[:, :split], x[:, split:]
    return torch.cat([fn(xa, sd[prefix + ".A." + leaf + "weight"], sd[prefix + ".A." + leaf + "bias"]),
                      fn(xb, sd[prefix + ".B." + leaf + "weight"], sd[prefix + ".B." + leaf + "bias"])], dim=1)


def _mw_linear(sd, prefix, x, split):
    return _mw(sd, prefix, "", x, split, F.linear)


def _mw_ln(sd, prefix, x, split, eps):
    return _mw(sd, prefix, "", x, split, lambda t, w, b: F.layer_norm(t, (t.shape[-1],), w, b, eps))


def multiway_attention(sd, prefix, x, key_padding_mask, split, H, eps):
    """torchscale MultiheadAttention.forward with subln (A.4); called at beit3_base.py:137-145."""
    B, L, D = x.shape
    dh = D // H
    q = _mw_linear(sd, prefix + ".q_proj", x, split) * (dh ** -0.5)
    k = _mw_linear(sd, prefix + ".k_proj", x, split)
    v = _mw_linear(sd, prefix + ".v_proj", x, split)
    q = q.view(B, L, H, dh).transpose(1, 2)
    k = k.view(B, L, H, dh).transpose(1, 2)
    v = v.view(B, L, H, dh).transpose(1, 2)
    w = q @ k.transpose(-1, -2)
    if key_padding_mask is not None:
        w = w.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2).to(torch.bool), float("-inf"))
    w = F.softmax(w, dim=-1, dtype=torch.float32).type_as(w)   # fp32 softmax whatever the activation dtype (A.4)
    a = (w @ v).transpose(1, 2).reshape(B, L, D)
    a = _mw_ln(sd, prefix + ".inner_attn_ln", a, split, eps)
    return _mw_linear(sd, prefix + ".out_proj", a, split)


def multiway_ffn(sd, prefix, x, split, eps):
    """torchscale FeedForwardNetwork.forward with subln (A.5), wrapped whole in MultiwayWrapper (beit3_base.py:57-63)."""
    outs = []
    for which, xx in (("A", x[:, :split]), ("B", x[:, split:])):
        p = "%s.%s." % (prefix, which)
        h = F.linear(xx, sd[p + "fc1.weight"], sd[p + "fc1.bias"])
        h = F.gelu(h.float()).type_as(h)                            # erf GELU evaluated in fp32 (A.5)
        h = F.layer_norm(h, (h.shape[-1],), sd[p + "ffn_layernorm.weight"], sd[p + "ffn_layernorm.bias"], eps)
        outs.append(F.linear(h, sd[p + "fc2.weight"], sd[p + "fc2.bias"]))
    return torch.cat(outs, dim=1)


def encoder_forward(sd, cfg, image, ids, pad_mask, prefix="", drop_path_masks=None):
    """BEiT3.forward + Encoder.forward (beit3_base.py:441-488, 336-407).  Returns encoder_out [B, Lv+Lt, D].

    drop_path_masks: optional list (per layer) of (attn_scale[B], ffn_scale[B]) — Bernoulli(keep)/keep draws (A.8);
    None = eval semantics."""
    p = prefix + "beit3."
    D, H, P, eps = cfg["D"], cfg["H"], cfg["patch_size"], cfg["eps"]
    B = image.shape[0]
    assert image.shape[2] == cfg["img_size"] and image.shape[3] == cfg["img_size"]
    # VisionEmbedding (A.6): Conv2d(k=P, s=P) -> flatten -> transpose -> prepend cls
    x1 = F.conv2d(image, sd[p + "vision_embed.proj.weight"], sd[p + "vision_embed.proj.bias"], stride=P)
    x1 = x1.flatten(2).transpose(1, 2)
    x1 = torch.cat([sd[p + "vision_embed.cls_token"].expand(B, -1, -1), x1], dim=1)
    split = x1.shape[1]
    x2 = F.embedding(ids, sd[p + "text_embed.weight"])                                   # TextEmbedding (A.7)
    x = torch.cat([x1, x2], dim=1)                                                         # beit3_base.py:464
    Lt = x2.shape[1]
    kpm = None
    if pad_mask is not None:                                                               # beit3_base.py:466-473
        kpm = torch.cat([torch.zeros(B, split, dtype=pad_mask.dtype, device=x.device), pad_mask], dim=1)
    # forward_embedding (beit3_base.py:317-334): positions arange(2, len+2) per expert (A.7), embed_scale = 1
    posA = sd[p + "encoder.embed_positions.A.weight"][2:2 + split]
    posB = sd[p + "encoder.embed_positions.B.weight"][2:2 + Lt]
    x = x + torch.cat([posA, posB], dim=0).unsqueeze(0)
    if kpm is not None:
        x = x * (1 - kpm.unsqueeze(-1).type_as(x))                                        # beit3_base.py:367
    for li in range(cfg["layers"]):                                                        # EncoderLayer.forward :127-172
        lp = "%sencoder.layers.%d." % (p, li)
        r = x
        h = _mw_ln(sd, lp + "self_attn_layer_norm", x, split, eps)
        h = multiway_attention(sd, lp + "self_attn", h, kpm, split, H, eps)
        if drop_path_masks is not None and drop_path_masks[li][0] is not None:
            h = h * drop_path_masks[li][0].view(B, 1, 1).type_as(h)
        x = r + h
        r = x
        h = _mw_ln(sd, lp + "final_layer_norm", x, split, eps)
        h = multiway_ffn(sd, lp + "ffn", h, split, eps)
        if drop_path_masks is not None and drop_path_masks[li][1] is not None:
            h = h * drop_path_masks[li][1].view(B, 1, 1).type_as(h)
        x = r + h
    return _mw_ln(sd, p + "encoder.layer_norm", x, split, eps)                             # :396-397


def beit3_forward(sd, cfg, image, ids, pad_mask, prefix="", drop_path_masks=None):
    """BEIT3.forward (beit3.py:176-185) -> (img_feat, text_feat, cls_feat)."""
    x = encoder_forward(sd, cfg, image, ids, pad_mask, prefix, drop_path_masks)
    Lt = ids.shape[-1]
    return x[:, 1:-Lt], x[:, -Lt:], x[:, 0]


# =============================================================================================== box ops (detrex, A.12)
def box_cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def box_xyxy_to_cxcywh(b):
    x0, y0, x1, y1 = b.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], dim=-1)


def box_iou(b1, b2):
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2 - inter
    return inter / union, union


def generalized_box_iou(b1, b2):
    assert (b1[:, 2:] >= b1[:, :2]).all() and (b2[:, 2:] >= b2[:, :2]).all()
    iou, union = box_iou(b1, b2)
    lt = torch.min(b1[:, None, :2], b2[:, :2])
    rb = torch.max(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[..., 0] * wh[..., 1]
    return iou - (area - union) / area


@torch.no_grad()
def hungarian_match(pred_logits, pred_boxes, targets, cost_class=1.0, cost_bbox=5.0, cost_giou=2.0):
    """detrex HungarianMatcher, cost_class_type="ce_cost" (A.12); built at tgqs_kd_detr_head.py:132-137."""
    B, nq = pred_logits.shape[:2]
    prob = pred_logits.flatten(0, 1).softmax(-1)
    boxes = pred_boxes.flatten(0, 1)
    tgt_ids = torch.cat([t["labels"] for t in targets])
    tgt_box = torch.cat([t["boxes"] for t in targets])
    C = cost_bbox * torch.cdist(boxes, tgt_box, p=1) + cost_class * (-prob[:, tgt_ids]) + cost_giou * (
        -generalized_box_iou(box_cxcywh_to_xyxy(boxes), box_cxcywh_to_xyxy(tgt_box)))
    C = C.view(B, nq, -1).cpu()
    sizes = [len(t["boxes"]) for t in targets]
    out = []
    for i, c in enumerate(C.split(sizes, -1)):
        r, col = linear_sum_assignment(c[i])
        out.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(col, dtype=torch.int64)))
    return out


# =============================================================================================== criterion
def _src_idx(indices):
    return (torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)]), torch.cat([s for s, _ in indices]))


def set_criterion(outputs, targets, num_classes=1, eos_coef=0.1, world_size=1):
    """SetCriterion.forward with loss_class_type="ce_loss" (criterion.py:226-271, 108-162, 164-204)."""
    logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
    empty_weight = torch.ones(num_classes + 1, dtype=logits.dtype, device=logits.device)
    empty_weight[-1] = eos_coef
    num_boxes = max(float(sum(len(t["labels"]) for t in targets)) / world_size, 1.0)

    def one(lg, bx, suffix):
        idx_list = hungarian_match(lg, bx, targets)
        idx = _src_idx(idx_list)
        tcls_o = torch.cat([t["labels"][j] for t, (_, j) in zip(targets, idx_list)])
        tcls = torch.full(lg.shape[:2], num_classes, dtype=torch.int64, device=lg.device)
        tcls[idx] = tcls_o
        out = {"loss_class" + suffix: F.cross_entropy(lg.transpose(1, 2), tcls, empty_weight)}
        src = bx[idx]
        tb = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, idx_list)], dim=0)
        out["loss_bbox" + suffix] = F.l1_loss(src, tb, reduction="none").sum() / num_boxes
        out["loss_giou" + suffix] = (1 - torch.diag(generalized_box_iou(box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tb)))).sum() / num_boxes
        return out

    losses = one(logits, boxes, "")
    for i, aux in enumerate(outputs.get("aux_outputs", [])):
        losses.update(one(aux["pred_logits"], aux["pred_boxes"], "_%d" % i))
    return losses


# =============================================================================================== head
def _linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def mlp(sd, prefix, x, num_layers, return_intermediate=False):
    """heads/utils.py:7-46."""
    inter = []
    for i in range(num_layers):
        x = _linear(sd, "%s.layers.%d" % (prefix, i), x)
        if i < num_layers - 1:
            x = F.relu(x)
        inter.append(x)
    return torch.stack(inter, dim=0) if return_intermediate else x


def pos_embed_sine_2d(mask, num_pos_feats=128, temperature=10000, scale=2 * math.pi, eps=1e-6):
    """detrex PositionEmbeddingSine(normalize=True) (A.11); built at tgqs_kd_detr_head.py:90-94."""
    not_mask = ~mask
    y = not_mask.cumsum(1, dtype=torch.float32)
    x = not_mask.cumsum(2, dtype=torch.float32)
    y = y / (y[:, -1:, :] + eps) * scale
    x = x / (x[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=mask.device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    px = x[:, :, :, None] / dim_t
    py = y[:, :, :, None] / dim_t
    Bm, Hm, Wm = mask.shape
    px = torch.stack((px[:, :, :, 0::2].sin(), px[:, :, :, 1::2].cos()), dim=4).view(Bm, Hm, Wm, -1)
    py = torch.stack((py[:, :, :, 0::2].sin(), py[:, :, :, 1::2].cos()), dim=4).view(Bm, Hm, Wm, -1)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def pos_embed_sine_1d(pos_len, dim):
    """PositionEmbeddingSine1D.forward incl. its `.to(torch.long)` quirk (heads/utils.py:72-100; Appendix C.2)."""
    emb = torch.zeros(pos_len, dim, dtype=torch.float)
    i = torch.arange(dim // 2, dtype=torch.float)
    i /= dim / 2
    i = (1 / torch.pow(10000, i)).to(torch.long)
    out = torch.arange(pos_len).to(torch.long)[:, None] @ i[None, :]
    emb[:, 0::2] = torch.sin(out)
    emb[:, 1::2] = torch.cos(out)
    return emb


def _mha(sd, prefix, query, key, value, identity, query_pos, key_pos, key_padding_mask, nheads=8):
    """detrex MultiheadAttention wrapper around nn.MultiheadAttention (A.9); dropout off (eval semantics)."""
    if key is None:
        key = query
    if value is None:
        value = key
    if identity is None:
        identity = query
    if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
        key_pos = query_pos
    if query_pos is not None:
        query = query + query_pos
    if key_pos is not None:
        key = key + key_pos
    E = query.shape[-1]
    out = F.multi_head_attention_forward(
        query, key, value, E, nheads, sd[prefix + ".attn.in_proj_weight"], sd[prefix + ".attn.in_proj_bias"], None, None,
        False, 0.0, sd[prefix + ".attn.out_proj.weight"], sd[prefix + ".attn.out_proj.bias"], training=False,
        key_padding_mask=key_padding_mask, need_weights=True, attn_mask=None)[0]
    return identity + out


def _decoder_layer(sd, prefix, query, key, value, query_pos, key_pos, key_padding_mask):
    """detrex BaseTransformerLayer, order (self_attn, norm, cross_attn, norm, ffn, norm) (A.10; transformer.py:107-125)."""
    E = query.shape[-1]
    ln = lambda i, t: F.layer_norm(t, (E,), sd["%s.norms.%d.weight" % (prefix, i)], sd["%s.norms.%d.bias" % (prefix, i)])  # noqa: E731
    query = _mha(sd, prefix + ".attentions.0", query, query, query, None, query_pos, query_pos, None)
    query = ln(0, query)
    query = _mha(sd, prefix + ".attentions.1", query, key, value, None, query_pos, key_pos, key_padding_mask)
    query = ln(1, query)
    h = F.relu(_linear(sd, prefix + ".ffns.0.layers.0.0", query))
    query = query + _linear(sd, prefix + ".ffns.0.layers.1", h)
    return ln(2, query)


def detr_decoder(sd, prefix, num_layers, query, key, value, query_pos, key_pos, key_padding_mask, return_intermediate):
    """DetrTransformerDecoder.forward (transformer.py:134-186)."""
    E = query.shape[-1]
    post = lambda t: F.layer_norm(t, (E,), sd[prefix + ".post_norm_layer.weight"], sd[prefix + ".post_norm_layer.bias"])  # noqa: E731
    inter = []
    for i in range(num_layers):
        query = _decoder_layer(sd, "%s.layers.%d" % (prefix, i), query, key, value, query_pos, key_pos, key_padding_mask)
        if return_intermediate:
            inter.append(post(query))
    if not return_intermediate:
        return post(query)[None]
    return torch.stack(inter)


def head_forward_general(sd, hc, x_mm, img_metas, cls_feat, text_feat, text_mask, prefix="head.", proj_hook=None):
    """TextGuidedQuerySelectKDDETRHead.forward_general (tgqs_kd_detr_head.py:375-454), text_guided_query_generation=True.

    proj_hook: optional callable applied to input_proj's output (oracle/bf16_emulation.py rounds its gradient there)."""
    nq = hc["num_queries"]
    B = x_mm.shape[0]
    x_mm = F.conv2d(x_mm, sd[prefix + "input_proj.weight"], sd[prefix + "input_proj.bias"])
    if proj_hook is not None:
        x_mm = proj_hook(x_mm)
    text_feat = _linear(sd, prefix + "input_text_proj", text_feat)
    cls_feat = _linear(sd, prefix + "input_cls_proj", cls_feat).unsqueeze(1)
    # x_mask_pos_enc (:322-338)
    try:
        ih, iw = img_metas[0]["batch_input_shape"]
    except Exception:  # noqa: BLE001
        ih, iw, _ = img_metas[0]["img_shape"]
    x_mask = x_mm.new_ones((B, ih, iw))
    for i in range(B):
        h, w, _ = img_metas[i]["img_shape"]
        x_mask[i, :h, :w] = 0
    x_mask = F.interpolate(x_mask.unsqueeze(1), size=x_mm.shape[-2:]).to(torch.bool).squeeze(1)
    pos = pos_embed_sine_2d(x_mask, x_mm.shape[1] // 2).to(x_mm.dtype)
    cls_feat = cls_feat.repeat((1, nq, 1))
    # TGQG (:385-399).  `~text_mask` on the int64 loader mask is a bitwise NOT -> integer row indexing (Appendix C.1).
    inv = ~text_mask
    tff = torch.cat([torch.max(f[m, :], dim=0, keepdim=True)[0] for f, m in zip(text_feat, inv)]).unsqueeze(1).repeat(1, nq, 1)
    qe_in = sd[prefix + "query_embed.weight"].unsqueeze(0).repeat(B, 1, 1).transpose(0, 1)
    target = torch.zeros_like(qe_in)
    tpos = pos_embed_sine_1d(text_feat.shape[1], text_feat.shape[2]).to(text_feat)
    tpos = tpos.unsqueeze(0).repeat(B, 1, 1).permute(1, 0, 2)
    tin = text_feat.transpose(0, 1)
    qe = detr_decoder(sd, prefix + "text_guided_query_generation_transformer", hc["num_tgqg_layers"], target, tin, tin, qe_in,
                      tpos, text_mask.bool(), return_intermediate=False)
    query_embed = qe[0].transpose(0, 1) + tff + qe_in.transpose(0, 1)
    cls_feat = query_embed + cls_feat
    out = {}
    if "decoder" in hc["branch_loss_weight"] and len(hc["branch_loss_weight"]) == 1:
        cls_tok = coord_tok = None
        out["token_branch_output"] = {"pred_logits": None, "pred_boxes": None}
    else:
        n_mlp = hc["num_token_mlp_layers"]
        cls_feat = mlp(sd, prefix + "mlp", cls_feat, n_mlp, return_intermediate=True) if n_mlp > 0 else cls_feat.unsqueeze(0)
        cls_tok = _linear(sd, prefix + "class_embed_token", cls_feat)
        coord_tok = mlp(sd, prefix + "bbox_embed_token", cls_feat, 3).sigmoid()
        out["token_branch_output"] = {"pred_logits": cls_tok[-1], "pred_boxes": coord_tok[-1]}
    # DetrTransformer.forward, only_decoder (transformer.py:205-235)
    Bc, C, hh, ww = x_mm.shape
    mem = x_mm.view(Bc, C, -1).permute(2, 0, 1)
    pos2 = pos.view(Bc, C, -1).permute(2, 0, 1)
    q = query_embed.transpose(0, 1)
    hs = detr_decoder(sd, prefix + "transformer.decoder", hc["num_decoder_layers"], torch.zeros_like(q), mem, mem, q, pos2,
                      x_mask.view(Bc, -1), return_intermediate=True).transpose(1, 2)
    cls_dec = _linear(sd, prefix + "class_embed_decoder", hs)
    coord_dec = mlp(sd, prefix + "bbox_embed_decoder", hs, 3).sigmoid()
    out["decoder_branch_output"] = {"pred_logits": cls_dec[-1], "pred_boxes": coord_dec[-1]}
    out.update(outputs_class_decoder_branch=cls_dec, outputs_coord_decoder_branch=coord_dec,
               outputs_class_token_branch=cls_tok, outputs_coord_token_branch=coord_tok, token_features=cls_feat,
               decoder_features=hs)
    return out


def _gt_targets(gt_bbox, img_metas):
    """prepare_soft_targets, GT half (tgqs_kd_detr_head.py:215-234), REC case (one box per sample)."""
    out = []
    for tb, meta in zip(gt_bbox, img_metas):
        h, w = meta["img_shape"][:2]
        scale = torch.as_tensor([w, h, w, h], dtype=torch.float, device=tb.device)
        assert tb.dim() == 1, "gREC multi-target samples are outside the oracle's scope"
        boxes = box_xyxy_to_cxcywh(tb.unsqueeze(0).float() / scale).float()
        out.append({"labels": torch.zeros(1, device=tb.device).long(), "boxes": boxes})
    return out


def _teacher_targets(dec_out, targets_gt):
    """prepare_soft_targets, "score_iou_weighted" half (tgqs_kd_detr_head.py:248-264): detached decoder boxes matched to GT,
    weight = softmax(logits)[..., 0] * IoU."""
    logits = dec_out["pred_logits"].detach()
    boxes = dec_out["pred_boxes"].detach()
    scores = F.softmax(logits, dim=-1)[:, :, 0:1]
    idx = hungarian_match(logits, boxes, targets_gt)
    out = []
    for (i_pred, i_tgt), pb, ps, tg in zip(idx, boxes, scores, targets_gt):
        pb_ = pb[i_pred]
        tg_ = tg["boxes"][i_tgt]
        ious = torch.diag(box_iou(box_cxcywh_to_xyxy(pb_), box_cxcywh_to_xyxy(tg_))[0])
        wgt = ps[i_pred].reshape(-1) * ious
        out.append({"labels": torch.zeros(pb_.shape[0], device=pb.device).long(), "boxes": pb_, "weight": wgt})
    return out


_WEIGHTS = {"loss_class": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0}


def _calc_loss(cls, coord, targets, world_size=1):
    """calc_loss (tgqs_kd_detr_head.py:340-350) with the aux weight dict of :174-180 (same weights for every aux layer)."""
    o = {"pred_logits": cls[-1], "pred_boxes": coord[-1],
         "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(cls[:-1], coord[:-1])]}
    ld = set_criterion(o, targets, world_size=world_size)
    return {k: v * _WEIGHTS[k.rsplit("_", 1)[0] if k[-1].isdigit() else k] for k, v in ld.items()}


def head_forward_train(sd, hc, x_mm, img_metas, cls_feat, text_feat, gt_bbox, text_mask, prefix="head.", world_size=1,
                       proj_hook=None):
    """forward_train (tgqs_kd_detr_head.py:456-572) for branch_loss_weight = {decoder[, balanced_distill]}."""
    out = head_forward_general(sd, hc, x_mm, img_metas, cls_feat, text_feat, text_mask, prefix, proj_hook=proj_hook)
    blw = hc["branch_loss_weight"]
    tg = _gt_targets(gt_bbox, img_metas)
    tp = _teacher_targets(out["decoder_branch_output"], tg)
    z = lambda: torch.zeros((), device=x_mm.device)  # noqa: E731
    l_dec, l_tok, l_kd = z(), z(), z()
    ld = {}
    if "decoder" in blw:
        l_dec = blw["decoder"] * sum(_calc_loss(out["outputs_class_decoder_branch"], out["outputs_coord_decoder_branch"], tg,
                                                world_size).values())
        ld["loss_dgt"] = l_dec
    if "balanced_distill" in blw:
        w = torch.mean(torch.cat([t["weight"] for t in tp]))
        ct, bt = out["outputs_class_token_branch"][-1:], out["outputs_coord_token_branch"][-1:]
        l_tok = blw["balanced_distill"]["token"] * sum(_calc_loss(ct, bt, tg, world_size).values()) * (1 - w)
        l_kd = blw["balanced_distill"]["distill"] * sum(_calc_loss(ct, bt, tp, world_size).values()) * w
        ld.update(loss_tgt=l_tok, loss_kd=l_kd, loss_distill_w=w)
    ld["loss_total"] = l_dec + l_tok + l_kd
    return ld, out


def get_predictions(output, img_metas):
    """MIXDETRMB.get_predictions + head.inference + detectron2 detector_postprocess (mix_detr_mb.py:127-159,
    tgqs_kd_detr_head.py:577-604, A.13), rescale=False."""
    if output["pred_logits"] is None:
        return dict(pred_bboxes=None, pred_masks=None, predict_classes=None)
    scores, labels = F.softmax(output["pred_logits"], dim=-1)[:, :, :-1].max(-1)
    boxes, classes = [], []
    for sc, lb, bp, meta in zip(scores, labels, output["pred_boxes"], img_metas):
        h, w = meta["img_shape"][:2]
        bx = box_cxcywh_to_xyxy(bp) * torch.as_tensor([w, h, w, h], dtype=bp.dtype, device=bp.device)
        bx = torch.stack([bx[:, 0].clamp(0, w), bx[:, 1].clamp(0, h), bx[:, 2].clamp(0, w), bx[:, 3].clamp(0, h)], dim=-1)
        keep = ((bx[:, 2] - bx[:, 0]) > 0) & ((bx[:, 3] - bx[:, 1]) > 0)
        bx, sc, lb = bx[keep], sc[keep], lb[keep]
        best = int(torch.argmax(sc))
        boxes.append(bx[best:best + 1])
        classes.append(lb)
    return dict(pred_bboxes=torch.cat(boxes, 0), pred_masks=None, predict_classes=torch.cat(classes, 0))


def get_predictions_grec(output, img_metas, rescale=False):
    """MIXDETRMB.get_predictions_grec (mix_detr_mb.py:161-190): every non-empty box of every image, per-image dicts."""
    if output["pred_logits"] is None:
        return dict(pred_bboxes=None, pred_masks=None, predict_classes=None)
    scores, labels = F.softmax(output["pred_logits"], dim=-1)[:, :, :-1].max(-1)
    preds = []
    for sc, lb, bp, meta in zip(scores, labels, output["pred_boxes"], img_metas):
        h, w = meta["img_shape"][:2]
        bx = box_cxcywh_to_xyxy(bp) * torch.as_tensor([w, h, w, h], dtype=bp.dtype, device=bp.device)
        bx = torch.stack([bx[:, 0].clamp(0, w), bx[:, 1].clamp(0, h), bx[:, 2].clamp(0, w), bx[:, 3].clamp(0, h)], dim=-1)
        keep = ((bx[:, 2] - bx[:, 0]) > 0) & ((bx[:, 3] - bx[:, 1]) > 0)
        bx, sc, lb = bx[keep], sc[keep], lb[keep]
        if rescale:
            bx = bx / bx.new_tensor(meta["scale_factor"])
        preds.append({"boxes": bx, "scores": sc, "labels": lb})
    return dict(pred_bboxes=preds, pred_masks=None)


# =============================================================================================== detector
class OracleModel:
    """MIXDETRMB (det_seg/mix_detr_mb.py:13-125) over a reference-keyed state dict."""

    def __init__(self, sd, vit_type, img_size, patch_size, head_cfg):
        self.sd = sd
        self.cfg = encoder_config(vit_type, img_size, patch_size)
        self.hc = head_cfg

    def _features(self, img, ids, text_mask):
        B, _, Hh, Ww = img.shape
        img_feat, text_feat, cls_feat = beit3_forward(self.sd, self.cfg, img, ids, text_mask, prefix="vis_enc.")
        P = self.cfg["patch_size"]
        return img_feat.transpose(-1, -2).reshape(B, -1, Hh // P, Ww // P), text_feat, cls_feat

    def forward_train(self, img, ids, img_metas, text_attention_mask, gt_bbox, world_size=1):
        for m in img_metas:
            m["batch_input_shape"] = tuple(img.shape[-2:])                                  # base.py:12-15
        x_mm, text_feat, cls_feat = self._features(img, ids, text_attention_mask)
        losses, out = head_forward_train(self.sd, self.hc, x_mm, img_metas, cls_feat, text_feat, gt_bbox, text_attention_mask,
                                         world_size=world_size)
        with torch.no_grad():
            preds = [get_predictions(out["decoder_branch_output"], img_metas),
                     get_predictions(out["token_branch_output"], img_metas)]
        return losses, preds, out

    @torch.no_grad()
    def forward_test(self, img, ids, img_metas, text_attention_mask):
        for m in img_metas:
            m["batch_input_shape"] = tuple(img.shape[-2:])
        x_mm, text_feat, cls_feat = self._features(img, ids, text_attention_mask)
        out = head_forward_general(self.sd, self.hc, x_mm, img_metas, cls_feat, text_feat, text_attention_mask)
        return [get_predictions(out["decoder_branch_output"], img_metas),
                get_predictions(out["token_branch_output"], img_metas)], out


def accuracy_at_05(pred_bboxes, gt_bbox):
    """apis/test.py:70-79: mean(IoU(gt, pred) >= 0.5) * 100 with mmdet's aligned bbox_overlaps (eps 1e-6)."""
    gt = torch.stack(gt_bbox).to(pred_bboxes)
    lt = torch.max(gt[:, :2], pred_bboxes[:, :2])
    rb = torch.min(gt[:, 2:], pred_bboxes[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    a1 = (gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])
    a2 = (pred_bboxes[:, 2] - pred_bboxes[:, 0]) * (pred_bboxes[:, 3] - pred_bboxes[:, 1])
    iou = inter / (a1 + a2 - inter).clamp(min=1e-6)
    return (iou >= 0.5).float().mean() * 100.0
