"""Synthetic RefCOCO-shaped inputs: image ~ N(0,1) fp32, XLM-R style token ids, int64 pad mask, xyxy boxes."""
import torch


def make_batch(B, S, Lt=20, seed=6666, device="cpu", vocab=64010):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, S, S, generator=g)
    ids = torch.ones(B, Lt, dtype=torch.int64)           # pad = 1
    mask = torch.ones(B, Lt, dtype=torch.int64)          # 1 = padded (int64, as the loader emits: loading.py:175-179)
    for b in range(B):
        k = int(torch.randint(3, Lt - 1, (1,), generator=g))
        ids[b, 0] = 0                                     # bos
        ids[b, 1:1 + k] = torch.randint(4, vocab, (k,), generator=g)
        ids[b, 1 + k] = 2                                 # eos
        mask[b, :k + 2] = 0
    c = torch.rand(B, 2, generator=g) * 0.6 + 0.2
    wh = torch.rand(B, 2, generator=g) * 0.4 + 0.1
    xyxy = torch.cat([(c - wh / 2), (c + wh / 2)], dim=1).clamp(0, 1) * (S - 1)
    gt_bbox = [xyxy[b].double().float() for b in range(B)]
    metas = [dict(img_shape=(S, S, 3), pad_shape=(S, S, 3), ori_shape=(S, S, 3), scale_factor=[1.0, 1.0, 1.0, 1.0],
                  filename="", expression="") for _ in range(B)]
    dev = torch.device(device)
    return dict(img=img.to(dev), ref_expr_inds=ids.to(dev), text_attention_mask=mask.to(dev),
                gt_bbox=[t.to(dev) for t in gt_bbox], img_metas=metas)


def model_cfg(vit_type="base", img_size=640, patch_size=16, num_decoder_layers=3, drop_path_rate=0.1, num_queries=1,
              branch_loss_weight=None):
    if branch_loss_weight is None:
        branch_loss_weight = {"decoder": 1.0, "balanced_distill": {"token": 2.0, "distill": 1.0}}
    return dict(
        type="MIXDETRMB",
        vis_enc=dict(type="BEIT3", img_size=img_size, patch_size=patch_size, vit_type=vit_type, drop_path_rate=drop_path_rate,
                     vocab_size=64010, freeze_layer=-1, vision_embed_proj_interpolate=True, pretrain=None),
        lan_enc=None, fusion=None,
        head=dict(type="TextGuidedQuerySelectKDDETRHead", num_queries=num_queries, text_max_token=20,
                  in_channels=768 if vit_type == "base" else 1024, embed_dim=256, decoder_freeze=False, num_classes=1,
                  aux_loss=True, num_encoder_layers=6, num_decoder_layers=num_decoder_layers, only_decoder=True,
                  text_embed_aug=False, branch_loss_weight=branch_loss_weight, distill_type="hard_weighted",
                  prepare_target_mode="score_iou_weighted", share_predicthead=False, num_token_mlp_layers=1,
                  mlp_aux_loss=False, text_guided_query_generation=True, num_tgqg_layers=2))


def synth_state_dict(template, seed):
    """Deterministic, well-conditioned weights for every floating tensor of `template` (name -> tensor), derived from
    (seed, name) only — so fixtures store no weights: the reference-over-shims run, the oracle and the CUDA product all
    regenerate identical parameters.  Matrices ~ N(0, 0.25/fan_in), LayerNorm gains ~ 1 + 0.1 N(0,1), biases ~ 0.05 N(0,1)."""
    import zlib
    out = {}
    for name in sorted(template):
        t = template[name]
        if not t.dtype.is_floating_point or "empty_weight" in name:
            out[name] = t.clone()
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if t.dim() >= 2:
            fan_in = 1
            for s in t.shape[1:]:
                fan_in *= s
            v = r * (0.5 / max(fan_in, 1) ** 0.5)
        elif name.endswith("weight"):
            v = 1.0 + 0.1 * r
        else:
            v = 0.05 * r
        out[name] = v.to(t.dtype)
    return out
