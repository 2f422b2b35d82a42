"""CPU-side checks: the drop-in boundary (registries, ctor kwargs, state-dict keys), flat parameter storage, the C ABI
(library loads and exports every symbol include/simvg_b200.h declares), attention tiling geometry, no-fallback policy."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "simvg_b200.h")).read()
    names = set(re.findall(r"\b(simvgb_[a-z0-9_]+)\s*\(", hdr))
    assert {"simvgb_gemm", "simvgb_attn_fwd", "simvgb_attn_bwd", "simvgb_ln_fwd", "simvgb_ln_bwd", "simvgb_adam_amsgrad"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), "libsimvg_b200.so does not export %s" % n
    assert lib.simvgb_version() == 200


def test_cabi_argument_validation_without_gpu(lib):
    """Error conventions: <0 + message, no exceptions, no device needed for argument errors."""
    from simvg_b200 import _lib as L
    a = L.GemmArgs()
    a.M, a.N, a.K = 0, 8, 8
    assert lib.simvgb_gemm(ctypes.byref(a), None) < 0
    assert b"bad shape" in lib.simvgb_last_error()
    assert lib.simvgb_gemm(None, None) < 0
    lib.simvgb_attn_lse_stride.restype = ctypes.c_int
    assert lib.simvgb_attn_lse_stride(1601, 20) == 13 * 128
    assert lib.simvgb_attn_lse_stride(401, 20) == 4 * 128
    assert lib.simvgb_attn_lse_stride(2305, 20) == 19 * 128
    assert lib.simvgb_attn_lse_stride(256, 20) == 3 * 128


def test_product_refuses_cpu_tensors():
    from simvg_b200 import kernels as K
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.ln_fwd(torch.zeros(4, 256), torch.ones(256), torch.zeros(256), 1e-5)
    from simvg_b200.models import build_model
    from tools.synth import make_batch, model_cfg
    m = build_model(model_cfg("base", 64, 32))
    b = make_batch(1, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(b["img"], b["ref_expr_inds"], b["img_metas"], return_loss=False, text_attention_mask=b["text_attention_mask"])


def test_registry_builds_reference_config_and_keys():
    from simvg_b200.models import HEADS, MODELS, VIS_ENCODERS, build_model
    from tools.synth import model_cfg
    assert "BEIT3" in VIS_ENCODERS and "TextGuidedQuerySelectKDDETRHead" in HEADS and "MIXDETRMB" in MODELS
    m = build_model(model_cfg("base", 640, 32))
    keys = set(m.state_dict().keys())
    # SURVEY Appendix D
    expect = ["vis_enc.beit3.text_embed.weight", "vis_enc.beit3.vision_embed.proj.weight", "vis_enc.beit3.vision_embed.mask_token",
              "vis_enc.beit3.vision_embed.cls_token", "vis_enc.beit3.encoder.embed_positions.A.weight",
              "vis_enc.beit3.encoder.embed_positions.B.weight", "vis_enc.beit3.encoder.layer_norm.B.bias",
              "head.input_proj.weight", "head.input_text_proj.bias", "head.input_cls_proj.weight", "head.query_embed.weight",
              "head.mlp.layers.0.weight", "head.class_embed_decoder.weight", "head.class_embed_token.bias",
              "head.bbox_embed_decoder.layers.2.weight", "head.bbox_embed_token.layers.0.bias",
              "head.transformer.decoder.post_norm_layer.weight", "head.text_guided_query_generation_transformer.post_norm_layer.bias",
              "head.criterion.empty_weight", "head.criterion_harddistill.empty_weight"]
    for i in (0, 11):
        for w in ("q_proj", "k_proj", "v_proj", "out_proj"):
            for ab in "AB":
                expect += ["vis_enc.beit3.encoder.layers.%d.self_attn.%s.%s.%s" % (i, w, ab, t) for t in ("weight", "bias")]
        for ab in "AB":
            expect += ["vis_enc.beit3.encoder.layers.%d.%s.%s.weight" % (i, n, ab) for n in
                       ("self_attn.inner_attn_ln", "self_attn_layer_norm", "final_layer_norm")]
            expect += ["vis_enc.beit3.encoder.layers.%d.ffn.%s.%s.%s" % (i, ab, n, t) for n in ("fc1", "fc2", "ffn_layernorm")
                       for t in ("weight", "bias")]
    for j in range(3):
        for a in (0, 1):
            expect += ["head.transformer.decoder.layers.%d.attentions.%d.attn.%s" % (j, a, t) for t in
                       ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias")]
        expect += ["head.transformer.decoder.layers.%d.ffns.0.layers.0.0.weight" % j,
                   "head.transformer.decoder.layers.%d.ffns.0.layers.1.bias" % j, "head.transformer.decoder.layers.%d.norms.2.weight" % j]
    for j in range(2):
        expect += ["head.text_guided_query_generation_transformer.layers.%d.ffns.0.layers.0.0.weight" % j]
    missing = [k for k in expect if k not in keys]
    assert not missing, missing
    sd = m.state_dict()
    assert sd["vis_enc.beit3.encoder.embed_positions.A.weight"].shape == (400 + 3, 768)
    assert sd["vis_enc.beit3.encoder.embed_positions.B.weight"].shape == (1024, 768)
    assert sd["vis_enc.beit3.vision_embed.proj.weight"].shape == (768, 3, 32, 32)
    assert sd["head.transformer.decoder.layers.0.ffns.0.layers.0.0.weight"].shape == (2048, 256)
    assert sd["head.text_guided_query_generation_transformer.layers.0.ffns.0.layers.0.0.weight"].shape == (512, 256)
    assert sd["head.transformer.decoder.layers.0.attentions.1.attn.in_proj_weight"].shape == (768, 256)
    assert m.vis_enc.hidden_size == 768 and m.vis_enc.get_num_layers() == 12 and m.fp16_enabled is False
    assert all("vis_enc" in n for n, _ in m.vis_enc.named_parameters(prefix="vis_enc"))
    # aux weight dict built from decoder depth (tgqs_kd_detr_head.py:174-180)
    assert set(m.head.criterion.weight_dict) == {"loss_class", "loss_bbox", "loss_giou", "loss_class_0", "loss_bbox_0",
                                                 "loss_giou_0", "loss_class_1", "loss_bbox_1", "loss_giou_1"}


def test_vit_large_config_and_droppath_quirk():
    from simvg_b200.models.vis_encs.beit.beit3 import BEIT3
    with pytest.raises(TypeError):
        BEIT3(img_size=64, patch_size=32, vit_type="huge", vocab_size=16)
    base = BEIT3(img_size=64, patch_size=32, vit_type="base", vocab_size=16, drop_path_rate=0.1)
    assert base.drop_path_probs[0] == 0.0 and abs(base.drop_path_probs[-1] - 0.1) < 1e-12 and len(base.drop_path_probs) == 12


def test_flat_buffer_views_grads_and_relocation():
    from simvg_b200.flat import FlatBuffer
    lin = torch.nn.Linear(5, 3)
    ln = torch.nn.LayerNorm(7)
    params = list(lin.named_parameters()) + list(ln.named_parameters())
    before = [p.detach().clone() for _, p in params]
    fb = FlatBuffer(params)
    assert fb.is_valid()
    for (n, p), b in zip(params, before):
        assert torch.equal(p, b)
        assert p.data_ptr() >= fb.data.data_ptr() and p.data_ptr() % 16 == 0
    fb.attach_grads()
    lin(torch.randn(2, 5)).sum().backward()
    assert lin.weight.grad.data_ptr() == fb.grad_of(0).data_ptr()
    assert float(fb.grad.abs().sum()) > 0
    fb.zero_grad()
    assert float(fb.grad.abs().sum()) == 0
    lin.weight.data = lin.weight.data.clone()      # what model.to(device) does
    assert not fb.is_valid()
    fb.ensure()
    assert fb.is_valid() and torch.equal(lin.weight, before[0])
    lin.zero_grad(set_to_none=True)
    ln.zero_grad(set_to_none=True)
    fb.attach_grads()
    assert lin.weight.grad is not None and lin.weight.grad.data_ptr() == fb.grad_of(0).data_ptr()


def test_encoder_flat_layout_makes_qkv_contiguous():
    from simvg_b200.models.vis_encs.beit.beit3 import _GROUP_FIELDS, BEIT3
    enc = BEIT3(img_size=64, patch_size=32, vit_type="base", vocab_size=32)
    fb = enc.flat()
    i = enc._n_global + _GROUP_FIELDS.index("q_w")
    D = 768
    assert fb.offsets[i + 1] - fb.offsets[i] == D * D and fb.offsets[i + 2] - fb.offsets[i + 1] == D * D
    sa = enc.beit3.encoder.layers[0].self_attn
    w = fb.data[fb.offsets[i]:fb.offsets[i] + 3 * D * D].view(3 * D, D)
    assert torch.equal(w[:D], sa.q_proj.A.weight) and torch.equal(w[D:2 * D], sa.k_proj.A.weight) and torch.equal(w[2 * D:], sa.v_proj.A.weight)
    # state-dict round trip keeps the views
    sd = {k: v.clone() + 1 for k, v in enc.state_dict().items()}
    enc.load_state_dict(sd)
    assert fb.is_valid() and torch.equal(w[:D], sd["beit3.encoder.layers.0.self_attn.q_proj.A.weight"])


def test_criterion_matches_oracle_on_cpu():
    """The product's SetCriterion / matcher / target preparation are plain PyTorch: check them against the oracle here."""
    from oracle import simvg_oracle as O
    from simvg_b200.core.criterion.criterion import HungarianMatcher, SetCriterion
    torch.manual_seed(0)
    for nq in (1, 7):
        B = 4
        logits = torch.randn(3, B, nq, 2)
        boxes = torch.rand(3, B, nq, 4) * 0.4 + 0.2
        tg = [{"labels": torch.zeros(1, dtype=torch.int64), "boxes": torch.rand(1, 4) * 0.3 + 0.3} for _ in range(B)]
        crit = SetCriterion(1, HungarianMatcher(1, 5.0, 2.0, "ce_cost"), {"loss_class": 1, "loss_bbox": 5.0, "loss_giou": 2.0},
                            loss_class_type="ce_loss", eos_coef=0.1)
        out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
               "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}
        got = crit(out, tg)
        want = O.set_criterion(out, tg)
        assert set(got) == set(want)
        for k in want:
            assert torch.allclose(got[k], want[k], rtol=1e-6, atol=1e-7), (nq, k)


def test_batched_rec_criterion_equals_generic_path():
    """The one-query / one-box fast path (no per-sample indexing, no matcher) must reproduce the generic path exactly."""
    from simvg_b200.core.criterion.criterion import BatchedTargets, HungarianMatcher, SetCriterion
    torch.manual_seed(3)
    B = 6
    logits, boxes = torch.randn(3, B, 1, 2), torch.rand(3, B, 1, 4) * 0.4 + 0.2
    gt = torch.rand(B, 4) * 0.3 + 0.3
    wgt = torch.rand(B)
    zeros = torch.zeros(B, 1, dtype=torch.int64)
    plain = [{"labels": zeros[i], "boxes": gt[i:i + 1], "weight": wgt[i:i + 1]} for i in range(B)]
    batched = BatchedTargets(plain, boxes=gt, labels=zeros[:, 0], weight=wgt)
    out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
           "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]}
    for kind in ("ce_loss", "weighted_ce_loss"):
        crit = SetCriterion(1, HungarianMatcher(1, 5.0, 2.0, "ce_cost"), {"loss_class": 1, "loss_bbox": 5.0, "loss_giou": 2.0},
                            loss_class_type=kind, eos_coef=0.1)
        a, b = crit(out, plain), crit(out, batched)
        assert set(a) == set(b)
        for k in a:
            assert torch.allclose(a[k], b[k], rtol=1e-6, atol=1e-7), (kind, k, float(a[k]), float(b[k]))


def test_get_predictions_matches_oracle_on_cpu():
    from oracle import simvg_oracle as O
    from simvg_b200.models.det_seg.mix_detr_mb import MIXDETRMB
    torch.manual_seed(1)
    metas = [dict(img_shape=(320, 480, 3), scale_factor=[1, 1, 1, 1]) for _ in range(5)]
    for nq in (1, 6):
        out = {"pred_logits": torch.randn(5, nq, 2), "pred_boxes": torch.rand(5, nq, 4) * 0.5 + 0.25}
        got = MIXDETRMB.get_predictions(None, out, metas)
        want = O.get_predictions(out, metas)
        assert torch.allclose(got["pred_bboxes"], want["pred_bboxes"], atol=1e-4)
        assert torch.equal(got["predict_classes"], want["predict_classes"])


def test_head_cpu_forward_matches_oracle_small():
    """Head wiring (TGQG, token branch, decoder branch, DWBD losses) on CPU tensors small enough to stay off the GEMM path."""
    import copy
    from oracle import simvg_oracle as O
    from simvg_b200.models.heads.tgqs_kd_detr_head import transformer as T
    from simvg_b200.models.heads.tgqs_kd_detr_head.tgqs_kd_detr_head import TextGuidedQuerySelectKDDETRHead
    from tools.synth import make_batch, model_cfg, synth_state_dict
    import simvg_b200.ops as ops
    hc = copy.deepcopy(model_cfg("base", 64, 32)["head"])
    hc["in_channels"] = 128
    hc.pop("type")
    head = TextGuidedQuerySelectKDDETRHead(**copy.deepcopy(hc)).eval()
    sd = synth_state_dict({k: v.float() for k, v in head.state_dict().items()}, seed=5)
    head.load_state_dict(sd)
    old = ops.linear
    ops.linear = lambda x, W, b=None: torch.nn.functional.linear(x, W, b)   # CPU stand-in for the GEMM in THIS TEST ONLY
    try:
        B = 3
        g = torch.Generator().manual_seed(2)
        x_mm = torch.randn(B, 128, 2, 2, generator=g)
        text, cls = torch.randn(B, 20, 128, generator=g), torch.randn(B, 128, generator=g)
        batch = make_batch(B, 64, seed=3)
        metas = batch["img_metas"]
        for m in metas:
            m["batch_input_shape"] = (64, 64)
        losses, out = head.forward_train(x_mm, metas, cls_feat=cls, text_feat=text, gt_bbox=batch["gt_bbox"],
                                         text_mask=batch["text_attention_mask"])
    finally:
        ops.linear = old
    ol, oo = O.head_forward_train({"head." + k: v for k, v in sd.items()}, hc, x_mm, copy.deepcopy(metas), cls, text,
                                  batch["gt_bbox"], batch["text_attention_mask"])
    for k in ol:
        assert torch.allclose(losses[k], ol[k], rtol=2e-5, atol=1e-6), (k, float(losses[k]), float(ol[k]))
    assert torch.allclose(out["outputs_coord_decoder_branch"], oo["outputs_coord_decoder_branch"], atol=1e-5)
    assert torch.allclose(out["outputs_coord_token_branch"], oo["outputs_coord_token_branch"], atol=1e-5)
    assert T._BIG_ROWS > 0


def test_absorbed_cross_attention_equals_projected_path():
    """The decoder's few-queries-vs-long-memory attention absorbs the key/value projections into the query/output side.  It is
    a reassociation of nn.MultiheadAttention's arithmetic: outputs and every gradient must match the projected form."""
    from simvg_b200.models.heads.tgqs_kd_detr_head import transformer as T
    torch.manual_seed(11)
    B, nq, nk, E, H = 3, 2, 300, 64, 4
    att = T._Attention(E, H, 0.0).double()
    with torch.no_grad():
        att.attn.in_proj_bias.normal_(0, 0.3)
        att.attn.out_proj.bias.normal_(0, 0.3)
    mask = torch.zeros(B, nk, dtype=torch.bool)
    mask[1, 250:] = True
    res = []
    for min_keys in (10 ** 9, 1):
        old = T._ABSORB_MIN_KEYS
        T._ABSORB_MIN_KEYS = min_keys
        try:
            g = torch.Generator().manual_seed(5)
            q = torch.randn(B, nq, E, generator=g, dtype=torch.double, requires_grad=True)
            mem = torch.randn(B, nk, E, generator=g, dtype=torch.double, requires_grad=True)
            qpos = torch.randn(B, nq, E, generator=g, dtype=torch.double)
            kpos = torch.randn(B, nk, E, generator=g, dtype=torch.double)
            att.zero_grad()
            out = att(q, mem, mem, query_pos=qpos, key_pos=kpos, key_padding_mask=mask)
            (out * torch.randn(out.shape, generator=g, dtype=torch.double)).sum().backward()
            res.append((out.detach(), q.grad, mem.grad, att.attn.in_proj_weight.grad.clone(), att.attn.in_proj_bias.grad.clone(),
                        att.attn.out_proj.weight.grad.clone()))
        finally:
            T._ABSORB_MIN_KEYS = old
    for a, b in zip(*res):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-11)


def test_wgrad_split_choice_fills_whole_waves():
    """Host-side split-K heuristic for the weight-gradient GEMMs (pure Python): the chosen factor keeps >= 32 k-blocks per
    split and never wastes most of a wave of the 74-cluster persistent grid on the in-step shapes."""
    from simvg_b200.kernels import wgrad_splits
    R = 64 * 1601
    want = {(768, 3072): 2, (3072, 768): 2, (2304, 768): 8, (768, 768): 8}     # measured best on B200 (tools/wgrad_ks_ab.py)
    for (M, N), ks in want.items():
        assert wgrad_splits(M, N, R) == ks
    for M, N, K_ in [(768, 3072, R), (1024, 4096, 32 * 1601), (2304, 768, 1280), (256, 256, 102400), (768, 72, 4096), (100, 768, 640)]:
        ks = wgrad_splits(M, N, K_)
        kb = (K_ + 63) // 64
        assert 1 <= ks <= max(1, kb // 32)
        if N > 128 and M >= 256 and ks > 1:
            tiles = ((M + 255) // 256) * ((N + 255) // 256)
            items = tiles * ks
            assert items / (74 * ((items + 73) // 74)) > 0.6     # at least 60 % of the last wave's slots are used on average


def test_head_absorbed_attention_and_batched_losses_match_oracle():
    """Head with a 16x16 feature map (256 image tokens: the decoder's cross-attention takes the absorbed-projection form) and
    a 6-layer decoder (batched auxiliary losses): losses, outputs and the gradient w.r.t. the image features against the oracle."""
    import copy
    from oracle import simvg_oracle as O
    from simvg_b200.models.heads.tgqs_kd_detr_head import transformer as T
    from simvg_b200.models.heads.tgqs_kd_detr_head.tgqs_kd_detr_head import TextGuidedQuerySelectKDDETRHead
    from tools.synth import make_batch, model_cfg, synth_state_dict
    import simvg_b200.ops as ops
    assert T._ABSORB_MIN_KEYS <= 256
    hc = copy.deepcopy(model_cfg("base", 512, 32, num_decoder_layers=6)["head"])
    hc["in_channels"] = 96
    hc.pop("type")
    head = TextGuidedQuerySelectKDDETRHead(**copy.deepcopy(hc)).eval()
    sd = synth_state_dict({k: v.float() for k, v in head.state_dict().items()}, seed=9)
    head.load_state_dict(sd)
    old = ops.linear
    ops.linear = lambda x, W, b=None: torch.nn.functional.linear(x, W, b)   # CPU stand-in for the GEMM in THIS TEST ONLY
    try:
        B = 2
        g = torch.Generator().manual_seed(4)
        x_mm = torch.randn(B, 96, 16, 16, generator=g, requires_grad=True)
        text, cls = torch.randn(B, 20, 96, generator=g), torch.randn(B, 96, generator=g)
        batch = make_batch(B, 512, seed=6)
        metas = batch["img_metas"]
        for m in metas:
            m["batch_input_shape"] = (512, 512)
        losses, out = head.forward_train(x_mm, metas, cls_feat=cls, text_feat=text, gt_bbox=batch["gt_bbox"],
                                         text_mask=batch["text_attention_mask"])
        losses["loss_total"].backward()
        g_ours = x_mm.grad.clone()
    finally:
        ops.linear = old
    x2 = x_mm.detach().clone().requires_grad_(True)
    ol, oo = O.head_forward_train({"head." + k: v for k, v in sd.items()}, hc, x2, copy.deepcopy(metas), cls, text,
                                  batch["gt_bbox"], batch["text_attention_mask"])
    ol["loss_total"].backward()
    assert set(ol) <= set(losses)
    for k in ol:
        assert torch.allclose(losses[k], ol[k], rtol=5e-5, atol=1e-6), (k, float(losses[k]), float(ol[k]))
    assert torch.allclose(out["outputs_coord_decoder_branch"], oo["outputs_coord_decoder_branch"], atol=2e-5)
    assert out["outputs_coord_decoder_branch"].shape[0] == 6
    rel_g = ((g_ours - x2.grad).norm() / x2.grad.norm()).item()
    assert rel_g < 1e-4, rel_g


def test_state_dict_keys_equal_the_reference_exactly(golden_dir):
    """Set-equality (names AND shapes) with the reference's own build_model state dict, ViT-B and ViT-L: checkpoints are
    exchanged by key, so one missing / extra / mis-shaped entry breaks the drop-in (fixture: oracle/make_golden.py)."""
    from simvg_b200.models import build_model
    from tools.synth import model_cfg
    want = torch.load(os.path.join(golden_dir, "state_dict_keys.pt"), weights_only=False)
    for vit in ("base", "large"):
        m = build_model(model_cfg(vit, 640, 16, num_decoder_layers=3))
        got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert set(got) == set(want[vit]), (sorted(set(got) ^ set(want[vit]))[:10])
        assert got == want[vit], [k for k in got if got[k] != want[vit][k]][:10]
        del m


def test_checkpoint_interpolation_matches_the_reference(golden_dir, tmp_path):
    """BEIT3.load_model_and_may_interpolate against the reference's own method (beit3.py:92-174): a 2x2-patch, P=32 checkpoint
    loaded into an 8x8-patch, P=16 model (bicubic position-embedding + patch-projection interpolation)."""
    from simvg_b200.models.vis_encs.beit.beit3 import BEIT3
    fx = torch.load(os.path.join(golden_dir, "interpolate.pt"), weights_only=False)
    g = torch.Generator().manual_seed(fx["seed"])
    ck = {"beit3.encoder.embed_positions.A.weight": torch.randn(2 * 2 + 3, 768, generator=g),
          "beit3.vision_embed.proj.weight": torch.randn(768, 3, 32, 32, generator=g) * 0.02,
          "beit3.vision_embed.proj.bias": torch.randn(768, generator=g) * 0.02}
    path = str(tmp_path / "ck.pth")
    torch.save({"model": ck}, path)
    enc = BEIT3(img_size=128, patch_size=16, vit_type="base", drop_path_rate=0.0, vision_embed_proj_interpolate=True, pretrain=None)
    enc.load_model_and_may_interpolate(path)
    sd = enc.state_dict()
    pe, pw = sd["beit3.encoder.embed_positions.A.weight"], sd["beit3.vision_embed.proj.weight"]
    assert pe.shape == (67, 768) and pw.shape == (768, 3, 16, 16)
    assert torch.allclose(pe[:, ::32], fx["pos_slice"], atol=1e-6) and abs(float(pe.double().norm()) - fx["pos_norm"]) < 1e-4 * fx["pos_norm"]
    assert torch.allclose(pw[::16, :, ::2, ::2], fx["proj_slice"], atol=1e-7)
    assert abs(float(pw.double().norm()) - fx["proj_norm"]) < 1e-4 * fx["proj_norm"]
    assert abs(float(sd["beit3.vision_embed.proj.bias"].double().norm()) - fx["bias_norm"]) < 1e-5
    # the constructor path (pretrain=<file>) goes through the same loader
    enc2 = BEIT3(img_size=128, patch_size=16, vit_type="base", vision_embed_proj_interpolate=True, pretrain=path)
    assert torch.equal(enc2.state_dict()["beit3.encoder.embed_positions.A.weight"], pe)


def test_token_branch_only_inference_skips_the_decoder():
    """head.only_token = True (the variable the reference hard-codes to False, tgqs_kd_detr_head.py:422): the token branch's
    prediction is unchanged, the decoder branch returns the reference's `None` outputs and neither input_proj nor the decoder
    runs."""
    from simvg_b200.models.heads.tgqs_kd_detr_head.tgqs_kd_detr_head import TextGuidedQuerySelectKDDETRHead
    from tools.synth import make_batch, model_cfg
    hc = dict(model_cfg("base", 64, 32)["head"])
    hc.pop("type")
    hc["in_channels"] = 64
    torch.manual_seed(0)
    head = TextGuidedQuerySelectKDDETRHead(**hc).eval()
    B = 2
    x_mm, text, cls = torch.randn(B, 64, 2, 2), torch.randn(B, 20, 64), torch.randn(B, 64)
    b = make_batch(B, 64)
    metas = b["img_metas"]
    for m in metas:
        m["batch_input_shape"] = (64, 64)
    import simvg_b200.ops as ops
    real = ops.linear
    ops.linear = lambda x, W, bias=None: torch.nn.functional.linear(x, W, bias)     # CPU stand-in for the tcgen05 GEMM
    try:
        with torch.no_grad():
            full = head.forward_test(x_mm, metas, text_feat=text, cls_feat=cls, text_mask=b["text_attention_mask"])
            head.only_token = True
            calls = []
            ops.linear = lambda *a, **k: calls.append(1)
            tok = head.forward_test(x_mm, metas, text_feat=text, cls_feat=cls, text_mask=b["text_attention_mask"])
    finally:
        ops.linear = real
    assert not calls
    assert tok["decoder_branch_output"] == {"pred_logits": None, "pred_boxes": None} and tok["decoder_features"] is None
    assert torch.equal(tok["token_branch_output"]["pred_boxes"], full["token_branch_output"]["pred_boxes"])
    assert torch.equal(tok["token_branch_output"]["pred_logits"], full["token_branch_output"]["pred_logits"])


def test_chunked_backward_plan_tiles_the_encoder_gradient_buffer():
    """The multi-GPU graph runtime cuts the encoder backward into chunks of layers and all-reduces each chunk's flat gradient
    range while the next chunk runs: the chunks must cover every layer exactly once, top-down, layer 0 last and alone, and
    their ranges (plus the embedding / final-LN parameters that travel with layer 0) must tile the buffer without overlap."""
    from simvg_b200.models.vis_encs.beit.beit3 import BEIT3, layer_flat_range
    from simvg_b200.runtime import GraphedTrainStep
    enc = BEIT3(img_size=64, patch_size=32, vit_type="base", vocab_size=32)
    nl = enc.cfg["layers"]
    fb = enc.flat()
    for cl in (1, 2, 3, 5, 12):
        step = GraphedTrainStep.__new__(GraphedTrainStep)
        step.chunk_layers = cl
        chunks = step._chunks(nl)
        assert chunks[-1] == (0, 0) and chunks[0][0] == nl - 1
        seen = [li for hi, lo in chunks for li in range(hi, lo - 1, -1)]
        assert seen == list(range(nl - 1, -1, -1)), (cl, chunks)
        ranges = [(layer_flat_range(enc, lo)[0], layer_flat_range(enc, hi)[1]) if lo > 0 else (0, layer_flat_range(enc, 0)[1])
                  for hi, lo in chunks]
        ranges.sort()
        assert ranges[0][0] == 0 and ranges[-1][1] == fb.numel
        assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:])), (cl, ranges)


def test_grec_predictions_batched_equal_the_per_image_path():
    """get_predictions_grec (all non-empty boxes per image, mix_detr_mb.py:161-190) batched on the device vs the oracle's
    per-image restatement: images of different sizes, rescale on / off, and a batch containing degenerate (empty after clipping)
    boxes, which must be dropped exactly where detectron2's `nonempty()` drops them."""
    from oracle import simvg_oracle as O
    from simvg_b200.models.det_seg.mix_detr_mb import MIXDETRMB
    g = torch.Generator().manual_seed(3)
    B, nq = 3, 10
    metas = [{"img_shape": (64, 96, 3), "scale_factor": [1.5, 1.0, 1.5, 1.0]}, {"img_shape": (80, 80, 3), "scale_factor": [2.0, 2.0, 2.0, 2.0]},
             {"img_shape": (50, 70, 3), "scale_factor": [0.5, 0.7, 0.5, 0.7]}]
    logits = torch.randn(B, nq, 2, generator=g)
    boxes = torch.rand(B, nq, 4, generator=g) * 0.5 + 0.2
    model = MIXDETRMB.__new__(MIXDETRMB)       # the method only touches its arguments
    for degenerate in (False, True):
        bx = boxes.clone()
        if degenerate:
            bx[0, 3] = torch.tensor([1.4, 0.5, 0.2, 0.2])      # entirely right of the image: zero width after clipping
            bx[2, 0] = torch.tensor([0.5, 0.5, 0.0, 0.3])      # zero width
        out = {"pred_logits": logits, "pred_boxes": bx}
        for rescale in (False, True):
            got = MIXDETRMB.get_predictions_grec(model, out, metas, rescale=rescale)["pred_bboxes"]
            want = O.get_predictions_grec(out, metas, rescale=rescale)["pred_bboxes"]
            assert len(got) == len(want) == B
            for a, b in zip(got, want):
                assert a["boxes"].shape == b["boxes"].shape, (degenerate, a["boxes"].shape, b["boxes"].shape)
                assert torch.allclose(a["boxes"], b["boxes"], atol=1e-5) and torch.equal(a["labels"], b["labels"])
                assert torch.allclose(a["scores"], b["scores"])
        if degenerate:
            assert got[0]["boxes"].shape[0] == nq - 1 and got[2]["boxes"].shape[0] == nq - 1 and got[1]["boxes"].shape[0] == nq


def test_cabi_round2_entry_points_validate_without_gpu(lib):
    """The entry points added in round 2 follow the same convention: argument errors return < 0 with a message before any
    device call (so they can be exercised on a box without a GPU); the cross-attention scratch size is a pure function."""
    from simvg_b200 import kernels as K
    K._lib_setup()
    # head linear / LayerNorm / attention: null pointers and unsupported shapes
    a = K.HeadLinArgs()
    assert lib.simvgb_head_lin_fwd(ctypes.byref(a), None) < 0 and b"null pointer" in lib.simvgb_last_error()
    assert lib.simvgb_head_lin_bwd(ctypes.byref(a), None) < 0
    ln = K.HeadLnArgs()
    buf = (ctypes.c_float * 8)()
    ln.a = ln.gamma = ctypes.cast(buf, ctypes.c_void_p)
    ln.R, ln.C = 4, 384
    assert lib.simvgb_head_lnres(ctypes.byref(ln), 0, None) < 0 and b"256 or 512" in lib.simvgb_last_error()
    at = K.HeadAttnArgs()
    at.q = at.k = at.v = ctypes.cast(buf, ctypes.c_void_p)
    at.B, at.nq, at.nk, at.H = 1, 1, 33, 8
    assert lib.simvgb_head_attn_small(ctypes.byref(at), 0, None) < 0 and b"nk <= 32" in lib.simvgb_last_error()
    xa = K.HeadXAttnArgs()
    for f in ("q", "kin", "val", "Wk", "bk", "Wv", "bv"):
        setattr(xa, f, ctypes.cast(buf, ctypes.c_void_p))
    xa.B, xa.nq, xa.N, xa.E, xa.H = 2, 1, 100, 128, 8
    assert lib.simvgb_head_xattn(ctypes.byref(xa), 0, None) < 0 and b"E = 256" in lib.simvgb_last_error()
    xa.E = 256
    assert lib.simvgb_head_xattn(ctypes.byref(xa), 0, None) < 0 and b"workspace" in lib.simvgb_last_error()
    # scratch: forward = u + c + NC partial sums; NC = min(16, key tiles of 32); backward adds dz, dps, dc, the score gradients, du
    ws = lib.simvgb_head_xattn_ws_floats
    R, N = 64, 1600
    nc = min(16, (N + 31) // 32)
    assert ws(64, 1, N, 0) == R * 2048 + R * 8 + nc * R * 2048
    assert ws(64, 1, N, 1) == 2 * (R * 2048 + R * 8) + 2 * R * 8 + R * 8 * N + nc * R * 2048 + R * 2048 - R * 8
    assert ws(2, 10, 40, 0) == 20 * 2048 + 20 * 8 + 2 * 20 * 2048
    assert ws(0, 1, 10, 0) < 0
    # Hungarian / uint8 patch gather
    assert lib.simvgb_hungarian(None, 1, 1, 1, None, None, None, 1, None) < 0
    i64 = (ctypes.c_int64 * 4)()
    i32 = (ctypes.c_int32 * 4)()
    assert lib.simvgb_hungarian(buf, 1, 33, 1, i32, i64, i64, 1, None) < 0 and b"bad shape" in lib.simvgb_last_error()
    mean = (ctypes.c_float * 3)(1, 1, 1)
    std0 = (ctypes.c_float * 3)(1, 0, 1)
    assert lib.simvgb_im2col_patch_u8(buf, buf, 1, 64, 12, mean, mean, 1, None) < 0      # P % 8 != 0
    assert lib.simvgb_im2col_patch_u8(buf, buf, 1, 64, 16, mean, std0, 1, None) < 0 and b"zero std" in lib.simvgb_last_error()
    h, m = ctypes.c_longlong(-1), ctypes.c_longlong(-1)
    lib.simvgb_tmap_cache_stats(ctypes.byref(h), ctypes.byref(m))
    assert h.value >= 0 and m.value >= 0
