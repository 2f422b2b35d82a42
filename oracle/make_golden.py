"""Generates tests/golden/*.pt by executing the REFERENCE's own files (imported verbatim from /root/reference over the
leaf-op shims in oracle/shims) on seeded synthetic inputs/weights, and cross-checks the oracle against them.

    python oracle/make_golden.py            (build container only: needs /root/reference)

Fixtures store inputs' seeds and the reference's outputs only; weights are regenerated from (seed, parameter name) by
tools/synth.synth_state_dict, so every fixture is a few KB.
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import simvg_oracle as O  # noqa: E402
from oracle.shims import load_reference  # noqa: E402
from tools.synth import make_batch, model_cfg, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.Tensor.cuda = lambda self, *a, **k: self   # tgqs_kd_detr_head.py:389 hard-codes .cuda()


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def golden_encoder_small():
    """Reference BEiT3 / Encoder / EncoderLayer (beit3_base.py) at small dims: D=128, H=2, F=512, 2 layers."""
    import importlib
    base = importlib.import_module("simvg.models.vis_encs.beit.beit3_base")
    from torchscale.architecture.config import EncoderConfig
    args = EncoderConfig(img_size=64, patch_size=16, vocab_size=97, multiway=True, layernorm_embedding=False,
                         normalize_output=True, no_output_layer=True, drop_path_rate=0.0, encoder_embed_dim=128,
                         encoder_attention_heads=2, encoder_ffn_embed_dim=512, encoder_layers=2, checkpoint_activations=None)
    torch.manual_seed(0)
    ref = base.BEiT3(args).double().eval()
    sd = synth_state_dict({k: v.float() for k, v in ref.state_dict().items()}, seed=11)
    ref.load_state_dict({k: v.double() for k, v in sd.items()})
    g = torch.Generator().manual_seed(5)
    img = torch.randn(2, 3, 64, 64, generator=g).double()   # fp32-representable so the stored copy is exact
    ids = torch.randint(0, 97, (2, 8), generator=g)
    pad = torch.tensor([[0, 0, 0, 0, 0, 1, 1, 1], [0, 0, 0, 0, 0, 0, 0, 0]], dtype=torch.int64)
    out = ref(textual_tokens=ids, visual_tokens=img, text_padding_position=pad)["encoder_out"]
    w = torch.randn(out.shape, generator=g).double()
    (out * w).sum().backward()
    grads = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    # oracle cross-check (fp64)
    osd = {"beit3." + k: v.double().requires_grad_(True) for k, v in sd.items()}
    cfg = dict(D=128, H=2, F=512, layers=2, img_size=64, patch_size=16, eps=1e-5, drop_path_rate=0.0)
    oo = O.encoder_forward(osd, cfg, img, ids, pad)
    (oo * w).sum().backward()
    e = rel(oo, out)
    ge = max(rel(osd["beit3." + k].grad, v) for k, v in grads.items())
    print("encoder_small: oracle vs reference  out %.2e  worst grad %.2e" % (e, ge))
    assert e < 1e-10 and ge < 1e-8
    keep = ["encoder.layers.0.self_attn.q_proj.A.weight", "encoder.layers.1.ffn.B.fc1.bias", "encoder.embed_positions.A.weight",
            "text_embed.weight", "vision_embed.cls_token", "encoder.layers.1.self_attn.inner_attn_ln.B.weight",
            "encoder.layer_norm.A.bias", "encoder.layers.0.ffn.A.ffn_layernorm.weight"]
    torch.save({"weight_seed": 11, "img": img.float(), "ids": ids, "pad": pad, "probe": w.float(), "encoder_out": out.detach(),
                "grads": {k: grads[k] for k in keep}, "grad_norms": {k: float(v.norm()) for k, v in grads.items()},
                "cfg": cfg}, os.path.join(OUT, "encoder_small.pt"))


def _run_reference_model(cfg, batch, weight_seed, dtype):
    _, _, _, build_model = load_reference()
    torch.manual_seed(6666)
    ref = build_model(copy.deepcopy(cfg)).eval()
    sd = synth_state_dict({k: v.float() for k, v in ref.state_dict().items()}, seed=weight_seed)
    ref.load_state_dict(sd)
    ref = ref.to(dtype)
    metas = copy.deepcopy(batch["img_metas"])
    losses, preds = ref(batch["img"].to(dtype), batch["ref_expr_inds"], metas, return_loss=True,
                        text_attention_mask=batch["text_attention_mask"], gt_bbox=batch["gt_bbox"], rescale=False)
    losses["loss_total"].backward()
    grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    with torch.no_grad():
        feats = ref.vis_enc(batch["img"].to(dtype), batch["ref_expr_inds"], batch["text_attention_mask"])
    return ref, sd, losses, preds, grads, feats


def golden_cfg1():
    """BASELINE configs[0]: ViT-B/16, 224x224, 20-token text, bs=2, CPU eager — full MIXDETRMB train step of the reference."""
    cfg = model_cfg("base", 224, 16, num_decoder_layers=3)
    batch = make_batch(2, 224, seed=6666)
    ref, sd, losses, preds, grads, feats = _run_reference_model(cfg, batch, weight_seed=21, dtype=torch.float32)
    # oracle cross-check (fp32, same weights)
    osd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "empty_weight" not in k) for k, v in sd.items()}
    om = O.OracleModel(osd, "base", 224, 16, cfg["head"])
    ol, op, oout = om.forward_train(batch["img"], batch["ref_expr_inds"], copy.deepcopy(batch["img_metas"]),
                                    batch["text_attention_mask"], batch["gt_bbox"])
    ol["loss_total"].backward()
    for k in losses:
        d = abs(float(ol[k]) - float(losses[k])) / max(abs(float(losses[k])), 1e-12)
        print("cfg1 %-16s ref %.6f oracle %.6f rel %.1e" % (k, float(losses[k]), float(ol[k]), d))
        assert d < 2e-4, k
    pe = rel(op[0]["pred_bboxes"], preds[0]["pred_bboxes"]), rel(op[1]["pred_bboxes"], preds[1]["pred_bboxes"])
    ge = sorted((rel(osd[k].grad, v), k) for k, v in grads.items() if v.norm() > 0)
    print("cfg1 preds rel", pe, " grads: median %.1e worst %.1e (%s)" % (ge[len(ge) // 2][0], ge[-1][0], ge[-1][1]))
    assert max(pe) < 1e-4 and ge[-1][0] < 5e-2 and ge[len(ge) // 2][0] < 1e-3
    torch.save({
        "weight_seed": 21, "batch_seed": 6666, "B": 2, "S": 224, "P": 16, "head_cfg": cfg["head"],
        "losses": {k: float(v) for k, v in losses.items()},
        "pred_dec": preds[0]["pred_bboxes"].detach(), "pred_tok": preds[1]["pred_bboxes"].detach(),
        "img_feat_slice": feats[0][:, ::28, ::64].detach().clone(), "text_feat_slice": feats[1][:, :, ::64].detach().clone(),
        "cls_feat": feats[2].detach().clone(),
        "grad_norms": {k: float(v.norm()) for k, v in grads.items()},
    }, os.path.join(OUT, "cfg1_train_step.pt"))


def golden_head(nq, blw, name):
    """Reference head alone (tgqs_kd_detr_head.py + transformer.py + criterion.py) on random features.  fp32: the
    reference's target preparation casts to .float() (tgqs_kd_detr_head.py:232-233), so the head cannot run in fp64."""
    _, Head, _, _ = load_reference()
    hc = dict(num_queries=nq, text_max_token=20, in_channels=128, embed_dim=256, decoder_freeze=False, num_classes=1,
              aux_loss=True, num_encoder_layers=6, num_decoder_layers=3, only_decoder=True, text_embed_aug=False,
              branch_loss_weight=blw, distill_type="hard_weighted", prepare_target_mode="score_iou_weighted",
              share_predicthead=False, num_token_mlp_layers=1, mlp_aux_loss=False, text_guided_query_generation=True,
              num_tgqg_layers=2)
    torch.manual_seed(1)
    head = Head(**copy.deepcopy(hc)).eval()
    sd = synth_state_dict({k: v.float() for k, v in head.state_dict().items()}, seed=31 + nq)
    head.load_state_dict(sd)
    g = torch.Generator().manual_seed(9)
    B, h, w = 3, 5, 5
    x_mm = torch.randn(B, 128, h, w, generator=g)
    text = torch.randn(B, 20, 128, generator=g)
    cls = torch.randn(B, 128, generator=g)
    tmask = torch.ones(B, 20, dtype=torch.int64)
    for b, k in enumerate((6, 20, 11)):
        tmask[b, :k] = 0
    batch = make_batch(B, 160, seed=77)
    metas = copy.deepcopy(batch["img_metas"])
    for m in metas:
        m["batch_input_shape"] = (160, 160)
    losses, out = head.forward_train(x_mm, metas, cls_feat=cls, text_feat=text, gt_bbox=batch["gt_bbox"], text_mask=tmask)
    osd = {"head." + k: v for k, v in sd.items()}
    ol, oout = O.head_forward_train(osd, hc, x_mm, copy.deepcopy(metas), cls, text, batch["gt_bbox"], tmask)
    for k in losses:
        d = abs(float(ol[k]) - float(losses[k])) / max(abs(float(losses[k])), 1e-12)
        print("%s %-16s ref %.8f oracle %.8f rel %.1e" % (name, k, float(losses[k]), float(ol[k]), d))
        assert d < 2e-5, k
    assert rel(oout["outputs_coord_decoder_branch"], out["outputs_coord_decoder_branch"]) < 1e-5
    fx = {"weight_seed": 31 + nq, "head_cfg": hc, "x_mm": x_mm.float(), "text": text.float(), "cls": cls.float(), "tmask": tmask,
          "batch_seed": 77, "S": 160, "losses": {k: float(v) for k, v in losses.items()},
          "dec_boxes": out["outputs_coord_decoder_branch"].detach(), "dec_logits": out["outputs_class_decoder_branch"].detach()}
    if out["outputs_coord_token_branch"] is not None:
        fx["tok_boxes"] = out["outputs_coord_token_branch"].detach()
        fx["tok_logits"] = out["outputs_class_token_branch"].detach()
    torch.save(fx, os.path.join(OUT, name + ".pt"))


def golden_interpolate():
    """Reference BEIT3.load_model_and_may_interpolate (beit3.py:92-174): a 2x2-patch / P=32 checkpoint loaded into an 8x8-patch /
    P=16 model -> bicubic position-embedding and patch-projection interpolation, run by the reference's own method."""
    import importlib
    import tempfile
    b3 = importlib.import_module("simvg.models.vis_encs.beit.beit3")
    torch.manual_seed(0)
    ref = b3.BEIT3(img_size=128, patch_size=16, vit_type="base", drop_path_rate=0.0, vocab_size=64010, freeze_layer=-1,
                   vision_embed_proj_interpolate=True, pretrain=None)
    g = torch.Generator().manual_seed(13)
    ck = {"beit3.encoder.embed_positions.A.weight": torch.randn(2 * 2 + 3, 768, generator=g),
          "beit3.vision_embed.proj.weight": torch.randn(768, 3, 32, 32, generator=g) * 0.02,
          "beit3.vision_embed.proj.bias": torch.randn(768, generator=g) * 0.02}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ck.pth")
        torch.save({"model": {k: v.clone() for k, v in ck.items()}}, path)
        ref.load_model_and_may_interpolate(path)
    sd = ref.state_dict()
    pe, pw = sd["beit3.encoder.embed_positions.A.weight"], sd["beit3.vision_embed.proj.weight"]
    assert pe.shape == (8 * 8 + 3, 768) and pw.shape == (768, 3, 16, 16)
    torch.save({"seed": 13, "pos_slice": pe[:, ::32].clone(), "pos_norm": float(pe.double().norm()),
                "proj_slice": pw[::16, :, ::2, ::2].clone(), "proj_norm": float(pw.double().norm()),
                "bias_norm": float(sd["beit3.vision_embed.proj.bias"].double().norm())}, os.path.join(OUT, "interpolate.pt"))
    print("interpolate: pos", tuple(pe.shape), "proj", tuple(pw.shape))


def golden_state_dict_keys():
    """The reference model's full {state-dict key: shape} map (ViT-B/16 640 config and ViT-L/16), from the reference's own
    build_model: the drop-in boundary requires set-equality (checkpoints are exchanged by key)."""
    _, _, _, build_model = load_reference()
    out = {}
    for vit in ("base", "large"):
        torch.manual_seed(0)
        m = build_model(copy.deepcopy(model_cfg(vit, 640, 16, num_decoder_layers=3)))
        out[vit] = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        del m
    torch.save(out, os.path.join(OUT, "state_dict_keys.pt"))
    print("state-dict keys:", {k: len(v) for k, v in out.items()})


def golden_grec_predictions():
    """gREC post-processing (all non-empty boxes per image) by the reference's own MIXDETRMB.get_predictions_grec +
    TextGuidedQuerySelectKDDETRHead.inference (mix_detr_mb.py:161-190, tgqs_kd_detr_head.py:577-604) over the detectron2 shims:
    images of different sizes, rescale on / off, boxes that are empty after clipping."""
    _, _, _, build_model = load_reference()
    torch.manual_seed(6666)
    ref = build_model(copy.deepcopy(model_cfg("base", 64, 32, num_decoder_layers=1, num_queries=10))).eval()
    g = torch.Generator().manual_seed(3)
    B, nq = 3, 10
    metas = [{"img_shape": (64, 96, 3), "scale_factor": [1.5, 1.0, 1.5, 1.0]}, {"img_shape": (80, 80, 3), "scale_factor": [2.0, 2.0, 2.0, 2.0]},
             {"img_shape": (50, 70, 3), "scale_factor": [0.5, 0.7, 0.5, 0.7]}]
    logits = torch.randn(B, nq, 2, generator=g)
    boxes = torch.rand(B, nq, 4, generator=g) * 0.5 + 0.2
    boxes[0, 3] = torch.tensor([1.4, 0.5, 0.2, 0.2])      # entirely right of the image: zero width after clipping
    boxes[2, 0] = torch.tensor([0.5, 0.5, 0.0, 0.3])      # zero width
    fixture = {"logits": logits, "boxes": boxes, "metas": metas}
    for rescale in (False, True):
        with torch.no_grad():
            got = ref.get_predictions_grec({"pred_logits": logits.clone(), "pred_boxes": boxes.clone()}, copy.deepcopy(metas),
                                           rescale=rescale)["pred_bboxes"]
        want = O.get_predictions_grec({"pred_logits": logits, "pred_boxes": boxes}, metas, rescale=rescale)["pred_bboxes"]
        for a, b in zip(got, want):
            assert a["boxes"].shape == b["boxes"].shape and torch.allclose(a["boxes"], b["boxes"], atol=1e-5)
            assert torch.equal(a["labels"], b["labels"]) and torch.allclose(a["scores"], b["scores"])
        fixture["rescale_%d" % int(rescale)] = [{k: v.detach().clone() for k, v in d.items()} for d in got]
    print("grec predictions: boxes kept per image", [len(d["boxes"]) for d in fixture["rescale_0"]])
    torch.save(fixture, os.path.join(OUT, "grec_predictions.pt"))


class _EmaToy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.LayerNorm(5)
        self.register_buffer("running", torch.zeros(5))


def golden_ema():
    """ExponentialMovingAverage by the reference's own class (simvg/models/utils.py:130-180; the file only needs torch / numpy, so it
    is loaded directly rather than through the shim package): five update_params() calls over seeded parameter / buffer changes,
    then apply_shadow() / restore()."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_models_utils", "/root/reference/simvg/models/utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(17)
    model = _EmaToy()
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ema = mod.ExponentialMovingAverage(model, 0.9)
    for step in range(5):
        g = torch.Generator().manual_seed(100 + step)
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            model.running.add_(torch.randn(5, generator=g))
        ema.update_params()
    live = {k: v.detach().clone() for k, v in model.state_dict().items()}
    shadow = {k: v.detach().clone() for k, v in ema.shadow.items()}
    ema.apply_shadow()
    applied = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ema.restore()
    assert all(torch.equal(model.state_dict()[k], live[k]) for k in live)
    torch.save({"init": init, "steps": 5, "alpha": 0.9, "live": live, "shadow": shadow, "applied": applied, "ema_step": ema.step},
               os.path.join(OUT, "ema.pt"))
    print("ema: step", ema.step, "shadow a.weight[0,:3]", shadow["a.weight"][0, :3].tolist())


def golden_known_answers():
    """Quirk known-answer vectors computed by the reference's own heads/utils.py."""
    import importlib
    hu = importlib.import_module("simvg.models.heads.utils")
    tab = hu.PositionEmbeddingSine1D(num_pos_feats=128, temperature=10000, normalize=True)(torch.zeros(1, 20, 256))
    assert torch.allclose(tab, O.pos_embed_sine_1d(20, 256))
    torch.manual_seed(3)
    m = hu.MLP(16, 16, 4, 3)
    x = torch.randn(5, 16)
    torch.save({"sine1d": tab, "mlp_sd": m.state_dict(), "mlp_x": x, "mlp_y": m(x).detach()}, os.path.join(OUT, "known_answers.pt"))
    print("known answers: sine1d row1[:6] =", tab[1, :6].tolist())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    load_reference()
    golden_known_answers()
    golden_encoder_small()
    golden_head(1, {"decoder": 1.0, "balanced_distill": {"token": 2.0, "distill": 1.0}}, "head_nq1_dwbd")
    golden_head(10, {"decoder": 1.0, "balanced_distill": {"token": 1.0, "distill": 0.4}}, "head_nq10_dwbd")
    golden_head(1, {"decoder": 1.0}, "head_nq1_decoder_only")
    golden_cfg1()
    golden_interpolate()
    golden_state_dict_keys()
    golden_grec_predictions()
    golden_ema()
    print("golden fixtures written to", OUT)
