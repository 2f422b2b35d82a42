"""Stand-ins for the third-party packages SimVG's hot-path files import but which are not installed here
(torchscale, fairscale, timm, mmcv, mmdet, pycocotools, detectron2, detrex) — TEST INFRASTRUCTURE ONLY.

`install()` registers them in sys.modules and mounts the reference tree (/root/reference/simvg) as *stub packages*
(package objects whose __path__ points at the real directories but whose __init__.py is not executed), so the reference's
own files — beit3_base.py, beit3.py, modeling_utils.py, tgqs_kd_detr_head.py, transformer.py, heads/utils.py,
criterion.py, det_seg/*.py, builder.py — are imported and executed VERBATIM, while only leaf ops whose source is not in
/root/reference are restated.  Used by oracle/make_golden.py in the build container to pin the oracle;
never by the product, the GPU tests, smoke() or bench.py (the reference tree does not exist on the GPU box).
"""
import importlib
import os
import sys
import types

REF = os.environ.get("SIMVG_REFERENCE", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _pkg(name, path=None):
    m = _mod(name)
    m.__path__ = [path] if path else []
    return m


def install():
    if "simvg" in sys.modules and getattr(sys.modules["simvg"], "_simvgb_shim", False):
        return
    from . import detrex_shim, misc_shim, torchscale_shim

    # ---- torchscale
    _pkg("torchscale")
    _pkg("torchscale.architecture")
    _mod("torchscale.architecture.config", EncoderConfig=torchscale_shim.EncoderConfig)
    _mod("torchscale.architecture.utils", init_bert_params=torchscale_shim.init_bert_params)
    _pkg("torchscale.component")
    _mod("torchscale.component.embedding", PositionalEmbedding=torchscale_shim.PositionalEmbedding,
         TextEmbedding=torchscale_shim.TextEmbedding, VisionEmbedding=torchscale_shim.VisionEmbedding)
    _mod("torchscale.component.multiway_network", MutliwayEmbedding=torchscale_shim.MutliwayEmbedding,
         MultiwayWrapper=torchscale_shim.MultiwayWrapper, MultiwayNetwork=torchscale_shim.MultiwayNetwork,
         set_split_position=torchscale_shim.set_split_position)
    _mod("torchscale.component.droppath", DropPath=torchscale_shim.DropPath)
    _mod("torchscale.component.feedforward_network", FeedForwardNetwork=torchscale_shim.FeedForwardNetwork,
         make_experts=torchscale_shim.make_experts)
    _mod("torchscale.component.multihead_attention", MultiheadAttention=torchscale_shim.MultiheadAttention)
    _mod("torchscale.component.relative_position_bias", RelativePositionBias=torchscale_shim.Unused)
    _pkg("torchscale.component.xmoe")
    _mod("torchscale.component.xmoe.moe_layer", MOELayer=torchscale_shim.Unused)
    _mod("torchscale.component.xmoe.routing", Top1Gate=torchscale_shim.Unused, Top2Gate=torchscale_shim.Unused)
    # ---- fairscale / timm
    _pkg("fairscale")
    _mod("fairscale.nn", checkpoint_wrapper=lambda m, *a, **k: m, wrap=lambda m, *a, **k: m)
    _pkg("timm")
    _pkg("timm.models")
    _mod("timm.models.layers", trunc_normal_=misc_shim.trunc_normal_)
    # ---- mmcv / mmdet / pycocotools
    _pkg("mmcv")
    _mod("mmcv.utils", Registry=misc_shim.Registry)
    _mod("mmcv.runner", BaseModule=misc_shim.BaseModule, auto_fp16=misc_shim.auto_fp16, get_dist_info=lambda: (0, 1))
    _pkg("mmdet")
    _mod("mmdet.core", BitmapMasks=object)
    _pkg("pycocotools")
    _mod("pycocotools.mask")
    # ---- detectron2
    _pkg("detectron2")
    _mod("detectron2.structures", Boxes=misc_shim.Boxes, Instances=misc_shim.Instances, ImageList=object)
    _mod("detectron2.modeling", detector_postprocess=misc_shim.detector_postprocess)
    # ---- detrex
    _pkg("detrex")
    box = _mod("detrex.layers.box_ops", box_cxcywh_to_xyxy=detrex_shim.box_cxcywh_to_xyxy,
               box_xyxy_to_cxcywh=detrex_shim.box_xyxy_to_cxcywh, box_iou=detrex_shim.box_iou,
               generalized_box_iou=detrex_shim.generalized_box_iou)
    lay = _pkg("detrex.layers")
    lay.__dict__.update(FFN=detrex_shim.FFN, BaseTransformerLayer=detrex_shim.BaseTransformerLayer,
                        MultiheadAttention=detrex_shim.MultiheadAttention,
                        TransformerLayerSequence=detrex_shim.TransformerLayerSequence,
                        box_cxcywh_to_xyxy=detrex_shim.box_cxcywh_to_xyxy, box_xyxy_to_cxcywh=detrex_shim.box_xyxy_to_cxcywh,
                        generalized_box_iou=detrex_shim.generalized_box_iou, box_iou=detrex_shim.box_iou, box_ops=box)
    sys.modules["detrex.layers.box_ops"] = box
    _mod("detrex.layers.position_embedding", PositionEmbeddingSine=detrex_shim.PositionEmbeddingSine,
         PositionEmbeddingLearned=detrex_shim.PositionEmbeddingLearned)
    _pkg("detrex.modeling")
    _pkg("detrex.modeling.matcher")
    _mod("detrex.modeling.matcher.matcher", HungarianMatcher=detrex_shim.HungarianMatcher)
    _mod("detrex.utils", get_world_size=lambda: 1, is_dist_avail_and_initialized=lambda: False)

    # ---- the reference tree as stub packages (no __init__.py is executed)
    sv = os.path.join(REF, "simvg")
    root = _pkg("simvg", sv)
    root._simvgb_shim = True
    models = _pkg("simvg.models", os.path.join(sv, "models"))
    builder = importlib.import_module("simvg.models.builder")          # reference file, verbatim
    for k in ("VIS_ENCODERS", "LAN_ENCODERS", "FUSIONS", "HEADS", "MODELS", "build_model", "build_vis_enc", "build_lan_enc",
              "build_fusion", "build_head"):
        setattr(models, k, getattr(builder, k))
    _pkg("simvg.models.vis_encs", os.path.join(sv, "models", "vis_encs"))
    _pkg("simvg.models.vis_encs.beit", os.path.join(sv, "models", "vis_encs", "beit"))
    _mod("simvg.models.vis_encs.beit.utils", load_state_dict=misc_shim.load_state_dict)  # 913-line upstream leftover
    _pkg("simvg.models.heads", os.path.join(sv, "models", "heads"))
    _pkg("simvg.models.heads.tgqs_kd_detr_head", os.path.join(sv, "models", "heads", "tgqs_kd_detr_head"))
    _pkg("simvg.models.det_seg", os.path.join(sv, "models", "det_seg"))
    _pkg("simvg.models.lan_encs")
    sys.modules["simvg.models.lan_encs"].LSTM = type("LSTM", (), {})
    _mod("simvg.models.utils", freeze_params=misc_shim.freeze_params)
    _pkg("simvg.core", os.path.join(sv, "core"))
    _pkg("simvg.core.criterion", os.path.join(sv, "core", "criterion"))
    _mod("simvg.core.criterion.distill_criterion", DistillCriterion=torchscale_shim.Unused)


def load_reference():
    """-> (BEIT3, TextGuidedQuerySelectKDDETRHead, MIXDETRMB, build_model) classes from the reference's own files."""
    install()
    beit3 = importlib.import_module("simvg.models.vis_encs.beit.beit3")
    head = importlib.import_module("simvg.models.heads.tgqs_kd_detr_head.tgqs_kd_detr_head")
    importlib.import_module("simvg.models.det_seg.base")
    importlib.import_module("simvg.models.det_seg.one_stage")
    det = importlib.import_module("simvg.models.det_seg.mix_detr_mb")
    return beit3.BEIT3, head.TextGuidedQuerySelectKDDETRHead, det.MIXDETRMB, sys.modules["simvg.models"].build_model
