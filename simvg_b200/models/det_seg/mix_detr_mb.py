"""MIXDETRMB — the detector object at the drop-in boundary (mirrors /root/reference/simvg/models/det_seg/mix_detr_mb.py:13-190).

forward(img, ref_expr_inds, img_metas, return_loss=True, text_attention_mask=, gt_bbox=, rescale=, with_bbox=, with_mask=)
-> (losses, [decoder_pred, token_pred]) / [decoder_pred, token_pred], each dict(pred_bboxes, pred_masks, predict_classes).
The image features reach the head as a channels-last *view* of the encoder output (same values and shape as the
reference's transpose+reshape at :52/:108, no copy), and the REC prediction path is batched (no per-sample host syncs).
"""
import torch
import torch.nn.functional as F

from simvg_b200.core.box_ops import box_cxcywh_to_xyxy, image_scale_tensor
from simvg_b200.models.builder import MODELS

from .one_stage import OneStageModel


@MODELS.register_module()
class MIXDETRMB(OneStageModel):
    def __init__(self, word_emb, num_token, vis_enc, lan_enc, head, fusion):
        super().__init__(word_emb, num_token, vis_enc, lan_enc, head, fusion)
        self.patch_size = vis_enc["patch_size"]

    def extract_visual_language(self, img, ref_expr_inds, text_attention_mask=None):
        return self.vis_enc(img, ref_expr_inds, text_attention_mask)

    def _features(self, img, ref_expr_inds, text_attention_mask):
        B = img.shape[0]
        H, W = self.image_hw(img)
        img_feat, text_feat, cls_feat = self.extract_visual_language(img, ref_expr_inds, text_attention_mask)
        h, w = H // self.patch_size, W // self.patch_size
        x_mm = img_feat.reshape(B, h, w, img_feat.shape[-1]).permute(0, 3, 1, 2)   # [B, D, h, w], channels-last strides
        return x_mm, text_feat, cls_feat

    def forward_train(self, img, ref_expr_inds, img_metas, text_attention_mask=None, gt_bbox=None, gt_mask_vertices=None,
                      rescale=False):
        x_mm, text_feat, cls_feat = self._features(img, ref_expr_inds, text_attention_mask)
        losses_dict, output = self.head.forward_train(x_mm, img_metas, cls_feat=cls_feat, gt_bbox=gt_bbox,
                                                      text_feat=text_feat, text_mask=text_attention_mask)
        with torch.no_grad():
            predictions = self._predict_both(output, img_metas, rescale)
        return losses_dict, predictions

    @torch.no_grad()
    def forward_test(self, img, ref_expr_inds, img_metas, text_attention_mask=None, with_bbox=False, with_mask=False,
                     rescale=False):
        x_mm, text_feat, cls_feat = self._features(img, ref_expr_inds, text_attention_mask)
        output = self.head.forward_test(x_mm, img_metas, text_feat=text_feat, cls_feat=cls_feat, with_bbox=with_bbox,
                                        with_mask=with_mask, text_mask=text_attention_mask)
        return self._predict_both(output, img_metas, rescale)

    def _predict_both(self, output, img_metas, rescale):
        grec = img_metas[0].get("target", None) is not None
        fn = self.get_predictions_grec if grec else self.get_predictions
        tok = fn(output["token_branch_output"], img_metas, rescale=rescale)
        dec = fn(output["decoder_branch_output"], img_metas, rescale=rescale)
        return [dec, tok]

    def get_predictions(self, output, img_metas, rescale=False):
        """Best-scoring box per image in pixels (mix_detr_mb.py:127-159 + tgqs_kd_detr_head.py:577-604 + detectron2
        detector_postprocess).  Batched: scale, clip, drop empty boxes, argmax — all on the device."""
        box_cls, box_pred = output["pred_logits"], output["pred_boxes"]
        if box_cls is None:
            return dict(pred_bboxes=None, pred_masks=None, predict_classes=None)
        B, nq = box_cls.shape[:2]
        scores, labels = F.softmax(box_cls, dim=-1)[:, :, :-1].max(-1)
        whwh = image_scale_tensor(img_metas, box_pred.device, box_pred.dtype, repeat=2).unsqueeze(1)   # [B, 1, (w, h, w, h)]
        boxes = box_cxcywh_to_xyxy(box_pred) * whwh
        boxes = torch.min(boxes.clamp(min=0), whwh)
        keep = ((boxes[..., 2] - boxes[..., 0]) > 0) & ((boxes[..., 3] - boxes[..., 1]) > 0)
        if nq == 1:
            # A sigmoid cxcywh box clipped to the image is never empty in exact arithmetic (w, h > 0, centre inside), so
            # the reference's per-sample filter + argmax (and its B host syncs) reduce to "take the only box".
            best = boxes[:, 0]
            classes = labels.reshape(-1)
        else:
            masked = scores.masked_fill(~keep, float("-inf"))
            idx = masked.argmax(dim=1)
            best = boxes[torch.arange(B, device=boxes.device), idx]
            classes = torch.cat([l[k] for l, k in zip(labels, keep)], dim=0)
        if rescale:
            sf = torch.tensor([m["scale_factor"] for m in img_metas], dtype=best.dtype).to(best.device)
            best = best / sf
        return dict(pred_bboxes=best, pred_masks=None, predict_classes=classes)

    def get_predictions_grec(self, output, img_metas, rescale=False):
        """Every non-empty box of every image as per-image dicts (mix_detr_mb.py:161-190 + tgqs_kd_detr_head.py:577-604 +
        detectron2 detector_postprocess).  Scores, scaling, clipping and the non-empty test are batched on the device; the
        per-image dicts are views of the batch tensors, and the only host synchronisation is ONE `keep.all()` for the batch
        (the reference: a filter + index per image) — when it holds, which it does unless a box degenerates, nothing is indexed."""
        box_cls, box_pred = output["pred_logits"], output["pred_boxes"]
        if box_cls is None:
            return dict(pred_bboxes=None, pred_masks=None, predict_classes=None)
        B = box_cls.shape[0]
        scores, labels = F.softmax(box_cls, dim=-1)[:, :, :-1].max(-1)
        whwh = image_scale_tensor(img_metas, box_pred.device, box_pred.dtype, repeat=2).unsqueeze(1)   # [B, 1, (w, h, w, h)]
        boxes = box_cxcywh_to_xyxy(box_pred) * whwh
        boxes = torch.min(boxes.clamp(min=0), whwh)
        keep = ((boxes[..., 2] - boxes[..., 0]) > 0) & ((boxes[..., 3] - boxes[..., 1]) > 0)
        if rescale:
            sf = torch.tensor([m["scale_factor"] for m in img_metas], dtype=boxes.dtype).to(boxes.device)
            boxes = boxes / sf.unsqueeze(1)
        all_kept = bool(keep.all())
        preds = []
        for b in range(B):
            if all_kept:
                preds.append({"boxes": boxes[b], "scores": scores[b], "labels": labels[b]})
            else:
                k = keep[b]
                preds.append({"boxes": boxes[b][k], "scores": scores[b][k], "labels": labels[b][k]})
        return dict(pred_bboxes=preds, pred_masks=None)
