"""Hungarian-matched set criterion (mirrors /root/reference/simvg/core/criterion/criterion.py:69-271 and detrex's
HungarianMatcher).

Same interface as the reference (`SetCriterion(num_classes, matcher, weight_dict, eos_coef, loss_class_type)(outputs,
targets)` with per-sample target dicts), but the REC case SimVG trains on — one query, one ground-truth box per sample —
is matched without leaving the device: the assignment is the identity, so the reference's 6 `C.cpu()` + scipy round
trips per step disappear.  Any other case (nq > 1, gREC multi-target) is solved by a device Hungarian kernel
(simvgb_hungarian, csrc/matcher.cu; scipy is the test oracle and the path for CPU tensors).
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from simvg_b200.core.box_ops import box_cxcywh_to_xyxy, generalized_box_iou
from simvg_b200.utils.distributed import get_world_size, is_dist_avail_and_initialized


class BatchedTargets(list):
    """Per-sample target dicts (the reference's format) that also carry their batched tensors: `boxes` [B,4], `labels` [B]
    and optionally `weight` [B] — present when every sample has exactly one target (REC).  Lets the criterion skip all
    per-sample indexing (hundreds of tiny kernel launches per step in the reference, criterion.py:114-121,170-172)."""

    def __init__(self, items, boxes, labels, weight=None):
        super().__init__(items)
        self.boxes, self.labels, self.weight = boxes, labels, weight


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class=1.0, cost_bbox=1.0, cost_giou=1.0, cost_class_type="ce_cost", alpha=0.25, gamma=2.0):
        super().__init__()
        assert cost_class_type == "ce_cost", "SimVG builds the matcher with ce_cost (tgqs_kd_detr_head.py:132-137)"
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou

    @torch.no_grad()
    def forward(self, outputs, targets):
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        B, nq = logits.shape[:2]
        sizes = [len(t["boxes"]) for t in targets]
        if nq == 1 and all(s == 1 for s in sizes):  # identity assignment, no host sync
            z = torch.zeros(1, dtype=torch.int64, device=logits.device)
            return [(z, z) for _ in range(B)]
        prob = logits.flatten(0, 1).softmax(-1)
        out_bbox = boxes.flatten(0, 1)
        tgt_ids = torch.cat([t["labels"] for t in targets])
        tgt_bbox = torch.cat([t["boxes"] for t in targets])
        C = self.cost_bbox * torch.cdist(out_bbox, tgt_bbox, p=1) + self.cost_class * (-prob[:, tgt_ids]) + \
            self.cost_giou * (-generalized_box_iou(box_cxcywh_to_xyxy(out_bbox), box_cxcywh_to_xyxy(tgt_bbox)))
        C = C.view(B, nq, -1)
        if C.is_cuda and nq <= 32 and max(sizes, default=0) <= 32:
            # the assignment is solved on the device (one thread per sample, simvgb_hungarian): no C.cpu() round trip
            from simvg_b200 import kernels as K
            return K.hungarian(C, sizes)
        from scipy.optimize import linear_sum_assignment     # CPU tensors (tests) / oversized problems: the reference's path
        C = C.cpu()
        idx = [linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))]
        dev = logits.device
        return [(torch.as_tensor(i, dtype=torch.int64, device=dev), torch.as_tensor(j, dtype=torch.int64, device=dev))
                for i, j in idx]


def sigmoid_focal_loss(inputs, targets, num_boxes, alpha=0.25, gamma=2.0):
    prob = inputs.sigmoid()
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.mean(1).sum() / num_boxes


class SetCriterion(nn.Module):
    def __init__(self, num_classes, matcher, weight_dict, losses: List[str] = ("class", "boxes"), eos_coef=0.1,
                 loss_class_type="focal_loss", alpha=0.25, gamma=2.0):
        super().__init__()
        assert loss_class_type in ["ce_loss", "focal_loss", "weighted_ce_loss"], \
            "only support ce loss and focal loss for computing classification loss"
        self.num_classes, self.matcher, self.weight_dict, self.losses = num_classes, matcher, weight_dict, list(losses)
        self.alpha, self.gamma, self.eos_coef, self.loss_class_type = alpha, gamma, eos_coef, loss_class_type
        if loss_class_type in ["ce_loss", "weighted_ce_loss"]:
            w = torch.ones(num_classes + 1)
            w[-1] = eos_coef
            self.register_buffer("empty_weight", w)

    @staticmethod
    def _src_idx(indices):
        return (torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)]), torch.cat([s for s, _ in indices]))

    def loss_labels(self, outputs, targets, indices, num_boxes):
        logits = outputs["pred_logits"]
        idx = self._src_idx(indices)
        tcls_o = torch.cat([t["labels"][j] for t, (_, j) in zip(targets, indices)])
        tcls = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        tcls[idx] = tcls_o
        if self.loss_class_type == "ce_loss":
            loss = F.cross_entropy(logits.transpose(1, 2), tcls, self.empty_weight)
        elif self.loss_class_type == "weighted_ce_loss":   # criterion.py:128-137
            wq = torch.full(logits.shape[:2], 0.1, device=logits.device)
            wq[idx] = 1.0
            loss = F.cross_entropy(logits.transpose(1, 2), tcls, self.empty_weight, reduction="none")
            loss = (wq * loss).mean(-1).sum()
        else:
            onehot = torch.zeros(logits.shape[0], logits.shape[1], logits.shape[2] + 1, dtype=logits.dtype, device=logits.device)
            onehot.scatter_(2, tcls.unsqueeze(-1), 1)
            loss = sigmoid_focal_loss(logits, onehot[:, :, :-1], num_boxes, self.alpha, self.gamma) * logits.shape[1]
        return {"loss_class": loss}

    def loss_boxes(self, outputs, targets, indices, num_boxes):
        idx = self._src_idx(indices)
        src = outputs["pred_boxes"][idx]
        tgt = torch.cat([t["boxes"][j] for t, (_, j) in zip(targets, indices)], dim=0)
        l1 = F.l1_loss(src, tgt, reduction="none")
        giou = 1 - torch.diag(generalized_box_iou(box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tgt)))
        if self.loss_class_type == "weighted_ce_loss":     # criterion.py:175-182,193-200
            tw = torch.cat([t["weight"][j] if len(t["weight"][j]) else torch.zeros(1, device=src.device)
                            for t, (_, j) in zip(targets, indices)]).squeeze()
            l1 = l1.sum(-1) * tw
            giou = giou * tw
        return {"loss_bbox": l1.sum() / num_boxes, "loss_giou": giou.sum() / num_boxes}

    def get_loss(self, loss, outputs, targets, indices, num_boxes):
        return {"class": self.loss_labels, "boxes": self.loss_boxes}[loss](outputs, targets, indices, num_boxes)

    def _forward_rec(self, outputs, targets):
        """One query, one target per sample: the assignment is the identity and every loss is a batched expression."""
        from simvg_b200.core.box_ops import aligned_iou_giou
        num_boxes = float(len(targets))
        if is_dist_avail_and_initialized() and get_world_size() > 1:
            # criterion.py:247-250 (all_reduce, / world, clamp(min=1)) kept on the device: no .item() sync.  The value only
            # depends on the ranks' batch sizes, so while a CUDA graph is being captured (simvg_b200.runtime: the step graphs
            # contain NO collective) the result of the last eager evaluation with this local batch size is reused — the
            # runtime's warm-up pass has run exactly that evaluation.
            dev = outputs["pred_logits"].device
            key = (int(num_boxes), str(dev))
            cache = self.__dict__.setdefault("_num_boxes_cache", {})
            if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
                if key not in cache:
                    raise RuntimeError("SetCriterion: capturing a step graph before any eager step ran with this batch size")
                num_boxes = cache[key]
            else:
                nb = torch.full((1,), num_boxes, dtype=torch.float, device=dev)
                torch.distributed.all_reduce(nb)
                nb = torch.clamp(nb / get_world_size(), min=1.0)
                num_boxes = nb[0]               # this (eager) evaluation uses its own tensor: autograd saves it
                with torch.no_grad():
                    if key in cache:
                        cache[key].copy_(nb[0])     # in place: a captured step graph may hold the cached tensor's address
                    else:
                        cache[key] = nb[0].clone()
        else:
            num_boxes = max(num_boxes, 1.0)
        # Main output and every auxiliary decoder layer in ONE batched expression ([Ld, B, ...]): with a 6-layer decoder the
        # per-layer evaluation was 18 x ~75 tiny forward/backward launches per step; the per-layer loss entries of the
        # returned dict are 0-dim views of the [Ld] results.
        aux = outputs.get("aux_outputs", [])
        layers_logits = [a_["pred_logits"] for a_ in aux] + [outputs["pred_logits"]]
        layers_boxes = [a_["pred_boxes"] for a_ in aux] + [outputs["pred_boxes"]]
        logits = torch.stack(layers_logits).squeeze(2)        # [Ld, B, C]   (one query: squeeze is a view, no SelectBackward fill)
        src = torch.stack(layers_boxes).squeeze(2)            # [Ld, B, 4]
        Ld, Bn, C = logits.shape
        tcls = targets.labels.view(-1)
        tbox = targets.boxes
        ce = F.cross_entropy(logits.reshape(Ld * Bn, C), tcls.repeat(Ld), self.empty_weight, reduction="none").view(Ld, Bn)
        if self.loss_class_type == "ce_loss":                 # weighted mean over the batch (criterion.py:127)
            lc = ce.sum(1) / self.empty_weight[tcls].sum()
        else:  # weighted_ce_loss with every query matched: per-query weight 1, mean over queries, sum over batch (:128-137)
            lc = ce.sum(1)
        l1 = (src - tbox).abs().sum(-1)                       # [Ld, B]
        giou = 1 - aligned_iou_giou(box_cxcywh_to_xyxy(src), box_cxcywh_to_xyxy(tbox).expand(Ld, Bn, 4))[1]
        if self.loss_class_type == "weighted_ce_loss":
            l1 = l1 * targets.weight
            giou = giou * targets.weight
        lb = l1.sum(1) / num_boxes
        lg = giou.sum(1) / num_boxes
        lc_l, lb_l, lg_l = lc.unbind(0), lb.unbind(0), lg.unbind(0)
        losses = {"loss_class": lc_l[-1], "loss_bbox": lb_l[-1], "loss_giou": lg_l[-1]}
        for i in range(len(aux)):
            losses.update({"loss_class_%d" % i: lc_l[i], "loss_bbox_%d" % i: lb_l[i], "loss_giou_%d" % i: lg_l[i]})
        return losses

    def forward(self, outputs, targets, return_indices=False):
        if (isinstance(targets, BatchedTargets) and not return_indices and outputs["pred_logits"].shape[1] == 1
                and self.loss_class_type != "focal_loss" and list(self.losses) == ["class", "boxes"]):
            return self._forward_rec(outputs, targets)
        main = {k: v for k, v in outputs.items() if k != "aux_outputs"}
        indices = self.matcher(main, targets)
        num_boxes = float(sum(len(t["labels"]) for t in targets))   # host-side count: no .item() sync
        if is_dist_avail_and_initialized():
            nb = torch.as_tensor([num_boxes], dtype=torch.float, device=outputs["pred_logits"].device)
            torch.distributed.all_reduce(nb)
            num_boxes = nb.item()
        num_boxes = max(num_boxes / get_world_size(), 1.0)
        losses = {}
        for name in self.losses:
            losses.update(self.get_loss(name, outputs, targets, indices, num_boxes))
        all_idx = [indices]
        for i, aux in enumerate(outputs.get("aux_outputs", [])):
            ind = self.matcher(aux, targets)
            all_idx.insert(-1, ind)
            for name in self.losses:
                losses.update({k + "_%d" % i: v for k, v in self.get_loss(name, aux, targets, ind, num_boxes).items()})
        if return_indices:
            return losses, all_idx
        return losses
