"""Box utilities (detrex.layers.box_ops semantics; identical math to the copy at
/root/reference/simvg/apis/test.py:30-67)."""
import torch


def box_cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def box_xyxy_to_cxcywh(b):
    x0, y0, x1, y1 = b.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], dim=-1)


def _area(b):
    return (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])


def box_iou(b1, b2):
    """Pairwise [N,M] IoU and union."""
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = _area(b1)[:, None] + _area(b2) - inter
    return inter / union, union


def generalized_box_iou(b1, b2):
    iou, union = box_iou(b1, b2)
    lt = torch.min(b1[:, None, :2], b2[:, :2])
    rb = torch.max(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[..., 0] * wh[..., 1]
    return iou - (area - union) / area


def aligned_iou_giou(b1, b2):
    """Element-wise (aligned) IoU and GIoU of two [..., 4] xyxy box tensors — the diagonal of the pairwise versions,
    without materialising the [N, N] matrix."""
    lt = torch.max(b1[..., :2], b2[..., :2])
    rb = torch.min(b1[..., 2:], b2[..., 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = _area(b1) + _area(b2) - inter
    iou = inter / union
    lt2 = torch.min(b1[..., :2], b2[..., :2])
    rb2 = torch.max(b1[..., 2:], b2[..., 2:])
    wh2 = (rb2 - lt2).clamp(min=0)
    area = wh2[..., 0] * wh2[..., 1]
    return iou, iou - (area - union) / area


_SCALE_CACHE = {}


def image_scale_tensor(img_metas, device, dtype=torch.float, repeat=2, key="img_shape"):
    """[B, 2*repeat] tensor of (w, h[, w, h]) per image, cached per (sizes, device): building it from Python lists costs a
    pageable host->device copy per call (a sync point, and illegal while a CUDA graph is being captured)."""
    sizes = tuple((int(m[key][1]), int(m[key][0])) for m in img_metas)
    k = (sizes, str(device), dtype, repeat)
    t = _SCALE_CACHE.get(k)
    if t is None:
        if len(_SCALE_CACHE) > 64:
            _SCALE_CACHE.clear()
        t = torch.tensor([list(wh) * repeat for wh in sizes], dtype=dtype).to(device)
        _SCALE_CACHE[k] = t
    return t
