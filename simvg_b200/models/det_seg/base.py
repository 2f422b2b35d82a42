"""BaseModel (mirrors /root/reference/simvg/models/det_seg/base.py:5-27; mmcv BaseModule/auto_fp16 are no-ops here since
use_fp16=False in every SimVG config)."""
from abc import ABCMeta

import torch.nn as nn


class BaseModel(nn.Module, metaclass=ABCMeta):
    def __init__(self):
        super().__init__()
        self.fp16_enabled = False

    @staticmethod
    def image_hw(img):
        """(H, W) of a batch given as float [B,3,H,W] (the reference's collated input) or as raw uint8 [B,H,W,3] (the GPU input
        path: Normalize + transpose fused into the patch embedding, see BEIT3.input_norm)."""
        import torch
        return tuple(img.shape[1:3]) if img.dtype == torch.uint8 else tuple(img.shape[-2:])

    def add_batch_input_shape(self, img, img_metas):
        shape = self.image_hw(img)
        for m in img_metas:
            m["batch_input_shape"] = shape

    def forward(self, img, ref_expr_inds, img_metas, return_loss=True, **kwargs):
        self.add_batch_input_shape(img, img_metas)
        if return_loss:
            return self.forward_train(img, ref_expr_inds, img_metas, **kwargs)
        return self.forward_test(img, ref_expr_inds, img_metas, **kwargs)
