"""BaseModel (mirrors /root/reference/simvg/models/det_seg/base.py:5-27; mmcv BaseModule/auto_fp16 are no-ops here since
use_fp16=False in every SimVG config)."""
from abc import ABCMeta

import torch.nn as nn


class BaseModel(nn.Module, metaclass=ABCMeta):
    def __init__(self):
        super().__init__()
        self.fp16_enabled = False

    def add_batch_input_shape(self, img, img_metas):
        shape = tuple(img.size()[-2:])
        for m in img_metas:
            m["batch_input_shape"] = shape

    def forward(self, img, ref_expr_inds, img_metas, return_loss=True, **kwargs):
        self.add_batch_input_shape(img, img_metas)
        if return_loss:
            return self.forward_train(img, ref_expr_inds, img_metas, **kwargs)
        return self.forward_test(img, ref_expr_inds, img_metas, **kwargs)
