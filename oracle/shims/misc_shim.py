"""mmcv / timm / detectron2 pieces the reference imports at its boundary, restated.
TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn as nn


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class Registry:
    def __init__(self, name):
        self._name, self._module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        return _reg(module) if module is not None else _reg

    def get(self, key):
        return self._module_dict.get(key)

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args is not None:
            for k, v in default_args.items():
                args.setdefault(k, v)
        cls = self.get(args.pop("type"))
        return cls(**args)


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


def auto_fp16(apply_to=None, out_fp32=False):
    def deco(fn):
        return fn
    return deco


def freeze_params(model):
    for p in model.parameters():
        p.requires_grad = False


def load_state_dict(model, state_dict, prefix="", ignore_missing="relative_position_index"):
    return model.load_state_dict(state_dict, strict=False)


class Boxes:
    def __init__(self, tensor):
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        self.tensor = tensor

    def scale(self, scale_x, scale_y):
        self.tensor[:, 0::2] *= scale_x
        self.tensor[:, 1::2] *= scale_y

    def clip(self, box_size):
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold=0.0):
        box = self.tensor
        return ((box[:, 2] - box[:, 0]) > threshold) & ((box[:, 3] - box[:, 1]) > threshold)

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self):
        return self.tensor.shape[0]


class Instances:
    def __init__(self, image_size, **kwargs):
        self._image_size = image_size
        self._fields = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name, value):
        self._fields[name] = value

    def get_fields(self):
        return self._fields

    def __getitem__(self, item):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret


def detector_postprocess(results, output_height, output_width, mask_threshold=0.5):
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    results = Instances((output_height, output_width), **results.get_fields())
    output_boxes = results.pred_boxes
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    return results[output_boxes.nonempty()]
