"""Minimal stand-ins for detectron2's Boxes / Instances / detector_postprocess — the reference
imports them at det_seg/mix_detr_mb.py:7,10 and tgqs_kd_detr_head.py:11; only the members SimVG touches exist here."""
import torch


class Boxes:
    def __init__(self, tensor):
        self.tensor = tensor.reshape(-1, 4) if tensor.numel() else tensor.reshape(0, 4)

    def scale(self, scale_x, scale_y):
        self.tensor = self.tensor * self.tensor.new_tensor([scale_x, scale_y, scale_x, scale_y])

    def clip(self, box_size):
        h, w = box_size
        t = self.tensor
        self.tensor = torch.stack([t[:, 0].clamp(0, w), t[:, 1].clamp(0, h), t[:, 2].clamp(0, w), t[:, 3].clamp(0, h)], dim=-1)

    def nonempty(self, threshold=0.0):
        t = self.tensor
        return ((t[:, 2] - t[:, 0]) > threshold) & ((t[:, 3] - t[:, 1]) > threshold)

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self):
        return self.tensor.shape[0]


class Instances:
    def __init__(self, image_size, **fields):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            self._fields[k] = v

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        self._fields[name] = val

    def __getattr__(self, name):
        fields = object.__getattribute__(self, "_fields")
        if name in fields:
            return fields[name]
        raise AttributeError("Cannot find field '%s' in the given Instances!" % name)

    def get_fields(self):
        return self._fields

    def __getitem__(self, item):
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out._fields[k] = v[item]
        return out


def detector_postprocess(results, output_height, output_width):
    """Rescale to the output resolution, clip, drop empty boxes."""
    sx = output_width / results.image_size[1]
    sy = output_height / results.image_size[0]
    out = Instances((output_height, output_width), **results.get_fields())
    boxes = Boxes(out.pred_boxes.tensor.clone())
    boxes.scale(sx, sy)
    boxes.clip((output_height, output_width))
    out._fields["pred_boxes"] = boxes
    return out[boxes.nonempty()]
