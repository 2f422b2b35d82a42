"""Head utilities (mirrors /root/reference/simvg/models/heads/utils.py): MLP and the 1-D sine position table."""
import torch
import torch.nn.functional as F
from torch import nn


class MLP(nn.Module):
    """Linear -> ReLU -> ... -> Linear; optional stack of every layer's output (heads/utils.py:7-46)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, return_intermediate=False):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.return_intermediate = return_intermediate

    def forward(self, x):
        outs = []
        native = None
        if x.is_cuda and x.dtype == torch.float32:
            from simvg_b200.models.heads.tgqs_kd_detr_head import native   # fused Linear(+ReLU) kernels (csrc/headops.cu)
        for i, layer in enumerate(self.layers):
            act = i < self.num_layers - 1
            if native is not None:
                x = native.linear(layer, x, relu=act)
            else:
                x = layer(x)
                if act:
                    x = F.relu(x)
            outs.append(x)
        return torch.stack(outs, dim=0) if self.return_intermediate else x


class PositionEmbeddingSine1D(nn.Module):
    """[pos_len, dim] table, reproducing the reference's integer-cast frequencies (heads/utils.py:85-89; SURVEY Appendix
    C.2): only the first channel pair varies with position -> rows are [sin p, cos p, 0, 1, 0, 1, ...].
    The table is input independent, so it is built once per (pos_len, dim, device) instead of on the CPU every forward."""

    def __init__(self, num_pos_feats=64, temperature=10000, scale=6.283185307179586, eps=1e-6, offset=0.0, normalize=False):
        super().__init__()
        self.num_pos_feats, self.temperature, self.normalize = num_pos_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset
        self._cache = {}

    def forward(self, text):
        pos_len, dim = text.shape[1:]
        assert dim % 2 == 0, "wrong dimension!"
        key = (pos_len, dim, text.device)
        if key not in self._cache:
            freq = torch.arange(dim // 2, dtype=torch.float) / (dim / 2)
            freq = (1 / torch.pow(10000, freq)).to(torch.long)
            ang = torch.arange(pos_len).to(torch.long)[:, None] @ freq[None, :]
            tab = torch.zeros(pos_len, dim, dtype=torch.float)
            tab[:, 0::2] = torch.sin(ang)
            tab[:, 1::2] = torch.cos(ang)
            self._cache[key] = tab.to(text.device)
        return self._cache[key]
