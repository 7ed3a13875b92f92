"""``SinePositionalEncodingHW`` -- mirror of detr_od/models/utils/positional_encoding.py:10-99
(cumsum-normalised sine/cosine embedding with one temperature per axis; DINO uses 20/20, normalize=True)."""
import math

import torch
from torch import nn

from ..registry import POSITIONAL_ENCODING


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncodingHW(nn.Module):
    def __init__(self, num_feats, temperatureH=10000, temperatureW=10000, normalize=False, scale=2 * math.pi,
                 eps=1e-6, offset=0.0, init_cfg=None):
        super().__init__()
        self.num_feats = num_feats
        self.temperatureH, self.temperatureW = temperatureH, temperatureW
        self.normalize, self.scale, self.eps, self.offset = normalize, scale, eps, offset

    def forward(self, mask):
        """mask (B, H, W), non-zero = padding -> (B, 2*num_feats, H, W)"""
        keep = 1 - mask.to(torch.int)
        y = keep.cumsum(1, dtype=torch.float32)
        x = keep.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y = (y + self.offset) / (y[:, -1:, :] + self.eps) * self.scale
            x = (x + self.offset) / (x[:, :, -1:] + self.eps) * self.scale
        idx = torch.arange(self.num_feats, dtype=torch.float32, device=mask.device)
        expo = 2 * (idx // 2) / self.num_feats
        px = x[..., None] / self.temperatureW ** expo
        py = y[..., None] / self.temperatureH ** expo
        B, H, W = mask.shape
        px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).view(B, H, W, -1)
        py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).view(B, H, W, -1)
        return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)
