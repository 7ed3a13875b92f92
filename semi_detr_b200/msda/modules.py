"""``MSDeformAttn`` module -- host-side mirror of
detr_od/models/utils/ops/modules/ms_deform_attn.py:30-126 (same constructor, parameter names, init and
forward signature, so reference checkpoints load and ``DINOTransformer`` layers can own it unchanged).

Differences, none of them numerical:
 * the sampling kernel behind ``MSDeformAttnFunction.apply`` is the sm_100a one (no im2col chunking);
 * the reference's shape assert ``(shapes[:,0]*shapes[:,1]).sum() == Len_in`` (ms_deform_attn.py:92) forces a
   device->host sync on every call; here it is checked once per distinct shapes tensor and cached;
 * bf16 inputs are up-cast around the op like the reference does for fp16 (ms_deform_attn.py:114-120).
"""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from . import MultiScaleDeformableAttention as MSDA
from ..layers import Linear
from .functions import MSDAPrologueFunction, MSDeformAttnFunction, MSDeformAttnFusedFunction


def _is_power_of_2(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return n != 0 and (n & (n - 1)) == 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("MSDeformAttn: a power-of-2 head dimension (32 in every shipped config) takes the "
                          "tuned sm_100a kernel; other sizes use the generic one.")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points

        self.sampling_offsets = Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = Linear(d_model, d_model)
        self.output_proj = Linear(d_model, d_model)
        self._checked_shapes = None
        # fold softmax + location arithmetic into the sampling kernels when the shape allows (same numerics)
        self.fused_prologue = True
        self._reset_parameters()

    def _reset_parameters(self):
        # ms_deform_attn.py:62-76: zero offset weights, bias = per-head unit direction scaled by point index
        nn.init.constant_(self.sampling_offsets.weight, 0.0)
        theta = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        direction = torch.stack([theta.cos(), theta.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        grid = direction.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        grid = grid * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.reshape(-1))
        nn.init.constant_(self.attention_weights.weight, 0.0)
        nn.init.constant_(self.attention_weights.bias, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.0)

    def _check_len(self, spatial_shapes, len_in):
        key = (spatial_shapes.data_ptr(), spatial_shapes._version, len_in)
        if self._checked_shapes != key:
            host = tuple(tuple(int(v) for v in hw) for hw in spatial_shapes.tolist())   # one read per geometry
            total = sum(h * w for h, w in host)
            assert total == len_in, f"sum(H*W)={total} does not match input length {len_in}"
            self._checked_shapes = key
            self._shapes_host = host
            self._checked_shapes_ref = spatial_shapes      # keeps the address from being reused while cached

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """query (N, Lq, C); reference_points (N, Lq, L, 2|4); input_flatten (N, sum HW, C);
        input_spatial_shapes (L, 2) [(H, W)]; input_level_start_index (L,); input_padding_mask (N, sum HW) bool
        -> (N, Lq, C)"""
        N, Lq, _ = query.shape
        _, S, _ = input_flatten.shape
        M, L, P = self.n_heads, self.n_levels, self.n_points
        self._check_len(input_spatial_shapes, S)

        # value_proj + masked_fill(padding) (ms_deform_attn.py:94-97): the mask is applied in the GEMM epilogue
        value = self.value_proj(input_flatten, row_mask=input_padding_mask)
        value = value.view(N, S, M, self.d_model // M)
        offsets = self.sampling_offsets(query).view(N, Lq, M, L, P, 2)
        logits = self.attention_weights(query).view(N, Lq, M, L * P)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        if (self.fused_prologue and not reference_points.requires_grad
                and MSDA.fused_supported(value, reference_points, L, P)):
            out = MSDeformAttnFusedFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                                  reference_points.contiguous(), offsets, logits)
            return self.output_proj(out)
        if (self.fused_prologue and value.is_cuda and not reference_points.requires_grad
                and MSDA.prologue_supported(offsets, reference_points, L, P)):
            # levels x points > 16 (the 5-level model): softmax + location arithmetic as one kernel each way, fp32
            # locations / weights for the MSDA kernels (bf16 `value` keeps its storage)
            locations, weights = MSDAPrologueFunction.apply(offsets, logits, reference_points, self._shapes_host)
            out = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, locations, weights,
                                             self.im2col_step)
            return self.output_proj(out)
        weights = F.softmax(logits, -1).view(N, Lq, M, L, P)
        ref = reference_points[:, :, None, :, None, :]
        if reference_points.shape[-1] == 2:
            wh = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            locations = ref + offsets / wh[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            locations = ref[..., :2] + offsets / P * ref[..., 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))

        if value.dtype == torch.bfloat16 and M == 8 and self.d_model // M == 32 and P == 4:
            # bf16 storage for value / output, fp32 sampling arithmetic (BASELINE.json configs[3]): the bf16 kernels
            out = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                             locations.float().contiguous(), weights.float().contiguous(),
                                             self.im2col_step)
        elif value.dtype in (torch.float16, torch.bfloat16):
            # the reference's own handling of half inputs: compute in fp32 (ms_deform_attn.py:114-121)
            out = MSDeformAttnFunction.apply(value.float(), input_spatial_shapes, input_level_start_index,
                                             locations.float(), weights.float(), self.im2col_step)
            out = out.to(value.dtype)
        else:
            out = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                             locations.contiguous(), weights, self.im2col_step)
        return self.output_proj(out)
