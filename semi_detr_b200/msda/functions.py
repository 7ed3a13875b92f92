"""``MSDeformAttnFunction`` -- same surface as the reference's autograd Function
(detr_od/models/utils/ops/functions/ms_deform_attn_func.py:21-38):

    MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
                               sampling_locations, attention_weights, im2col_step) -> (N, Lq, M*D)

saves the same five tensors, is once-differentiable, and returns ``None`` gradients for the shapes,
the level start index and ``im2col_step``.  There is deliberately no pure-PyTorch core here: the
CPU restatement lives under ``oracle/`` and is test infrastructure only.
"""
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = MSDA.ms_deform_attn_backward(
            value, shapes, level_start, sampling_locations, attention_weights, grad_output.contiguous(),
            ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


class MSDeformAttnFusedFunction(Function):
    """Same operator with the module's softmax and sampling-location arithmetic folded into the kernels
    (``sdb_msda_fused_forward/backward_f32``): takes the raw ``sampling_offsets`` / ``attention_weights`` linear
    outputs and the reference points, so neither ``sampling_locations`` nor the softmax weights are materialised.
    The reference points get no gradient (DINO detaches them, transformer.py:1030-1036)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points, sampling_offsets,
                attention_logits):
        output = MSDA.ms_deform_attn_fused_forward(value, value_spatial_shapes, value_level_start_index,
                                                   reference_points, sampling_offsets, attention_logits)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, reference_points,
                              sampling_offsets, attention_logits)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, ref, offsets, logits = ctx.saved_tensors
        if ctx.needs_input_grad[3]:
            raise RuntimeError("MSDeformAttnFusedFunction: reference_points must not require grad "
                               "(use MSDeformAttnFunction for that case)")
        grad_value, grad_off, grad_logits = MSDA.ms_deform_attn_fused_backward(
            value, shapes, level_start, ref, offsets, logits, grad_output.contiguous())
        return grad_value, None, None, None, grad_off, grad_logits


class MSDAPrologueFunction(Function):
    """softmax + sampling-location arithmetic of the module (ms_deform_attn.py:98-112) as one kernel each way
    (``sdb_msda_prologue_forward/backward_{f32,bf16}``), for shapes the fused MSDA kernels do not take (levels x points >
    16): raw ``sampling_offsets`` (N, Lq, M, L, P, 2) and ``attention_logits`` (N, Lq, M, L*P) in fp32 or bf16 ->
    fp32 ``sampling_locations`` and ``attention_weights`` (N, Lq, M, L, P).  ``shapes_host``: tuple of (H, W) per level.
    No gradient to the reference points (DINO detaches them)."""

    @staticmethod
    def forward(ctx, sampling_offsets, attention_logits, reference_points, shapes_host):
        loc, attn = MSDA.msda_prologue_forward(sampling_offsets, attention_logits, reference_points, shapes_host)
        ctx.save_for_backward(attn, reference_points)
        ctx.shapes_host, ctx.raw_dtype = shapes_host, sampling_offsets.dtype
        return loc, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loc, grad_attn):
        attn, ref = ctx.saved_tensors
        if ctx.needs_input_grad[2]:
            raise RuntimeError("MSDAPrologueFunction: reference_points must not require grad")
        g_off, g_logit = MSDA.msda_prologue_backward(grad_loc.contiguous(), grad_attn.contiguous(), attn, ref,
                                                     ctx.shapes_host, ctx.raw_dtype)
        return g_off, g_logit, None, None
