"""``Linear`` -- an ``nn.Linear`` (same parameters / state dict) whose three products run on the hand-written tcgen05
GEMM (``csrc/gemm_tf32.cu``) and whose bias gradient is the streaming column-sum kernel (``csrc/colsum.cu``).

Used for the token-wise linears of the transformer (FFN, MSDA projections, enc_output: ``transformer.py:596-630,
765-791``, ``ms_deform_attn.py:52-55`` of the reference).  ``forward(x, relu=..., row_mask=...)`` folds the FFN's ReLU
and ``value.masked_fill(padding_mask[..., None], 0)`` (``ms_deform_attn.py:96-97``) into the GEMM epilogue.

The tcgen05 kernel computes in TF32, so it is taken exactly when torch's own switch for TF32 matrix products
(``torch.backends.cuda.matmul.allow_tf32``) is on -- with the switch off the products are full-fp32 library GEMMs,
which is what the parity tests against the fp32 CPU oracle use.  Small or oddly shaped products (a few hundred rows,
4-wide box heads) and -- until the 2-CTA variant of the kernel lands -- the 2048-wide FFN products also stay library
GEMMs: plumbing, like the convolutions.  ``SDB_LINEAR=cublas`` sends everything to the library and
``SDB_LINEAR=tcgen05_all`` everything eligible to the kernel (A/B comparisons only).
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib
from . import gemm

# products smaller than this many rows are latency-bound either way; keep them on the library path
MIN_ROWS = 1024


# widest layer the kernel takes by default: the projection family (value / offsets / weights / output, enc_output)
MAX_FEATURES = 512


def use_tcgen05():
    return os.environ.get("SDB_LINEAR", "tcgen05") != "cublas" and torch.backends.cuda.matmul.allow_tf32


def column_sum(x2d):
    """(rows, cols) fp32 CUDA -> (cols,)"""
    rows, cols = x2d.shape
    out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().sdb_colsum_f32(_lib.current_stream(x2d.device), x2d.data_ptr(), rows, cols, out.data_ptr())
    _lib.check(rc, "colsum")
    _lib.LAUNCHES["colsum"] += 1
    return out


class _LinearFn(torch.autograd.Function):
    """Library GEMMs + column-sum bias gradient (the small-shape path)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (g2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            gw = g2.t() @ x.reshape(-1, x.shape[-1])
        if ctx.needs_input_grad[2]:
            gb = column_sum(g2)
        return gx, gw, gb


class _TensorCoreLinearFn(torch.autograd.Function):
    """y = [mask rows](relu)(x W^T + b) with all three products on sdb_gemm_tf32."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, row_mask):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        mask = None if row_mask is None else row_mask.reshape(-1).contiguous()
        y = gemm.linear_forward(x2, weight, bias, relu=relu, row_mask=mask)
        ctx.save_for_backward(x2, weight, y if relu else None, mask)
        ctx.relu = relu
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        x2, weight, y, mask = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        if ctx.relu:
            g2 = torch.ops.aten.threshold_backward(g2, y, 0.0)   # also zeroes masked rows (their y is 0)
        elif mask is not None:
            g2 = g2.masked_fill(mask.view(torch.bool)[:, None], 0.0)
        elif not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm.linear_grad_input(g2, weight).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            gw = gemm.linear_grad_weight(g2, x2)
        if ctx.needs_input_grad[2]:
            gb = column_sum(g2)
        return gx, gw, gb, None, None


class Linear(nn.Linear):
    def _tensor_core_ok(self, x):
        return (x.is_cuda and x.dtype == torch.float32 and self.bias is not None and self.weight.dtype == torch.float32
                and self.in_features % 4 == 0 and self.out_features % 4 == 0 and self.in_features >= 64
                and self.out_features >= 64 and x.numel() // self.in_features >= MIN_ROWS and use_tcgen05()
                and (max(self.in_features, self.out_features) <= MAX_FEATURES
                     or os.environ.get("SDB_LINEAR") == "tcgen05_all"))

    def forward(self, x, relu=False, row_mask=None):
        """relu / row_mask (bool, one entry per row of x): applied to the output, inside the GEMM epilogue when the
        product runs on the tcgen05 kernel."""
        if self._tensor_core_ok(x):
            return _TensorCoreLinearFn.apply(x, self.weight, self.bias, relu, row_mask)
        if (x.is_cuda and x.dtype == torch.float32 and self.bias is not None and self.out_features % 4 == 0
                and x.numel() >= (1 << 16) and torch.is_grad_enabled()):
            y = _LinearFn.apply(x, self.weight, self.bias)
        else:
            y = F.linear(x, self.weight, self.bias)
        if relu:
            y = F.relu(y)
        if row_mask is not None:
            y = y.masked_fill(row_mask[..., None], 0.0)
        return y
