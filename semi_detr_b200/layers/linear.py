"""``Linear`` -- an ``nn.Linear`` (same parameters / state dict) whose three products run on the hand-written tcgen05
GEMM (``csrc/gemm_tf32.cu``) and whose bias gradient is the streaming column-sum kernel (``csrc/colsum.cu``).

Used for the token-wise linears of the transformer (FFN, MSDA projections, enc_output: ``transformer.py:596-630,
765-791``, ``ms_deform_attn.py:52-55`` of the reference).  ``forward(x, relu=..., row_mask=...)`` folds the FFN's ReLU
and ``value.masked_fill(padding_mask[..., None], 0)`` (``ms_deform_attn.py:96-97``) into the GEMM epilogue.

The tcgen05 kernel computes in TF32, so it is taken exactly when torch's own switch for TF32 matrix products
(``torch.backends.cuda.matmul.allow_tf32``) is on -- with the switch off every product is a full-fp32 library GEMM,
which is what the parity tests against the fp32 CPU oracle use.

Which of a layer's three products run on the kernel is a measured choice (``profiles/gemm_r1_*.jsonl``, B200,
44 446 tokens), ``SDB_LINEAR`` selects the policy:

* ``auto`` (default) -- every product where the kernel is at least as fast as the library: the split-K weight
  gradients of the 256-wide projection family (1.7-1.9x faster than cuBLAS), every forward whose epilogue absorbs a
  separate elementwise pass (value_proj + padding mask, FFN linear1 + ReLU), and nothing else;
* ``tcgen05`` -- all three products of the projection family (value / offsets / weights / output projections,
  enc_output); ``tcgen05_all`` -- additionally the 2048-wide FFN products; ``cublas`` -- library only.

Small or oddly shaped products (a few hundred rows, 4-wide box heads) always stay library GEMMs: plumbing, like the
convolutions.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib
from . import gemm

# products smaller than this many rows are latency-bound either way; keep them on the library path
MIN_ROWS = 1024
# widest layer of the projection family
MAX_FEATURES = 512


def policy():
    p = os.environ.get("SDB_LINEAR", "auto")
    if p not in ("auto", "tcgen05", "tcgen05_all", "cublas"):
        raise RuntimeError(f"SDB_LINEAR={p!r}: expected auto, tcgen05, tcgen05_all or cublas")
    return p


def product_plan(in_features, out_features, fused_epilogue, has_mask=None):
    """-> (forward, grad_input, grad_weight) on the tcgen05 kernel?  None: the whole layer stays on the library.
    ``fused_epilogue``: the forward has a ReLU and / or a row mask to absorb; ``has_mask`` (default: same) says whether
    a row mask is among them -- a ReLU alone is also a library epilogue (cuBLASLt RELU_BIAS through
    ``torch._addmm_activation``: 88 us against our 170 us at 44 446 x 256 -> 2048, tools/time_linear1.py), so under
    ``auto`` only masked forwards take the tcgen05 kernel; the layer still goes through ``_TensorCoreLinearFn`` for its
    fused ReLU-backward + bias-gradient pass."""
    if has_mask is None:
        has_mask = fused_epilogue
    p = policy()
    if p == "cublas" or not torch.backends.cuda.matmul.allow_tf32:
        return None
    family = max(in_features, out_features) <= MAX_FEATURES
    if p == "tcgen05_all" or (p == "tcgen05" and family):
        return True, True, True
    if p == "tcgen05":
        return None
    plan = (bool(fused_epilogue and has_mask), False, family)          # auto
    return plan if (any(plan) or fused_epilogue) else None


def column_sum(x2d):
    """(rows, cols) fp32 CUDA -> (cols,)"""
    rows, cols = x2d.shape
    out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().sdb_colsum_f32(_lib.current_stream(x2d.device), x2d.data_ptr(), rows, cols, out.data_ptr())
    _lib.check(rc, "colsum")
    _lib.LAUNCHES["colsum"] += 1
    return out


def relu_backward_colsum(dy2d, y2d):
    """-> (dy * (y > 0), its column sums): ReLU backward and the bias gradient of the layer before it in one pass"""
    rows, cols = y2d.shape
    g = torch.empty_like(y2d)
    out = torch.empty(cols, dtype=torch.float32, device=y2d.device)
    with torch.cuda.device(y2d.device):
        rc = _lib.lib().sdb_relu_backward_colsum_f32(_lib.current_stream(y2d.device), dy2d.data_ptr(), y2d.data_ptr(),
                                                     rows, cols, g.data_ptr(), out.data_ptr())
    _lib.check(rc, "relu_backward_colsum")
    _lib.LAUNCHES["relu_backward_colsum"] += 1
    return g, out


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


def _colsum_bf16(x2d):
    rows, cols = x2d.shape
    out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().sdb_colsum_bf16(_lib.current_stream(x2d.device), x2d.data_ptr(), rows, cols, out.data_ptr())
    _lib.check(rc, "colsum_bf16")
    _lib.LAUNCHES["colsum"] += 1
    return out


def _relu_backward_colsum_bf16(dy2d, y2d):
    rows, cols = y2d.shape
    g = torch.empty_like(y2d)
    out = torch.empty(cols, dtype=torch.float32, device=y2d.device)
    with torch.cuda.device(y2d.device):
        rc = _lib.lib().sdb_relu_backward_colsum_bf16(_lib.current_stream(y2d.device), dy2d.data_ptr(), y2d.data_ptr(),
                                                      rows, cols, g.data_ptr(), out.data_ptr())
    _lib.check(rc, "relu_backward_colsum_bf16")
    _lib.LAUNCHES["relu_backward_colsum"] += 1
    return g, out


class _AutocastLinearFn(torch.autograd.Function):
    """The layer under ``torch.autocast(bfloat16)`` (BASELINE.json configs[3]): library bf16 GEMMs (ReLU in the library
    epilogue), and the bias gradient -- with the ReLU backward of FFN linear1 in the same pass -- on this library's bf16
    column-sum kernels with fp32 sums.  Plain autocast leaves those to a generic bf16 reduction plus a separate
    threshold pass: 9.9 of the 75 ms of the 5-scale step (profiles/step_profile_r2_sup5.txt)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        with torch.autocast("cuda", enabled=False):
            xb = x.reshape(-1, x.shape[-1]).to(torch.bfloat16)
            if not xb.is_contiguous():
                xb = xb.contiguous()
            wb, bb = weight.to(torch.bfloat16), bias.to(torch.bfloat16)
            if relu and hasattr(torch, "_addmm_activation"):
                y = torch._addmm_activation(bb, xb, wb.t(), use_gelu=False)
            else:
                y = torch.addmm(bb, xb, wb.t())
                if relu:
                    y = y.relu_()
        ctx.save_for_backward(xb, wb, y if relu else None)
        ctx.relu, ctx.x_shape, ctx.x_dtype = relu, x.shape, x.dtype
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        xb, wb, y = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            g2 = g.reshape(-1, g.shape[-1])
            if g2.dtype != torch.bfloat16:
                g2 = g2.to(torch.bfloat16)
            if not g2.is_contiguous():
                g2 = g2.contiguous()
            if ctx.relu:
                g2, gb = _relu_backward_colsum_bf16(g2, y)
            else:
                gb = _colsum_bf16(g2) if ctx.needs_input_grad[2] else None
            gx = (g2 @ wb).view(ctx.x_shape).to(ctx.x_dtype) if ctx.needs_input_grad[0] else None
            gw = (g2.t() @ xb).float() if ctx.needs_input_grad[1] else None
        return gx, gw, gb, None


def linear(x, weight, bias):
    """``F.linear`` for layers that are not ``Linear`` modules (the in / out projections of nn.MultiheadAttention): on the
    device the bias gradient comes from this library's column-sum kernel instead of a generic reduction."""
    if (x.is_cuda and x.dtype == torch.float32 and bias is not None and weight.shape[0] % 4 == 0
            and x.numel() >= (1 << 16) and torch.is_grad_enabled() and not torch.is_autocast_enabled()):
        return _LinearFn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class _TensorCoreLinearFn(torch.autograd.Function):
    """y = [mask rows](relu)(x W^T + b); ``plan`` says which of the three products run on sdb_gemm_tf32."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, row_mask, plan):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        mask = None if row_mask is None else row_mask.reshape(-1).contiguous()
        if plan[0]:
            y = gemm.linear_forward(x2, weight, bias, relu=relu, row_mask=mask)
        else:
            if relu and hasattr(torch, "_addmm_activation"):
                y = torch._addmm_activation(bias, x2, weight.t(), use_gelu=False)   # library GEMM, ReLU in its epilogue
            else:
                y = torch.addmm(bias, x2, weight.t())
                if relu:
                    y = y.relu_()
            if mask is not None:
                y = y.masked_fill_(mask.view(torch.bool)[:, None], 0.0)
        ctx.save_for_backward(x2, weight, y if relu else None, mask)
        ctx.relu, ctx.plan, ctx.x_shape = relu, plan, x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        x2, weight, y, mask = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        gb_fused = None
        if ctx.relu:
            if ctx.needs_input_grad[2] and g2.is_contiguous() and g2.shape[1] % 4 == 0:
                g2, gb_fused = relu_backward_colsum(g2, y)       # masked rows have y == 0: zeroed as well
            else:
                g2 = torch.ops.aten.threshold_backward(g2, y, 0.0)
        elif mask is not None:
            # in place: the incoming gradient is the MSDA backward's own grad_value buffer (this layer's output has one
            # consumer), so the out-of-place form's 45 MB copy per call buys nothing
            g2 = (g2 if g2.is_contiguous() else g2.contiguous()).masked_fill_(mask.view(torch.bool)[:, None], 0.0)
        elif not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (gemm.linear_grad_input(g2, weight) if ctx.plan[1] else g2 @ weight).view(ctx.x_shape)
        if gb_fused is not None:
            gb = gb_fused
            if ctx.needs_input_grad[1]:
                gw = gemm.linear_grad_weight(g2, x2) if ctx.plan[2] else g2.t() @ x2
        elif ctx.needs_input_grad[1] and ctx.needs_input_grad[2] and ctx.plan[2]:
            gw, gb = gemm.linear_grad_weight(g2, x2, with_bias_grad=True)   # bias gradient from the same launch
        else:
            if ctx.needs_input_grad[1]:
                gw = gemm.linear_grad_weight(g2, x2) if ctx.plan[2] else g2.t() @ x2
            if ctx.needs_input_grad[2]:
                gb = column_sum(g2)
        return gx, gw, gb, None, None, None


class Linear(nn.Linear):
    def _plan(self, x, fused_epilogue, has_mask=None):
        if not (x.is_cuda and x.dtype == torch.float32 and self.bias is not None and self.weight.dtype == torch.float32
                and self.in_features % 4 == 0 and self.out_features % 4 == 0 and self.in_features >= 64
                and self.out_features >= 64 and x.numel() // self.in_features >= MIN_ROWS):
            return None
        return product_plan(self.in_features, self.out_features, fused_epilogue, has_mask)

    def forward(self, x, relu=False, row_mask=None):
        """relu / row_mask (bool, one entry per row of x): applied to the output, inside the GEMM epilogue when the
        forward product runs on the tcgen05 kernel."""
        # bf16 autocast (BASELINE.json configs[3]): the product is a library bf16 GEMM managed by torch.autocast -- the
        # hand-written GEMM of this package is TF32 on fp32 storage
        if (torch.is_autocast_enabled() and x.is_cuda and row_mask is None and self.bias is not None
                and torch.get_autocast_dtype("cuda") == torch.bfloat16 and self.out_features % 4 == 0
                and x.numel() // self.in_features >= MIN_ROWS and torch.is_grad_enabled()):
            return _AutocastLinearFn.apply(x, self.weight, self.bias, relu)
        plan = None if torch.is_autocast_enabled() else self._plan(x, relu or row_mask is not None, row_mask is not None)
        if plan is not None:
            return _TensorCoreLinearFn.apply(x, self.weight, self.bias, relu, row_mask, plan)
        if (x.is_cuda and x.dtype == torch.float32 and self.bias is not None and self.out_features % 4 == 0
                and x.numel() >= (1 << 16) and torch.is_grad_enabled() and not torch.is_autocast_enabled()):
            y = _LinearFn.apply(x, self.weight, self.bias)
        else:
            y = F.linear(x, self.weight, self.bias)
        if relu:
            y = F.relu(y)
        if row_mask is not None:
            y = y.masked_fill(row_mask[..., None], 0.0)
        return y
