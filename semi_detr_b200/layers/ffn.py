"""The FFN sub-block of every DINO layer, ``norm(x + linear2(relu(linear1(x))))`` (detr_od/models/utils/transformer.py
:626-630 encoder ``forward_ffn``, :878-882 decoder ``forward_ffn``; dropout is 0 in the DINO configs), as ONE autograd
node, so that its backward can be arranged around the kernels instead of around autograd's per-op graph:

* the grad-input product of ``linear2`` runs on the tcgen05 GEMM with the ReLU backward and ``linear1``'s bias gradient
  in its epilogue (``sdb_gemm_tf32_relu_grad``) -- autograd's ``mm -> threshold_backward -> sum`` is three passes over the
  (tokens, d_ffn) gradient, 364 MB each at the encoder shape;
* the residual's gradient d(x + y) rides into ``linear1``'s grad-input product as the library GEMM's C operand
  (``addmm``, beta = 1) instead of a standalone 45 MB add issued by the autograd engine;
* the residual add, the LayerNorm and the next encoder layer's ``+ pos`` are the fused LayerNorm kernels either way.

Forward products are the library's (cuBLASLt ReLU epilogue for linear1: measured faster than our kernel,
``layers/linear.py``); the weight gradients of the 2048-wide layers are library GEMMs as before.
"""
import os

import torch

from . import gemm
from .layernorm import add_layernorm_backward, add_layernorm_forward
from .linear import MIN_ROWS, column_sum, policy


class _FFNBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, eps, pos):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        h = torch._addmm_activation(b1, x2, w1.t(), use_gelu=False)       # relu(x W1^T + b1), ReLU in the library epilogue
        y = torch.addmm(b2, h, w2.t())
        p2 = None if pos is None else pos.reshape(x2.shape).contiguous()
        out, q, mean, rstd = add_layernorm_forward(x2, y, gamma, beta, eps, p2)
        ctx.save_for_backward(x2, h, y, w1, w2, gamma, mean, rstd)
        ctx.x_shape, ctx.has_q = x.shape, q is not None
        if q is None:
            return out.view(x.shape)
        return out.view(x.shape), q.view(x.shape)

    @staticmethod
    def backward(ctx, dout, dq=None):
        x2, h, y, w1, w2, gamma, mean, rstd = ctx.saved_tensors
        dpos = dq if (ctx.has_q and ctx.needs_input_grad[8]) else None
        if dout is None:
            dout, dq = dq, None
        dout = dout.reshape(x2.shape).contiguous()
        dq = dq.reshape(x2.shape).contiguous() if dq is not None else None
        g, dgamma, dbeta = add_layernorm_backward(dout, dq, x2, y, gamma, mean, rstd)     # d(x + y)
        db2 = column_sum(g)
        dw2 = g.t() @ h
        dh, db1 = gemm.linear_grad_input_relu(g, w2, h)                   # (g W2) * (h > 0) and its column sums
        dw1 = dh.t() @ x2
        dx = torch.addmm(g, dh, w1).view(ctx.x_shape) if ctx.needs_input_grad[0] else None   # g + dh W1
        return dx, dw1, db1, dw2, db2, dgamma, dbeta, None, dpos


def fused_ok(x, linear1, linear2, norm, dropouts):
    """The one-node FFN serves the shipped device configuration: fp32 CUDA tokens, TF32 products allowed, no active
    dropout, d_model = 256 LayerNorm kernel, gradients wanted; anything else takes the layer-by-layer route."""
    if not (x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and not torch.is_autocast_enabled()):
        return False
    if os.environ.get("SDB_FFN_BLOCK", "1") == "0":            # A/B switch: the layer-by-layer route
        return False
    if not torch.backends.cuda.matmul.allow_tf32 or policy() == "cublas":
        return False
    if any(d.training and d.p > 0 for d in dropouts):
        return False
    if not hasattr(torch, "_addmm_activation") or not norm._kernel_ok(x):
        return False
    if x.numel() // x.shape[-1] < MIN_ROWS or linear1.out_features % 4 or linear1.in_features % 4:
        return False
    return all(p is not None and p.dtype == torch.float32 and p.requires_grad
               for p in (linear1.weight, linear1.bias, linear2.weight, linear2.bias, norm.weight, norm.bias))


def ffn_block(x, linear1, linear2, norm, pos=None):
    """-> norm(x + linear2(relu(linear1(x)))) or, with ``pos``, (that, that + pos)"""
    return _FFNBlockFn.apply(x, linear1.weight, linear1.bias, linear2.weight, linear2.bias, norm.weight, norm.bias,
                             norm.eps, pos)
