"""The second half of every DINO layer, ``norm_b(x + linear2(relu(linear1(x))))`` with ``x = norm_a(a + r)`` -- the
post-attention LayerNorm, the FFN and the closing LayerNorm (detr_od/models/utils/transformer.py:606-642 encoder layer,
:878-882 / :762-791 decoder ``forward_ffn`` behind the cross-attention norm; dropout is 0 in the DINO configs) -- as ONE
autograd node, so that its backward is arranged around the kernels instead of around autograd's per-op graph:

* the grad-input product of ``linear2`` runs on the tcgen05 GEMM with the ReLU backward and ``linear1``'s bias gradient
  in its epilogue (``sdb_gemm_tf32_relu_grad``): autograd's ``mm -> threshold_backward -> sum`` is three passes over the
  (tokens, d_ffn) gradient, 364 MB each at the encoder shape;
* ``x`` feeds both ``linear1`` and the residual of ``norm_b``; its two gradients -- ``dh W1`` and d(x + y) -- go to
  ``norm_a``'s backward kernel as its two gradient inputs (it already folds the gradient of a second output) instead of
  through a standalone 45 MB add issued by the autograd engine.  (Handing d(x + y) to the grad-input GEMM as its C
  operand, ``addmm`` with beta = 1, was tried first and showed no gain in the step.)

Forward products are the library's (cuBLASLt ReLU epilogue for linear1: measured faster than our kernel,
``layers/linear.py``); the weight gradients of the 2048-wide layers are library GEMMs as before.
"""
import os

import torch

from . import gemm
from .layernorm import add_layernorm_backward, add_layernorm_forward
from .linear import MIN_ROWS, column_sum, policy


class _PostAttentionBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, r, gamma_a, beta_a, eps_a, w1, b1, w2, b2, gamma_b, beta_b, eps_b, pos):
        shape = a.shape
        a2, r2 = a.reshape(-1, shape[-1]), r.reshape(-1, shape[-1])
        if not a2.is_contiguous():
            a2 = a2.contiguous()
        if not r2.is_contiguous():
            r2 = r2.contiguous()
        x, _, mean_a, rstd_a = add_layernorm_forward(a2, r2, gamma_a, beta_a, eps_a, None)
        h = torch._addmm_activation(b1, x, w1.t(), use_gelu=False)        # relu(x W1^T + b1), ReLU in the library epilogue
        y = torch.addmm(b2, h, w2.t())
        p2 = None if pos is None else pos.reshape(a2.shape).contiguous()
        out, q, mean_b, rstd_b = add_layernorm_forward(x, y, gamma_b, beta_b, eps_b, p2)
        ctx.save_for_backward(a2, r2, x, h, y, w1, w2, gamma_a, mean_a, rstd_a, gamma_b, mean_b, rstd_b)
        ctx.shape, ctx.has_q = shape, q is not None
        if q is None:
            return out.view(shape)
        return out.view(shape), q.view(shape)

    @staticmethod
    def backward(ctx, dout, dq=None):
        a2, r2, x, h, y, w1, w2, gamma_a, mean_a, rstd_a, gamma_b, mean_b, rstd_b = ctx.saved_tensors
        dpos = dq if (ctx.has_q and ctx.needs_input_grad[12]) else None
        if dout is None:
            dout, dq = dq, None
        dout = dout.reshape(a2.shape).contiguous()
        dq = dq.reshape(a2.shape).contiguous() if dq is not None else None
        g, dgamma_b, dbeta_b = add_layernorm_backward(dout, dq, x, y, gamma_b, mean_b, rstd_b)   # d(x + y)
        db2 = column_sum(g)
        dw2 = g.t() @ h
        dh, db1 = gemm.linear_grad_input_relu(g, w2, h)                   # (g W2) * (h > 0) and its column sums
        dw1 = dh.t() @ x
        # dx = dh W1 + g: the sum is formed inside norm_a's backward kernel (its second gradient input)
        d_ar, dgamma_a, dbeta_a = add_layernorm_backward(dh @ w1, g, a2, r2, gamma_a, mean_a, rstd_a)
        d_ar = d_ar.view(ctx.shape)
        return (d_ar if ctx.needs_input_grad[0] else None, d_ar if ctx.needs_input_grad[1] else None, dgamma_a, dbeta_a,
                None, dw1, db1, dw2, db2, dgamma_b, dbeta_b, None, dpos)


def fused_ok(a, norm_a, linear1, linear2, norm_b, dropouts):
    """The one-node block serves the shipped device configuration: fp32 CUDA tokens, TF32 products allowed, no active
    dropout, d_model = 256 LayerNorm kernels, gradients wanted; anything else takes the layer-by-layer route."""
    if not (a.is_cuda and a.dtype == torch.float32 and torch.is_grad_enabled() and not torch.is_autocast_enabled()):
        return False
    if os.environ.get("SDB_FFN_BLOCK", "1") == "0":            # A/B switch: the layer-by-layer route
        return False
    if not torch.backends.cuda.matmul.allow_tf32 or policy() != "auto":    # the other policies pin every product's owner
        return False
    if any(d.training and d.p > 0 for d in dropouts):
        return False
    if not hasattr(torch, "_addmm_activation") or not norm_a._kernel_ok(a) or not norm_b._kernel_ok(a):
        return False
    if a.numel() // a.shape[-1] < MIN_ROWS or linear1.out_features % 4 or linear1.in_features % 4:
        return False
    return all(p is not None and p.dtype == torch.float32 and p.requires_grad
               for p in (linear1.weight, linear1.bias, linear2.weight, linear2.bias, norm_a.weight, norm_a.bias,
                         norm_b.weight, norm_b.bias))


def post_attention_block(a, r, norm_a, linear1, linear2, norm_b, pos=None):
    """x = norm_a(a + r) -> norm_b(x + linear2(relu(linear1(x)))) or, with ``pos``, (that, that + pos)"""
    if r.dtype != a.dtype or r.shape != a.shape:
        raise RuntimeError("post_attention_block: the residual must match the input's shape and dtype")
    return _PostAttentionBlockFn.apply(a, r, norm_a.weight, norm_a.bias, norm_a.eps, linear1.weight, linear1.bias,
                                       linear2.weight, linear2.bias, norm_b.weight, norm_b.bias, norm_b.eps, pos)
