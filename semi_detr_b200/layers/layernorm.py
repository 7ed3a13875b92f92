"""``LayerNorm`` -- an ``nn.LayerNorm`` (same parameters / state dict, transformer.py:606-642, 762-791, 1039) whose
CUDA path is the sm_100a kernel pair of ``csrc/layernorm.cu`` when normalising fp32 over d_model = 256.  Other
widths / dtypes use the library LayerNorm (plumbing outside the shipped configs)."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib

_WS = {}


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x2 = x.contiguous()
        rows = x2.numel() // x2.shape[-1]
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().sdb_layernorm_forward_f32(_lib.current_stream(x.device), x2.data_ptr(), weight.data_ptr(),
                                                      bias.data_ptr(), rows, x2.shape[-1], eps, y.data_ptr(),
                                                      mean.data_ptr(), rstd.data_ptr())
        _lib.check(rc, "layernorm_forward")
        _lib.LAUNCHES["layernorm_forward"] += 1
        ctx.save_for_backward(x2, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        rows = x.numel() // x.shape[-1]
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(weight)
        dbeta = torch.empty_like(weight)
        ws = torch.empty(_lib.lib().sdb_layernorm_bwd_workspace_floats(), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().sdb_layernorm_backward_f32(_lib.current_stream(x.device), dy.data_ptr(), x.data_ptr(),
                                                       weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows,
                                                       x.shape[-1], dx.data_ptr(), dgamma.data_ptr(),
                                                       dbeta.data_ptr(), ws.data_ptr())
        _lib.check(rc, "layernorm_backward")
        _lib.LAUNCHES["layernorm_backward"] += 2
        return dx, dgamma, dbeta, None


def add_layernorm_forward(x2, r2, weight, bias, eps, pos=None):
    """Raw launch of ``sdb_add_layernorm_forward_f32`` on contiguous operands: -> (y, q or None, mean, rstd) with
    y = LN(x2 + r2), q = y + pos."""
    rows = x2.numel() // x2.shape[-1]
    y = torch.empty_like(x2)
    q = torch.empty_like(x2) if pos is not None else None
    mean = torch.empty(rows, dtype=torch.float32, device=x2.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x2.device)
    with torch.cuda.device(x2.device):
        rc = _lib.lib().sdb_add_layernorm_forward_f32(
            _lib.current_stream(x2.device), x2.data_ptr(), r2.data_ptr(), weight.data_ptr(), bias.data_ptr(), rows,
            x2.shape[-1], eps, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _lib.ptr(pos), _lib.ptr(q))
    _lib.check(rc, "add_layernorm_forward")
    _lib.LAUNCHES["layernorm_forward"] += 1
    return y, q, mean, rstd


def add_layernorm_backward(dy, dq, x, r, weight, mean, rstd):
    """Raw launch of ``sdb_add_layernorm_backward_f32``: -> (d(x + r), dgamma, dbeta); ``dq`` (gradient of y + pos, or
    None) is folded into ``dy`` inside the kernel."""
    rows = x.numel() // x.shape[-1]
    dx = torch.empty_like(x)
    dgamma, dbeta = torch.empty_like(weight), torch.empty_like(weight)
    ws = torch.empty(_lib.lib().sdb_layernorm_bwd_workspace_floats(), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().sdb_add_layernorm_backward_f32(
            _lib.current_stream(x.device), dy.data_ptr(), _lib.ptr(dq), x.data_ptr(), r.data_ptr(),
            weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, x.shape[-1], dx.data_ptr(), dgamma.data_ptr(),
            dbeta.data_ptr(), ws.data_ptr())
    _lib.check(rc, "add_layernorm_backward")
    _lib.LAUNCHES["layernorm_backward"] += 2
    return dx, dgamma, dbeta


class _AddLayerNormFn(torch.autograd.Function):
    """y = LN(x + residual) [, q = y + pos] in one pass (``sdb_add_layernorm_forward_f32``); backward hands the same
    d(x + residual) to both addends and folds dq into dy inside the kernel -- no standalone add in either direction."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps, pos):
        x2, r2 = x.contiguous(), residual.contiguous()
        p2 = pos.contiguous() if pos is not None else None
        y, q, mean, rstd = add_layernorm_forward(x2, r2, weight, bias, eps, p2)
        ctx.save_for_backward(x2, r2, weight, mean, rstd)
        ctx.has_q = q is not None
        return y if q is None else (y, q)

    @staticmethod
    def backward(ctx, dy, dq=None):
        x, r, weight, mean, rstd = ctx.saved_tensors
        dpos = dq if (ctx.has_q and ctx.needs_input_grad[5]) else None    # q = y + pos: pos gets dq as it is
        if dy is None:                      # only the query output was used downstream
            dy, dq = dq, None
        dy = dy.contiguous()
        dq = dq.contiguous() if dq is not None else None
        dx, dgamma, dbeta = add_layernorm_backward(dy, dq, x, r, weight, mean, rstd)
        return dx, dx, dgamma, dbeta, None, dpos


class LayerNorm(nn.LayerNorm):
    def _kernel_ok(self, x):
        return (x.is_cuda and x.dtype == torch.float32 and self.normalized_shape == (256,) and self.elementwise_affine
                and self.weight.dtype == torch.float32)

    def add_norm(self, x, residual, pos=None):
        """``self(x + residual)`` -- the post-norm step of every DINO layer -- and, with ``pos``, also ``out + pos`` (the
        next encoder layer's query): -> out or (out, out + pos).  One kernel on the device path."""
        if self._kernel_ok(x) and residual.dtype == x.dtype and residual.shape == x.shape:
            return _AddLayerNormFn.apply(x, residual, self.weight, self.bias, self.eps, pos)
        out = self.forward(x + residual)
        return out if pos is None else (out, out + pos)

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and self.normalized_shape == (256,) and self.elementwise_affine
                and self.weight.dtype == torch.float32):
            return _LayerNormFn.apply(x, self.weight, self.bias, self.eps)
        return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
