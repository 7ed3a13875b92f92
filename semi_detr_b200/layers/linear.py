"""``Linear`` -- an ``nn.Linear`` (same parameters / state dict) whose backward computes the bias gradient with the
streaming column-sum kernel (``csrc/colsum.cu``) instead of torch's generic reduction; the two matrix products
stay library GEMMs.  Used for the token-wise linears of the transformer (FFN, MSDA projections, enc_output)."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib


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


class Linear(nn.Linear):
    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and self.bias is not None and self.out_features % 4 == 0
                and x.numel() >= (1 << 16) and torch.is_grad_enabled()):
            return _LinearFn.apply(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)
