"""Decoder self-attention core on the sm_100a kernels of ``csrc/attention.cu`` (``sdb_mha_forward/backward_f32``).

``fused_self_attention(qk, v, mask_add, mask_add_t, num_heads)`` computes, per (image, head),
``softmax((q d^-1/2) k^T + mask) v`` -- the inside of ``nn.MultiheadAttention`` as the DINO decoder layer calls it
(detr_od/models/utils/transformer.py:795-812) -- from the in-projection outputs, without the (N*H, T, T) score tensor:
``qk`` is the (T, N, 2C) output of the shared q / k projection (both halves are read in place through strides), ``v``
the (T, N, C) value projection, the result is (T, N, C) ready for ``out_proj``.

The kernels' products are TF32 tensor-core contractions, so the path is taken exactly when torch's TF32 matmul switch
is on (the mode bench.py times; with the switch off the layer keeps the fp32 library products it is compared against
the CPU oracle with).  ``SDB_ATTENTION=eager`` forces the written-out library path.
"""
import os

import torch

from .. import _lib


def attention_enabled(x, head_dim):
    return (x.is_cuda and x.dtype == torch.float32 and head_dim == 32 and torch.backends.cuda.matmul.allow_tf32
            and not torch.is_autocast_enabled() and os.environ.get("SDB_ATTENTION", "auto") != "eager")


_FLAGS = {"mask": None, "version": None, "flags": None}      # one entry: the decoder hands the same mask to all layers


def mask_tile_flags(mask_add):
    """(T, T) additive mask -> uint8 (ceil(T/64), ceil(T/64)): bit 0 = the 64 x 64 tile holds a non-zero mask value,
    bit 1 = every element of the tile (past-the-end rows / columns count as masked) is -inf.  Computed once per mask
    tensor (the cache keeps the tensor alive, so its address cannot be reused by another mask)."""
    if _FLAGS["mask"] is mask_add and _FLAGS["version"] == mask_add._version:
        return _FLAGS["flags"]
    T = mask_add.shape[0]
    nb = (T + 63) // 64
    pad = nb * 64 - T
    tiles = lambda m: m.view(nb, 64, nb, 64).permute(0, 2, 1, 3).reshape(nb, nb, 64 * 64)
    some = tiles(torch.nn.functional.pad(mask_add != 0, (0, pad, 0, pad), value=False)).any(-1)
    every = tiles(torch.nn.functional.pad(mask_add == float("-inf"), (0, pad, 0, pad), value=True)).all(-1)
    flags = (some.to(torch.uint8) + 2 * every.to(torch.uint8)).contiguous()
    _FLAGS.update(mask=mask_add, version=mask_add._version, flags=flags)
    return flags


class _SelfAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qk, v, mask_add, mask_add_t, num_heads):
        T, N, C2 = qk.shape
        C = C2 // 2
        D = C // num_heads
        qk, v = qk.contiguous(), v.contiguous()
        out = torch.empty((T, N, C), dtype=torch.float32, device=qk.device)
        lse = torch.empty((N * num_heads, T), dtype=torch.float32, device=qk.device)
        scale = float(D) ** -0.5
        mp = mask_add.data_ptr() if mask_add is not None else None
        flags = mask_tile_flags(mask_add) if mask_add is not None else None
        with torch.cuda.device(qk.device):
            rc = _lib.lib().sdb_mha_forward_f32(
                _lib.current_stream(qk.device), qk.data_ptr(), N * C2, C2, qk.data_ptr() + 4 * C, N * C2, C2,
                v.data_ptr(), N * C, C, mp, _lib.ptr(flags), T, N, num_heads, D, scale, out.data_ptr(), lse.data_ptr())
        _lib.check(rc, "mha_forward")
        _lib.LAUNCHES["mha_forward"] += 1
        ctx.save_for_backward(qk, v, out, lse, mask_add, mask_add_t, flags)
        ctx.num_heads, ctx.scale = num_heads, scale
        return out

    @staticmethod
    def backward(ctx, dout):
        qk, v, out, lse, mask_add, mask_add_t, flags = ctx.saved_tensors
        T, N, C2 = qk.shape
        C = C2 // 2
        H = ctx.num_heads
        dout = dout.contiguous()
        dqk = torch.empty_like(qk)
        dv = torch.empty_like(v)
        delta = torch.empty_like(lse)
        mp = mask_add.data_ptr() if mask_add is not None else None
        mtp = mask_add_t.data_ptr() if mask_add_t is not None else None
        with torch.cuda.device(qk.device):
            rc = _lib.lib().sdb_mha_backward_f32(
                _lib.current_stream(qk.device), qk.data_ptr(), N * C2, C2, qk.data_ptr() + 4 * C, N * C2, C2,
                v.data_ptr(), N * C, C, mp, mtp, _lib.ptr(flags), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), T, N, H,
                C // H,
                ctx.scale, dqk.data_ptr(), N * C2, C2, dqk.data_ptr() + 4 * C, N * C2, C2, dv.data_ptr(), N * C, C,
                delta.data_ptr())
        _lib.check(rc, "mha_backward")
        _lib.LAUNCHES["mha_backward"] += 2
        return dqk, dv, None, None, None


def fused_self_attention(qk, v, mask_add, mask_add_t, num_heads):
    """qk (T, N, 2C) = [q | k] projections, v (T, N, C); mask_add (T, T) additive float mask (row = query) or None and
    its transpose (row = key) -> (T, N, C)"""
    if (mask_add is None) != (mask_add_t is None):
        raise ValueError("fused_self_attention: pass the additive mask together with its transpose")
    if mask_add is not None:
        T = qk.shape[0]
        if mask_add.shape != (T, T) or mask_add.dtype != torch.float32 or not mask_add.is_contiguous() \
                or mask_add_t.shape != (T, T) or not mask_add_t.is_contiguous():
            raise ValueError("fused_self_attention: the mask must be a contiguous float32 (T, T) tensor")
    return _SelfAttentionFn.apply(qk, v, mask_add, mask_add_t, num_heads)
