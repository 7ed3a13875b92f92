"""Host side of ``sdb_gemm_tf32`` (``csrc/gemm_tf32.cu``): the three products of an ``nn.Linear`` on tcgen05 tensor
cores, TF32 arithmetic on fp32 storage.

    y = x W^T + b   -> linear_forward(x, W, b)            (optional ReLU / zeroed rows in the epilogue)
    dx = dy W       -> linear_grad_input(dy, W)
    dW = dy^T x     -> linear_grad_weight(dy, x)          (split along the token axis, reduce-added by TMA)

Reference call sites: ``ms_deform_attn.py:61-65, 94-112`` (value_proj with ``masked_fill``, sampling_offsets,
attention_weights, output_proj) and the FFNs of ``transformer.py:626-630, 878-882``.  No CPU path: CPU tensors raise.
"""
import os

import torch

from .. import _lib

_SMS = {}


def _sm_count(device):
    i = device.index if device.index is not None else torch.cuda.current_device()
    if i not in _SMS:
        _SMS[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _SMS[i]


def _check(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"gemm_tf32: {name} is not a CUDA tensor (Not implemented on the CPU)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError(f"gemm_tf32: {name} must be contiguous float32")


def gemm_tf32(a, a_mn_major, b, b_mn_major, m, n, k, bias=None, row_mask=None, relu=False, out=None, k_splits=1,
              a_column_sums=None, round_mode=3):
    """Raw entry: see include/semidetr_b200.h.  ``out`` (m, n) is allocated when None (zero-filled if k_splits > 1).
    round_mode 3 (default): both operands rounded to nearest TF32 -- the unbiased product cuBLAS-TF32 computes."""
    _check(a, "a")
    _check(b, "b")
    if out is None:
        out = (torch.zeros if k_splits > 1 else torch.empty)((m, n), dtype=torch.float32, device=a.device)
    else:
        _check(out, "out")
    if bias is not None:
        _check(bias, "bias")
    if row_mask is not None:
        if row_mask.dtype == torch.bool:
            row_mask = row_mask.view(torch.uint8)
        if not row_mask.is_cuda or row_mask.dtype != torch.uint8 or not row_mask.is_contiguous() or row_mask.numel() != m:
            raise RuntimeError("gemm_tf32: row_mask must be a contiguous (m,) bool / uint8 CUDA tensor")
    with torch.cuda.device(a.device):
        rc = _lib.lib().sdb_gemm_tf32(_lib.current_stream(a.device), a.data_ptr(), int(a_mn_major), b.data_ptr(),
                                      int(b_mn_major), out.data_ptr(), m, n, k, _lib.ptr(bias), _lib.ptr(row_mask),
                                      int(relu), int(k_splits), int(round_mode), _lib.ptr(a_column_sums))
    _lib.check(rc, "gemm_tf32")
    _lib.LAUNCHES["gemm_tf32"] += 1
    return out


def linear_forward(x2d, weight, bias=None, relu=False, row_mask=None):
    """x2d (tokens, in) . weight (out, in)^T + bias -> (tokens, out)"""
    m, k = x2d.shape
    n = weight.shape[0]
    return gemm_tf32(x2d, 0, weight, 0, m, n, k, bias=bias, row_mask=row_mask, relu=relu)


def linear_grad_input(g2d, weight):
    """g2d (tokens, out) . weight (out, in) -> (tokens, in)"""
    m, k = g2d.shape
    n = weight.shape[1]
    return gemm_tf32(g2d, 0, weight, 1, m, n, k)


def linear_grad_input_relu(g2d, weight, h2d, round_mode=None):
    """The grad-input product of the layer behind a ReLU with that ReLU's backward in its epilogue:
    -> ((g2d . weight) * (h2d > 0), its column sums) = (d pre-activation, d bias of the layer in front).
    g2d (tokens, out), weight (out, in), h2d (tokens, in) = relu output that fed the layer.
    round_mode (default ``SDB_GEMM_RELU_GRAD_ROUND`` or 2): as in ``gemm_tf32``.  2 = the weights are rounded to the
    nearest TF32 once per CTA (they are resident), the streamed gradient operand is left to the tensor core's own
    truncation -- what the library's TF32 GEMM does with both operands -- which takes the per-stage rounding pass over
    shared memory out of the main loop (181 vs 217 us at the encoder FFN shape, tools/time_ffn_dgrad.py)."""
    _check(g2d, "g2d")
    _check(weight, "weight")
    _check(h2d, "h2d")
    m, k = g2d.shape
    n = weight.shape[1]
    if h2d.shape != (m, n) or weight.shape[0] != k:
        raise RuntimeError(f"linear_grad_input_relu: shapes {tuple(g2d.shape)} {tuple(weight.shape)} {tuple(h2d.shape)}")
    if round_mode is None:
        round_mode = int(os.environ.get("SDB_GEMM_RELU_GRAD_ROUND", "2"))
    out = torch.empty((m, n), dtype=torch.float32, device=g2d.device)
    sums = torch.empty(n, dtype=torch.float32, device=g2d.device)
    with torch.cuda.device(g2d.device):
        rc = _lib.lib().sdb_gemm_tf32_relu_grad(_lib.current_stream(g2d.device), g2d.data_ptr(), 0, weight.data_ptr(), 1,
                                                out.data_ptr(), m, n, k, h2d.data_ptr(), sums.data_ptr(),
                                                int(round_mode))
    _lib.check(rc, "gemm_tf32_relu_grad")
    _lib.LAUNCHES["gemm_tf32"] += 1
    return out, sums


def linear_grad_weight(g2d, x2d, with_bias_grad=False):
    """g2d (tokens, out)^T . x2d (tokens, in) -> (out, in); the token axis is split over the SMs.
    with_bias_grad: also return g2d.sum(0), accumulated by the same launch from the tiles of g2d it stages."""
    k, m = g2d.shape
    n = x2d.shape[1]
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    splits = max(1, min((k + 31) // 32, (2 * _sm_count(g2d.device)) // tiles))
    # (a floor on the k-blocks per split -- fewer, longer splits for the decoder-sized products, whose 69 splits each
    # reduce-add a full 64 KB tile -- was measured: 1.52 / 1.49 / 1.58 / 1.85 ms per step for a floor of 0 / 4 / 8 / 16)
    if not with_bias_grad:
        return gemm_tf32(g2d, 1, x2d, 1, m, n, k, k_splits=splits)
    buf = torch.zeros(m * n + m, dtype=torch.float32, device=g2d.device)      # one fill for both outputs
    gw, gb = buf[:m * n].view(m, n), buf[m * n:]
    gemm_tf32(g2d, 1, x2d, 1, m, n, k, out=gw, k_splits=splits, a_column_sums=gb)
    return gw, gb
