"""Drop-in for the reference's compiled extension module ``MultiScaleDeformableAttention``.

The reference builds a pybind11 module of that name exporting ``ms_deform_attn_forward`` and
``ms_deform_attn_backward`` (detr_od/models/utils/ops/src/vision.cpp:13-16, signatures
src/ms_deform_attn.h:20-61) and imports it as ``MSDA`` in functions/ms_deform_attn_func.py:18.
This module has the same two functions with the same argument order, checks and error type, and
forwards to the sm_100a kernels through the C ABI (include/semidetr_b200.h).  Installing it under
that name (``sys.modules['MultiScaleDeformableAttention'] = this module``, see
``semi_detr_b200.install_as_reference_extension``) lets the reference's own ``MSDeformAttnFunction``
run on it unchanged.
"""
import ctypes

import torch

from .. import _lib

_FLOAT = (torch.float32, torch.float64)
# bf16 storage for value / output / grad_output with fp32 locations, weights and gradients (BASELINE.json configs[3]).
# The reference op has no such path (AT_DISPATCH_FLOATING_TYPES); this extends the surface, it replaces nothing.
_BF16 = torch.bfloat16

# TMA-staged forward for encoder self-attention (num_query == spatial_size): see csrc/msda_forward_tma.cu.
# Bit-identical to the L1-gather kernel but measured slower on B200 (187 vs 138 us on the N=2 microbench: both are
# bound by the same 128 B/clk/SM load data pipe and the staged variant keeps fewer warps resident), so it is opt-in.
USE_TMA = False
_HOST_INDEX = {}


def _host_index(spatial_shapes, level_start_index):
    """Host copies of the two index tensors (the TMA descriptors are encoded on the host): one device->host read
    per distinct tensor, then cached.  The cache entry keeps the device tensors alive so their addresses cannot be
    handed to a different tensor while the entry exists (``_version`` covers in-place writes)."""
    key = (spatial_shapes.data_ptr(), spatial_shapes._version, level_start_index.data_ptr(), level_start_index._version)
    hit = _HOST_INDEX.get(key)
    if hit is None:
        if len(_HOST_INDEX) > 256:
            _HOST_INDEX.clear()
        hit = (spatial_shapes.cpu().contiguous(), level_start_index.cpu().contiguous(), spatial_shapes,
               level_start_index)
        _HOST_INDEX[key] = hit
    return hit


def _tma_ok(value, num_query, num_levels, num_point):
    b, s, m, d = value.shape
    return (USE_TMA and value.dtype == torch.float32 and d == 32 and m == 8 and num_point == 4 and num_levels <= 4
            and num_query == s and s * 256 < (1 << 30) and value.data_ptr() % 128 == 0)


def _forward_tma(fused, value, spatial_shapes, level_start_index, ref, loc, attn, out, dims):
    b, s, m, d, l, q, p = dims
    hs, hl = _host_index(spatial_shapes, level_start_index)[:2]
    with torch.cuda.device(value.device):
        rc = _timed("fwd", b, s, q, lambda: _lib.lib().sdb_msda_forward_tma_f32(
            _lib.current_stream(value.device), 1 if fused else 0, value.data_ptr(), spatial_shapes.data_ptr(),
            level_start_index.data_ptr(), hs.data_ptr(), hl.data_ptr(), _lib.ptr(ref),
            ref.shape[-1] if ref is not None else 0, loc.data_ptr(), attn.data_ptr(), b, s, m, d, l, q, p,
            out.data_ptr()))
    _lib.check(rc, "ms_deform_attn_forward_tma")
    _lib.LAUNCHES["msda_forward_tma"] += 1
    return out


# Optional live timing (bench.py): when a list is installed here every launch is bracketed by CUDA events on the
# launching stream and (kind, batch, spatial_size, num_query, start_event, end_event) is appended.
EVENT_LOG = None


def _timed(kind, b, s, q, launch):
    if EVENT_LOG is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = launch()
    e1.record()
    EVENT_LOG.append((kind, b, s, q, e0, e1))
    return rc


def _check_inputs(named):
    # ms_deform_attn_cuda.cu:28-38 / :93-105 -- contiguity and device asserts, same messages
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    for name, t in named:
        if not t.is_cuda:
            if name == "value":
                raise RuntimeError("Not implemented on the CPU")      # src/ms_deform_attn.h:38,60
            raise RuntimeError(f"{name} must be a CUDA tensor")


def _dims(value, spatial_shapes, sampling_loc, im2col_step):
    batch, spatial_size, num_heads, channels = value.shape
    num_levels = spatial_shapes.shape[0]
    num_query, num_point = sampling_loc.shape[1], sampling_loc.shape[4]
    step = min(batch, im2col_step)
    if batch > 0 and (step <= 0 or batch % step != 0):
        # ms_deform_attn_cuda.cu:50-52
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")
    return batch, spatial_size, num_heads, channels, num_levels, num_query, num_point


def _index_tensors(spatial_shapes, level_start_index):
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 (torch.long)")
    return spatial_shapes, level_start_index


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """-> output (batch, num_query, num_heads*channels).  ms_deform_attn_cuda.cu:20-80."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)])
    _index_tensors(spatial_shapes, level_start_index)
    if value.dtype == _BF16:
        return _forward_bf16(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    if value.dtype not in _FLOAT:
        raise RuntimeError(f'"ms_deform_attn_forward_cuda" not implemented for \'{value.dtype}\'')
    if sampling_loc.dtype != value.dtype or attn_weight.dtype != value.dtype:
        raise RuntimeError("value, sampling_loc and attn_weight must share one floating dtype")
    b, s, m, d, l, q, p = _dims(value, spatial_shapes, sampling_loc, im2col_step)
    out = torch.empty((b, q, m * d), dtype=value.dtype, device=value.device)   # every element is written
    if _tma_ok(value, q, l, p):
        return _forward_tma(False, value, spatial_shapes, level_start_index, None, sampling_loc, attn_weight, out,
                            (b, s, m, d, l, q, p))
    fn = _lib.lib().sdb_msda_forward_f32 if value.dtype == torch.float32 else _lib.lib().sdb_msda_forward_f64
    with torch.cuda.device(value.device):
        rc = _timed("fwd", b, s, q, lambda: fn(
            _lib.current_stream(value.device), value.data_ptr(), spatial_shapes.data_ptr(),
            level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
            b, s, m, d, l, q, p, out.data_ptr()))
    _lib.check(rc, "ms_deform_attn_forward")
    _lib.LAUNCHES["msda_forward"] += 1
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight].  ms_deform_attn_cuda.cu:83-153."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)])
    _index_tensors(spatial_shapes, level_start_index)
    if value.dtype == _BF16:
        return _backward_bf16(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                              im2col_step)
    if value.dtype not in _FLOAT:
        raise RuntimeError(f'"ms_deform_attn_backward_cuda" not implemented for \'{value.dtype}\'')
    if not (sampling_loc.dtype == attn_weight.dtype == grad_output.dtype == value.dtype):
        raise RuntimeError("value, sampling_loc, attn_weight and grad_output must share one floating dtype")
    b, s, m, d, l, q, p = _dims(value, spatial_shapes, sampling_loc, im2col_step)
    grad_value = torch.empty_like(value)            # zero-filled by the call, stream-ordered
    grad_loc = torch.empty_like(sampling_loc)       # fully overwritten
    grad_attn = torch.empty_like(attn_weight)       # fully overwritten
    fn = _lib.lib().sdb_msda_backward_f32 if value.dtype == torch.float32 else _lib.lib().sdb_msda_backward_f64
    with torch.cuda.device(value.device):
        rc = _timed("bwd", b, s, q, lambda: fn(
            _lib.current_stream(value.device), grad_output.data_ptr(), value.data_ptr(),
            spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), b, s, m, d, l, q, p, grad_value.data_ptr(), grad_loc.data_ptr(),
            grad_attn.data_ptr()))
    _lib.check(rc, "ms_deform_attn_backward")
    _lib.LAUNCHES["msda_backward"] += 1
    return [grad_value, grad_loc, grad_attn]


# ---- bf16 storage (not part of the reference extension surface) -------------------------------------------

def _bf16_dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step, what):
    if sampling_loc.dtype != torch.float32 or attn_weight.dtype != torch.float32:
        raise RuntimeError(f"{what}: with a bfloat16 value, sampling_loc and attn_weight must be float32 "
                           "(the sampling arithmetic stays fp32)")
    dims = _dims(value, spatial_shapes, sampling_loc, im2col_step)
    b, s, m, d, l, q, p = dims
    if not (d == 32 and m == 8 and p == 4):
        raise RuntimeError(f"{what}: the bfloat16 kernels are built for 8 heads x 32 channels x 4 points "
                           f"(got {m} x {d} x {p}); use float32")
    return dims


def _forward_bf16(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    b, s, m, d, l, q, p = _bf16_dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step,
                                     "ms_deform_attn_forward")
    out = torch.empty((b, q, m * d), dtype=_BF16, device=value.device)
    with torch.cuda.device(value.device):
        rc = _timed("fwd", b, s, q, lambda: _lib.lib().sdb_msda_forward_bf16(
            _lib.current_stream(value.device), value.data_ptr(), spatial_shapes.data_ptr(),
            level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
            b, s, m, d, l, q, p, out.data_ptr()))
    _lib.check(rc, "ms_deform_attn_forward (bf16)")
    _lib.LAUNCHES["msda_forward_bf16"] += 1
    return out


def _backward_bf16(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    """-> [grad_value (bf16, narrowed from the fp32 accumulation), grad_sampling_loc fp32, grad_attn_weight fp32]"""
    if grad_output.dtype != _BF16:
        raise RuntimeError("ms_deform_attn_backward: grad_output must be bfloat16 like value")
    b, s, m, d, l, q, p = _bf16_dims(value, spatial_shapes, sampling_loc, attn_weight, im2col_step,
                                     "ms_deform_attn_backward")
    grad_value = torch.empty(value.shape, dtype=torch.float32, device=value.device)   # zero-filled by the call
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device):
        rc = _timed("bwd", b, s, q, lambda: _lib.lib().sdb_msda_backward_bf16(
            _lib.current_stream(value.device), grad_output.data_ptr(), value.data_ptr(),
            spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), b, s, m, d, l, q, p, grad_value.data_ptr(), grad_loc.data_ptr(),
            grad_attn.data_ptr()))
    _lib.check(rc, "ms_deform_attn_backward (bf16)")
    _lib.LAUNCHES["msda_backward_bf16"] += 1
    return [grad_value.to(_BF16), grad_loc, grad_attn]


# ---- fused prologue (not part of the reference extension surface) ----------------------------------------

def fused_supported(value, reference_points, n_levels, n_points):
    b, s, m, d = value.shape
    return (value.is_cuda and value.dtype == torch.float32 and d == 32 and m == 8 and n_points == 4
            and n_levels * n_points <= 16 and reference_points.shape[-1] in (2, 4))


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                 attn_logits):
    """value (N,S,M,D); reference_points (N,Lq,L,2|4); sampling_offsets (N,Lq,M,L,P,2) RAW; attn_logits
    (N,Lq,M,L*P) RAW (pre-softmax) -> (N, Lq, M*D).  Softmax + location arithmetic happen inside the kernel."""
    _check_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("reference_points", reference_points), ("sampling_offsets", sampling_offsets),
                   ("attn_logits", attn_logits)])
    b, s, m, d = value.shape
    l = spatial_shapes.shape[0]
    q, p = sampling_offsets.shape[1], sampling_offsets.shape[4]
    out = torch.empty((b, q, m * d), dtype=value.dtype, device=value.device)
    if _tma_ok(value, q, l, p):
        return _forward_tma(True, value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                            attn_logits, out, (b, s, m, d, l, q, p))
    with torch.cuda.device(value.device):
        rc = _timed("fwd", b, s, q, lambda: _lib.lib().sdb_msda_fused_forward_f32(
            _lib.current_stream(value.device), value.data_ptr(), spatial_shapes.data_ptr(),
            level_start_index.data_ptr(), reference_points.data_ptr(), reference_points.shape[-1],
            sampling_offsets.data_ptr(), attn_logits.data_ptr(), b, s, m, d, l, q, p, out.data_ptr()))
    _lib.check(rc, "ms_deform_attn_fused_forward")
    _lib.LAUNCHES["msda_fused_forward"] += 1
    return out


def ms_deform_attn_fused_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                  attn_logits, grad_output):
    """-> [grad_value, grad_sampling_offsets, grad_attn_logits]"""
    _check_inputs([("value", value), ("reference_points", reference_points), ("sampling_offsets", sampling_offsets),
                   ("attn_logits", attn_logits), ("grad_output", grad_output)])
    b, s, m, d = value.shape
    l = spatial_shapes.shape[0]
    q, p = sampling_offsets.shape[1], sampling_offsets.shape[4]
    grad_value = torch.empty_like(value)
    grad_off = torch.empty_like(sampling_offsets)
    grad_logits = torch.empty_like(attn_logits)
    with torch.cuda.device(value.device):
        rc = _timed("bwd", b, s, q, lambda: _lib.lib().sdb_msda_fused_backward_f32(
            _lib.current_stream(value.device), grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(),
            level_start_index.data_ptr(), reference_points.data_ptr(), reference_points.shape[-1],
            sampling_offsets.data_ptr(), attn_logits.data_ptr(), b, s, m, d, l, q, p, grad_value.data_ptr(),
            grad_off.data_ptr(), grad_logits.data_ptr()))
    _lib.check(rc, "ms_deform_attn_fused_backward")
    _lib.LAUNCHES["msda_fused_backward"] += 1
    return [grad_value, grad_off, grad_logits]


# ---- standalone prologue (levels x points > 16: the fused kernels do not apply) ---------------------------------------

def prologue_supported(sampling_offsets, reference_points, n_levels, n_points):
    return (sampling_offsets.is_cuda and sampling_offsets.dtype in (torch.float32, _BF16) and n_points == 4
            and n_levels <= 8 and n_levels * n_points <= 32 and reference_points.shape[-1] in (2, 4))


def _shapes_array(shapes_host):
    flat = [int(v) for hw in shapes_host for v in hw]
    return (ctypes.c_int64 * len(flat))(*flat)


def msda_prologue_forward(sampling_offsets, attn_logits, reference_points, shapes_host):
    """-> (sampling_locations fp32 (N, Lq, M, L, P, 2), attention_weights fp32 (N, Lq, M, L, P))"""
    n, q, m, l, p, _ = sampling_offsets.shape
    off, lg = sampling_offsets.contiguous(), attn_logits.to(sampling_offsets.dtype).contiguous()
    ref = reference_points.float().contiguous()
    loc = torch.empty((n, q, m, l, p, 2), dtype=torch.float32, device=off.device)
    attn = torch.empty((n, q, m, l, p), dtype=torch.float32, device=off.device)
    fn = _lib.lib().sdb_msda_prologue_forward_bf16 if off.dtype == _BF16 else _lib.lib().sdb_msda_prologue_forward_f32
    with torch.cuda.device(off.device):
        rc = fn(_lib.current_stream(off.device), off.data_ptr(), lg.data_ptr(), ref.data_ptr(), ref.shape[-1],
                _shapes_array(shapes_host), n, q, m, l, p, loc.data_ptr(), attn.data_ptr())
    _lib.check(rc, "msda_prologue_forward")
    _lib.LAUNCHES["msda_prologue_forward"] += 1
    return loc, attn


def msda_prologue_backward(grad_loc, grad_attn, attn, reference_points, shapes_host, raw_dtype):
    """-> (grad_sampling_offsets, grad_attn_logits) in ``raw_dtype``"""
    n, q, m, l, p, _ = grad_loc.shape
    ref = reference_points.float().contiguous()
    g_off = torch.empty((n, q, m, l, p, 2), dtype=raw_dtype, device=grad_loc.device)
    g_lg = torch.empty((n, q, m, l * p), dtype=raw_dtype, device=grad_loc.device)
    fn = _lib.lib().sdb_msda_prologue_backward_bf16 if raw_dtype == _BF16 else _lib.lib().sdb_msda_prologue_backward_f32
    with torch.cuda.device(grad_loc.device):
        rc = fn(_lib.current_stream(grad_loc.device), grad_loc.float().data_ptr() if grad_loc.dtype != torch.float32
                else grad_loc.data_ptr(), grad_attn.data_ptr(), attn.data_ptr(), ref.data_ptr(), ref.shape[-1],
                _shapes_array(shapes_host), n, q, m, l, p, g_off.data_ptr(), g_lg.data_ptr())
    _lib.check(rc, "msda_prologue_backward")
    _lib.LAUNCHES["msda_prologue_backward"] += 1
    return g_off, g_lg
