"""ctypes access to oracle/_ref/libref_msda.so -- the REFERENCE's own CUDA launchers compiled for sm_100a
(built by `make -C oracle ref` in the authoring container; the prebuilt .so travels to the GPU box).
Test infrastructure only."""
import ctypes
import os

import torch

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_msda.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        _lib.ref_msda_forward_f32.argtypes = [vp] * 6 + [ci] * 7 + [vp]
        _lib.ref_msda_backward_f32.argtypes = [vp] * 7 + [ci] * 7 + [vp] * 3
    return _lib


def forward(value, shapes, start, loc, attn):
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    out = torch.empty(N, Lq, M * D, device=value.device)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib().ref_msda_forward_f32(st, value.data_ptr(), shapes.data_ptr(), start.data_ptr(), loc.data_ptr(),
                                    attn.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr())
    assert rc == 0, rc
    return out


def backward(value, shapes, start, loc, attn, gout):
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(attn)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib().ref_msda_backward_f32(st, gout.data_ptr(), value.data_ptr(), shapes.data_ptr(), start.data_ptr(),
                                     loc.data_ptr(), attn.data_ptr(), N, S, M, D, L, Lq, P, gv.data_ptr(),
                                     gl.data_ptr(), ga.data_ptr())
    assert rc == 0, rc
    return gv, gl, ga
