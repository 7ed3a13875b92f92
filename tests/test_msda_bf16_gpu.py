"""bf16-storage MSDA kernels (BASELINE.json configs[3]: bf16 value / output, fp32 sampling math) against the oracle.

First hardware run: round 2 (11 passed on a B200); part of the default `pytest -m gpu` suite since.

Tolerances: the output is the exact sum rounded to bf16 once (device: fp32 accumulation), so it must sit within one
bf16 ulp (2^-8 relative) of the float64 oracle value plus the fp32 accumulation error; gradients are produced in fp32
from bf16-exact inputs, so they keep the fp32 bound of the north star (1e-3 relative), grad_value after its final
narrowing to bf16 one bf16 ulp.
"""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu

BF16_ULP = 2.0 ** -8
LEVELS4 = [(19, 27), (10, 14), (5, 7), (3, 4)]
LEVELS5 = [(38, 54), (19, 27), (10, 14), (5, 7), (3, 4)]


def _inputs(levels, mode, Lq, seed):
    from semi_detr_b200.synthetic import msda_inputs
    x = msda_inputs(levels, N=2, Lq=Lq, mode=mode, seed=seed)
    x["value_bf16"] = x["value"].to(torch.bfloat16)
    x["gout_bf16"] = x["gout"].to(torch.bfloat16)
    return x


def _relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _smooth_mask(loc, levels):
    """Samples within 1e-3 px of an integer pixel coordinate are left out of the grad_loc comparison (the derivative
    jumps there and fp32 / fp64 can round to different sides) -- same rule as tests/test_msda_gpu.py."""
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float64)[None, None, None, :, None, :]
    px = loc.double().cpu() * wh - 0.5
    near = (px - px.round()).abs() < 1e-3
    return ~(near.any(-1, keepdim=True).expand_as(px))


def _np(x, *keys):
    return [x[k].float().cpu().numpy() for k in keys]


@pytest.mark.parametrize("levels", [LEVELS4, LEVELS5], ids=["4lvl", "5lvl"])
@pytest.mark.parametrize("mode,Lq", [("encoder", None), ("wide", 300), ("uniform", 77)])
def test_forward_vs_oracle(levels, mode, Lq):
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    x = _inputs(levels, mode, Lq, seed=3)
    out = MSDA.ms_deform_attn_forward(x["value_bf16"], x["shapes"], x["start"], x["loc"], x["attn"], 64)
    assert out.dtype == torch.bfloat16 and out.shape == (2, x["loc"].shape[1], 256)
    value, start, loc, attn = _np(x, "value_bf16", "start", "loc", "attn")
    rounded, exact = O.msda_forward_bf16(value, levels, start.astype(np.int64), loc, attn)
    got = out.float().cpu().numpy()
    # one rounding to bf16 of a value that differs from `exact` only by fp32 accumulation error
    np.testing.assert_allclose(got, exact, rtol=BF16_ULP, atol=1e-4)
    assert (got == rounded).mean() > 0.99     # nearly every element is the correctly rounded bf16 value


@pytest.mark.parametrize("levels", [LEVELS4, LEVELS5], ids=["4lvl", "5lvl"])
@pytest.mark.parametrize("mode,Lq", [("encoder", None), ("wide", 300)])
def test_backward_vs_oracle(levels, mode, Lq):
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    x = _inputs(levels, mode, Lq, seed=4)
    gv, gl, ga = MSDA.ms_deform_attn_backward(x["value_bf16"], x["shapes"], x["start"], x["loc"], x["attn"],
                                              x["gout_bf16"], 64)
    assert gv.dtype == torch.bfloat16 and gl.dtype == torch.float32 and ga.dtype == torch.float32
    value, start, loc, attn, gout = _np(x, "value_bf16", "start", "loc", "attn", "gout_bf16")
    rv, rl, ra = O.msda_backward_bf16(value, levels, start.astype(np.int64), loc, attn, gout)
    mask = _smooth_mask(x["loc"], levels)
    assert _relerr(gl.cpu() * mask, torch.from_numpy(rl) * mask) < 1e-5
    assert _relerr(ga.cpu(), torch.from_numpy(ra)) < 1e-5
    np.testing.assert_allclose(ga.cpu().numpy(), ra, rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose((gl.cpu() * mask).numpy(), (torch.from_numpy(rl) * mask).numpy(), rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(gv.float().cpu().numpy(), rv, rtol=BF16_ULP, atol=2e-3)
    assert _relerr(gv.float().cpu(), torch.from_numpy(rv)) < BF16_ULP


def test_autograd_function_round_trip_and_errors():
    from semi_detr_b200.msda import MSDeformAttnFunction
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    x = _inputs(LEVELS4, "wide", 64, seed=5)
    v = x["value_bf16"].clone().requires_grad_(True)
    l = x["loc"].clone().requires_grad_(True)
    a = x["attn"].clone().requires_grad_(True)
    out = MSDeformAttnFunction.apply(v, x["shapes"], x["start"], l, a, 64)
    out.backward(x["gout_bf16"])
    assert v.grad.dtype == torch.bfloat16 and l.grad.dtype == torch.float32 and a.grad.dtype == torch.float32
    # same op on the same (bf16-exact) numbers in fp32 storage
    v32 = x["value_bf16"].float().requires_grad_(True)
    l32 = x["loc"].clone().requires_grad_(True)
    a32 = x["attn"].clone().requires_grad_(True)
    out32 = MSDeformAttnFunction.apply(v32, x["shapes"], x["start"], l32, a32, 64)
    out32.backward(x["gout_bf16"].float())
    assert torch.equal(out, out32.to(torch.bfloat16))                       # identical fp32 arithmetic, one rounding
    assert torch.allclose(l.grad, l32.grad, rtol=1e-4, atol=1e-5) and torch.allclose(a.grad, a32.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(v.grad.float(), v32.grad, rtol=2 * BF16_ULP, atol=1e-3)
    with pytest.raises(RuntimeError, match="float32"):
        MSDA.ms_deform_attn_forward(x["value_bf16"], x["shapes"], x["start"], x["loc"].to(torch.bfloat16), x["attn"], 64)
    with pytest.raises(RuntimeError, match="8 heads"):
        MSDA.ms_deform_attn_forward(x["value_bf16"][:, :, :4].contiguous(), x["shapes"], x["start"],
                                    x["loc"][:, :, :4].contiguous(), x["attn"][:, :, :4].contiguous(), 64)
    # empty query set: nothing launched, grad_value all zero
    e = MSDA.ms_deform_attn_forward(x["value_bf16"], x["shapes"], x["start"], x["loc"][:, :0].contiguous(),
                                    x["attn"][:, :0].contiguous(), 64)
    assert e.shape == (2, 0, 256)
