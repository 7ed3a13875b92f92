"""The numpy MSDA oracle against golden vectors produced by the reference's own
ms_deform_attn_core_pytorch (functions/ms_deform_attn_func.py:41-61) -- see tests/golden/make_golden.py."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

CASES = ["testpy", "small_d32", "wide_d32", "odd_d30", "d64_l5", "one_point"]


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference_fp64(msda_golden, name):
    c = msda_golden[name]
    out = O.msda_forward(c["value"], c["shapes"], c["start"], c["loc"], c["attn"])
    np.testing.assert_allclose(out, c["out"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_backward_matches_reference_fp64(msda_golden, name):
    c = msda_golden[name]
    gv, gl, ga = O.msda_backward(c["value"], c["shapes"], c["start"], c["loc"], c["attn"], c["gout"])
    np.testing.assert_allclose(gv, c["grad_value"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ga, c["grad_attn"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gl, c["grad_loc"], rtol=1e-8, atol=1e-10)


def test_testpy_case_fp32_tolerance(msda_golden):
    """ops/test.py:47-60: fp32 result within rtol 1e-2 / atol 1e-3 of the python fallback."""
    c = msda_golden["testpy"]
    out = O.msda_forward(c["value"], c["shapes"], c["start"], c["loc"], c["attn"], dtype=np.float32)
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, c["out"], rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("name", ["small_d32", "wide_d32"])
def test_torch_cpu_baseline_matches(msda_golden, name):
    c = msda_golden[name]
    v = torch.from_numpy(c["value"]).double().requires_grad_(True)
    l = torch.from_numpy(c["loc"]).double().requires_grad_(True)
    a = torch.from_numpy(c["attn"]).double().requires_grad_(True)
    y = O.msda_forward_torch(v, c["shapes"], l, a)
    y.backward(torch.from_numpy(c["gout"]).double())
    np.testing.assert_allclose(y.detach().numpy(), c["out"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(v.grad.numpy(), c["grad_value"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(l.grad.numpy(), c["grad_loc"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(a.grad.numpy(), c["grad_attn"], rtol=1e-9, atol=1e-12)


def test_empty_query_and_out_of_range():
    shapes = np.array([[3, 4]], np.int64)
    start = np.array([0], np.int64)
    value = np.random.default_rng(0).standard_normal((1, 12, 2, 4))
    loc = np.full((1, 3, 2, 1, 2, 2), 5.0)          # far outside: contributes exactly zero
    attn = np.full((1, 3, 2, 1, 2), 0.5)
    out = O.msda_forward(value, shapes, start, loc, attn)
    assert out.shape == (1, 3, 8) and not out.any()
    gv, gl, ga = O.msda_backward(value, shapes, start, loc, attn, np.ones((1, 3, 8)))
    assert not gv.any() and not gl.any() and not ga.any()
    out0 = O.msda_forward(value, shapes, start, loc[:, :0], attn[:, :0])
    assert out0.shape == (1, 0, 8)


def test_bf16_rounding_is_torch_round_to_nearest_even():
    """The bf16-storage oracle (configs[3]) rounds exactly like ``tensor.to(torch.bfloat16)`` / ``cvt.rn.bf16.f32``."""
    import torch
    rng = np.random.default_rng(0)
    x = np.concatenate([(rng.standard_normal(100000) * 10.0 ** rng.uniform(-30, 30, 100000)).astype(np.float32),
                        np.array([0., -0., np.inf, -np.inf, 1.00390625, 1.01171875, 3.3895314e38, 1e-40, -1e-45],
                                 np.float32)])
    got = O.bf16_round(x)
    want = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.isnan(O.bf16_round(np.array([np.nan], np.float32))[0])


def test_bf16_variant_is_the_fp_op_on_rounded_inputs(msda_golden):
    """bf16 storage changes what is stored, not the arithmetic: forward = round(exact op on rounded value),
    backward = exact adjoint on rounded value / grad_output."""
    c = msda_golden["small_d32"]
    levels = [tuple(int(v) for v in r) for r in c["shapes"]]
    rounded, exact = O.msda_forward_bf16(c["value"], levels, c["start"], c["loc"], c["attn"])
    want = O.msda_forward(O.bf16_round(c["value"]), levels, c["start"], c["loc"], c["attn"])
    assert np.array_equal(exact, want)
    assert np.allclose(rounded, exact, rtol=2.0 ** -8, atol=1e-30)
    gv, gl, ga = O.msda_backward_bf16(c["value"], levels, c["start"], c["loc"], c["attn"], c["gout"])
    wv, wl, wa = O.msda_backward(O.bf16_round(c["value"]), levels, c["start"], c["loc"], c["attn"],
                                 O.bf16_round(c["gout"]))
    assert np.array_equal(gv, wv) and np.array_equal(gl, wl) and np.array_equal(ga, wa)


def test_bf16_widening_bit_logic_of_the_kernels():
    """``Chan4<__nv_bfloat16>::widen`` (csrc/common.cuh): four bf16 channels arrive as two little-endian 32-bit words;
    element 0 is the LOW half of word 0, and bf16 -> fp32 is a 16-bit left shift (x << 16, x & 0xffff0000)."""
    import torch
    v = torch.randn(64, 4).to(torch.bfloat16)
    words = v.view(torch.int16).numpy().astype(np.uint16).reshape(-1, 2, 2)            # [.., word, half]
    rx = words[:, 0, 0].astype(np.uint32) | (words[:, 0, 1].astype(np.uint32) << 16)
    ry = words[:, 1, 0].astype(np.uint32) | (words[:, 1, 1].astype(np.uint32) << 16)
    widened = np.stack([(rx << np.uint32(16)).view(np.float32), (rx & np.uint32(0xFFFF0000)).view(np.float32),
                        (ry << np.uint32(16)).view(np.float32), (ry & np.uint32(0xFFFF0000)).view(np.float32)], 1)
    assert np.array_equal(widened, v.float().numpy())
