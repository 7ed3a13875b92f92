"""LayerNorm kernels against torch's fp64 LayerNorm (forward, input gradient, affine-parameter gradients)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 256), (7, 256), (2, 22223, 256), (1100, 2, 256), (3, 0, 256)])
def test_layernorm_matches_torch(shape):
    from semi_detr_b200.layers import LayerNorm
    torch.manual_seed(0)
    ln = LayerNorm(256).cuda()
    with torch.no_grad():
        ln.weight.copy_(torch.randn(256) * 0.5 + 1)
        ln.bias.copy_(torch.randn(256) * 0.1)
    x = (torch.randn(*shape, device="cuda") * 3 + 0.7).requires_grad_(True)
    gout = torch.randn(*shape, device="cuda")
    y = ln(x)
    y.backward(gout)
    xd = x.detach().double().requires_grad_(True)
    wd = ln.weight.detach().double().requires_grad_(True)
    bd = ln.bias.detach().double().requires_grad_(True)
    yd = F.layer_norm(xd, (256,), wd, bd, ln.eps)
    yd.backward(gout.double())
    assert y.shape == x.shape
    if x.numel() == 0:
        assert not ln.weight.grad.any() and not ln.bias.grad.any()
        return
    assert torch.allclose(y.double(), yd, rtol=1e-5, atol=1e-5)
    assert torch.allclose(x.grad.double(), xd.grad, rtol=1e-4, atol=1e-5)
    rel = lambda a, b: ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert rel(ln.weight.grad, wd.grad) < 1e-5 and rel(ln.bias.grad, bd.grad) < 1e-5


def test_other_widths_use_library_layernorm():
    from semi_detr_b200.layers import LayerNorm
    ln = LayerNorm(64).cuda()
    x = torch.randn(5, 64, device="cuda")
    assert torch.allclose(ln(x), F.layer_norm(x, (64,), ln.weight, ln.bias, ln.eps))


@pytest.mark.parametrize("rows,cols", [(1, 128), (777, 256), (44446, 2048), (44446, 256), (2200, 2048), (5, 4)])
def test_column_sum_matches_torch(rows, cols):
    from semi_detr_b200.layers import column_sum
    torch.manual_seed(rows + cols)
    x = torch.randn(rows, cols, device="cuda")
    got, want = column_sum(x), x.double().sum(0)
    assert torch.allclose(got.double(), want, rtol=1e-4, atol=1e-3 * (rows ** 0.5))


def test_linear_backward_matches_nn_linear():
    from semi_detr_b200.layers import Linear
    torch.manual_seed(0)
    a, b = Linear(256, 2048).cuda(), torch.nn.Linear(256, 2048).cuda()
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 9000, 256, device="cuda", requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    g = torch.randn(2, 9000, 2048, device="cuda")
    a(x).backward(g)
    b(x2).backward(g)
    rel = lambda u, v: ((u - v).norm() / v.norm()).item()
    assert rel(x.grad, x2.grad) < 1e-5 and rel(a.weight.grad, b.weight.grad) < 1e-5
    assert rel(a.bias.grad, b.bias.grad) < 1e-5


@pytest.mark.parametrize("with_pos", [False, True])
def test_add_norm_matches_layer_norm_of_the_sum(with_pos):
    """``LayerNorm.add_norm(x, r[, pos])`` = LN(x + r) [and LN(x + r) + pos] in one kernel: values and every gradient
    (x, r, gamma, beta, pos) against the library expression; fp32 summation-order differences only."""
    from semi_detr_b200.layers.layernorm import LayerNorm
    torch.manual_seed(3)
    ln = LayerNorm(256).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
    x = torch.randn(2, 1301, 256, device="cuda", requires_grad=True)
    r = (torch.randn(2, 1301, 256, device="cuda") * 0.7).requires_grad_(True)
    pos = torch.randn(2, 1301, 256, device="cuda", requires_grad=True) if with_pos else None
    gy, gq = torch.randn(2, 1301, 256, device="cuda"), torch.randn(2, 1301, 256, device="cuda")
    out = ln.add_norm(x, r, pos)
    loss = ((out[0] * gy).sum() + (out[1] * gq).sum()) if with_pos else (out * gy).sum()
    got = torch.autograd.grad(loss, [x, r, ln.weight, ln.bias] + ([pos] if with_pos else []))
    want_y = torch.nn.functional.layer_norm(x + r, (256,), ln.weight, ln.bias, ln.eps)
    want_loss = ((want_y * gy).sum() + ((want_y + pos) * gq).sum()) if with_pos else (want_y * gy).sum()
    want = torch.autograd.grad(want_loss, [x, r, ln.weight, ln.bias] + ([pos] if with_pos else []))
    y = out[0] if with_pos else out
    assert torch.allclose(y, want_y, rtol=1e-5, atol=1e-5)
    if with_pos:
        assert torch.allclose(out[1], want_y + pos, rtol=1e-5, atol=1e-5)
    for g, w, name in zip(got, want, ("x", "residual", "gamma", "beta", "pos")):
        assert float((g - w).abs().max()) <= 2e-5 * float(w.abs().max()) + 1e-6, name
    assert got[0].data_ptr() == got[1].data_ptr() or torch.equal(got[0], got[1])
