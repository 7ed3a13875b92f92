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
