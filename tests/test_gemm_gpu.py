"""tcgen05 TF32 GEMM (sdb_gemm_tf32) against an fp64 product of the TF32-rounded operands.

The kernel rounds both operands to the nearest TF32 value (cvt.rna) before the tensor core reads them; products are
exact in fp32 and accumulated in fp32, so against the fp64 product of the rounded operands only the accumulation
order differs (round_mode 0 -- the raw truncating tensor-core product -- is checked the same way).  Tolerance: 2e-5 of
the largest output magnitude (k <= 44448 fp32 accumulations), far inside the 1e-3 relative the north star asks of the
attention tensors; cuBLAS-TF32 (what the reference's nn.Linear runs on) is checked against the same bound."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trunc(t):
    """cvt.rna.tf32.f32: nearest TF32, ties away from zero"""
    return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _close(y, ref, tol=2e-5):
    scale = ref.abs().max().item() + 1e-30
    err = (y.double() - ref).abs().max().item() / scale
    assert err < tol, err


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (256, 256, 64), (1000, 384, 256), (77, 132, 36), (4446, 256, 256),
                                   (4446, 2048, 256), (4446, 256, 2048), (44446, 256, 256)])
@pytest.mark.parametrize("epilogue", ["plain", "bias", "bias_relu_mask"])
def test_forward(m, n, k, epilogue):
    from semi_detr_b200.layers import gemm as G
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    x = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.1
    b = torch.randn(n, device="cuda", generator=g) if epilogue != "plain" else None
    mask = (torch.rand(m, device="cuda", generator=g) < 0.2) if epilogue == "bias_relu_mask" else None
    y = G.linear_forward(x, w, b, relu=epilogue == "bias_relu_mask", row_mask=mask)
    ref = _trunc(x).double() @ _trunc(w).double().t()
    if b is not None:
        ref = ref + b.double()
    if epilogue == "bias_relu_mask":
        ref = torch.relu(ref).masked_fill(mask[:, None], 0.0)
        assert (y[mask] == 0).all()
    _close(y, ref)


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (1000, 256, 384), (4446, 256, 2048), (4446, 2048, 256), (44446, 256, 256),
                                   (900, 4, 256)])
def test_grad_input(m, n, k):
    from semi_detr_b200.layers import gemm as G
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    dy = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(k, n, device="cuda", generator=g) * 0.1
    _close(G.linear_grad_input(dy, w), _trunc(dy).double() @ _trunc(w).double())


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 256, 1000), (256, 256, 44446), (2048, 256, 4446), (256, 2048, 4446),
                                   (384, 256, 44448), (132, 36, 76)])
def test_grad_weight(m, n, k):
    from semi_detr_b200.layers import gemm as G
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    dy = torch.randn(k, m, device="cuda", generator=g)
    x = torch.randn(k, n, device="cuda", generator=g)
    _close(G.linear_grad_weight(dy, x), _trunc(dy).double().t() @ _trunc(x).double())


def test_split_product_accumulates_into_out():
    from semi_detr_b200.layers import gemm as G
    g = torch.Generator(device="cuda").manual_seed(5)
    dy = torch.randn(3000, 256, device="cuda", generator=g)
    x = torch.randn(3000, 128, device="cuda", generator=g)
    init = torch.randn(256, 128, device="cuda", generator=g)
    out = init.clone()
    G.gemm_tf32(dy, 1, x, 1, 256, 128, 3000, out=out, k_splits=7)
    _close(out, init.double() + _trunc(dy).double().t() @ _trunc(x).double())


def test_truncating_mode_is_the_raw_tensor_core_product():
    from semi_detr_b200.layers import gemm as G
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(1000, 256, device="cuda", generator=g)
    w = torch.randn(256, 256, device="cuda", generator=g)
    chop = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    y = G.gemm_tf32(x, 0, w, 0, 1000, 256, 256, round_mode=0)
    _close(y, chop(x).double() @ chop(w).double().t())


def test_rejects_cpu_and_bad_shapes():
    from semi_detr_b200.layers import gemm as G
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        G.linear_forward(torch.randn(8, 8), torch.randn(8, 8).cuda())
    with pytest.raises(RuntimeError, match="multiple of 4"):
        G.linear_forward(torch.randn(8, 8).cuda(), torch.randn(6, 8).cuda())
    with pytest.raises(RuntimeError, match="epilogue"):
        G.gemm_tf32(torch.randn(64, 8).cuda(), 1, torch.randn(64, 8).cuda(), 1, 8, 8, 64, bias=torch.randn(8).cuda(), k_splits=2)


def test_linear_layer_matches_torch_autograd(monkeypatch):
    monkeypatch.setenv("SDB_LINEAR", "tcgen05")
    """Linear (forward + all three gradients through the tcgen05 GEMM) against nn.Linear in fp32 (no TF32):
    1e-3 relative of the tensor's scale, the north star's bound."""
    from semi_detr_b200.layers.linear import Linear
    torch.manual_seed(0)
    lin = Linear(256, 384).cuda()
    ref = torch.nn.Linear(256, 384).cuda()
    ref.load_state_dict(lin.state_dict())
    x = torch.randn(2, 3000, 256, device="cuda", requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    gy = torch.randn(2, 3000, 384, device="cuda")
    from semi_detr_b200 import _lib
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        yr = ref(xr)
        yr.backward(gy)
        torch.backends.cuda.matmul.allow_tf32 = True      # the switch that selects the tcgen05 kernel
        before = _lib.LAUNCHES["gemm_tf32"]
        colsum_before = _lib.LAUNCHES["colsum"]
        y = lin(x)
        y.backward(gy)
        assert _lib.LAUNCHES["gemm_tf32"] - before == 3, "forward, grad-input and grad-weight run on the tcgen05 kernel"
        assert _lib.LAUNCHES["colsum"] == colsum_before, "the bias gradient comes out of the grad-weight launch"
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    for a, b in ((y, yr), (x.grad, xr.grad), (lin.weight.grad, ref.weight.grad), (lin.bias.grad, ref.bias.grad)):
        assert (a - b).abs().max().item() <= 1e-3 * b.abs().max().item()


def test_auto_policy_mixes_kernel_and_library(monkeypatch):
    """SDB_LINEAR=auto: masked value projection -> forward + grad-weight on the kernel, grad-input on the library;
    same numbers as the all-kernel policy to TF32 accuracy."""
    from semi_detr_b200 import _lib
    from semi_detr_b200.layers.linear import Linear, product_plan
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        monkeypatch.setenv("SDB_LINEAR", "auto")
        assert product_plan(256, 256, True) == (True, False, True) and product_plan(256, 256, False) == (False, False, True)
        assert product_plan(256, 2048, True) == (True, False, False) and product_plan(2048, 256, False) is None
        torch.manual_seed(1)
        lin = Linear(256, 256).cuda()
        x = torch.randn(2, 2000, 256, device="cuda", requires_grad=True)
        mask = torch.rand(2, 2000, device="cuda") < 0.3
        gy = torch.randn(2, 2000, 256, device="cuda")
        before = _lib.LAUNCHES["gemm_tf32"]
        y = lin(x, row_mask=mask)
        y.backward(gy)
        assert _lib.LAUNCHES["gemm_tf32"] - before == 2
        assert (y[mask] == 0).all()
        got = (y.detach().clone(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
        monkeypatch.setenv("SDB_LINEAR", "cublas")
        x.grad = None
        lin.zero_grad()
        y2 = lin(x, row_mask=mask)
        y2.backward(gy)
        for a, b in zip(got, (y2, x.grad, lin.weight.grad, lin.bias.grad)):
            assert (a - b).abs().max().item() <= 2e-3 * b.abs().max().item()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 256, 44446), (384, 256, 4448), (132, 36, 76), (2048, 256, 4446)])
def test_grad_weight_with_fused_bias_gradient(m, n, k):
    """The column sums of dy (bias gradient) taken by the grad-weight launch from the tiles it stages: exact fp32
    values (before the TF32 rounding), every column exactly once whatever the number of n-blocks and k-splits."""
    from semi_detr_b200.layers import gemm as G
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    dy = torch.randn(k, m, device="cuda", generator=g) + 0.25
    x = torch.randn(k, n, device="cuda", generator=g)
    gw, gb = G.linear_grad_weight(dy, x, with_bias_grad=True)
    _close(gw, _trunc(dy).double().t() @ _trunc(x).double())
    _close(gb, dy.double().sum(0), tol=1e-5)


@pytest.mark.parametrize("kind", ["fwd", "dx"])
def test_weights_in_tmem_variant(kind):
    """The opt-in variant that parks the weight block in tensor memory and computes y^T = W . x^T (TS-form
    tcgen05.mma, tcgen05.st, transposing epilogue): same result as the default path.  Runs in a subprocess because the
    variant is selected once per process (SDB_GEMM_WRES=1)."""
    import os
    import subprocess
    import sys
    code = f"""
import torch
from semi_detr_b200.layers import gemm as G
g = torch.Generator(device="cuda").manual_seed(3)
rna = lambda t: ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
m, n, k = 40000, 384, 256
x = torch.randn(m, k, device="cuda", generator=g)
if "{kind}" == "fwd":
    w = torch.randn(n, k, device="cuda", generator=g) * 0.1
    b = torch.randn(n, device="cuda", generator=g)
    mask = torch.rand(m, device="cuda", generator=g) < 0.2
    y = G.linear_forward(x, w, b, relu=True, row_mask=mask)
    ref = torch.relu(rna(x).double() @ rna(w).double().t() + b.double()).masked_fill(mask[:, None], 0.0)
else:
    w = torch.randn(k, n, device="cuda", generator=g) * 0.1
    y = G.linear_grad_input(x, w)
    ref = rna(x).double() @ rna(w).double()
err = float((y.double() - ref).abs().max() / ref.abs().max())
assert err < 2e-5, err
print("ok", err)
"""
    env = dict(os.environ, SDB_GEMM_WRES="1", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("rows,cols", [(44446, 2048), (2184, 2048), (1000, 256), (7, 4)])
def test_relu_backward_colsum(rows, cols):
    from semi_detr_b200.layers.linear import relu_backward_colsum
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    dy = torch.randn(rows, cols, device="cuda", generator=g)
    y = torch.relu(torch.randn(rows, cols, device="cuda", generator=g))
    got, gb = relu_backward_colsum(dy, y)
    want = torch.ops.aten.threshold_backward(dy, y, 0.0)
    assert torch.equal(got, want)
    ref = want.double().sum(0)
    assert float((gb.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-6


def test_ffn_layer_uses_fused_relu_backward(monkeypatch):
    """FFN linear1 + ReLU under the default policy: forward = the library GEMM with the ReLU in its epilogue (cuBLASLt is
    twice as fast as our kernel at this shape, tools/time_linear1.py), backward = ONE fused ReLU-backward + bias-gradient
    pass of this library + library products; same numbers as nn.Linear + relu."""
    from semi_detr_b200 import _lib
    from semi_detr_b200.layers.linear import Linear
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        monkeypatch.setenv("SDB_LINEAR", "auto")
        torch.manual_seed(2)
        lin = Linear(256, 2048).cuda()
        x = torch.randn(2, 1500, 256, device="cuda", requires_grad=True)
        gy = torch.randn(2, 1500, 2048, device="cuda")
        before = dict(_lib.LAUNCHES)
        y = lin(x, relu=True)
        y.backward(gy)
        assert _lib.LAUNCHES["gemm_tf32"] == before["gemm_tf32"]
        assert _lib.LAUNCHES["relu_backward_colsum"] - before["relu_backward_colsum"] == 1
        assert _lib.LAUNCHES["colsum"] == before["colsum"]
        xr = x.detach().clone().requires_grad_(True)
        yr = torch.relu(torch.nn.functional.linear(xr, lin.weight.detach(), lin.bias.detach()))
        w2 = lin.weight.detach().clone().requires_grad_(True)
        b2 = lin.bias.detach().clone().requires_grad_(True)
        yr = torch.relu(torch.nn.functional.linear(xr, w2, b2))
        yr.backward(gy)
        # hidden units within TF32 noise of zero may sit on different sides of the ReLU kink in the two
        # implementations, so compare in the Frobenius norm rather than element by element
        for a, b in ((y, yr), (x.grad, xr.grad), (lin.weight.grad, w2.grad), (lin.bias.grad, b2.grad)):
            assert float((a - b).norm() / b.norm()) <= 5e-3
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("kind", ["fwd", "dx"])
def test_cta_pair_variant(kind):
    """The opt-in cta_group::2 kernel (one 256 x 256 tile per CTA pair, leader-issued tcgen05.mma.cta_group::2,
    multicast commit, remote barrier arrives): same result as the default path, ragged shapes included.  Subprocess:
    the variant is selected once per process (SDB_GEMM_2CTA=1)."""
    import os
    import subprocess
    import sys
    code = f"""
import torch
from semi_detr_b200.layers import gemm as G
g = torch.Generator(device="cuda").manual_seed(4)
rna = lambda t: ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
for m, n, k in ((77, 132, 36), (1000, 384, 256), (20000, 256, 512)):
    x = torch.randn(m, k, device="cuda", generator=g)
    if "{kind}" == "fwd":
        w = torch.randn(n, k, device="cuda", generator=g) * 0.1
        b = torch.randn(n, device="cuda", generator=g)
        mask = torch.rand(m, device="cuda", generator=g) < 0.2
        y = G.linear_forward(x, w, b, relu=True, row_mask=mask)
        ref = torch.relu(rna(x).double() @ rna(w).double().t() + b.double()).masked_fill(mask[:, None], 0.0)
    else:
        w = torch.randn(k, n, device="cuda", generator=g) * 0.1
        y = G.linear_grad_input(x, w)
        ref = rna(x).double() @ rna(w).double()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, (m, n, k, err)
print("ok")
"""
    env = dict(os.environ, SDB_GEMM_2CTA="1", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("rows,cols", [(178046, 256), (5000, 2048), (1024, 384), (9, 4)])
def test_bf16_column_sum_kernels(rows, cols):
    """sdb_colsum_bf16 / sdb_relu_backward_colsum_bf16: bf16 storage, fp32 sums.  The masked gradient is a copy of bf16
    inputs (bit-equal); the sums against float64 sums of the same bf16 values."""
    from semi_detr_b200.layers.linear import _colsum_bf16, _relu_backward_colsum_bf16
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    dy = torch.randn(rows, cols, device="cuda", generator=g).to(torch.bfloat16)
    y = torch.relu(torch.randn(rows, cols, device="cuda", generator=g)).to(torch.bfloat16)
    s = _colsum_bf16(dy)
    ref = dy.double().sum(0)
    assert s.dtype == torch.float32
    assert float((s.double() - ref).abs().max()) <= 2e-5 * float(dy.double().abs().sum(0).max()) + 1e-6
    got, gb = _relu_backward_colsum_bf16(dy, y)
    want = torch.where(y > 0, dy, torch.zeros_like(dy))
    assert torch.equal(got, want)
    ref = want.double().sum(0)
    assert float((gb.double() - ref).abs().max()) <= 2e-5 * float(want.double().abs().sum(0).max()) + 1e-6


def test_autocast_linear_matches_plain_autocast():
    """`Linear` under bf16 autocast (library bf16 GEMMs + this library's bf16 bias-gradient / ReLU-backward pass) against
    torch's own autocast of the same layer: same bf16 products, so outputs are equal and gradients agree to bf16
    rounding of the bias-gradient sum (ours is summed in fp32)."""
    from semi_detr_b200 import _lib
    from semi_detr_b200.layers.linear import Linear
    torch.manual_seed(4)
    for relu in (False, True):
        lin = Linear(256, 512).cuda()
        x = torch.randn(3, 700, 256, device="cuda", requires_grad=True)
        gy = torch.randn(3, 700, 512, device="cuda")
        before = _lib.LAUNCHES["colsum"] + _lib.LAUNCHES["relu_backward_colsum"]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = lin(x, relu=relu)
        y.backward(gy.to(y.dtype))
        assert _lib.LAUNCHES["colsum"] + _lib.LAUNCHES["relu_backward_colsum"] == before + 1
        got = y.detach().float(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()
        x2 = x.detach().clone().requires_grad_(True)
        w2, b2 = lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y2 = torch.nn.functional.linear(x2, w2, b2)
            if relu:
                y2 = torch.relu(y2)
        y2.backward(gy.to(y2.dtype))
        assert y.dtype == torch.bfloat16 and torch.equal(got[0], y2.detach().float())
        for a, b, tol in ((got[1], x2.grad, 1e-2), (got[2], w2.grad, 1e-2), (got[3], b2.grad, 2e-2)):
            assert a.dtype == b.dtype and float((a - b).abs().max()) <= tol * float(b.abs().max())


@pytest.mark.parametrize("m,n,k", [(4446, 2048, 256), (44446, 2048, 256), (2184, 2048, 256), (300, 1000, 256), (4999, 132, 64),
                                   (129, 36, 256)])
@pytest.mark.parametrize("round_mode", [3, 2])
def test_grad_input_with_relu_backward_epilogue(m, n, k, round_mode):
    """sdb_gemm_tf32_relu_grad: (dy W) * (h > 0) and its column sums -- the resident-weight variant (first two shapes),
    the streaming variant, ragged row / column tiles.  round_mode 2 leaves the gradient operand to the tensor core's
    own truncation (what cuBLAS-TF32 does with both operands)."""
    from semi_detr_b200.layers import gemm as G
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    dy = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(k, n, device="cuda", generator=g) * 0.1
    h = torch.relu(torch.randn(m, n, device="cuda", generator=g))
    y, sums = G.linear_grad_input_relu(dy, w, h, round_mode=round_mode)
    a = _trunc(dy) if round_mode & 1 else (dy.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref = (a.double() @ _trunc(w).double()) * (h > 0)
    assert (y[h <= 0] == 0).all()
    _close(y, ref)
    # column sums: fp32 atomics over m / 32 partial sums; compared on the scale of the column's absolute sum
    ref_sums = y.double().sum(0)
    scale = y.double().abs().sum(0).max().item() + 1e-30
    assert ((sums.double() - ref_sums).abs().max().item() / scale) < 1e-5


def test_post_attention_block_matches_layer_by_layer_route(monkeypatch):
    """layers/ffn.py: norm_a(a + r) -> FFN -> norm_b as one autograd node against the layer-by-layer modules (the route
    SDB_FFN_BLOCK=0 takes): outputs identical (same forward kernels), every gradient within TF32 product error."""
    from semi_detr_b200.layers import LayerNorm, Linear, ffn
    torch.manual_seed(0)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        l1, l2 = Linear(256, 2048).cuda(), Linear(2048, 256).cuda()
        na, nb = LayerNorm(256).cuda(), LayerNorm(256).cuda()
        with torch.no_grad():
            for n_ in (na, nb):
                n_.weight.add_(torch.randn_like(n_.weight) * 0.1)
                n_.bias.add_(torch.randn_like(n_.bias) * 0.1)
        mods = (l1, l2, na, nb)
        a0 = torch.randn(2, 3000, 256, device="cuda")
        r0 = torch.randn(2, 3000, 256, device="cuda")
        pos = torch.randn(2, 3000, 256, device="cuda")
        gy, gq = torch.randn_like(a0), torch.randn_like(a0)
        drops = (torch.nn.Dropout(0.0),)

        def run(fused, with_pos):
            a, r = a0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
            p = pos.clone().requires_grad_(True) if with_pos else None
            for m_ in mods:
                m_.zero_grad()
            if fused:
                assert ffn.fused_ok(a, na, l1, l2, nb, drops)
                out = ffn.post_attention_block(a, r, na, l1, l2, nb, p)
            else:
                x = na.add_norm(a, r)
                out = nb.add_norm(x, l2(l1(x, relu=True)), p)
            if with_pos:
                (out[0] * gy).sum().add((out[1] * gq).sum()).backward()
                outs = [o.detach() for o in out]
            else:
                (out * gy).sum().backward()
                outs = [out.detach()]
            grads = [a.grad, r.grad] + [p_.grad.clone() for m_ in mods for p_ in m_.parameters()]
            if with_pos:
                grads.append(p.grad)
            return outs, grads

        for with_pos in (False, True):
            o_f, g_f = run(True, with_pos)
            o_r, g_r = run(False, with_pos)
            for x_, y_ in zip(o_f, o_r):
                assert torch.equal(x_, y_)
            for x_, y_ in zip(g_f, g_r):
                scale = y_.abs().max().item() + 1e-30
                assert (x_ - y_).abs().max().item() / scale < 2e-3, (x_.shape, (x_ - y_).abs().max().item() / scale)
        monkeypatch.setenv("SDB_FFN_BLOCK", "0")
        assert not ffn.fused_ok(a0.requires_grad_(True), na, l1, l2, nb, ())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
