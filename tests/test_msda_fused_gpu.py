"""Fused-prologue MSDA (softmax + location arithmetic inside the kernels) against the unfused path and autograd
through the module's own torch arithmetic."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _inputs(levels, N, Lq, ref_dim, seed):
    from semi_detr_b200.synthetic import level_tensors
    g = torch.Generator().manual_seed(seed)
    L, M, D, P = len(levels), 8, 32, 4
    S = sum(h * w for h, w in levels)
    shapes, start = level_tensors(levels, "cuda")
    value = torch.randn(N, S, M, D, generator=g).cuda()
    if ref_dim == 2:
        ref = torch.rand(N, Lq, L, 2, generator=g).cuda()
        off = (torch.randn(N, Lq, M, L, P, 2, generator=g) * 3).cuda()
    else:
        ref = torch.cat([torch.rand(N, Lq, L, 2, generator=g), torch.rand(N, Lq, L, 2, generator=g) * 0.4 + 0.02], -1).cuda()
        off = (torch.randn(N, Lq, M, L, P, 2, generator=g) * 2).cuda()
    logits = (torch.randn(N, Lq, M, L * P, generator=g) * 2).cuda()
    gout = torch.randn(N, Lq, M * D, generator=g).cuda()
    return value, shapes, start, ref, off, logits, gout


def _unfused(value, shapes, start, ref, off, logits, P=4):
    from semi_detr_b200.msda import MSDeformAttnFunction
    N, Lq, M, L = off.shape[:4]
    w = F.softmax(logits, -1).view(N, Lq, M, L, P)
    if ref.shape[-1] == 2:
        wh = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        loc = ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + off / P * ref[:, :, None, :, None, 2:] * 0.5
    return MSDeformAttnFunction.apply(value, shapes, start, loc.contiguous(), w, 64)


@pytest.mark.parametrize("levels,Lq,ref_dim", [([(19, 27), (10, 14), (5, 7), (3, 4)], None, 2),
                                               ([(19, 27), (10, 14), (5, 7), (3, 4)], 211, 4),
                                               ([(9, 8), (4, 5)], 37, 4), ([(6, 5)], None, 2)])
def test_fused_matches_unfused(levels, Lq, ref_dim):
    from semi_detr_b200.msda.functions import MSDeformAttnFusedFunction
    S = sum(h * w for h, w in levels)
    value, shapes, start, ref, off, logits, gout = _inputs(levels, 2, Lq or S, ref_dim, seed=len(levels) + ref_dim)
    leaves = [t.clone().requires_grad_(True) for t in (value, off, logits)]
    out = MSDeformAttnFusedFunction.apply(leaves[0], shapes, start, ref, leaves[1], leaves[2])
    out.backward(gout)
    ref_leaves = [t.clone().requires_grad_(True) for t in (value, off, logits)]
    want = _unfused(ref_leaves[0], shapes, start, ref, ref_leaves[1], ref_leaves[2])
    want.backward(gout)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
    assert rel(out, want) < 1e-5
    assert torch.allclose(out, want, rtol=1e-3, atol=1e-4)
    for a, b, name in zip(leaves, ref_leaves, ("value", "offsets", "logits")):
        assert rel(a.grad, b.grad) < 2e-4, name


def test_module_uses_fused_path_and_matches_unfused_module():
    from semi_detr_b200 import _lib
    from semi_detr_b200.msda import MSDeformAttn
    from semi_detr_b200.synthetic import level_tensors
    torch.manual_seed(0)
    levels = [(20, 27), (10, 14), (5, 7), (3, 4)]
    S = sum(h * w for h, w in levels)
    shapes, start = level_tensors(levels, "cuda")
    mod = MSDeformAttn().cuda()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.05)
        mod.sampling_offsets.bias.add_(torch.randn_like(mod.sampling_offsets.bias) * 0.3)
    q = torch.randn(2, S, 256, device="cuda")
    src = torch.randn(2, S, 256, device="cuda", requires_grad=True)
    ref = torch.rand(2, S, 4, 2, device="cuda")
    mask = torch.zeros(2, S, dtype=torch.bool, device="cuda")
    mask[1, -50:] = True
    fused_launches = lambda: _lib.LAUNCHES["msda_fused_forward"] + _lib.LAUNCHES["msda_forward_tma"]
    before = fused_launches()
    y1 = mod(q, ref, src, shapes, start, mask)
    assert fused_launches() == before + 1          # one fused-prologue kernel (L1-gather or TMA-staged variant)
    g1 = torch.autograd.grad(y1.square().sum(), [src] + list(mod.parameters()))
    mod.fused_prologue = False
    y2 = mod(q, ref, src, shapes, start, mask)
    g2 = torch.autograd.grad(y2.square().sum(), [src] + list(mod.parameters()))
    assert torch.allclose(y1, y2, rtol=1e-4, atol=1e-5)
    for a, b in zip(g1, g2):
        assert ((a - b).norm() / b.norm().clamp_min(1e-20)) < 1e-3


# ---- TMA-staged forward (encoder self-attention) ---------------------------------------------------------

@pytest.mark.parametrize("levels", [[(19, 27), (10, 14), (5, 7), (3, 4)], [(100, 134), (50, 67), (25, 34), (13, 17)],
                                    [(9, 8), (4, 5)], [(6, 5)], [(33, 9), (17, 5), (9, 3), (5, 2)]])
@pytest.mark.parametrize("mode", ["encoder", "wide"])
def test_tma_forward_matches_l1_kernel(levels, mode):
    """Same inputs through the TMA-staged kernel and the L1-gather kernel; 'wide' locations leave the staged boxes
    (and the image), exercising the global fallback, the zero fill and the out-of-range test."""
    from semi_detr_b200 import _lib
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    from semi_detr_b200.synthetic import msda_inputs
    S = sum(h * w for h, w in levels)
    x = msda_inputs(levels, N=2, mode="encoder", seed=7)
    if mode == "wide":
        g = torch.Generator().manual_seed(1)
        x["loc"] = (torch.rand(x["loc"].shape, generator=g) * 1.6 - 0.3).cuda()
    a = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
    before = _lib.LAUNCHES["msda_forward_tma"]
    MSDA.USE_TMA = True
    try:
        got = MSDA.ms_deform_attn_forward(*a, 64)
    finally:
        MSDA.USE_TMA = False
    assert _lib.LAUNCHES["msda_forward_tma"] == before + 1
    want = MSDA.ms_deform_attn_forward(*a, 64)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), (got - want).abs().max()


def test_tma_fused_forward_matches_fused_l1_kernel():
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    levels = [(40, 54), (20, 27), (10, 14), (5, 7)]
    S = sum(h * w for h, w in levels)
    value, shapes, start, ref, off, logits, _ = _inputs(levels, 2, S, 2, seed=3)
    from semi_detr_b200.synthetic import encoder_reference_points
    ref = encoder_reference_points(levels, "cuda")[None, :, None, :].expand(2, S, 4, 2).contiguous()
    MSDA.USE_TMA = True
    try:
        got = MSDA.ms_deform_attn_fused_forward(value, shapes, start, ref, off, logits)
    finally:
        MSDA.USE_TMA = False
    want = MSDA.ms_deform_attn_fused_forward(value, shapes, start, ref, off, logits)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_standalone_prologue_matches_the_module_arithmetic(ref_dim, dtype):
    """`MSDAPrologueFunction` (5 levels x 4 points: the fused MSDA kernels do not apply) against the module's own tensor
    arithmetic (ms_deform_attn.py:98-112) in float32 on the same -- for bf16, already rounded -- raw tensors: locations
    and softmax weights to 1e-6, and the gradients of the raw offsets / logits to fp32 (bf16: one bf16 ulp) accuracy."""
    from semi_detr_b200.msda.functions import MSDAPrologueFunction
    levels = [(40, 52), (20, 26), (10, 13), (5, 7), (3, 4)]
    N, Lq, M, L, P = 2, 333, 8, 5, 4
    g = torch.Generator(device="cuda").manual_seed(ref_dim)
    off = (torch.randn(N, Lq, M, L, P, 2, device="cuda", generator=g) * 3).to(dtype)
    lg = torch.randn(N, Lq, M, L * P, device="cuda", generator=g).to(dtype)
    ref = torch.rand(N, Lq, L, ref_dim, device="cuda", generator=g) * 0.8 + 0.1
    gl = torch.randn(N, Lq, M, L, P, 2, device="cuda", generator=g)
    ga = torch.randn(N, Lq, M, L, P, device="cuda", generator=g)
    o1, l1 = off.clone().requires_grad_(True), lg.clone().requires_grad_(True)
    loc, w = MSDAPrologueFunction.apply(o1, l1, ref, tuple(levels))
    (loc * gl).sum().backward(retain_graph=True)
    (w * ga).sum().backward()
    o2, l2 = off.float().clone().requires_grad_(True), lg.float().clone().requires_grad_(True)
    w2 = torch.softmax(l2, -1).view(N, Lq, M, L, P)
    r = ref[:, :, None, :, None, :]
    if ref_dim == 2:
        wh = torch.tensor([[wd, ht] for ht, wd in levels], dtype=torch.float32, device="cuda")
        loc2 = r + o2 / wh[None, None, None, :, None, :]
    else:
        loc2 = r[..., :2] + o2 / P * r[..., 2:] * 0.5
    ((loc2 * gl).sum() + (w2 * ga).sum()).backward()
    assert loc.dtype == torch.float32 and w.dtype == torch.float32
    assert float((loc - loc2).abs().max()) < 2e-6 and float((w - w2).abs().max()) < 1e-6
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert o1.grad.dtype == dtype and l1.grad.dtype == dtype
    assert float((o1.grad.float() - o2.grad).abs().max()) <= tol * float(o2.grad.abs().max())
    assert float((l1.grad.float() - l2.grad).abs().max()) <= tol * float(l2.grad.abs().max()) + 1e-6
