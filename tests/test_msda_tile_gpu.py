"""The tile-combining encoder backward (csrc/msda_backward_tile.cu, the default when num_query == spatial_size)
against the numpy oracle at small sizes and against the per-corner reduction kernel at the train-step size.

Cases walk every path of the kernel: points inside the per-level windows (bucketed, one reduction per touched pixel),
points outside them (direct reductions: far offsets, uniform locations), out-of-range samples, border tiles, levels
smaller than a tile, 1..5 levels (16- and 32-slot builds), several images, and the fused-prologue form."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _smooth_mask(loc, levels):
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float64)[None, None, None, :, None, :]
    px = loc.double().cpu() * wh - 0.5
    near = (px - px.round()).abs() < 1e-3
    return ~(near.any(-1, keepdim=True).expand_as(px))


def _encoder_inputs(levels, N, spread, seed, frac_far=0.0):
    """loc = own grid point + U(-spread, spread) px on every level; `frac_far` of the points are U(-0.2, 1.2) instead."""
    from semi_detr_b200.synthetic import encoder_reference_points, level_tensors
    g = torch.Generator().manual_seed(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    shapes, start = level_tensors(levels, "cuda")
    ref = encoder_reference_points(levels, "cpu")
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)
    off = (torch.rand(N, S, 8, L, 4, 2, generator=g) * 2 - 1) * spread / wh[None, None, None, :, None, :]
    loc = ref[None, :, None, None, None, :] + off
    if frac_far > 0:
        far = torch.rand(N, S, 8, L, 4, 1, generator=g) < frac_far
        loc = torch.where(far, torch.rand(N, S, 8, L, 4, 2, generator=g) * 1.4 - 0.2, loc)
    attn = torch.softmax(torch.randn(N, S, 8, L * 4, generator=g), -1).view(N, S, 8, L, 4)
    return dict(value=torch.randn(N, S, 8, 32, generator=g).cuda(), shapes=shapes, start=start,
                loc=loc.cuda().contiguous(), attn=attn.cuda().contiguous(), gout=torch.randn(N, S, 256, generator=g).cuda())


def _check_vs_oracle(x, levels):
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    gv, gl, ga = O.msda_backward(x["value"].cpu().numpy(), levels, x["start"].cpu().numpy(), x["loc"].cpu().numpy(),
                                 x["attn"].cpu().numpy(), x["gout"].cpu().numpy())
    got = MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["gout"], 64)
    mask = _smooth_mask(x["loc"], levels)
    for g, r, k in zip(got, (gv, gl, ga), ("grad_value", "grad_loc", "grad_attn")):
        g, r = g.cpu(), torch.from_numpy(r)
        if k == "grad_loc":
            g, r = g * mask, r * mask
        assert _relerr(g, r) < 1e-5, k
        np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=RTOL, atol=2e-3, err_msg=k)


CASES = {
    "four_levels_near": ([(19, 27), (10, 14), (5, 7), (3, 4)], 2, 4.0, 0.0),
    "four_levels_mixed_far": ([(19, 27), (10, 14), (5, 7), (3, 4)], 2, 4.0, 0.3),
    "beyond_the_halo": ([(17, 23), (9, 12), (5, 6), (3, 3)], 1, 9.0, 0.0),
    "all_far": ([(16, 16), (8, 8)], 1, 4.0, 1.0),
    "one_level": ([(21, 13)], 2, 3.0, 0.05),
    "two_levels": ([(24, 31), (12, 16)], 1, 4.0, 0.1),
    "three_levels_three_images": ([(18, 18), (9, 9), (5, 5)], 3, 4.0, 0.1),
    "five_levels": ([(33, 40), (17, 20), (9, 10), (5, 5), (3, 3)], 1, 4.0, 0.1),
    "tiny_levels": ([(5, 3), (3, 2), (2, 1), (1, 1)], 2, 2.0, 0.2),
    "exact_tiles": ([(16, 24), (8, 16)], 1, 4.0, 0.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_tile_backward_vs_oracle(name):
    levels, N, spread, far = CASES[name]
    _check_vs_oracle(_encoder_inputs(levels, N, spread, seed=len(name), frac_far=far), levels)


def test_tile_backward_integer_offsets_on_the_pixel_lattice():
    """The benchmark's random-init model samples at integer pixel offsets (lh = lw = 0: the right / lower corners carry
    weight exactly 0).  Sums must still come out exact up to fp32 order."""
    from semi_detr_b200.synthetic import encoder_reference_points, level_tensors
    levels = [(20, 28), (10, 14), (5, 7), (3, 4)]
    S = sum(h * w for h, w in levels)
    g = torch.Generator().manual_seed(7)
    shapes, start = level_tensors(levels, "cuda")
    ref = encoder_reference_points(levels, "cpu")
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)
    dirs = torch.tensor([[1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1], [1, -1]], dtype=torch.float32)
    off = dirs[None, None, :, None, None, :] * torch.arange(1, 5, dtype=torch.float32)[None, None, None, None, :, None]
    loc = (ref[None, :, None, None, None, :] + off / wh[None, None, None, :, None, :]).expand(2, S, 8, 4, 4, 2)
    x = dict(value=torch.randn(2, S, 8, 32, generator=g).cuda(), shapes=shapes, start=start, loc=loc.contiguous().cuda(),
             attn=torch.softmax(torch.randn(2, S, 8, 16, generator=g), -1).view(2, S, 8, 4, 4).cuda(),
             gout=torch.randn(2, S, 256, generator=g).cuda())
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    gv, gl, ga = O.msda_backward(x["value"].cpu().numpy(), levels, x["start"].cpu().numpy(), x["loc"].cpu().numpy(),
                                 x["attn"].cpu().numpy(), x["gout"].cpu().numpy())
    got = MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["gout"], 64)
    # on the lattice the fp32 coordinate may round to either side of the integer: value and attention gradients are
    # continuous there, the location gradient is not (see _smooth_mask) and is left out
    assert _relerr(got[0].cpu(), torch.from_numpy(gv)) < 1e-4
    assert _relerr(got[2].cpu(), torch.from_numpy(ga)) < 1e-4


@pytest.mark.parametrize("shape", ["train_step", "microbench"])
def test_tile_backward_equals_per_corner_kernel_at_full_size(shape):
    """N=2 at the 800x1333 (S = 22 223) and microbench (S = 17 821) level pyramids: the tile kernel and the validated
    per-corner reduction kernel (variant 5) agree; grad_loc / grad_attn bit-for-bit apart from summation order."""
    from semi_detr_b200 import _lib
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    from semi_detr_b200.synthetic import COCO_4SCALE_LEVELS, MICROBENCH_LEVELS, msda_inputs
    levels = COCO_4SCALE_LEVELS if shape == "train_step" else MICROBENCH_LEVELS
    x = msda_inputs(levels, N=2, mode="encoder", seed=1)
    a = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["gout"], 64)
    new = MSDA.ms_deform_attn_backward(*a)
    _lib.lib().sdb_msda_set_variant(0, 5)
    try:
        old = MSDA.ms_deform_attn_backward(*a)
    finally:
        _lib.lib().sdb_msda_set_variant(0, 0)
    for g, r, k in zip(new, old, ("grad_value", "grad_loc", "grad_attn")):
        assert _relerr(g, r) < 2e-6, k
        assert float((g - r).abs().max()) <= 1e-3 * float(r.abs().max()), k
    # linearity in grad_output (a size-independent property): backward(2 g) == 2 backward(g)
    twice = MSDA.ms_deform_attn_backward(*a[:5], 2 * x["gout"], 64)
    for g, t in zip(new, twice):
        assert _relerr(t, 2 * g) < 2e-6


def test_fused_tile_backward_vs_unfused_chain():
    """Fused-prologue form (raw offsets / logits + reference points): gradients equal the unfused op chained through
    autograd's softmax and location arithmetic."""
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    from semi_detr_b200.synthetic import encoder_reference_points, level_tensors
    levels = [(19, 27), (10, 14), (5, 7), (3, 4)]
    S = sum(h * w for h, w in levels)
    g = torch.Generator().manual_seed(11)
    shapes, start = level_tensors(levels, "cuda")
    ref = encoder_reference_points(levels, "cpu")[None, :, None, :].expand(2, S, 4, 2).contiguous().cuda()
    value = torch.randn(2, S, 8, 32, generator=g).cuda()
    offs = ((torch.rand(2, S, 8, 4, 4, 2, generator=g) * 2 - 1) * 4.5).cuda().requires_grad_(True)
    logits = torch.randn(2, S, 8, 16, generator=g).cuda().requires_grad_(True)
    gout = torch.randn(2, S, 256, generator=g).cuda()
    gv, goff, glog = MSDA.ms_deform_attn_fused_backward(value, shapes, start, ref, offs.detach(), logits.detach(), gout)
    # unfused chain
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32, device="cuda")
    loc = ref[:, :, None, :, None, :] + offs / wh[None, None, None, :, None, :]
    attn = torch.softmax(logits, -1).view(2, S, 8, 4, 4)
    from semi_detr_b200 import _lib
    _lib.lib().sdb_msda_set_variant(0, 5)
    try:
        rgv, rgl, rga = MSDA.ms_deform_attn_backward(value, shapes, start, loc.detach().contiguous(),
                                                     attn.detach().contiguous(), gout, 64)
    finally:
        _lib.lib().sdb_msda_set_variant(0, 0)
    torch.autograd.backward([loc, attn], [rgl, rga])
    assert _relerr(gv, rgv) < 1e-5
    assert _relerr(goff, offs.grad) < 1e-5
    assert _relerr(glog, logits.grad) < 1e-4
