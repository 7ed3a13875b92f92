"""Experimental MSDA backward variants against the oracle and against the validated default kernel:

  7       4 lanes x 8 channels per (query, head) pair (csrc/msda_backward_x8.cu): 0.56x the instructions per pair
  8       the same with the loop over a chunk's 4 batches kept rolled (29 KB of code instead of 60 KB)
  10, 11  the 8-lane kernel with the corner loads of 2 points in flight per warp (4 / 3 CTAs per SM)
  12      ... of 4 points in flight (3 CTAs per SM)
  15, 16  8 x 8 pixel tiles (the forward's tile shape) with 256 threads, 1 / 2 points in flight

Motivation (profiles/ncu_msda_stalls_r1.txt, ncu source view of the encoder backward): the default kernel has 4 loads
in flight per warp and waits one L2 latency per point (32 % of the stall samples sit on the first FMUL after each
point's loads), issues 46 % of its slots, and misses the instruction cache (hit rate 85.7 %).

NOT YET RUN ON HARDWARE (written after round 1's GPU budget was spent); runs only with SDB_RUN_UNVALIDATED=1:

    SDB_RUN_UNVALIDATED=1 python -m pytest tests/test_msda_bwd_experimental_gpu.py -m gpu -q && python tools/bwd_variants.py
"""
import os

import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SDB_RUN_UNVALIDATED") != "1",
                                 reason="experimental backward variants have not run on hardware yet "
                                        "(set SDB_RUN_UNVALIDATED=1)")]

VARIANTS = [7, 8, 10, 11, 12, 15, 16]


def _relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _smooth_mask(loc, levels):
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float64)[None, None, None, :, None, :]
    px = loc.double().cpu() * wh - 0.5
    near = (px - px.round()).abs() < 1e-3
    return ~(near.any(-1, keepdim=True).expand_as(px))


class _variant:
    def __init__(self, bwd):
        self.bwd = bwd

    def __enter__(self):
        from semi_detr_b200 import _lib
        _lib.lib().sdb_msda_set_variant(0, self.bwd)

    def __exit__(self, *exc):
        from semi_detr_b200 import _lib
        _lib.lib().sdb_msda_set_variant(0, 0)


@pytest.mark.parametrize("levels", [[(19, 27), (10, 14), (5, 7), (3, 4)], [(38, 54), (19, 27), (10, 14), (5, 7), (3, 4)],
                                    [(9, 8), (4, 5)], [(6, 5)]], ids=["4lvl", "5lvl", "2lvl", "1lvl"])
@pytest.mark.parametrize("mode,Lq", [("encoder", None), ("wide", 300), ("uniform", 77)])
@pytest.mark.parametrize("variant", VARIANTS)
def test_unfused_vs_oracle_and_default_kernel(variant, levels, mode, Lq):
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    from semi_detr_b200.synthetic import msda_inputs
    x = msda_inputs(levels, N=2, Lq=Lq, mode=mode, seed=21)
    args = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["gout"], 64)
    with _variant(variant):
        got = MSDA.ms_deform_attn_backward(*args)
    base = MSDA.ms_deform_attn_backward(*args)
    gv, gl, ga = O.msda_backward(x["value"].cpu().numpy(), levels, x["start"].cpu().numpy(), x["loc"].cpu().numpy(),
                                 x["attn"].cpu().numpy(), x["gout"].cpu().numpy())
    mask = _smooth_mask(x["loc"], levels)
    for g, b, r, k in zip(got, base, (gv, gl, ga), ("grad_value", "grad_loc", "grad_attn")):
        g, b, r = g.cpu(), b.cpu(), torch.from_numpy(r)
        if k == "grad_loc":
            g, b, r = g * mask, b * mask, r * mask
        assert _relerr(g, r) < 1e-5, k
        np.testing.assert_allclose(g.numpy(), r.numpy(), rtol=1e-3, atol=2e-3, err_msg=k)
        assert _relerr(g, b) < 1e-5, k + " vs the default kernel"     # same arithmetic, different summation order


@pytest.mark.parametrize("levels,Lq,ref_dim", [([(19, 27), (10, 14), (5, 7), (3, 4)], None, 2),
                                               ([(19, 27), (10, 14), (5, 7), (3, 4)], 211, 4),
                                               ([(9, 8), (4, 5)], 37, 4), ([(6, 5)], None, 2)])
@pytest.mark.parametrize("variant", VARIANTS)
def test_fused_vs_default_fused_kernel(variant, levels, Lq, ref_dim):
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    from fused_fixture import fused_inputs
    S = sum(h * w for h, w in levels)
    value, shapes, start, ref, off, logits, gout = fused_inputs(levels, 2, Lq or S, ref_dim, seed=len(levels) + ref_dim)
    with _variant(variant):
        got = MSDA.ms_deform_attn_fused_backward(value, shapes, start, ref, off, logits, gout)
    base = MSDA.ms_deform_attn_fused_backward(value, shapes, start, ref, off, logits, gout)
    for g, b, k in zip(got, base, ("grad_value", "grad_offsets", "grad_logits")):
        assert _relerr(g, b) < 2e-5, k
        assert torch.allclose(g, b, rtol=1e-3, atol=2e-3 * float(b.abs().max())), k


@pytest.mark.parametrize("variant", VARIANTS)
def test_full_size_properties(variant):
    """Train-step encoder shape: grad_value conserves mass (sum over pixels = sum_q a * w * grad_out over valid
    corners is the same number the 8-lane kernel produces), every output finite."""
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    from semi_detr_b200.synthetic import COCO_4SCALE_LEVELS, msda_inputs
    x = msda_inputs(COCO_4SCALE_LEVELS, N=2, mode="encoder", seed=1)
    args = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["gout"], 64)
    with _variant(variant):
        got = MSDA.ms_deform_attn_backward(*args)
    base = MSDA.ms_deform_attn_backward(*args)
    for g, b in zip(got, base):
        assert torch.isfinite(g).all()
        assert _relerr(g, b) < 1e-5
    assert abs(float(got[0].double().sum()) - float(base[0].double().sum())) < 1e-6 * float(base[0].double().abs().sum())
