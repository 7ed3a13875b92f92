"""Torch-CPU restatement of the matching cost + assign() against goldens from the real mmdet
HungarianAssigner (hungarian_assigner.py:55-188), loaded from /root/reference by make_golden.py."""
import numpy as np
import pytest
import torch

from oracle import hungarian_oracle as H

CASES = ["q900_g7", "q300_g1", "q400_g30", "q900_g100", "q100_g0", "q50_g60", "q300_g13"]


def _t(c, k):
    return torch.from_numpy(c[k])


@pytest.mark.parametrize("name", CASES)
def test_assign_matches_reference(hungarian_golden, name):
    c = hungarian_golden[name]
    ih, iw = (int(x) for x in c["img_hw"])
    gi, lb = H.hungarian_assign(_t(c, "bbox_pred"), _t(c, "cls_pred"), _t(c, "gt_bboxes"), _t(c, "gt_labels"), ih, iw)
    assert np.array_equal(gi.numpy(), c["gt_inds"])
    assert np.array_equal(lb.numpy(), c["labels"])


@pytest.mark.parametrize("name", [n for n in CASES if "_g0" not in n])
def test_cost_bit_exact(hungarian_golden, name):
    c = hungarian_golden[name]
    ih, iw = (int(x) for x in c["img_hw"])
    cost = H.match_cost(_t(c, "bbox_pred"), _t(c, "cls_pred"), _t(c, "gt_bboxes"), _t(c, "gt_labels"), ih, iw)
    assert np.array_equal(cost.numpy(), c["cost"])      # same torch ops in the same order -> same bits


def test_ioucost_doctest_kat(hungarian_golden):
    # match_cost.py:155-162
    b = torch.FloatTensor([[1, 1, 2, 2], [2, 2, 3, 4]])
    g = torch.FloatTensor([[0, 0, 2, 4], [1, 2, 3, 4]])
    out = H.giou_cost(b, g, weight=1.0).numpy()
    np.testing.assert_allclose(out, [[-0.1250, 0.1667], [0.1667, -0.5000]], atol=5e-5)
    assert np.array_equal(out, hungarian_golden["ioucost_doctest"]["out"])
