"""Device Hungarian matching against the real mmdet assigner's golden outputs, scipy's golden answers and the
C oracle: assignment indices must be bit-identical."""
import numpy as np
import pytest
import torch

from oracle import hungarian_oracle as H
from oracle.lsap_oracle import linear_sum_assignment as oracle_lsa

pytestmark = pytest.mark.gpu

CASES = ["q900_g7", "q300_g1", "q400_g30", "q900_g100", "q100_g0", "q50_g60", "q300_g13"]


def _assigner():
    from semi_detr_b200.matching import HungarianAssigner
    return HungarianAssigner(cls_cost=dict(type="FocalLossCost", weight=2.0),
                             reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
                             iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))


@pytest.mark.parametrize("name", CASES)
def test_assign_matches_reference_assigner(hungarian_golden, name):
    c = hungarian_golden[name]
    ih, iw = (int(x) for x in c["img_hw"])
    a = _assigner()
    res = a.assign(torch.from_numpy(c["bbox_pred"]).cuda(), torch.from_numpy(c["cls_pred"]).cuda(),
                   torch.from_numpy(c["gt_bboxes"]).cuda(), torch.from_numpy(c["gt_labels"]).cuda(),
                   dict(img_shape=(ih, iw, 3)))
    a.check_status()
    assert res.gt_inds.dtype == torch.int64 and res.labels.dtype == torch.int64
    assert np.array_equal(res.gt_inds.cpu().numpy(), c["gt_inds"])
    assert np.array_equal(res.labels.cpu().numpy(), c["labels"])
    assert res.num_gts == c["gt_bboxes"].shape[0]


@pytest.mark.parametrize("name", [n for n in CASES if "_g0" not in n])
def test_cost_matches_reference(hungarian_golden, name):
    from semi_detr_b200.matching import MatchTargets
    c = hungarian_golden[name]
    ih, iw = (int(x) for x in c["img_hw"])
    t = MatchTargets([torch.from_numpy(c["gt_bboxes"])], [torch.from_numpy(c["gt_labels"])], [(iw, ih)], "cuda")
    _, _, costs = _assigner().assign_batch(torch.from_numpy(c["bbox_pred"]).cuda()[None],
                                           torch.from_numpy(c["cls_pred"]).cuda()[None], t, return_cost=True)
    got = costs[0].cpu().numpy()
    # fp32, same operation order as the torch kernels; libm-level differences only
    np.testing.assert_allclose(got, c["cost"], rtol=2e-5, atol=2e-6)


def test_solver_bit_identical_on_golden_costs(lsap_golden):
    from semi_detr_b200.matching.hungarian_assigner import linear_sum_assignment as dev_lsa
    for name, c in lsap_golden.items():
        if name == "scipy_version":
            continue
        r, cc = dev_lsa(torch.from_numpy(c["cost"]).cuda())
        assert np.array_equal(r.cpu().numpy(), c["rows"]), name
        assert np.array_equal(cc.cpu().numpy(), c["cols"]), name


def test_solver_vs_oracle_random_and_ties():
    from semi_detr_b200.matching.hungarian_assigner import linear_sum_assignment as dev_lsa
    rng = np.random.default_rng(2)
    for Q in (1, 5, 33, 200, 900):
        mats = []
        for t in range(12):
            G = int(rng.integers(1, 120))
            kind = t % 4
            if kind == 0:
                m = rng.standard_normal((Q, G))
            elif kind == 1:
                m = rng.integers(0, 3, (Q, G))
            elif kind == 2:
                m = np.round(rng.standard_normal((Q, G)) * 2)
            else:
                m = rng.integers(0, 2, (Q, G))
            mats.append(m.astype(np.float32))
        got = dev_lsa([torch.from_numpy(m).cuda() for m in mats])      # one launch, 12 problems
        for m, (r, c) in zip(mats, got):
            wr, wc = oracle_lsa(m)
            assert np.array_equal(r.cpu().numpy(), wr) and np.array_equal(c.cpu().numpy(), wc), (Q, m.shape)


def test_solver_invalid_and_infeasible():
    from semi_detr_b200.matching.hungarian_assigner import linear_sum_assignment as dev_lsa
    with pytest.raises(ValueError, match="invalid numeric"):
        dev_lsa(torch.tensor([[float("nan"), 1.0]], device="cuda"))
    with pytest.raises(ValueError, match="invalid numeric"):
        dev_lsa(torch.tensor([[-float("inf"), 1.0]], device="cuda"))
    with pytest.raises(ValueError, match="infeasible"):
        dev_lsa(torch.full((2, 2), float("inf"), device="cuda"))


def test_batched_step_matches_oracle():
    """A supervised step's worth of problems: 7 'layers' x 2 images, ragged GT counts incl. an empty image."""
    from semi_detr_b200.matching import MatchTargets
    g = torch.Generator().manual_seed(0)
    Q, C, layers = 900, 80, 7
    counts = [7, 0, 23]
    sizes = [(1333, 800), (1201, 800), (1000, 750)]
    gtb, gtl = [], []
    for n, (w, h) in zip(counts, sizes):
        xy = torch.rand(n, 2, generator=g) * 0.6
        wh = torch.rand(n, 2, generator=g) * 0.35 + 0.03
        gtb.append(torch.cat([xy, xy + wh], 1) * torch.tensor([w, h, w, h]))
        gtl.append(torch.randint(0, C, (n,), generator=g))
    P = layers * len(counts)
    bbox = torch.rand(P, Q, 4, generator=g) * torch.tensor([1, 1, 0.5, 0.5]) + torch.tensor([0, 0, 0.01, 0.01])
    cls = torch.randn(P, Q, C, generator=g) * 2 - 3
    t = MatchTargets(gtb, gtl, sizes, "cuda")
    a = _assigner()
    gi, lb = a.assign_batch(bbox.cuda(), cls.cuda(), t)
    a.check_status()
    for p in range(P):
        i = p % len(counts)
        w, h = sizes[i]
        wi, wl = H.hungarian_assign(bbox[p], cls[p], gtb[i], gtl[i], h, w)
        assert np.array_equal(gi[p].cpu().numpy(), wi.numpy()), p
        assert np.array_equal(lb[p].cpu().numpy(), wl.numpy()), p


def test_out_of_range_label_is_reported_like_an_invalid_cost():
    """A ground-truth label outside [0, num_classes) would index the class scores out of bounds: the cost kernel turns
    it into a NaN entry instead, the solver flags the problem, and `check_status()` raises what scipy raises for a cost
    matrix with invalid entries (hungarian_assigner.py:136); `FusedSupervisedTrainStep.check()` calls it off the hot
    path."""
    g = torch.Generator().manual_seed(0)
    bbox = (torch.rand(50, 4, generator=g) * 0.5 + 0.1).cuda()
    cls = torch.randn(50, 80, generator=g).cuda()
    gtb = torch.tensor([[10., 10., 60., 90.], [100., 50., 400., 300.]]).cuda()
    a = _assigner()
    a.assign(bbox, cls, gtb, torch.tensor([3, 17]).cuda(), dict(img_shape=(800, 1333, 3)))
    a.check_status()                                   # in-range labels: fine
    a.assign(bbox, cls, gtb, torch.tensor([3, 80]).cuda(), dict(img_shape=(800, 1333, 3)))
    with pytest.raises(ValueError, match="invalid numeric entries"):
        a.check_status()
