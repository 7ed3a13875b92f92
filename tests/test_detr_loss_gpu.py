"""Fused head-loss kernels (csrc/detr_loss.cu: targets from the assignment + focal + L1 + GIoU for every (layer, image)
problem in one launch, and the backward) against the oracle's restatement of mmdet's expressions (oracle/loss_oracle.py,
itself pinned to the reference head's 35 loss values by tests/test_dino_reference_golden.py on the CPU).
Bounds: sums 1e-5 relative (fp32 summation order), gradients 1e-5 of their scale -- far inside the north star's 1e-3."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle

pytestmark = pytest.mark.gpu


def _problem(P, Q, C, counts, seed, wh=((1333.0, 800.0), (1201.0, 777.0), (640.0, 480.0))):
    g = torch.Generator().manual_seed(seed)
    nseg = len(counts)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    G = int(offs[-1])
    img_wh = torch.tensor([wh[s % len(wh)] for s in range(nseg)], dtype=torch.float32)
    seg_of_gt = np.repeat(np.arange(nseg), counts)
    xy = torch.rand(G, 2, generator=g) * 0.6
    sz = torch.rand(G, 2, generator=g) * 0.35 + 0.02
    whg = img_wh[seg_of_gt] if G else torch.zeros(0, 2)
    gt_bboxes = torch.cat([xy, xy + sz], 1) * torch.cat([whg, whg], 1)
    gt_labels = torch.randint(0, C, (G,), generator=g)
    prob_seg = torch.tensor([p % nseg for p in range(P)], dtype=torch.int32)
    gt_inds = torch.zeros(P, Q, dtype=torch.int64)
    for p in range(P):                          # a random partial assignment: each GT of the segment to a distinct query
        c = counts[int(prob_seg[p])]
        if c:
            rows = torch.randperm(Q, generator=g)[:min(c, Q)]
            gt_inds[p, rows] = torch.arange(1, len(rows) + 1)
    cls = torch.randn(P, Q, C, generator=g) * 2 - 2
    box = torch.rand(P, Q, 4, generator=g) * torch.tensor([1, 1, 0.6, 0.6]) + 0.01
    return dict(cls=cls, box=box, gt_inds=gt_inds, prob_seg=prob_seg, seg_offsets=torch.from_numpy(offs),
                gt_bboxes=gt_bboxes, gt_labels=gt_labels, img_wh=img_wh)


def _run_both(x, cls_weight=None, alpha=0.25, gamma=2.0, eps=1e-6):
    from semi_detr_b200.dino import fused_loss
    gsum = torch.randn(x["cls"].shape[0], 5, generator=torch.Generator().manual_seed(99)).double()
    # oracle in float64
    c64 = x["cls"].double().requires_grad_(True)
    b64 = x["box"].double().requires_grad_(True)
    want = loss_oracle.detr_loss_sums(c64, b64, x["gt_inds"], x["prob_seg"], x["seg_offsets"], x["gt_bboxes"].double(),
                                      x["gt_labels"], x["img_wh"].double(),
                                      None if cls_weight is None else cls_weight.double(), alpha, gamma, eps)
    (want * gsum).sum().backward()
    cd = x["cls"].cuda().requires_grad_(True)
    bd = x["box"].cuda().requires_grad_(True)
    got = fused_loss.detr_loss_sums(cd, bd, x["gt_inds"].cuda(), x["prob_seg"].cuda(), x["seg_offsets"].cuda(),
                                    x["gt_bboxes"].cuda(), x["gt_labels"].cuda(), x["img_wh"].cuda(),
                                    None if cls_weight is None else cls_weight.cuda(), alpha, gamma, eps)
    (got * gsum.float().cuda()).sum().backward()
    return got.detach().cpu().double(), want.detach(), cd.grad.cpu().double(), c64.grad, bd.grad.cpu().double(), b64.grad


CASES = {
    "train_step_shape": dict(P=14, Q=900, C=80, counts=[7, 4, 7, 4]),
    "denoising_shape": dict(P=12, Q=200, C=80, counts=[9, 3]),
    "an_image_without_boxes": dict(P=6, Q=300, C=80, counts=[0, 5]),
    "no_boxes_at_all": dict(P=4, Q=100, C=80, counts=[0, 0]),
    "crowded": dict(P=3, Q=900, C=80, counts=[100, 60, 1]),
    "odd_sizes": dict(P=5, Q=37, C=91, counts=[3, 1, 2]),
    "twenty_classes": dict(P=7, Q=450, C=20, counts=[6]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_loss_sums_and_gradients_match_the_oracle(name):
    x = _problem(seed=len(name), **CASES[name])
    got, want, gc, wc, gb, wb = _run_both(x)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (got - want).abs().max()
    assert float((gc - wc).abs().max()) <= 1e-5 * max(float(wc.abs().max()), 1e-12)
    assert float((gb - wb).abs().max()) <= 1e-5 * max(float(wb.abs().max()), 1e-12)
    if sum(CASES[name]["counts"]) == 0:
        assert float(got[:, 1:].abs().max()) == 0.0 and float(gb.abs().max()) == 0.0


def test_class_weights_and_other_focal_parameters():
    x = _problem(P=6, Q=128, C=80, counts=[4, 0, 2], seed=5)
    cw = torch.tensor([1.0, 0.0, 1.0, 1.0, 0.0, 0.5])
    for alpha, gamma in ((0.25, 2.0), (0.4, 1.5), (0.25, 0.0)):
        got, want, gc, wc, gb, wb = _run_both(x, cls_weight=cw, alpha=alpha, gamma=gamma)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
        assert float((gc - wc).abs().max()) <= 1e-5 * float(wc.abs().max())
        assert float((gb - wb).abs().max()) <= 1e-5 * float(wb.abs().max())


def test_disjoint_nested_and_degenerate_boxes():
    """GIoU branches: no overlap (enclosure term only), nested boxes, a zero-area prediction (union clamp)."""
    x = _problem(P=1, Q=4, C=5, counts=[4], seed=1)
    x["gt_bboxes"] = torch.tensor([[100., 100., 300., 300.], [100., 100., 300., 300.], [100., 100., 300., 300.],
                                   [0., 0., 1e-4, 1e-4]])
    x["box"] = torch.tensor([[[0.8, 0.8, 0.1, 0.1], [0.15, 0.25, 0.05, 0.1], [0.15, 0.25, 0.6, 0.9],
                              [0.5, 0.5, 0.0, 0.0]]])
    x["gt_inds"] = torch.tensor([[1, 2, 3, 4]])
    got, want, gc, wc, gb, wb = _run_both(x)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert float((gb - wb).abs().max()) <= 1e-5 * float(wb.abs().max())


def test_sums_are_bitwise_reproducible():
    from semi_detr_b200.dino import fused_loss
    x = {k: v.cuda() for k, v in _problem(P=14, Q=900, C=80, counts=[7, 4], seed=2).items()}
    a = fused_loss.detr_loss_sums(x["cls"], x["box"], x["gt_inds"], x["prob_seg"], x["seg_offsets"], x["gt_bboxes"],
                                  x["gt_labels"], x["img_wh"])
    for _ in range(3):
        b = fused_loss.detr_loss_sums(x["cls"], x["box"], x["gt_inds"], x["prob_seg"], x["seg_offsets"], x["gt_bboxes"],
                                      x["gt_labels"], x["img_wh"])
        assert torch.equal(a, b)


def test_cpu_tensors_raise():
    from semi_detr_b200.dino import fused_loss
    x = _problem(P=2, Q=10, C=5, counts=[1], seed=3)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        fused_loss.detr_loss_sums(x["cls"], x["box"], x["gt_inds"], x["prob_seg"], x["seg_offsets"], x["gt_bboxes"],
                                  x["gt_labels"], x["img_wh"])
