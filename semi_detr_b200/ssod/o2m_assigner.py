"""``O2MAssigner`` (warm-up phase, iter < warm_up_step) -- mirror of
detr_od/core/bbox/assigners/o2m_assigner.py:18-170: alignment metric score^alpha * IoU^beta, top-13 candidates per
GT, a query claimed by several GTs goes to the one with the largest IoU.  Same result fields as
``O2MAssignResult`` (o2m_assign_result.py).  Vectorised on the device: the reference loops over GTs in python
(:138-139) and again in the head to normalise the metrics (dino_detr_ssod_head.py:1152-1157)."""
import torch

from ..consts import device_const
from ..matching.match_cost import bbox_cxcywh_to_xyxy
from ..registry import BBOX_ASSIGNERS

INF = 100000000


def pairwise_iou(a, b, eps=1e-6):
    """mmdet bbox_overlaps(mode='iou') (iou2d_calculator.py:218-251)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, None, 2:], b[None, :, 2:]) - torch.max(a[:, None, :2], b[None, :, :2])).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    return overlap / (area_a[:, None] + area_b[None, :] - overlap).clamp(min=eps)


class O2MAssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, assign_metrics, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps = num_gts, gt_inds, max_overlaps
        self.assign_metrics, self.labels = assign_metrics, labels


@BBOX_ASSIGNERS.register_module()
class O2MAssigner:
    def __init__(self, candidate_topk=13, debug=False):
        self.candidate_topk = candidate_topk
        self.debug = debug

    def assign(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, img_meta, gt_bboxes_ignore=None, alpha=1, beta=6):
        """bbox_pred (Q,4) cxcywh in [0,1]; cls_pred (Q,C) *sigmoid scores*; gt_bboxes (G,4) xyxy pixels."""
        assert gt_bboxes_ignore is None
        num_gts, Q = gt_bboxes.size(0), bbox_pred.size(0)
        gt_labels = gt_labels.long()
        if num_gts == 0 or Q == 0:
            gt_inds = bbox_pred.new_full((Q,), 0 if num_gts == 0 else -1, dtype=torch.long)
            return O2MAssignResult(num_gts, gt_inds, bbox_pred.new_zeros((Q,)), bbox_pred.new_zeros((Q,)),
                                   labels=bbox_pred.new_full((Q,), -1, dtype=torch.long))
        h, w, _ = img_meta["img_shape"]
        factor = device_const(bbox_pred.device, "whwh", (w, h), lambda: torch.tensor([w, h, w, h], dtype=torch.float32))
        pred = bbox_cxcywh_to_xyxy(bbox_pred) * factor
        overlaps = pairwise_iou(pred, gt_bboxes).detach()
        metrics = cls_pred[:, gt_labels].detach() ** alpha * overlaps ** beta
        k = min(self.candidate_topk, Q)
        cand = metrics.topk(k, dim=0, largest=True)[1]                       # (k, G)
        is_pos = metrics.gather(0, cand) > 0
        claimed = torch.zeros_like(overlaps, dtype=torch.bool).scatter_(0, cand, is_pos)
        masked = torch.where(claimed, overlaps, overlaps.new_full((), -INF))
        max_overlaps, argmax = masked.max(dim=1)
        pos = max_overlaps != -INF
        gt_inds = torch.where(pos, argmax + 1, torch.zeros_like(argmax))
        assign_metrics = torch.where(pos, metrics.gather(1, argmax[:, None])[:, 0], metrics.new_zeros(()))
        labels = torch.where(pos, gt_labels[argmax], torch.full_like(argmax, -1))
        return O2MAssignResult(num_gts, gt_inds, max_overlaps, assign_metrics, labels=labels)


def assign_layers(assigner, bbox_pred, cls_prob, gt_bboxes, gt_labels, img_meta, alpha=1, beta=6):
    """``O2MAssigner.assign`` for ALL decoder layers of one image in one pass: bbox_pred (Lyr, Q, 4), cls_prob
    (Lyr, Q, C) sigmoid scores, gt_bboxes (G, 4) shared, gt_labels (Lyr, G) (the encoder-proposal row is matched against
    class 0, dino_detr_ssod_head.py:574-577) -> gt_inds, max_overlaps, assign_metrics, normalised metrics, each (Lyr, Q).
    Every operation is the per-layer one with a leading layer dimension, so the numbers are those of Lyr separate
    ``assign`` + ``normalized_alignment_metrics`` calls (the reference: 7 x batch python-level calls per step)."""
    Lyr, Q = bbox_pred.shape[:2]
    G = gt_bboxes.size(0)
    if G == 0 or Q == 0:
        z = bbox_pred.new_zeros((Lyr, Q))
        return z.long(), z, z, z
    h, w, _ = img_meta["img_shape"]
    factor = device_const(bbox_pred.device, "whwh", (w, h), lambda: torch.tensor([w, h, w, h], dtype=torch.float32))
    pred = bbox_cxcywh_to_xyxy(bbox_pred) * factor                                   # (Lyr, Q, 4)
    area_a = (pred[..., 2] - pred[..., 0]) * (pred[..., 3] - pred[..., 1])
    area_b = (gt_bboxes[:, 2] - gt_bboxes[:, 0]) * (gt_bboxes[:, 3] - gt_bboxes[:, 1])
    wh = (torch.min(pred[:, :, None, 2:], gt_bboxes[None, None, :, 2:]) -
          torch.max(pred[:, :, None, :2], gt_bboxes[None, None, :, :2])).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    overlaps = (overlap / (area_a[..., None] + area_b[None, None, :] - overlap).clamp(min=1e-6)).detach()
    sel = cls_prob.gather(2, gt_labels.long()[:, None, :].expand(Lyr, Q, G)).detach()
    metrics = sel ** alpha * overlaps ** beta                                          # (Lyr, Q, G)
    k = min(assigner.candidate_topk, Q)
    cand = metrics.topk(k, dim=1, largest=True)[1]                                     # (Lyr, k, G)
    is_pos = metrics.gather(1, cand) > 0
    claimed = torch.zeros_like(overlaps, dtype=torch.bool).scatter_(1, cand, is_pos)
    masked = torch.where(claimed, overlaps, overlaps.new_full((), -INF))
    max_overlaps, argmax = masked.max(dim=2)                                           # (Lyr, Q)
    pos = max_overlaps != -INF
    gt_inds = torch.where(pos, argmax + 1, torch.zeros_like(argmax))
    assign_metrics = torch.where(pos, metrics.gather(2, argmax[..., None])[..., 0], metrics.new_zeros(()))
    # per-(layer, GT) normalisation of the positives' metrics (dino_detr_ssod_head.py:1146-1157)
    g = (gt_inds - 1).clamp(min=0)
    ious = torch.where(pos, max_overlaps, torch.zeros_like(max_overlaps))
    zero = assign_metrics.new_zeros((Lyr, G))
    max_metric = zero.scatter_reduce(1, g, torch.where(pos, assign_metrics, zero[0, 0]), "amax", include_self=True)
    max_iou = zero.scatter_reduce(1, g, ious, "amax", include_self=True)
    norm = assign_metrics / (max_metric.gather(1, g) + 10e-8) * max_iou.gather(1, g)
    return gt_inds, max_overlaps, assign_metrics, torch.where(pos, norm, torch.zeros_like(norm))


def normalized_alignment_metrics(res):
    """Per-GT normalisation of the positives' metrics (dino_detr_ssod_head.py:1146-1157):
    metric / (max metric of that GT + 1e-7) * (max IoU of that GT), zero for negatives."""
    pos = res.gt_inds > 0
    g = (res.gt_inds - 1).clamp(min=0)
    n = max(res.num_gts, 1)
    ious = torch.where(res.max_overlaps == -INF, torch.zeros_like(res.max_overlaps), res.max_overlaps)
    zero = res.assign_metrics.new_zeros(n)
    max_metric = zero.scatter_reduce(0, g, torch.where(pos, res.assign_metrics, zero[0]), "amax", include_self=True)
    max_iou = zero.scatter_reduce(0, g, torch.where(pos, ious, zero[0]), "amax", include_self=True)
    out = res.assign_metrics / (max_metric[g] + 10e-8) * max_iou[g]
    return torch.where(pos, out, torch.zeros_like(out))
