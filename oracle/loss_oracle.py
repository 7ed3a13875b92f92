"""Oracle: the DINO head's per-problem loss sums with plain torch ops on the CPU.  TEST INFRASTRUCTURE ONLY.

Restates, element for element, what the reference's ``loss_single`` feeds its loss modules
(detr_od/models/dense_heads/dino_detr_head.py:634-736 with targets from :895-980):
 * py_sigmoid_focal_loss      thirdparty/mmdetection/mmdet/models/losses/focal_loss.py:12-57
 * l1_loss                    .../smooth_l1_loss.py:34-46 on normalised cxcywh, plus the xy / hw parts (:726-734)
 * giou_loss                  .../iou_loss.py:101-116 through bbox_overlaps(aligned) iou2d_calculator.py:204-260
Same signature as ``semi_detr_b200.dino.fused_loss.detr_loss_sums``; differentiable through autograd.
"""
import torch
import torch.nn.functional as F


def _cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def _xyxy_to_cxcywh(b):
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], -1)


def sigmoid_focal(pred, labels, num_classes, gamma, alpha):
    target = (labels.unsqueeze(-1) == torch.arange(num_classes, device=labels.device)).type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * target + p * (1 - target)
    focal_weight = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    return F.binary_cross_entropy_with_logits(pred, target, reduction="none") * focal_weight


def giou_aligned(a, b, eps):
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    wh = (torch.min(a[..., 2:], b[..., 2:]) - torch.max(a[..., :2], b[..., :2])).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = (area_a + area_b - overlap).clamp(min=eps)
    ewh = (torch.max(a[..., 2:], b[..., 2:]) - torch.min(a[..., :2], b[..., :2])).clamp(min=0)
    earea = (ewh[..., 0] * ewh[..., 1]).clamp(min=eps)
    return overlap / union - (earea - union) / earea


def detr_loss_sums(cls, box, gt_inds, prob_seg, seg_offsets, gt_bboxes, gt_labels, img_wh, cls_weight=None, alpha=0.25,
                   gamma=2.0, eps=1e-6):
    P, Q, C = cls.shape
    seg = prob_seg.long()
    wh4 = torch.cat([img_wh, img_wh], 1).to(cls.dtype)                 # (nseg, 4)
    pos = gt_inds > 0
    if gt_bboxes is not None and gt_bboxes.numel():
        counts = (seg_offsets[1:] - seg_offsets[:-1]).long()
        seg_of_gt = torch.repeat_interleave(torch.arange(counts.numel(), device=counts.device), counts)
        gt_norm = _xyxy_to_cxcywh(gt_bboxes.to(cls.dtype) / wh4[seg_of_gt])
        gidx = (seg_offsets.long()[seg][:, None] + gt_inds - 1).clamp(min=0)
        box_t = gt_norm[gidx] * pos.unsqueeze(-1)
        labels = torch.where(pos, gt_labels.long()[gidx], torch.full_like(gt_inds, C))
    else:
        box_t = torch.zeros_like(box)
        labels = torch.full_like(gt_inds, C)
    w = pos.unsqueeze(-1).to(box.dtype)
    focal = sigmoid_focal(cls, labels, C, gamma, alpha)
    if cls_weight is not None:
        focal = focal * cls_weight[:, None, None]
    l1 = (box - box_t).abs() * w
    factor = wh4[seg][:, None, :]
    giou = giou_aligned(_cxcywh_to_xyxy(box) * factor, _cxcywh_to_xyxy(box_t) * factor, eps)
    return torch.stack([focal.sum((1, 2)), l1.sum((1, 2)), l1[..., :2].sum((1, 2)), l1[..., 2:].sum((1, 2)),
                        ((1 - giou) * w.squeeze(-1)).sum(1)], 1)
