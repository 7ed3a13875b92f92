"""Oracle: DETR matching cost + Hungarian assignment on the CPU.  TEST INFRASTRUCTURE ONLY.

Restates (torch-CPU, fp32, same operation order as the reference so the cost matrix is the
one the reference would hand to scipy):

* FocalLossCost   thirdparty/mmdetection/mmdet/core/bbox/match_costs/match_cost.py:83-99
* BBoxL1Cost      .../match_cost.py:33-50  (box_format='xywh': gt converted, torch.cdist p=1)
* IoUCost (giou)  .../match_cost.py:169-185 -> bbox_overlaps
                  thirdparty/mmdetection/mmdet/core/bbox/iou_calculators/iou2d_calculator.py:192-260
* box transforms  thirdparty/mmdetection/mmdet/core/bbox/transforms.py:222-247
* HungarianAssigner.assign  .../assigners/hungarian_assigner.py:96-148

Weights default to configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:40-44 (cls 2, L1 5, giou 2).
Pinning: the reference's only KAT on this path is the IoUCost doctest (match_cost.py:155-162),
checked in tests/test_oracle_hungarian.py; the assign() result is otherwise unpinned by the
reference's tests (mmdet's tests/ are not vendored) and is pinned by our own goldens from scipy.
"""
import torch

from .lsap_oracle import linear_sum_assignment


def cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def xyxy_to_cxcywh(b):
    x1, y1, x2, y2 = b.unbind(-1)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], -1)


def focal_cost(cls_pred, gt_labels, weight=2.0, alpha=0.25, gamma=2, eps=1e-12):
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    return (pos[:, gt_labels] - neg[:, gt_labels]) * weight


def l1_cost(bbox_pred, gt_norm_xyxy, weight=5.0):
    return torch.cdist(bbox_pred, xyxy_to_cxcywh(gt_norm_xyxy), p=1) * weight


def giou_matrix(a, b, eps=1e-6):
    """Pairwise GIoU of xyxy boxes a (Q,4), b (G,4)  (iou2d_calculator.py:218-260)."""
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None, :] - overlap
    e_lt = torch.min(a[:, None, :2], b[None, :, :2])
    e_rb = torch.max(a[:, None, 2:], b[None, :, 2:])
    epsv = union.new_tensor([eps])
    union = torch.max(union, epsv)
    ious = overlap / union
    e_wh = (e_rb - e_lt).clamp(min=0)
    e_area = torch.max(e_wh[..., 0] * e_wh[..., 1], epsv)
    return ious - (e_area - union) / e_area


def giou_cost(bboxes_xyxy, gt_xyxy, weight=2.0):
    return -giou_matrix(bboxes_xyxy, gt_xyxy) * weight


def match_cost(bbox_pred, cls_pred, gt_bboxes, gt_labels, img_h, img_w,
               w_cls=2.0, w_l1=5.0, w_iou=2.0):
    """(Q,G) fp32 cost exactly as hungarian_assigner.py:115-129 composes it."""
    factor = gt_bboxes.new_tensor([img_w, img_h, img_w, img_h]).unsqueeze(0)
    c = focal_cost(cls_pred, gt_labels.long(), w_cls)
    r = l1_cost(bbox_pred, gt_bboxes / factor, w_l1)
    i = giou_cost(cxcywh_to_xyxy(bbox_pred) * factor, gt_bboxes, w_iou)
    return c + r + i


def hungarian_assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, img_h, img_w, solver=None, **kw):
    """Returns (assigned_gt_inds, assigned_labels) int64 (Q,): 0 = background, k+1 = GT k;
    labels -1 where unmatched (hungarian_assigner.py:100-148).  ``solver`` defaults to the C oracle;
    tests also pass scipy's to pin the two against each other."""
    Q, G = bbox_pred.shape[0], gt_bboxes.shape[0]
    gt_inds = torch.full((Q,), -1, dtype=torch.long)
    labels = torch.full((Q,), -1, dtype=torch.long)
    if G == 0 or Q == 0:
        if G == 0:
            gt_inds[:] = 0
        return gt_inds, labels
    cost = match_cost(bbox_pred, cls_pred, gt_bboxes, gt_labels, img_h, img_w, **kw)
    rows, cols = (solver or linear_sum_assignment)(cost.detach().cpu().numpy())
    rows = torch.from_numpy(rows)
    cols = torch.from_numpy(cols)
    gt_inds[:] = 0
    gt_inds[rows] = cols + 1
    labels[rows] = gt_labels.long()[cols]
    return gt_inds, labels
