"""Matching-cost descriptors with the reference's registry names and constructor arguments
(thirdparty/mmdetection/mmdet/core/bbox/match_costs/match_cost.py:10-185).

On the hot path they are *parameters* of the fused cost kernel (``sdb_match_cost_f32``): the assigner reads
their weights and launches one kernel for all layers x images.  ``__call__`` keeps the reference's per-term
call surface (used directly by detr_ssod/models/dino_detr_ssod.py:265-271); it is plain torch elementwise
plumbing on whatever device the inputs live on.
"""
import torch

from ..registry import MATCH_COST


def bbox_cxcywh_to_xyxy(bbox):
    cx, cy, w, h = bbox.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def bbox_xyxy_to_cxcywh(bbox):
    x1, y1, x2, y2 = bbox.unbind(-1)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], -1)


def bbox_overlaps_giou(a, b, eps=1e-6):
    """Pairwise GIoU (Q,G) of xyxy boxes; mmdet iou2d_calculator.py:218-260 with mode='giou'."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, None, 2:], b[None, :, 2:]) - torch.max(a[:, None, :2], b[None, :, :2])).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = (area_a[:, None] + area_b[None, :] - overlap).clamp(min=eps)
    ewh = (torch.max(a[:, None, 2:], b[None, :, 2:]) - torch.min(a[:, None, :2], b[None, :, :2])).clamp(min=0)
    earea = (ewh[..., 0] * ewh[..., 1]).clamp(min=eps)
    return overlap / union - (earea - union) / earea


@MATCH_COST.register_module()
class BBoxL1Cost:
    def __init__(self, weight=1.0, box_format="xyxy"):
        assert box_format in ("xyxy", "xywh")
        self.weight = weight
        self.box_format = box_format

    def __call__(self, bbox_pred, gt_bboxes):
        if self.box_format == "xywh":
            gt_bboxes = bbox_xyxy_to_cxcywh(gt_bboxes)
        else:
            bbox_pred = bbox_cxcywh_to_xyxy(bbox_pred)
        return torch.cdist(bbox_pred, gt_bboxes, p=1) * self.weight


@MATCH_COST.register_module()
class FocalLossCost:
    def __init__(self, weight=1.0, alpha=0.25, gamma=2, eps=1e-12):
        self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps

    def __call__(self, cls_pred, gt_labels):
        p = cls_pred.sigmoid()
        neg = -(1 - p + self.eps).log() * (1 - self.alpha) * p.pow(self.gamma)
        pos = -(p + self.eps).log() * self.alpha * (1 - p).pow(self.gamma)
        return (pos[:, gt_labels] - neg[:, gt_labels]) * self.weight


@MATCH_COST.register_module()
class IoUCost:
    def __init__(self, iou_mode="giou", weight=1.0):
        self.weight, self.iou_mode = weight, iou_mode

    def __call__(self, bboxes, gt_bboxes):
        if self.iou_mode != "giou":
            raise NotImplementedError("semi_detr_b200 implements the 'giou' matching cost the shipped configs use")
        return -bbox_overlaps_giou(bboxes, gt_bboxes) * self.weight


def build_match_cost(cfg, default_args=None):
    return MATCH_COST.build(cfg, default_args)
