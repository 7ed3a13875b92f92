"""Losses of the DINO head with the reference's registry names, arguments and element-wise definitions:

* ``FocalLoss``   thirdparty/mmdetection/mmdet/models/losses/focal_loss.py:12-57,107-170 (sigmoid focal; on CUDA the
                  reference calls mmcv.ops.sigmoid_focal_loss, same arithmetic)
* ``L1Loss``      .../smooth_l1_loss.py:34-46
* ``GIoULoss``    .../iou_loss.py:101-116, 357-393 (eps 1e-6, (n,4) weights averaged to (n,))
* ``weight_reduce_loss``  .../utils.py:29-55

Besides the reference ``forward(pred, target, weight, avg_factor)`` surface each loss has an ``elementwise``
method: the head evaluates all decoder layers x images in one batched call and reduces per layer itself
(13 ``loss_single`` calls with ~20 launches and 2-3 host syncs each in the reference).
"""
import torch
import torch.nn.functional as F
from torch import nn

from ..matching.match_cost import bbox_cxcywh_to_xyxy  # noqa: F401  (re-export for the head)
from ..registry import LOSSES


def weight_reduce_loss(loss, weight=None, reduction="mean", avg_factor=None):
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return {"mean": loss.mean, "sum": loss.sum, "none": lambda: loss}[reduction]()
    if reduction == "mean":
        return loss.sum() / avg_factor
    if reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def sigmoid_focal_elementwise(pred, labels, num_classes, gamma=2.0, alpha=0.25):
    """pred (..., C) logits, labels (...) with ``num_classes`` = background -> (..., C) focal loss terms."""
    target = (labels.unsqueeze(-1) == torch.arange(num_classes, device=labels.device)).type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * target + p * (1 - target)
    focal_weight = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    return F.binary_cross_entropy_with_logits(pred, target, reduction="none") * focal_weight


def giou_aligned(a, b, eps=1e-6):
    """Row-wise GIoU of xyxy boxes (iou2d_calculator.py:204-260, is_aligned=True)."""
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    wh = (torch.min(a[..., 2:], b[..., 2:]) - torch.max(a[..., :2], b[..., :2])).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = (area_a + area_b - overlap).clamp(min=eps)
    ewh = (torch.max(a[..., 2:], b[..., 2:]) - torch.min(a[..., :2], b[..., :2])).clamp(min=0)
    earea = (ewh[..., 0] * ewh[..., 1]).clamp(min=eps)
    return overlap / union - (earea - union) / earea


@LOSSES.register_module()
class FocalLoss(nn.Module):
    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, "Only sigmoid focal loss supported now."
        self.use_sigmoid, self.gamma, self.alpha = use_sigmoid, gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight

    def elementwise(self, pred, labels):
        return sigmoid_focal_elementwise(pred, labels, pred.shape[-1], self.gamma, self.alpha)

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        loss = self.elementwise(pred, target)
        if weight is not None and weight.shape != loss.shape:
            weight = weight.view(-1, 1) if weight.size(0) == loss.size(0) else weight.view(loss.size(0), -1)
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction_override or self.reduction, avg_factor)


@LOSSES.register_module()
class L1Loss(nn.Module):
    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        loss = (pred - target).abs() if target.numel() else pred.sum() * 0
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction_override or self.reduction, avg_factor)


@LOSSES.register_module()
class GIoULoss(nn.Module):
    def __init__(self, eps=1e-6, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        if weight is not None and weight.dim() > 1:
            weight = weight.mean(-1)
        loss = 1 - giou_aligned(pred, target, self.eps)
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction_override or self.reduction, avg_factor)
