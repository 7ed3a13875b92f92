"""``TaskAlignedFocalLoss`` -- mirror of detr_od/models/losses/task_aligned_focal_loss.py:35-65, 139-205:
soft label = alignment metric on the assigned class, loss = |soft - p|^gamma * BCE(p, soft)."""
import torch
import torch.nn.functional as F
from torch import nn

from ..dino.losses import weight_reduce_loss
from ..registry import LOSSES


def task_aligned_focal_loss(prob, target, alignment_metric, weight=None, gamma=2.0, reduction="mean", avg_factor=None):
    C = prob.shape[-1]
    one_hot = (target.unsqueeze(-1) == torch.arange(C, device=prob.device)).to(prob.dtype)
    soft = alignment_metric.unsqueeze(-1) * one_hot
    loss = (soft - prob).abs().pow(gamma) * F.binary_cross_entropy(prob, soft, reduction="none")
    return weight_reduce_loss(loss, weight, reduction, avg_factor)


@LOSSES.register_module()
class TaskAlignedFocalLoss(nn.Module):
    def __init__(self, use_sigmoid=True, gamma=2.0, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True
        self.use_sigmoid, self.gamma, self.reduction, self.loss_weight = use_sigmoid, gamma, reduction, loss_weight

    def forward(self, prob, target, alignment_metric, weight=None, avg_factor=None, reduction_override=None):
        return self.loss_weight * task_aligned_focal_loss(prob, target, alignment_metric, weight, self.gamma,
                                                          reduction_override or self.reduction, avg_factor)
