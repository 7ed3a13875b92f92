"""Train-step plumbing around the detector: optimizer construction with the configs' param-wise rule and one
supervised step (zero_grad -> forward/loss -> backward -> grad clip -> AdamW), i.e. what mmcv's
``EpochBasedRunner`` + ``OptimizerHook`` do per iteration for
configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:122-128 (AdamW lr 1e-4, wd 1e-4, backbone lr x0.1,
grad_clip max_norm 0.1).  torch's fused AdamW / foreach clip are library plumbing here; fusing clip+AdamW+EMA
over flat buffers is a "next" row."""
import torch
from torch import nn


def build_optimizer(model, lr=1e-4, weight_decay=1e-4, backbone_lr_mult=0.1, fused=None):
    backbone, rest = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (backbone if "backbone" in name else rest).append(p)
    groups = [dict(params=rest, lr=lr, weight_decay=weight_decay),
              dict(params=backbone, lr=lr * backbone_lr_mult, weight_decay=weight_decay)]
    if fused is None:
        fused = all(p.is_cuda for p in rest + backbone)
    return torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, fused=fused)


class SupervisedTrainStep:
    """One data-parallel rank's step.  ``model`` may be wrapped in DistributedDataParallel (gradient all-reduce
    over NCCL/NVLink overlaps backward)."""

    def __init__(self, model, optimizer, max_grad_norm=0.1):
        self.model, self.optimizer, self.max_grad_norm = model, optimizer, max_grad_norm
        self.params = [p for g in optimizer.param_groups for p in g["params"]]

    def __call__(self, data):
        self.optimizer.zero_grad(set_to_none=True)
        losses = self.model(**data)
        inner = self.model.module if hasattr(self.model, "module") else self.model
        loss, log_vars = inner._parse_losses(losses)
        loss.backward()
        if self.max_grad_norm is not None:
            nn.utils.clip_grad_norm_(self.params, self.max_grad_norm, norm_type=2, foreach=True)
        self.optimizer.step()
        return loss.detach(), log_vars
