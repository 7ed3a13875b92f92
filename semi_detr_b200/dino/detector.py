"""``DINODETR`` detector and the supervised train step -- host-side mirror of detr_od/models/dino_detr.py:11-24 on
mmdet's ``SingleStageDetector`` / ``BaseDetector`` (thirdparty/mmdetection/mmdet/models/detectors/
single_stage.py:57-90, base.py:176-244): backbone -> head.forward_train -> loss dict -> ``_parse_losses``.

``_parse_losses`` keeps the reference's rule (every entry whose key contains 'loss' is summed into the total,
base.py:198-199) but does not all-reduce + ``.item()`` each of the 66 scalars every step (base.py:202-207): the
log vector stays on the device and is reduced in one collective only when somebody asks for it.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist
from torch import nn

from ..registry import BACKBONES, DETECTORS, HEADS
from . import backbone as _bb  # noqa: F401  (registers ResNet)
from . import head as _head  # noqa: F401  (registers DINODETRHead)


@DETECTORS.register_module()
class DINODETR(nn.Module):
    def __init__(self, backbone, bbox_head, neck=None, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        assert neck is None, "DINODETR has no neck (dino_detr.py:14-24)"
        self.backbone = BACKBONES.build(backbone)
        bbox_head = dict(bbox_head)
        bbox_head.update(train_cfg=train_cfg, test_cfg=test_cfg)
        self.bbox_head = HEADS.build(bbox_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.backbone.to(memory_format=torch.channels_last)
        self.bbox_head.input_proj.to(memory_format=torch.channels_last)

    def extract_feat(self, img, cut_before_last_stage=None):
        # NHWC end to end through the convolutional backbone: cuDNN's tensor-core kernels are channels-last, so this
        # removes the NCHW<->NHWC transposes around every convolution, and (N, C, H, W) channels-last flattens to the
        # transformer's (N, HW, C) token layout for free
        if img.is_cuda:
            img = img.contiguous(memory_format=torch.channels_last)
        if cut_before_last_stage is not None:
            return self.backbone(img, cut_before_last_stage=cut_before_last_stage)
        return self.backbone(img)

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore=None, **kwargs):
        batch_input_shape = tuple(img.shape[-2:])
        for m in img_metas:
            m["batch_input_shape"] = batch_input_shape        # single_stage.py:84-86 / base.py forward_train
        x = self.extract_feat(img)
        return self.bbox_head.forward_train(x, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore, **kwargs)

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        raise NotImplementedError("inference path is outside the train-step hot path")

    @staticmethod
    def _parse_losses(losses, reduce_log_vars=False):
        """-> (total loss tensor, log_vars): ``log_vars`` maps names to 0-dim device tensors; pass
        ``reduce_log_vars=True`` to average them over ranks (one all-reduce) and get python floats."""
        log_vars = OrderedDict()
        total = getattr(losses, "total", None)
        if total is not None:
            # the head already summed its entries (dino.head.LossDict): log the scalars as they are
            assert all(torch.is_tensor(v) and v.dim() == 0 and "loss" in k for k, v in losses.items())
            log_vars.update((k, v.detach()) for k, v in losses.items())
            loss = total
        else:
            for name, value in losses.items():
                if torch.is_tensor(value):
                    log_vars[name] = value.mean()
                elif isinstance(value, (list, tuple)):
                    log_vars[name] = sum(v.mean() for v in value)
                else:
                    raise TypeError(f"{name} is not a tensor or list of tensors")
            loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        if reduce_log_vars:
            vec = torch.stack([v.detach() for v in log_vars.values()])
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(vec.div_(dist.get_world_size()))
            log_vars = OrderedDict(zip(log_vars.keys(), vec.tolist()))
        return loss, log_vars

    def train_step(self, data, optimizer=None):
        """base.py:211-244: returns dict(loss, log_vars, num_samples); the caller runs backward + optimizer."""
        losses = self(**data)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data["img_metas"]))
