"""Oracle: run the hot path's host code on the CPU with the reference's CPU algorithms.  TEST INFRASTRUCTURE ONLY.

``reference_cpu_ops()`` swaps the three device entry points of ``semi_detr_b200`` for their CPU restatements:

* MSDA            -> ``msda_oracle.msda_forward_torch`` (the reference's pure-PyTorch fallback,
                     functions/ms_deform_attn_func.py:41-61; differentiable through autograd)
* Hungarian       -> per-problem cost on torch-CPU + ``cost.cpu()`` + LSAP on the host
                     (hungarian_assigner.py:115-148), i.e. the reference's actual control flow
* EMA             -> python loop of ``mul_`` / ``add_`` (mean_teacher.py:60-64)
* teacher decode  -> ``ssod_oracle.pseudo_label_nms`` (torchvision batched_nms per image + the mean / std filter)
* GMM threshold   -> ``ssod_oracle.gmm_threshold`` (float64 numpy EM, ``gmm_oracle``)
* head loss sums  -> ``loss_oracle.detr_loss_sums`` (mmdet's focal / L1 / GIoU expressions in plain torch, targets
                     gathered from the assignment like ``_get_target_single``, dino_detr_head.py:895-980)

Used by ``tests/`` (to check the host logic without a GPU) and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg (to time the reference's CPU path on the box's host cores).  The product never imports
this module; outside this context manager the package has no CPU path at all.
"""
import contextlib

import torch

from . import hungarian_oracle, loss_oracle, msda_oracle, ssod_oracle
from .ema_oracle import ema_update


class _CpuMSDeformAttnFunction:
    @staticmethod
    def apply(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
        shapes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
        return msda_oracle.msda_forward_torch(value, shapes, sampling_locations, attention_weights)


def _cpu_assign_batch(self, bbox_preds, cls_preds, targets, prob_img=None, return_cost=False):
    P, Q, C = cls_preds.shape
    n_img = targets.num_imgs
    if prob_img is None:
        prob_img = [p % n_img for p in range(P)]
    gt_inds = torch.zeros((P, Q), dtype=torch.long)
    labels = torch.full((P, Q), -1, dtype=torch.long)
    offs = targets.offsets_host
    wh = targets.img_wh.cpu()
    costs = []
    for p in range(P):
        s = prob_img[p]
        gb = targets.gt_bboxes[offs[s]:offs[s + 1]].cpu() if offs[s + 1] > offs[s] else torch.zeros(0, 4)
        gl = targets.gt_labels[offs[s]:offs[s + 1]].cpu() if offs[s + 1] > offs[s] else torch.zeros(0, dtype=torch.long)
        w, h = float(wh[s, 0]), float(wh[s, 1])
        gi, lb = hungarian_oracle.hungarian_assign(bbox_preds[p].detach().float().cpu(), cls_preds[p].detach().float().cpu(),
                                                   gb, gl, h, w, w_cls=float(self.cls_cost.weight),
                                                   w_l1=float(self.reg_cost.weight), w_iou=float(self.iou_cost.weight))
        gt_inds[p], labels[p] = gi, lb
        if return_cost:
            costs.append(hungarian_oracle.match_cost(bbox_preds[p].detach().float().cpu(), cls_preds[p].detach().float().cpu(),
                                                     gb, gl, h, w) if len(gb) else torch.zeros(Q, 0))
    gt_inds, labels = gt_inds.to(bbox_preds.device), labels.to(bbox_preds.device)
    return (gt_inds, labels, costs) if return_cost else (gt_inds, labels)


def _cpu_plan_step(self, momentum):
    ema_update(self.teacher, self.student, momentum)


@contextlib.contextmanager
def reference_cpu_ops():
    from semi_detr_b200.matching import hungarian_assigner as ha
    from semi_detr_b200.msda import modules as msda_modules
    from semi_detr_b200.dino import fused_loss
    from semi_detr_b200.ssod import device_ops
    from semi_detr_b200.teacher import mean_teacher as mt
    saved = (msda_modules.MSDeformAttnFunction, ha.HungarianAssigner.assign_batch, mt.EmaPlan.step,
             fused_loss.detr_loss_sums, device_ops.pseudo_label_nms, device_ops.gmm_threshold)
    msda_modules.MSDeformAttnFunction = _CpuMSDeformAttnFunction
    ha.HungarianAssigner.assign_batch = _cpu_assign_batch
    mt.EmaPlan.step = _cpu_plan_step
    fused_loss.detr_loss_sums = loss_oracle.detr_loss_sums
    device_ops.pseudo_label_nms = ssod_oracle.pseudo_label_nms
    device_ops.gmm_threshold = ssod_oracle.gmm_threshold
    try:
        yield
    finally:
        (msda_modules.MSDeformAttnFunction, ha.HungarianAssigner.assign_batch, mt.EmaPlan.step,
         fused_loss.detr_loss_sums, device_ops.pseudo_label_nms, device_ops.gmm_threshold) = saved
