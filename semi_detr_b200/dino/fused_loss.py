"""Per-problem loss sums of the DINO head on the device: one kernel for every (decoder layer, image) problem --
targets gathered from the assignment, sigmoid focal loss, L1 (+ xy / hw parts) and GIoU loss -- and one for the
backward (``csrc/detr_loss.cu``: ``sdb_detr_loss_forward_f32`` / ``sdb_detr_loss_backward_f32``).

Replaces the reference's 13 x ``loss_single`` / ``get_targets`` / ``_get_target_single`` chain
(detr_od/models/dense_heads/dino_detr_head.py:634-736, 895-980; mmdet losses/focal_loss.py:12-57,
smooth_l1_loss.py:34-46, iou_loss.py:101-116).  There is no CPU path: the CPU restatement used by the parity tests
lives in ``oracle/loss_oracle.py``.
"""
import torch
from torch.autograd.function import once_differentiable

from .. import _lib

LOSS_SUM_NAMES = ("loss_cls", "loss_bbox", "loss_bbox_xy", "loss_bbox_hw", "loss_iou")


class _DetrLossSums(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls, box, gt_inds, prob_seg, seg_offsets, gt_bboxes, gt_labels, img_wh, cls_weight, alpha, gamma,
                eps):
        P, Q, C = cls.shape
        sums = torch.empty((P, 5), dtype=torch.float32, device=cls.device)
        with torch.cuda.device(cls.device):
            rc = _lib.lib().sdb_detr_loss_forward_f32(
                _lib.current_stream(cls.device), cls.data_ptr(), box.data_ptr(), gt_inds.data_ptr(),
                prob_seg.data_ptr(), seg_offsets.data_ptr(), _lib.ptr(gt_bboxes), _lib.ptr(gt_labels),
                img_wh.data_ptr(), _lib.ptr(cls_weight), P, Q, C, alpha, gamma, eps, sums.data_ptr())
        _lib.check(rc, "detr_loss_forward")
        _lib.LAUNCHES["detr_loss_forward"] += 1
        ctx.save_for_backward(cls, box, gt_inds, prob_seg, seg_offsets, gt_bboxes, gt_labels, img_wh, cls_weight)
        ctx.hyper = (alpha, gamma, eps)
        return sums

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_sums):
        cls, box, gt_inds, prob_seg, seg_offsets, gt_bboxes, gt_labels, img_wh, cls_weight = ctx.saved_tensors
        alpha, gamma, eps = ctx.hyper
        P, Q, C = cls.shape
        grad_sums = grad_sums.contiguous()
        g_cls, g_box = torch.empty_like(cls), torch.empty_like(box)
        with torch.cuda.device(cls.device):
            rc = _lib.lib().sdb_detr_loss_backward_f32(
                _lib.current_stream(cls.device), cls.data_ptr(), box.data_ptr(), gt_inds.data_ptr(),
                prob_seg.data_ptr(), seg_offsets.data_ptr(), _lib.ptr(gt_bboxes), _lib.ptr(gt_labels),
                img_wh.data_ptr(), _lib.ptr(cls_weight), grad_sums.data_ptr(), P, Q, C, alpha, gamma, eps,
                g_cls.data_ptr(), g_box.data_ptr())
        _lib.check(rc, "detr_loss_backward")
        _lib.LAUNCHES["detr_loss_backward"] += 1
        return (g_cls, g_box) + (None,) * 10


def detr_loss_sums(cls, box, gt_inds, prob_seg, seg_offsets, gt_bboxes, gt_labels, img_wh, cls_weight=None, alpha=0.25,
                   gamma=2.0, eps=1e-6):
    """cls (P, Q, C) logits, box (P, Q, 4) normalised cxcywh, gt_inds (P, Q) int64 (0 = background, k + 1 = GT k of the
    problem's segment), prob_seg (P,) int32, seg_offsets (nseg + 1,) int32, gt_bboxes (G, 4) pixel xyxy, gt_labels (G,)
    int64, img_wh (nseg, 2) -> (P, 5) sums in the order of ``LOSS_SUM_NAMES`` (un-normalised, un-weighted)."""
    if not cls.is_cuda:
        raise RuntimeError("detr_loss_sums: Not implemented on the CPU (semi_detr_b200 has no CPU path)")
    if cls.dtype != torch.float32 or box.dtype != torch.float32:
        raise RuntimeError("detr_loss_sums: float32 predictions expected (the head casts, force_fp32)")
    if gt_inds.dtype != torch.int64 or prob_seg.dtype != torch.int32 or seg_offsets.dtype != torch.int32:
        raise RuntimeError("detr_loss_sums: gt_inds int64, prob_seg / seg_offsets int32 expected")
    if gt_bboxes is not None and gt_bboxes.numel() == 0:
        gt_bboxes = gt_labels = None
    if gt_bboxes is not None:
        gt_bboxes = gt_bboxes.contiguous().float()
        gt_labels = gt_labels.contiguous().long()
    return _DetrLossSums.apply(cls.contiguous(), box.contiguous(), gt_inds.contiguous(), prob_seg, seg_offsets,
                               gt_bboxes, gt_labels, img_wh.contiguous().float(), cls_weight, float(alpha), float(gamma),
                               float(eps))
