"""Device entry points of the pseudo-label side path (``csrc/ssod.cu``): class-wise NMS + mean/std filter of the
teacher's detections, and the GMM cost threshold.  No CPU path; the CPU restatements used by the parity tests live in
``oracle/ssod_oracle.py``."""
import torch

from .. import _lib


def pseudo_label_nms(scores, boxes_xyxy, score_thr=0.01, iou_thr=0.6, max_per_img=300, mean_std_filter=True):
    """scores (B, Q, C) sigmoid class scores, boxes_xyxy (B, Q, 4) pixels ->
    (boxes (B, max_per_img, 4), scores (B, max_per_img), labels (B, max_per_img) int64, count (B,) int32,
     nms_count (B,) int32), all on the device; rows past ``count`` are zero.
    dino_detr_ssod_head.py:1371-1395 (multiclass_nms) + dino_detr_ssod.py:921-939 (mean + std, w / h > 0)."""
    if not scores.is_cuda:
        raise RuntimeError("pseudo_label_nms: Not implemented on the CPU (semi_detr_b200 has no CPU path)")
    B, Q, C = scores.shape
    dev = scores.device
    ss, ii = scores.reshape(B, Q * C).float().sort(dim=1, descending=True)
    boxes_xyxy = boxes_xyxy.float().contiguous()
    ob = torch.empty((B, max_per_img, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((B, max_per_img), dtype=torch.float32, device=dev)
    ol = torch.empty((B, max_per_img), dtype=torch.int64, device=dev)
    cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    ncnt = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().sdb_pseudo_label_nms_f32(
            _lib.current_stream(dev), ss.data_ptr(), ii.data_ptr(), boxes_xyxy.data_ptr(), B, Q * C, Q, C,
            float(score_thr), float(iou_thr), int(max_per_img), 1 if mean_std_filter else 0, ob.data_ptr(),
            os_.data_ptr(), ol.data_ptr(), cnt.data_ptr(), ncnt.data_ptr())
    _lib.check(rc, "pseudo_label_nms")
    _lib.LAUNCHES["pseudo_label_nms"] += 1
    return ob, os_, ol, cnt, ncnt


def gmm_threshold(costs, seg_counts=None, seg_stride=None, tol=1e-3, max_iter=100, reg_covar=1e-5):
    """Pooled matched costs -> (2,) device tensor [threshold, number of costs] (dino_detr_ssod.py:832-890).
    ``costs`` 1-D float32; either one segment (``seg_counts`` None: all of it) or the padded all-gather layout:
    ``seg_counts`` (nseg,) int32 on the device, ``seg_stride`` floats per segment."""
    if not costs.is_cuda:
        raise RuntimeError("gmm_threshold: Not implemented on the CPU (semi_detr_b200 has no CPU path)")
    dev = costs.device
    costs = costs.float().contiguous()
    if seg_counts is None:
        seg_counts = torch.full((1,), costs.numel(), dtype=torch.int32, device=dev)
        seg_stride = costs.numel()
    if costs.numel() == 0:
        costs = torch.zeros(1, dtype=torch.float32, device=dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().sdb_gmm_threshold_f32(_lib.current_stream(dev), costs.data_ptr(), seg_counts.data_ptr(),
                                              int(seg_counts.numel()), int(seg_stride), float(tol), int(max_iter),
                                              float(reg_covar), out.data_ptr())
    _lib.check(rc, "gmm_threshold")
    _lib.LAUNCHES["gmm_threshold"] += 1
    return out
