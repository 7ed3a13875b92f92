"""Oracle: the pseudo-label side path with plain torch / torchvision ops on the CPU.  TEST INFRASTRUCTURE ONLY.

``pseudo_label_nms`` restates what the reference does per image with mmdet / mmcv ops:
 * multiclass_nms (dino_detr_ssod_head.py:1371-1395 -> mmdet core/post_processing/bbox_nms.py): every (query, class)
   pair with score > score_thr competes; class-wise NMS (mmcv batched_nms: boxes shifted by class * (max + 1), one
   greedy pass, IoU > thr suppresses), survivors in descending score order, first max_per_img kept;
 * extract_teacher_info's filter (dino_detr_ssod.py:921-939): score >= mean + std (unbiased), w > 0, h > 0.
Same signature and fixed-capacity outputs as ``semi_detr_b200.ssod.device_ops.pseudo_label_nms``.
``gmm_threshold`` wraps gmm_oracle.fit_gmm_threshold with the device op's signature.
"""
import numpy as np
import torch
from torchvision.ops import batched_nms

from .gmm_oracle import fit_gmm_threshold


def pseudo_label_nms(scores, boxes_xyxy, score_thr=0.01, iou_thr=0.6, max_per_img=300, mean_std_filter=True):
    B, Q, C = scores.shape
    ob = torch.zeros(B, max_per_img, 4)
    os_ = torch.zeros(B, max_per_img)
    ol = torch.zeros(B, max_per_img, dtype=torch.int64)
    cnt = torch.zeros(B, dtype=torch.int32)
    ncnt = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        s = scores[b].float().cpu()
        q_idx, c_idx = (s > score_thr).nonzero(as_tuple=True)
        bb, ss = boxes_xyxy[b].float().cpu()[q_idx], s[q_idx, c_idx]
        keep = batched_nms(bb, ss, c_idx, iou_thr)[:max_per_img] if bb.numel() else torch.zeros(0, dtype=torch.long)
        bb, ss, ll = bb[keep], ss[keep], c_idx[keep]
        ncnt[b] = len(keep)
        if mean_std_filter:
            thr = ss.mean() + ss.std() if len(keep) else torch.tensor(0.0)
            ok = (ss >= thr) & (bb[:, 2] - bb[:, 0] > 0) & (bb[:, 3] - bb[:, 1] > 0)
            bb, ss, ll = bb[ok], ss[ok], ll[ok]
        n = bb.shape[0]
        ob[b, :n], os_[b, :n], ol[b, :n], cnt[b] = bb, ss, ll, n
    dev = scores.device
    return ob.to(dev), os_.to(dev), ol.to(dev), cnt.to(dev), ncnt.to(dev)


def gmm_threshold(costs, seg_counts=None, seg_stride=None, tol=1e-3, max_iter=100, reg_covar=1e-5):
    c = costs.detach().float().cpu().numpy().reshape(-1)
    if seg_counts is not None:
        counts = seg_counts.cpu().numpy()
        c = np.concatenate([c[s * seg_stride:s * seg_stride + int(n)] for s, n in enumerate(counts)] + [np.zeros(0, np.float32)])
    return torch.tensor([fit_gmm_threshold(c), float(c.size)], dtype=torch.float32, device=costs.device)
