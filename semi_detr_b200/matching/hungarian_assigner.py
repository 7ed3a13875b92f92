"""``HungarianAssigner`` on the device -- host-side mirror of
thirdparty/mmdetection/mmdet/core/bbox/assigners/hungarian_assigner.py:16-188.

``assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, img_meta, gt_bboxes_ignore=None, eps=1e-7)`` keeps the
reference signature and returns an ``AssignResult`` (gt_inds: 0 = background, k+1 = GT k; labels -1 / class).
``assign_batch`` is what the DINO head calls: every (decoder layer, image) problem of a step goes through ONE
cost-build launch and ONE solver launch (``sdb_hungarian_assign_f32``) with no ``.cpu()`` and no scipy call --
the reference does a device->host sync + host solve + two host->device copies per problem
(hungarian_assigner.py:131-140), 7 x batch problems per supervised step.

No fallback: cost configurations other than the (FocalLossCost, BBoxL1Cost[xywh], IoUCost[giou]) family the
shipped configs use (configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:40-44) raise NotImplementedError.
"""
import torch

from .. import _lib
from ..consts import device_const
from ..registry import BBOX_ASSIGNERS
from .match_cost import BBoxL1Cost, FocalLossCost, IoUCost, build_match_cost


class AssignResult:
    """mmdet assign_result.py: num_gts, gt_inds, max_overlaps, labels."""

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts = num_gts
        self.gt_inds = gt_inds
        self.max_overlaps = max_overlaps
        self.labels = labels

    @property
    def num_preds(self):
        return len(self.gt_inds)


class MatchTargets:
    """Device-side description of the ground truth of one step, built once and shared by all layers:
    concatenated boxes / labels, int32 segment offsets, per-image (w, h)."""

    def __init__(self, gt_bboxes_list, gt_labels_list, img_wh_list, device):
        self.num_imgs = len(gt_bboxes_list)
        self.counts = [int(b.shape[0]) for b in gt_bboxes_list]          # host ints, no sync
        self.max_gt = max(self.counts) if self.counts else 0
        offs = [0]
        for c in self.counts:
            offs.append(offs[-1] + c)
        self.offsets_host = offs
        if offs[-1] > 0:
            self.gt_bboxes = torch.cat([b.reshape(-1, 4) for b in gt_bboxes_list]).to(device, torch.float32).contiguous()
            self.gt_labels = torch.cat([l.reshape(-1) for l in gt_labels_list]).to(device, torch.int64).contiguous()
        else:
            self.gt_bboxes = torch.zeros((1, 4), dtype=torch.float32, device=device)
            self.gt_labels = torch.zeros((1,), dtype=torch.int64, device=device)
        self.seg_offsets = device_const(device, "seg_offsets", tuple(offs),
                                        lambda: torch.tensor(offs, dtype=torch.int32))
        wh = tuple((float(w), float(h)) for (w, h) in img_wh_list)
        self.img_wh = device_const(device, "img_wh", wh,
                                   lambda: torch.tensor(wh, dtype=torch.float32).reshape(-1, 2))


def linear_sum_assignment(cost):
    """Device drop-in for ``scipy.optimize.linear_sum_assignment`` on a (Q, G) float32 CUDA cost matrix (or a
    list of them, solved in one launch): returns (row_ind ascending, col_ind) int64 CUDA tensors, bit-identical
    to scipy's, raising ValueError for NaN / -inf / infeasible input like scipy does (that check synchronises).
    Call sites in the reference: hungarian_assigner.py:136, detr_ssod/models/dino_detr_ssod.py:279."""
    single = not isinstance(cost, (list, tuple))
    costs = [cost] if single else list(cost)
    dev = costs[0].device
    if not costs[0].is_cuda:
        raise RuntimeError("linear_sum_assignment: CUDA tensors expected (no CPU path)")
    P = len(costs)
    shapes = [tuple(c.shape) for c in costs]
    Q = shapes[0][0]
    assert all(s[0] == Q for s in shapes), "all problems of one call share the number of rows"
    counts = [s[1] for s in shapes]
    offs, seg = [0], [0]
    for g in counts:
        offs.append(offs[-1] + Q * g)
        seg.append(seg[-1] + g)
    flat = [(c.detach().to(torch.float32).t() if Q > c.shape[1] else c.detach().to(torch.float32)).reshape(-1)
            for c in costs]
    work = torch.cat(flat).contiguous() if offs[-1] > 0 else torch.zeros(1, dtype=torch.float32, device=dev)
    prob_seg = torch.arange(P, dtype=torch.int32).to(dev)
    seg_offsets = torch.tensor(seg, dtype=torch.int32).to(dev)
    cost_offsets = torch.tensor(offs, dtype=torch.int64).to(dev)
    gt_inds = torch.empty((P, max(Q, 1)), dtype=torch.int64, device=dev)
    status = torch.empty((P,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().sdb_lsap_solve_f32(_lib.current_stream(dev), work.data_ptr(), cost_offsets.data_ptr(),
                                           prob_seg.data_ptr(), seg_offsets.data_ptr(), None, P, Q,
                                           max(counts) if counts else 0, gt_inds.data_ptr(), None,
                                           status.data_ptr())
    _lib.check(rc, "lsap_solve")
    _lib.LAUNCHES["lsap_solve"] += 1
    st = status.cpu()
    if (st == 1).any():
        raise ValueError("matrix contains invalid numeric entries")
    if (st == 2).any():
        raise ValueError("cost matrix is infeasible")
    out = []
    for p in range(P):
        rows = torch.nonzero(gt_inds[p, :Q] > 0).reshape(-1)
        out.append((rows, gt_inds[p, rows] - 1))
    return out[0] if single else out


@BBOX_ASSIGNERS.register_module()
class HungarianAssigner:
    def __init__(self, cls_cost=dict(type="ClassificationCost", weight=1.0),
                 reg_cost=dict(type="BBoxL1Cost", weight=1.0),
                 iou_cost=dict(type="IoUCost", iou_mode="giou", weight=1.0), debug=False):
        self.cls_cost = build_match_cost(cls_cost)
        self.reg_cost = build_match_cost(reg_cost)
        self.iou_cost = build_match_cost(iou_cost)
        self.debug = debug
        c, r, i = self.cls_cost, self.reg_cost, self.iou_cost
        self._fused_ok = (isinstance(c, FocalLossCost) and c.alpha == 0.25 and c.gamma == 2 and c.eps == 1e-12
                          and isinstance(r, BBoxL1Cost) and r.box_format == "xywh"
                          and isinstance(i, IoUCost) and i.iou_mode == "giou")
        self.last_status = None

    # -- batched hot path ---------------------------------------------------------------------------
    def assign_batch(self, bbox_preds, cls_preds, targets, prob_img=None, return_cost=False):
        """bbox_preds (P, Q, 4) cxcywh in [0,1]; cls_preds (P, Q, C) logits; ``targets`` a MatchTargets;
        ``prob_img`` (P,) image index of each problem (default p % num_imgs, i.e. layer-major stacking).
        Returns gt_inds (P, Q) int64, labels (P, Q) int64 [, list of (Q, G_p) cost views]."""
        if not self._fused_ok:
            raise NotImplementedError(
                "HungarianAssigner: only FocalLossCost(alpha .25, gamma 2) + BBoxL1Cost(xywh) + IoUCost(giou) "
                "is implemented on the device (there is no host fallback)")
        if not bbox_preds.is_cuda:
            raise RuntimeError("HungarianAssigner: predictions must be CUDA tensors (no CPU path)")
        P, Q, C = cls_preds.shape
        dev = bbox_preds.device
        bbox_preds = bbox_preds.detach().to(torch.float32).contiguous()
        cls_preds = cls_preds.detach().to(torch.float32).contiguous()
        n_img = targets.num_imgs
        if prob_img is None:
            prob_img = [p % n_img for p in range(P)]
        cost_offs = [0]
        for p in range(P):
            cost_offs.append(cost_offs[-1] + Q * targets.counts[prob_img[p]])
        prob_img = tuple(int(x) for x in prob_img)
        prob_seg = device_const(dev, "prob_seg", prob_img, lambda: torch.tensor(prob_img, dtype=torch.int32))
        cost_offsets = device_const(dev, "cost_offsets", tuple(cost_offs),
                                    lambda: torch.tensor(cost_offs, dtype=torch.int64))
        total = max(cost_offs[-1], 1)
        workspace = torch.empty(total, dtype=torch.float32, device=dev)
        cost_qg = torch.empty(total, dtype=torch.float32, device=dev) if return_cost else None
        gt_inds = torch.empty((P, Q), dtype=torch.int64, device=dev)
        labels = torch.empty((P, Q), dtype=torch.int64, device=dev)
        status = torch.empty((P,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().sdb_hungarian_assign_f32(
                _lib.current_stream(dev), cls_preds.data_ptr(), bbox_preds.data_ptr(), targets.gt_bboxes.data_ptr(),
                targets.gt_labels.data_ptr(), prob_seg.data_ptr(), targets.seg_offsets.data_ptr(),
                targets.img_wh.data_ptr(), cost_offsets.data_ptr(), P, Q, C, targets.max_gt,
                float(self.cls_cost.weight), float(self.reg_cost.weight), float(self.iou_cost.weight),
                workspace.data_ptr(), _lib.ptr(cost_qg), gt_inds.data_ptr(), labels.data_ptr(), status.data_ptr())
        _lib.check(rc, "hungarian_assign")
        _lib.LAUNCHES["match_cost"] += 1 if targets.max_gt > 0 else 0
        _lib.LAUNCHES["lsap_solve"] += 1
        self.last_status = status
        if return_cost:
            costs = [cost_qg[cost_offs[p]:cost_offs[p + 1]].view(Q, targets.counts[prob_img[p]]) for p in range(P)]
            return gt_inds, labels, costs
        return gt_inds, labels

    def check_status(self):
        """Raise what scipy would have raised (hungarian_assigner.py:136).  Synchronises; off the hot path."""
        if self.last_status is None:
            return
        st = self.last_status.cpu()
        if (st == 1).any():
            raise ValueError("matrix contains invalid numeric entries")
        if (st == 2).any():
            raise ValueError("cost matrix is infeasible")

    # -- reference surface --------------------------------------------------------------------------
    def assign(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, img_meta, gt_bboxes_ignore=None, eps=1e-7):
        assert gt_bboxes_ignore is None, "Only case when gt_bboxes_ignore is None is supported."
        num_gts, num_bboxes = gt_bboxes.size(0), bbox_pred.size(0)
        if num_gts == 0 or num_bboxes == 0:
            gt_inds = bbox_pred.new_full((num_bboxes,), 0 if num_gts == 0 else -1, dtype=torch.long)
            labels = bbox_pred.new_full((num_bboxes,), -1, dtype=torch.long)
            return AssignResult(num_gts, gt_inds, None, labels=labels)
        img_h, img_w, _ = img_meta["img_shape"]
        targets = MatchTargets([gt_bboxes], [gt_labels.long()], [(img_w, img_h)], bbox_pred.device)
        gt_inds, labels = self.assign_batch(bbox_pred[None], cls_pred[None], targets)
        return AssignResult(num_gts, gt_inds[0], None, labels=labels[0])
