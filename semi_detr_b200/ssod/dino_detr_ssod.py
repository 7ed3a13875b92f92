"""``DinoDetrSSOD`` -- the teacher-student step, host-side mirror of detr_ssod/models/dino_detr_ssod.py:75-978
(+ ``MultiSteamDetector``, multi_stream_detector.py:5-33).

Per iteration: supervised loss on the labelled images; on the unlabelled pairs the frozen
EMA teacher decodes pseudo boxes from the weak view (NMS, keep score >= mean + std), they are warped into the
strong view, the student's (no-grad) predictions are Hungarian-matched against them, the matched costs of all
images and ranks are pooled and a 2-component GMM gives the cost threshold; boxes with score >= 0.4 supervise the
student (classification / box / denoising losses), that set united with the low-cost ones seeds the cross-view
consistency queries: RoIAlign(teacher features) -> Projector -> 5 groups of queries decoded by the student
(strong view) and the teacher (weak view) and compared layer by layer.

What is different from the reference is where the work runs:
 * the per-image Hungarian matching (``cost.cpu()`` + scipy per image, :249-293) is one batched device call;
   only the matched costs (a few hundred floats, needed on the host because they decide tensor shapes) come back,
   in ONE device->host copy together with the per-image counts;
 * the teacher backbone runs once per step -- the reference recomputes ``teacher.extract_feat`` on the same weak
   images three times (:364, :598, :897);
 * the teacher's detections are decoded, NMS-ed and mean+std-filtered in one launch and the GMM cost threshold is
   fitted by a float64 EM kernel (``device_ops``, csrc/ssod.cu): two device->host reads per step remain -- the counts
   of pseudo boxes and of reliable / high-recall boxes, which decide tensor shapes.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn
from torchvision.ops import roi_align

from ..consts import device_const
from ..dino.transformer import inverse_sigmoid
from ..matching import MatchTargets
from ..matching.match_cost import bbox_xyxy_to_cxcywh
from ..registry import DETECTORS
from . import ssod_head as _ssod_head  # noqa: F401  (registers DINODETRSSODHead)
from .bbox_utils import Transform2D
from . import device_ops


class _ConvHeuristicAlgo(torch.autograd.Function):
    """3 x 3 / stride 1 / pad 1 convolution without bias whose cuDNN algorithms -- forward AND both gradients, which run
    later inside the autograd engine -- come from cuDNN's heuristics even when ``cudnn.benchmark`` is on.  For inputs whose
    batch size changes every step (the RoIs of the pseudo boxes): autotuning each new size costs hundreds of ms and
    ends with the autotuner emptying the caching allocator (measured: 0.4 - 2.2 s steps, tools/ssod_steps.py)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        with torch.backends.cudnn.flags(enabled=True, benchmark=False):
            return F.conv2d(x, w, None, 1, 1)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        with torch.backends.cudnn.flags(enabled=True, benchmark=False):
            gx, gw, _ = torch.ops.aten.convolution_backward(
                g.contiguous(), x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1,
                (ctx.needs_input_grad[0], ctx.needs_input_grad[1], False))
        return gx, gw


def _conv3x3(conv, x):
    if x.is_cuda and torch.backends.cudnn.benchmark:
        return _ConvHeuristicAlgo.apply(x, conv.weight)
    return conv(x)


class Projector(nn.Module):
    """RoI feature (256, 7, 7) -> query content (256): conv-BN-ReLU x2, FC 12544->1024, BN, ReLU, FC 1024->256,
    ReLU (dino_detr_ssod.py:33-72; same parameter names)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(256, 256, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(256)
        self.ac1 = nn.ReLU()
        self.conv2 = nn.Conv2d(256, 256, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(256)
        self.ac2 = nn.ReLU()
        self.flatten = nn.Flatten()
        self.fc1 = nn.Linear(12544, 1024)
        self.fc_relu1 = nn.ReLU()
        self.bn = nn.BatchNorm1d(1024)
        self.fc2 = nn.Linear(1024, 256)
        self.fc_relu2 = nn.ReLU()

    def forward(self, x):
        # The number of RoIs follows the number of pseudo boxes and changes from step to step: under
        # cudnn.benchmark every new batch size would be autotuned (hundreds of ms, and the autotuner empties the
        # caching allocator afterwards), so these two convolutions take cuDNN's heuristic choice (_ConvHeuristicAlgo).
        x = self.ac1(self.bn1(_conv3x3(self.conv1, x)))
        x = self.ac2(self.bn2(_conv3x3(self.conv2, x)))
        x = self.fc_relu1(self.bn(self.fc1(self.flatten(x))))
        return self.fc_relu2(self.fc2(x))


def single_level_roi_extract(feats, rois, strides=(8, 16, 32, 64), out=7, finest_scale=56):
    """mmdet ``SingleRoIExtractor`` + mmcv ``RoIAlign(output_size=7, sampling_ratio=0, aligned=True)``
    (dino_detr_ssod.py:97-101): each RoI is pooled from the level picked by its scale."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvls = torch.floor(torch.log2(scale / finest_scale + 1e-6)).clamp(0, len(feats) - 1).long()
    res = feats[0].new_zeros(rois.size(0), feats[0].size(1), out, out)
    for i, (f, s) in enumerate(zip(feats, strides)):
        pooled = roi_align(f, rois, (out, out), spatial_scale=1.0 / s, sampling_ratio=0, aligned=True)
        res = torch.where((lvls == i)[:, None, None, None], pooled, res)
    return res


def pooled_cost_segments(t, max_len=4096):
    """``concat_all_gather`` (detr_ssod/models/utils/dist_utils.py:5-30) of a 1-D tensor without a shape exchange and
    without reading anything back: ONE fixed-size padded all-gather with each rank's length in slot 0.
    -> (costs, seg_counts, seg_stride) as ``device_ops.gmm_threshold`` takes them: rank r's values sit at
    ``costs[r * seg_stride : r * seg_stride + seg_counts[r]]`` (single process: ``(t, None, None)``).
    A rank with more than ``max_len`` values is an error, not a silent truncation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t, None, None
    if t.numel() > max_len:
        raise ValueError(f"{t.numel()} matched costs on this rank exceed the all-gather buffer ({max_len})")
    world = dist.get_world_size()
    buf = t.new_zeros(max_len + 1)
    buf[0] = t.numel()
    buf[1:1 + t.numel()] = t
    out = t.new_zeros(world * (max_len + 1))
    dist.all_gather_into_tensor(out, buf)
    counts = out.view(world, max_len + 1)[:, 0].to(torch.int32)
    return out[1:], counts, max_len + 1


def concat_all_gather_1d(t, max_len=4096):
    """The pooled values as one compact tensor (reads the per-rank counts back: diagnostics and tests only)."""
    costs, counts, stride = pooled_cost_segments(t, max_len)
    if counts is None:
        return costs
    return torch.cat([costs[r * stride:r * stride + n] for r, n in enumerate(counts.tolist())])


def dict_split(data, tags):
    """structure_utils.py:49-53: split every entry of ``data`` by the per-sample tag."""
    groups = {}
    for tag in sorted(set(tags)):
        idx = [i for i, t in enumerate(tags) if t == tag]
        g = {}
        for k, v in data.items():
            if torch.is_tensor(v):
                g[k] = v[torch.as_tensor(idx, device=v.device)] if len(idx) != v.shape[0] else v
            else:
                g[k] = [v[i] for i in idx]
        groups[tag] = g
    return groups


def weighted_loss(loss, weight):
    """structure_utils.py:132-150 for a scalar weight: entries whose key contains 'loss' are scaled."""
    return {k: (v * weight if "loss" in k else v) for k, v in loss.items()}


@DETECTORS.register_module()
class DinoDetrSSOD(nn.Module):
    def __init__(self, model, train_cfg=None, test_cfg=None):
        super().__init__()
        self.teacher = DETECTORS.build(model)
        self.student = DETECTORS.build(model)
        self.submodules = ["teacher", "student"]
        self.train_cfg = dict(train_cfg or {})
        self.test_cfg = dict(test_cfg or {})
        self.inference_on = self.test_cfg.get("inference_on", "teacher")
        if train_cfg is not None:
            self.freeze("teacher")
            self.unsup_weight = self.train_cfg["unsup_weight"]
        self.covariance_type = "diag"
        self.curr_step = 0
        self.projector = Projector()
        self.featmap_strides = (8, 16, 32, 64)

    def freeze(self, name):
        m = getattr(self, name)
        m.eval()
        for p in m.parameters():
            p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        if self.train_cfg:
            self.teacher.eval()          # the EMA teacher never runs in training mode
        return self

    # ------------------------------------------------------------------------------------------------
    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if not return_loss:
            raise NotImplementedError("inference path is outside the train-step hot path")
        return self.forward_train(img, img_metas, **kwargs)

    def forward_train(self, img, img_metas, **kwargs):
        """dino_detr_ssod.py:112-152"""
        batch_input_shape = tuple(img.shape[-2:])
        for m in img_metas:
            m["batch_input_shape"] = batch_input_shape
        data = dict(kwargs)
        data.update(img=img, img_metas=img_metas)
        groups = dict_split(data, [m["tag"] for m in img_metas])
        loss = {}
        if "sup" in groups:
            g = groups["sup"]
            x = self.student.extract_feat(g["img"])
            sup = self.student.bbox_head.forward_train(x, g["img_metas"], g["gt_bboxes"], g["gt_labels"],
                                                       curr_step=self.curr_step)
            loss.update({"sup_" + k: v for k, v in sup.items()})
        if "unsup_student" in groups:
            unsup = weighted_loss(self.foward_unsup_train(groups["unsup_teacher"], groups["unsup_student"]),
                                  self.unsup_weight)
            loss.update({"unsup_" + k: v for k, v in unsup.items()})
        return loss

    _parse_losses = staticmethod(lambda losses, reduce_log_vars=False: DETECTORS.get("DINODETR")._parse_losses(
        losses, reduce_log_vars))

    # ------------------------------------------------------------------------------------------------
    def foward_unsup_train(self, teacher_data, student_data):
        """dino_detr_ssod.py:154-201 (name kept, typo included)."""
        tnames = [m["filename"] for m in teacher_data["img_metas"]]
        snames = [m["filename"] for m in student_data["img_metas"]]
        tidx = [tnames.index(n) for n in snames]
        with torch.no_grad():
            timg = teacher_data["img"]
            if tidx != list(range(len(tidx))):
                timg = timg[torch.as_tensor(tidx, device=timg.device)]
            teacher_info = self.extract_teacher_info(timg, [teacher_data["img_metas"][i] for i in tidx])
        student_info = self.extract_student_info(**student_data)
        M = [bt @ at.inverse() for bt, at in zip(student_info["transform_matrix"], teacher_info["transform_matrix"])]
        pseudo_bboxes = Transform2D.transform_bboxes(teacher_info["det_bboxes"], M,
                                                     [m["img_shape"] for m in student_info["img_metas"]])
        return self.unsup_loss(student_info, teacher_info, pseudo_bboxes, teacher_info["det_labels"],
                               teacher_info["det_scores"])

    @torch.no_grad()
    def extract_teacher_info(self, img, img_metas):
        """Teacher pseudo labels (:893-951): class-wise NMS detections, keep score >= mean + std and w, h > 0."""
        info = dict(img=img, img_metas=img_metas)
        head = self.teacher.bbox_head
        det_bboxes, det_labels, det_scores = [], [], []
        if hasattr(head, "pseudo_label_detections"):
            # backbone + transformer + decode + class-wise NMS + mean/std filter for the whole batch: shapes depend on
            # the batch geometry only, so the section replays from a CUDA graph (engine.GraphedNoGrad); the per-image
            # counts decide tensor shapes from here on: ONE device->host read
            key = ("teacher", tuple((tuple(m["img_shape"]), tuple(m.get("batch_input_shape", ()))) for m in img_metas),
                   bool(self.curr_step < head.warm_up_step) if self.curr_step is not None else None)
            if self.curr_step is not None:
                head.in_warm_up = self.curr_step < head.warm_up_step     # host state the replayed section would not set
            feat, boxes, scores, labels, counts = self._graphed_teacher()(key, [img], img_metas, self.curr_step)
            info["backbone_feature"] = feat
            for i, n in enumerate(counts.tolist()):
                det_bboxes.append(boxes[i, :n])
                det_labels.append(labels[i, :n])
                det_scores.append(scores[i, :n])
        else:   # any head that only offers the reference's list interface
            feat = self.teacher.extract_feat(img)
            info["backbone_feature"] = feat
            proposals = head.simple_test_bboxes(feat, img_metas, rescale=False, curr_step=self.curr_step,
                                                for_pseudo_label=True)
            for boxes, labels in proposals:
                if boxes.shape[0] == 0:
                    boxes = boxes.new_zeros(0, 5)
                s = boxes[:, -1]
                thr = s.mean() + s.std()
                keep = (s >= thr) & (boxes[:, 2] - boxes[:, 0] > 0) & (boxes[:, 3] - boxes[:, 1] > 0)
                det_bboxes.append(boxes[keep, :4])
                det_labels.append(labels[keep])
                det_scores.append(boxes[keep, 4])
        info.update(det_bboxes=det_bboxes, det_labels=det_labels, det_scores=det_scores)
        info["transform_matrix"] = [torch.as_tensor(np.asarray(m["transform_matrix"]), dtype=torch.float32,
                                                    device=img.device) for m in img_metas]
        return info

    def _graphed_teacher(self):
        g = getattr(self, "_teacher_graphs", None)
        if g is None:
            from ..engine import GraphedNoGrad
            g = GraphedNoGrad(self._teacher_pass)
            object.__setattr__(self, "_teacher_graphs", g)
        return g

    def _teacher_pass(self, x, metas, curr_step):
        head = self.teacher.bbox_head
        f = self.teacher.extract_feat(x)
        return (tuple(f),) + tuple(head.pseudo_label_detections(f, metas, curr_step=curr_step))

    def extract_student_info(self, img, img_metas, **kwargs):
        """:813-830 -- student features (with grad) and its no-grad predictions on the strong view."""
        info = dict(img=img, img_metas=img_metas)
        feat = self.student.extract_feat(img)
        info["backbone_feature"] = feat
        with torch.no_grad():
            # the student's own predictions on the strong view (no gradient, shapes fixed by the geometry): graph replay
            g = getattr(self, "_student_graphs", None)
            if g is None:
                from ..engine import GraphedNoGrad
                g = GraphedNoGrad(lambda *a: self.student.bbox_head.forward(tuple(a[:-1]), a[-1]))
                object.__setattr__(self, "_student_graphs", g)
            key = ("student", tuple((tuple(m["img_shape"]), tuple(m.get("batch_input_shape", ()))) for m in img_metas))
            info["outs"] = g(key, [f.detach() for f in feat], img_metas)
        info["transform_matrix"] = [torch.as_tensor(np.asarray(m["transform_matrix"]), dtype=torch.float32,
                                                    device=img.device) for m in img_metas]
        return info

    def _fit_gmm(self, costs, seg_counts=None, seg_stride=None):
        """Cost threshold of dino_detr_ssod.py:832-890 as a 0-d DEVICE tensor (``sdb_gmm_threshold_f32``)."""
        return device_ops.gmm_threshold(costs, seg_counts, seg_stride)[0]

    # ------------------------------------------------------------------------------------------------
    def unsup_loss(self, student_info, teacher_info, pseudo_bboxes, pseudo_labels, pseudo_scores):
        """dino_detr_ssod.py:204-482"""
        head = self.student.bbox_head
        img_metas_v1, img_metas_v2 = student_info["img_metas"], teacher_info["img_metas"]
        dev = student_info["img"].device
        bs = len(img_metas_v1)
        cls_last, box_last = student_info["outs"][0][-1], student_info["outs"][1][-1]

        # 1. Hungarian-match the student's predictions to the pseudo boxes: one batched device call (:249-293)
        counts = [int(b.shape[0]) for b in pseudo_bboxes]
        with torch.no_grad():
            if sum(counts) > 0:
                t = MatchTargets(pseudo_bboxes, pseudo_labels, [(m["img_shape"][1], m["img_shape"][0]) for m in img_metas_v1], dev)
                gt_inds, _, costs = head.assigner2.assign_batch(box_last, cls_last, t, prob_img=list(range(bs)),
                                                                return_cost=True)
                matched_cost, matched_gt = [], []
                for i in range(bs):
                    # the min(count, Q) matched rows in ascending order, like scipy's row_ind -- their number is known
                    # on the host, so a stable sort of the mask replaces nonzero() (no synchronisation)
                    k = min(counts[i], gt_inds.shape[1])
                    rows = torch.sort((gt_inds[i] > 0).to(torch.uint8), descending=True, stable=True).indices[:k]
                    cols = gt_inds[i][rows] - 1
                    matched_gt.append(cols)
                    matched_cost.append(costs[i][rows, cols] if counts[i] else box_last.new_zeros(0))
                cost_all = torch.cat(matched_cost)
            else:
                matched_cost = [box_last.new_zeros(0) for _ in range(bs)]
                matched_gt = [box_last.new_zeros(0, dtype=torch.long) for _ in range(bs)]
                cost_all = box_last.new_zeros(0)
            thr = self._fit_gmm(*pooled_cost_segments(cost_all.detach()))

        # 2. double filter (:324-353): reliable = score >= 0.4; high-recall = reliable U {matched cost <= thr}
        base_thr = self.train_cfg["pseudo_label_initial_score_thr"]
        assert isinstance(base_thr, float), "Dynamic Threshold is not implemented yet."
        gt_b, gt_l, gt_s, hr_b, hr_l, det_b, det_l = [], [], [], [], [], [], []
        rel_masks, keep_masks = [], []
        for i in range(bs):
            reliable = pseudo_scores[i] >= base_thr
            low_cost = torch.zeros_like(reliable)
            if matched_gt[i].numel():
                low_cost[matched_gt[i]] = matched_cost[i] <= thr          # distinct indices: a plain scatter
            rel_masks.append(reliable)
            keep_masks.append(reliable | low_cost)
        # the sizes of the two sets decide tensor shapes: ONE device->host read for all images, then order-preserving
        # compaction with a stable sort instead of boolean indexing (which would synchronise per use)
        sizes = torch.stack([m.sum() for m in rel_masks + keep_masks]).tolist() if bs else []

        def take(mask, k, *tensors):
            idx = torch.sort(mask.to(torch.uint8), descending=True, stable=True).indices[:k]
            return [t[idx] for t in tensors]
        for i in range(bs):
            b, l, sc = take(rel_masks[i], int(sizes[i]), pseudo_bboxes[i][:, :4], pseudo_labels[i], pseudo_scores[i])
            gt_b.append(b); gt_l.append(l); gt_s.append(sc)
            b, l, tb, tl = take(keep_masks[i], int(sizes[bs + i]), pseudo_bboxes[i][:, :4], pseudo_labels[i],
                                teacher_info["det_bboxes"][i][:, :4], teacher_info["det_labels"][i])
            hr_b.append(b); hr_l.append(l); det_b.append(tb); det_l.append(tl)

        head.in_warm_up = self.curr_step < head.warm_up_step
        self.teacher.bbox_head.in_warm_up = head.in_warm_up
        teacher_feat = teacher_info["backbone_feature"]          # computed once; the reference recomputes it twice

        # 3. student pass on the strong view with [consistency | denoising | matching] queries (:377-411)
        q1_label, q1_bbox, q2_label, q2_bbox, mask1, meta1 = self.prepare_unsup_cdn(
            teacher_info, student_info, hr_b, hr_l, det_b, det_l, self._dn_args(gt_b, gt_l, img_metas_v1),
            teacher_feat=teacher_feat)
        outs1 = head.forward_dummy(student_info["backbone_feature"], img_metas_v1,
                                   torch.cat([q1_label, q2_label], 1), torch.cat([q1_bbox, q2_bbox], 1), mask1, meta1)
        hs_v1 = outs1[0]
        losses = head.loss(outs1[1], outs1[2], outs1[3], outs1[4], dn_cls_scores=outs1[7], dn_bbox_preds=outs1[8],
                           gt_bboxes_list=gt_b, gt_labels_list=gt_l, gt_scores_list=gt_s, img_metas=img_metas_v1,
                           dn_metas=meta1, is_pseudo_label=True)

        # 4. teacher pass on the weak view with the SAME consistency content (:413-456)
        with torch.no_grad():
            prior = dict(loss_weights=meta1["loss_weights"], input_query_label_1=q1_label)
            p1_label, p1_bbox, p2_label, p2_bbox, mask2, meta2 = self.prepare_unsup_cdn(
                teacher_info, teacher_info, det_b, det_l, det_b, det_l, self._dn_args(det_b, det_l, img_metas_v2),
                prior_info=prior, teacher_feat=teacher_feat)
            outs2 = self.teacher.bbox_head.forward_dummy(teacher_feat, img_metas_v2,
                                                         torch.cat([p1_label, p2_label], 1),
                                                         torch.cat([p1_bbox, p2_bbox], 1), mask2, meta2)
        hs_v2 = outs2[0]

        # 5. cross-view consistency on the first pad_size_1 queries of every decoder layer (:458-481)
        assert meta1["pad_size_1"] == meta2["pad_size_1"]
        pad1 = meta1["pad_size_1"]
        bid, slot = meta1["known_bid_1"], meta1["map_known_indice_1"]
        w = meta1["loss_weights"]
        if self.curr_step >= head.warm_up_step:
            w = torch.zeros_like(w)
        for l in range(len(hs_v1)):
            h1 = hs_v1[l][:, :pad1][bid, slot]
            h2 = hs_v2[l][:, :pad1][bid, slot]
            mse = F.mse_loss(F.normalize(h1, p=2, dim=-1), F.normalize(h2, p=2, dim=-1).detach(), reduction="none")
            losses[f"consis_loss.d{l}"] = 10 * (mse * w.unsqueeze(-1)).mean()
        return losses

    def _dn_args(self, boxes, labels, img_metas):
        head = self.student.bbox_head
        norm = []
        for meta, b in zip(img_metas, boxes):
            h, w, _ = meta["img_shape"]
            fac = device_const(b.device, "whwh", (w, h), lambda: torch.tensor([w, h, w, h], dtype=torch.float32))
            norm.append(bbox_xyxy_to_cxcywh(b) / fac)
        return (dict(labels=labels, boxes=norm), head.dn_number, head.dn_label_noise_ratio, head.dn_box_noise_scale)

    # ------------------------------------------------------------------------------------------------
    def prepare_unsup_cdn(self, teacher_info, student_info, pseudo_bboxes, pseudo_labels, det_bboxes, det_labels,
                          dn_args=None, hidden_dim=256, num_queries=900, num_classes=80, prior_info=None,
                          teacher_feat=None):
        """Consistency queries + CDN queries + attention mask (dino_detr_ssod.py:484-760).

        Part 1 (5 groups, no noise): anchors = pseudo boxes in the target view; content =
        Projector(RoIAlign(teacher features, boxes in the teacher view)), reused for the teacher pass through
        ``prior_info``.  Part 2: the usual contrastive denoising queries of the reliable boxes.  Mask layout:
        [part 1 | part 2 | matching]; matching sees neither part; every group of either part sees only itself and
        the matching part (:723-744)."""
        from ..dino.dn_components import prepare_for_cdn
        head = self.student.bbox_head
        num_queries, hidden_dim, num_classes = head.num_query, head.embed_dims, head.num_classes
        metas_tgt = student_info["img_metas"]
        dev = student_info["img"].device
        bs = len(metas_tgt)
        # ---- part 1: consistency queries ---------------------------------------------------------------
        norm_boxes, counts = [], []
        for meta, b in zip(metas_tgt, pseudo_bboxes):
            h, w, _ = meta["img_shape"]
            if b.size(0) == 0:
                b = device_const(dev, "center_box", (w, h), lambda: torch.tensor([[w / 4, h / 4, 3 * w / 4, 3 * h / 4]],
                                                                                  dtype=torch.float32))
            fac = device_const(dev, "whwh", (w, h), lambda: torch.tensor([w, h, w, h], dtype=torch.float32))
            norm_boxes.append((bbox_xyxy_to_cxcywh(b) / fac).clamp(0.0, 1.0))
            counts.append(int(b.size(0)))
        groups1 = 5
        single1 = max(counts)
        pad1 = single1 * groups1

        def build_idx():
            bid = np.concatenate([np.full(c, i, dtype=np.int64) for i, c in enumerate(counts)])
            within = np.concatenate([np.arange(c, dtype=np.int64) for c in counts])
            return np.stack([np.tile(bid, groups1), np.concatenate([within + single1 * g for g in range(groups1)])])
        idx = device_const(dev, "consis_idx", (tuple(counts), groups1), build_idx)
        bid1, slot1 = idx[0], idx[1]
        q1_bbox = torch.zeros(bs, pad1, 4, device=dev)
        q1_bbox[bid1, slot1] = inverse_sigmoid(torch.cat(norm_boxes).repeat(groups1, 1))
        if prior_info is None:
            props, lw = [], []
            for i, b in enumerate(det_bboxes):
                if b.size(0) == 0:
                    h, w, _ = teacher_info["img_metas"][i]["img_shape"]
                    props.append(device_const(dev, "center_box", (w, h), lambda: torch.tensor(
                        [[w / 4, h / 4, 3 * w / 4, 3 * h / 4]], dtype=torch.float32)))
                    lw.append(b.new_zeros(1))
                else:
                    props.append(b[:, :4])
                    lw.append(b.new_ones(b.size(0)))
            loss_weights = torch.cat(lw).unsqueeze(-1).repeat(groups1, 1)
            rois = torch.cat([bid1.unsqueeze(-1).float(), torch.cat(props).repeat(groups1, 1)], dim=-1)
            with torch.no_grad():
                feats = teacher_feat if teacher_feat is not None else self.teacher.extract_feat(teacher_info["img"])
                srcs = self._project_feats(feats)
                roi_feat = single_level_roi_extract(srcs, rois, self.featmap_strides)
            q1_label = torch.zeros(bs, pad1, hidden_dim, device=dev)
            q1_label[bid1, slot1] = self.projector(roi_feat)
        else:
            loss_weights = prior_info["loss_weights"]
            q1_label = prior_info["input_query_label_1"]
        # ---- part 2: contrastive denoising queries (empty images get a dummy box, :613-626) -------------------
        q2_label, q2_bbox, mask2, m2 = prepare_for_cdn(dn_args, True, num_queries, num_classes, hidden_dim,
                                                       head.label_enc, fill_empty=True)
        pad2, groups2 = m2["pad_size"], m2["num_dn_group"]
        attn_mask = self._ssod_mask(single1, groups1, pad2, groups2, num_queries, dev)
        in_warm = self.curr_step < head.warm_up_step
        dn_meta = dict(pad_size_1=pad1, pad_size_2=pad2, num_dn_group_1=groups1, num_dn_group_2=groups2,
                       known_bid_1=bid1, map_known_indice_1=slot1,
                       loss_weights=loss_weights if in_warm else torch.zeros_like(loss_weights))
        return q1_label, q1_bbox, q2_label, q2_bbox, attn_mask, dn_meta

    def _project_feats(self, feats):
        """Teacher ``input_proj`` over the backbone levels (+ the stride-64 level), the feature pyramid the RoI
        extractor pools from (``prepare_feats``, :762-802)."""
        head = self.teacher.bbox_head
        srcs = [head.input_proj[l](f) for l, f in enumerate(feats)]
        for l in range(len(srcs), head.num_feature_levels):
            srcs.append(head.input_proj[l](feats[-1] if l == len(feats) else srcs[-1]))
        return srcs

    @staticmethod
    def _ssod_mask(single1, groups1, pad2, groups2, num_queries, device):
        def build():
            pad1 = single1 * groups1
            size = pad1 + pad2 + num_queries
            pos = torch.arange(size)
            part = torch.where(pos < pad1, 0, torch.where(pos < pad1 + pad2, 1, 2))
            g1 = pos // max(single1, 1)
            g2 = (pos - pad1) // max(pad2 // max(groups2, 1), 1)
            m = torch.zeros(size, size, dtype=torch.bool)
            P0, P1, P2 = (part == 0), (part == 1), (part == 2)
            m |= P2[:, None] & ~P2[None, :]                                   # matching part sees no dn query
            m |= P0[:, None] & P0[None, :] & (g1[:, None] != g1[None, :])     # consistency groups are isolated
            m |= P0[:, None] & P1[None, :]                                    # ... and do not see part 2
            m |= P1[:, None] & P0[None, :]                                    # part 2 does not see part 1
            m |= P1[:, None] & P1[None, :] & (g2[:, None] != g2[None, :])     # denoising groups are isolated
            return m
        return device_const(device, "ssod_mask", (single1, groups1, pad2, groups2, num_queries), build)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        """A plain detector checkpoint initialises both teacher and student (:953-978)."""
        if not any("student" in k or "teacher" in k for k in state_dict.keys()):
            keys = list(state_dict.keys())
            state_dict.update({"teacher." + k: state_dict[k] for k in keys})
            state_dict.update({"student." + k: state_dict[k] for k in keys})
            for k in keys:
                state_dict.pop(k)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
