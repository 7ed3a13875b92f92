"""``DINODETRSSODHead`` -- host-side mirror of detr_od/models/dense_heads/dino_detr_ssod_head.py:42-1581.

Two training phases keyed on ``curr_step`` (``warm_up_step`` = 60000 in the shipped config):
 * warm-up   : ``assigner1`` = O2MAssigner (top-13 one-to-many) + ``loss_cls1`` = TaskAlignedFocalLoss, box losses
               weighted by the normalised alignment metric (:665-738, :1110-1160);
 * afterwards: ``assigner2`` = HungarianAssigner + ``loss_cls2`` = FocalLoss -- the batched device path of
               ``DINODETRHead.loss`` (one cost launch + one solver launch for all layers x images).
Also here: ``forward_dummy`` (returns the decoder states and splits off the consistency / denoising parts,
:420-505), the pseudo-label decoding used on the teacher (sigmoid -> class-wise NMS 0.6, score > 0.01, top 300;
:1332-1395), and the CDN variant that tolerates images without boxes (dn_components.py:128-274).
"""
import torch

from ..consts import device_const
from ..dino.dn_components import prepare_for_cdn
from ..dino.head import LOSS_PARTS, DINODETRHead, _clamp_min1, reduce_mean_scalar
from ..dino.losses import giou_aligned
from ..matching.match_cost import bbox_cxcywh_to_xyxy, bbox_xyxy_to_cxcywh
from . import device_ops
from ..registry import BBOX_ASSIGNERS, HEADS, LOSSES
from . import o2m_assigner as _o2m  # noqa: F401  (registers O2MAssigner)
from . import task_aligned_focal_loss as _tal  # noqa: F401
from .o2m_assigner import assign_layers, normalized_alignment_metrics


def _reduce_mean_tensor(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t.div_(dist.get_world_size()))
    return t


@HEADS.register_module()
class DINODETRSSODHead(DINODETRHead):
    def __init__(self, *args, loss_cls1=dict(type="TaskAlignedFocalLoss", use_sigmoid=True, gamma=2.0, loss_weight=2.0),
                 loss_cls2=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
                 train_cfg=None, test_cfg=dict(max_per_img=300, warm_up_step=60000), **kwargs):
        train_cfg = dict(train_cfg or {})
        assigner1 = train_cfg.get("assigner1", dict(type="O2MAssigner"))
        assigner2 = train_cfg.get("assigner2", dict(type="HungarianAssigner",
                                                    cls_cost=dict(type="FocalLossCost", weight=2.0),
                                                    reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
                                                    iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0)))
        super().__init__(*args, loss_cls=loss_cls2, train_cfg=dict(assigner=assigner2), test_cfg=test_cfg, **kwargs)
        self.loss_cls1 = LOSSES.build(loss_cls1)
        self.loss_cls2 = self.loss_cls
        self.assigner1 = BBOX_ASSIGNERS.build(assigner1)
        self.assigner2 = self.assigner
        self.warm_up_step = train_cfg.get("warm_up_step", (test_cfg or {}).get("warm_up_step", 60000))
        self.in_warm_up = True
        self.train_cfg = train_cfg

    # ---------------------------------------------------------------------------------------------
    def forward_dummy(self, mlvl_feats, img_metas, input_query_label=None, input_query_bbox=None, attn_mask=None,
                      dn_meta=None):
        """-> hs, outputs_class, outputs_coord, interm_class, interm_coord, consistency_class, consistency_coord,
        dn_class, dn_coord   (dino_detr_ssod_head.py:420-505)"""
        hs, cls_all, coord_all, interm_class, interm_coord = self._decode(mlvl_feats, img_metas, input_query_label,
                                                                          input_query_bbox, attn_mask)
        if self.dn_number > 0 and dn_meta is not None:
            p1, p2 = dn_meta["pad_size_1"], dn_meta["pad_size_2"]
            return (hs, cls_all[:, :, p1 + p2:], coord_all[:, :, p1 + p2:], interm_class, interm_coord,
                    cls_all[:, :, :p1], coord_all[:, :, :p1], cls_all[:, :, p1:p1 + p2], coord_all[:, :, p1:p1 + p2])
        return hs, cls_all, coord_all, interm_class, interm_coord, None, None, None, None

    # ---------------------------------------------------------------------------------------------
    def _warmup_terms(self, cls_stack, box_stack, gt_bboxes_list, gt_labels_list, img_metas, labels_override=None):
        """O2M phase (:665-738): one-to-many assignment of every (layer, image) problem, TaskAlignedFocalLoss and box
        losses weighted by the normalised alignment metric; returns per-layer losses.  cls_stack (layers, bs, Q, C),
        box_stack (layers, bs, Q, 4).  All layers of an image are assigned in one vectorised pass and the losses of
        all layers are formed together (the reference: layers x images python-level calls, ~60 launches each)."""
        layers, bs, Q, C = cls_stack.shape
        dev = cls_stack.device
        prob = cls_stack.sigmoid()
        labels, box_t, metrics, facs = [], [], [], []
        for i in range(bs):
            gl = gt_labels_list[i].long()
            G = gt_bboxes_list[i].shape[0]
            h, w, _ = img_metas[i]["img_shape"]
            fac = device_const(dev, "whwh", (w, h), lambda: torch.tensor([w, h, w, h], dtype=torch.float32))
            if G > 0:
                gl_layers = gl[None].expand(layers, G)
                if labels_override is not None:      # the last row (encoder proposals) is matched against class 0
                    gl_layers = torch.cat([gl_layers[:-1], labels_override[i].long()[None]])
                gt_inds, _, _, nm = assign_layers(self.assigner1, box_stack[:, i].detach(), prob[:, i].detach(),
                                                  gt_bboxes_list[i], gl_layers, img_metas[i])
                pos = gt_inds > 0
                g = (gt_inds - 1).clamp(min=0)
                lab = torch.where(pos, gl_layers.gather(1, g), torch.full_like(g, self.num_classes))
                bt = bbox_xyxy_to_cxcywh(gt_bboxes_list[i] / fac)[g] * pos[..., None]
            else:
                lab = torch.full((layers, Q), self.num_classes, dtype=torch.long, device=dev)
                bt = torch.zeros((layers, Q, 4), device=dev)
                nm = torch.zeros((layers, Q), device=dev)
            labels.append(lab); box_t.append(bt); metrics.append(nm)
            facs.append(fac[None, None].expand(layers, Q, 4))
        labels, box_t = torch.stack(labels, 1), torch.stack(box_t, 1)            # (layers, bs, Q[, 4])
        metrics, facs = torch.stack(metrics, 1), torch.stack(facs, 1)
        sum_metrics = _reduce_mean_tensor(metrics.sum((1, 2))).clamp(min=1.0)    # (layers,): one all-reduce
        tal = self.loss_cls1(prob.reshape(layers, bs * Q, C), labels.reshape(layers, bs * Q),
                             metrics.reshape(layers, bs * Q), reduction_override="none")
        w = metrics[..., None]
        giou = giou_aligned(bbox_cxcywh_to_xyxy(box_stack) * facs, bbox_cxcywh_to_xyxy(box_t) * facs, self.loss_iou.eps)
        l1 = (box_stack - box_t).abs() * w
        reg_avg = sum_metrics                                                     # the same sum, the same clamp
        return dict(loss_cls=tal.sum((1, 2)) / sum_metrics,
                    loss_iou=((1 - giou) * metrics).sum((1, 2)) / reg_avg * self.loss_iou.loss_weight,
                    loss_bbox=l1.sum((1, 2, 3)) / reg_avg * self.loss_bbox.loss_weight,
                    loss_bbox_xy=l1[..., :2].sum((1, 2, 3)) / reg_avg * self.loss_bbox.loss_weight,
                    loss_bbox_hw=l1[..., 2:].sum((1, 2, 3)) / reg_avg * self.loss_bbox.loss_weight)

    def loss(self, all_cls_scores, all_bbox_preds, enc_cls_scores, enc_bbox_preds, dn_cls_scores, dn_bbox_preds,
             gt_bboxes_list, gt_labels_list, gt_scores_list=None, img_metas=None, dn_metas=None,
             gt_bboxes_ignore=None, is_pseudo_label=False):
        """dino_detr_ssod_head.py:508-625.  In the Hungarian phase ``gt_scores_list`` is accepted and -- like the
        reference's ``_get_target_single`` (:1170-1205) -- not used for the weights."""
        if not self.in_warm_up:
            dn_meta = dn_metas
            if dn_metas is not None and is_pseudo_label:
                dn_meta = dict(pad_size=dn_metas["pad_size_2"], num_dn_group=dn_metas["num_dn_group_2"])
            return super().loss(all_cls_scores, all_bbox_preds, enc_cls_scores, enc_bbox_preds, dn_cls_scores,
                                dn_bbox_preds, gt_bboxes_list, gt_labels_list, None, img_metas=img_metas,
                                dn_metas=dn_meta, gt_bboxes_ignore=gt_bboxes_ignore, zero_weight_empty_dn=True)
        # ---- warm-up: O2M matching part; DN part as in the Hungarian phase (skipped for pseudo labels, :548-553)
        L = all_cls_scores.shape[0]
        cls_stack = torch.cat([all_cls_scores.float(), enc_cls_scores.float()[None]])
        box_stack = torch.cat([all_bbox_preds.float(), enc_bbox_preds.float()[None]])
        zero_labels = [torch.zeros_like(l) for l in gt_labels_list]
        main = self._warmup_terms(cls_stack, box_stack, gt_bboxes_list, gt_labels_list, img_metas, zero_labels)
        dev = cls_stack.device
        if is_pseudo_label or dn_cls_scores is None:
            dn = {k: torch.zeros(L, device=dev) for k in LOSS_PARTS}
        else:
            dn = self._dn_terms(dn_cls_scores, dn_bbox_preds, gt_bboxes_list, gt_labels_list, img_metas, dn_metas,
                                zero_weight_empty_dn=True)
        return self._assemble(main, dn, L, True)

    # ---------------------------------------------------------------------------------------------
    def forward_train(self, x, img_metas, gt_bboxes, gt_labels, gt_scores=None, gt_bboxes_ignore=None,
                      proposal_cfg=None, curr_step=None, is_pseudo_label=False, **kwargs):
        """dino_detr_ssod_head.py:1208-1279 (CDN that tolerates empty images, phase switch on curr_step)."""
        assert proposal_cfg is None
        if curr_step is not None:
            self.in_warm_up = curr_step < self.warm_up_step
        if self.dn_number > 0:
            boxes = []
            for meta, b in zip(img_metas, gt_bboxes):
                h, w, _ = meta["img_shape"]
                fac = device_const(b.device, "whwh", (w, h), lambda: torch.tensor([w, h, w, h], dtype=torch.float32))
                boxes.append(bbox_xyxy_to_cxcywh(b) / fac)
            q_label, q_bbox, attn_mask, dn_meta = prepare_for_cdn(
                dn_args=(dict(labels=gt_labels, boxes=boxes), self.dn_number, self.dn_label_noise_ratio,
                         self.dn_box_noise_scale),
                training=True, num_queries=self.num_query, num_classes=self.num_classes, hidden_dim=self.embed_dims,
                label_enc=self.label_enc, fill_empty=True)
        else:
            q_label = q_bbox = attn_mask = dn_meta = None
        outs = self(x, img_metas, q_label, q_bbox, attn_mask, dn_meta)
        return self.loss(*outs, gt_bboxes, gt_labels, gt_scores, img_metas=img_metas, dn_metas=dn_meta,
                         gt_bboxes_ignore=gt_bboxes_ignore, is_pseudo_label=is_pseudo_label)

    # ---------------------------------------------------------------------------------------------
    def _decode_last_layer(self, feats, img_metas, rescale=False):
        """Last decoder layer -> (sigmoid scores (B, Q, C), xyxy boxes in pixels clamped to the image (B, Q, 4))
        (dino_detr_ssod_head.py:1332-1370)."""
        cls_all, coord_all = self.forward(feats, img_metas)[:2]
        cls, box = cls_all[-1], coord_all[-1]
        dev = cls.device
        whs = tuple((float(m["img_shape"][1]), float(m["img_shape"][0])) for m in img_metas)
        wh = device_const(dev, "img_wh", whs, lambda: torch.tensor(whs, dtype=torch.float32).reshape(-1, 2))
        whwh = torch.cat([wh, wh], 1)[:, None, :]                                  # (B, 1, 4)
        b = torch.minimum((bbox_cxcywh_to_xyxy(box) * whwh).clamp(min=0), whwh)
        if rescale:
            sf = tuple(tuple(float(v) for v in m["scale_factor"]) for m in img_metas)
            b = b / device_const(dev, "scale_factor", sf, lambda: torch.tensor(sf, dtype=torch.float32))[:, None, :]
        return cls.sigmoid(), b

    @torch.no_grad()
    def pseudo_label_detections(self, feats, img_metas, curr_step=None, mean_std_filter=True):
        """The teacher's pseudo boxes without leaving the device: class-wise NMS (score > 0.01, IoU 0.6, top
        ``max_per_img``; :1371-1395) and, by default, the ``score >= mean + std`` / non-degenerate filter of
        detr_ssod/models/dino_detr_ssod.py:921-939, in ONE launch for the whole batch (``sdb_pseudo_label_nms_f32``).
        -> boxes (B, max_per_img, 4), scores (B, max_per_img), labels (B, max_per_img), count (B,) int32 on the device."""
        if curr_step is not None:
            self.in_warm_up = curr_step < self.warm_up_step
        scores, b = self._decode_last_layer(feats, img_metas)
        max_per_img = (self.test_cfg or {}).get("max_per_img", self.num_query)
        ob, os_, ol, cnt, _ = device_ops.pseudo_label_nms(scores, b, 0.01, 0.6, max_per_img, mean_std_filter)
        return ob, os_, ol, cnt

    @torch.no_grad()
    def simple_test_bboxes(self, feats, img_metas, rescale=False, curr_step=None, for_pseudo_label=False):
        """Teacher decoding (:1281-1400): last decoder layer, sigmoid, class-wise NMS (warm-up / pseudo labels) or plain
        top-k.  -> list of (det_bboxes (n, 5) [x1 y1 x2 y2 score], det_labels (n,)); the list form costs one
        device->host read of the per-image counts."""
        if curr_step is not None:
            self.in_warm_up = curr_step < self.warm_up_step
        scores, b = self._decode_last_layer(feats, img_metas, rescale)
        max_per_img = (self.test_cfg or {}).get("max_per_img", self.num_query)
        results = []
        if self.in_warm_up or for_pseudo_label:
            ob, os_, ol, cnt, _ = device_ops.pseudo_label_nms(scores, b, 0.01, 0.6, max_per_img, False)
            for i, n in enumerate(cnt.tolist()):
                results.append((torch.cat([ob[i, :n], os_[i, :n, None]], 1), ol[i, :n]))
        else:
            for i in range(scores.shape[0]):
                s, idx = scores[i].reshape(-1).topk(max_per_img)
                results.append((torch.cat([b[i][idx // self.num_classes], s[:, None]], 1), idx % self.num_classes))
        return results
