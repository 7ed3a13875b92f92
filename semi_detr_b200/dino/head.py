"""``DINODETRHead`` -- host-side mirror of detr_od/models/dense_heads/dino_detr_head.py:39-1046.

Module tree / parameter names (``input_proj``, ``fc_cls``, ``fc_reg``, ``fc_enc_cls``, ``fc_enc_reg``, ``label_enc``,
``transformer.*``), ``forward`` / ``forward_train`` / ``loss`` signatures and the 65-key loss dict are the
reference's.  The body of ``loss`` is re-designed for the device:

 * all (decoder layer + encoder) x image Hungarian problems of the step go through ONE cost kernel and ONE solver
   kernel (``HungarianAssigner.assign_batch``) -- the reference runs ``multi_apply(_get_target_single)`` with a
   ``.cpu()`` + scipy call per problem (dino_detr_head.py:895-980, 937);
 * targets are gathered and the focal / L1 / GIoU terms evaluated for all layers in one batched pass and reduced
   per layer, instead of 13 ``loss_single`` calls (dino_detr_head.py:546-582, 634-736);
 * normalisers come from the host-known GT counts (every GT is matched when num_query >= num_gt), so no ``.item()``;
   under data parallelism the one cross-rank mean (dino_detr_head.py:698-699, 720-723) is a single tiny all-reduce
   whose result stays on the device.
The arithmetic per element and per normaliser is the reference's.
"""
import copy
import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from ..consts import device_const
from ..matching import HungarianAssigner, MatchTargets  # noqa: F401  (registers the assigner)
from ..matching.match_cost import bbox_cxcywh_to_xyxy, bbox_xyxy_to_cxcywh
from ..registry import BBOX_ASSIGNERS, HEADS, LOSSES, POSITIONAL_ENCODING, TRANSFORMER
from . import losses as _losses  # noqa: F401  (registers the losses)
from . import positional_encoding as _pe  # noqa: F401
from .dn_components import dn_post_process, prepare_for_cdn
from . import fused_loss
from .transformer import MLP, inverse_sigmoid

LOSS_PARTS = ("loss_cls", "loss_bbox", "loss_iou", "loss_bbox_xy", "loss_bbox_hw")


class LossDict(dict):
    """The reference's loss dict (name -> scalar tensor) plus ``total``: the sum of every entry whose name contains
    "loss" -- what mmdet's ``_parse_losses`` (base.py:176-209) computes entry by entry -- precomputed by the head.
    Adding or replacing an entry afterwards drops ``total`` (the slow entry-by-entry sum is then used)."""
    total = None

    def __setitem__(self, key, value):
        self.total = None
        super().__setitem__(key, value)

    def update(self, *args, **kwargs):
        self.total = None
        super().update(*args, **kwargs)


_SCALAR_ALLREDUCE = None


def set_scalar_allreduce(fn):
    """``fn(tensor, slot) -> tensor`` (in-place sum over ranks of a 1-element device tensor) replaces the NCCL call of
    ``reduce_mean_scalar`` -- the engine installs the peer-memory kernel (``sdb_dp_small_allreduce_f32``) when the
    gradient exchange runs over NVLink multicast; ``None`` restores NCCL."""
    global _SCALAR_ALLREDUCE
    _SCALAR_ALLREDUCE = fn


def reduce_mean_scalar(value, device, slot=0):
    """mmdet core/utils/dist_utils.py:67-73 for a host scalar: mean over ranks, kept on the device (no .item()).
    ``slot`` tells independent call sites of one step apart (they may be in flight on different ranks at once)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    v = float(value)
    t = device_const(device, "scalar", v, lambda: torch.tensor([v], dtype=torch.float32)).clone()   # graph-safe
    t.div_(dist.get_world_size())
    if _SCALAR_ALLREDUCE is not None:
        return _SCALAR_ALLREDUCE(t, slot)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def _clamp_min1(x):
    return x.clamp(min=1) if torch.is_tensor(x) else max(x, 1)


@HEADS.register_module()
class DINODETRHead(nn.Module):
    def __init__(self, num_classes=80, in_channels=2048, num_query=900, num_reg_fcs=2, transformer=None,
                 num_feature_levels=4, num_backbone_outs=3, backbone_channels=(512, 1024, 2048),
                 sync_cls_avg_factor=False, iter_update=True, dn_number=100, dn_box_noise_scale=0.4,
                 dn_label_noise_ratio=0.5, dn_labelbook_size=81, query_dim=4, dec_pred_class_embed_share=True,
                 dec_pred_bbox_embed_share=True, two_stage_bbox_embed_share=False, two_stage_class_embed_share=False,
                 bbox_embed_diff_each_layer=False, random_refpoints_xy=False,
                 positional_encoding=dict(type="SinePositionalEncodingHW", num_feats=128, normalize=True,
                                          temperatureH=20, temperatureW=20),
                 loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
                 loss_bbox=dict(type="L1Loss", loss_weight=5.0), loss_iou=dict(type="GIoULoss", loss_weight=2.0),
                 train_cfg=dict(assigner=dict(type="HungarianAssigner",
                                              cls_cost=dict(type="FocalLossCost", weight=2.0),
                                              reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
                                              iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))),
                 test_cfg=dict(max_per_img=300), init_cfg=None, **kwargs):
        super().__init__()
        assert dec_pred_class_embed_share and dec_pred_bbox_embed_share and not bbox_embed_diff_each_layer
        assert not two_stage_bbox_embed_share and not two_stage_class_embed_share and not sync_cls_avg_factor
        self.bg_cls_weight = 0
        if train_cfg:
            assigner = train_cfg["assigner"]
            # the reference asserts loss and matcher weights agree (dino_detr_head.py:130-141)
            assert loss_cls["loss_weight"] == assigner["cls_cost"]["weight"]
            assert loss_bbox["loss_weight"] == assigner["reg_cost"]["weight"]
            assert loss_iou["loss_weight"] == assigner["iou_cost"]["weight"]
            self.assigner = BBOX_ASSIGNERS.build(assigner)
        self.num_query, self.num_classes, self.in_channels = num_query, num_classes, in_channels
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.loss_cls, self.loss_bbox, self.loss_iou = (LOSSES.build(c) for c in (loss_cls, loss_bbox, loss_iou))
        self.num_feature_levels, self.num_backbone_outs = num_feature_levels, num_backbone_outs
        self.backbone_channels = list(backbone_channels)
        self.query_dim = query_dim
        self.dn_number, self.dn_box_noise_scale = dn_number, dn_box_noise_scale
        self.dn_label_noise_ratio, self.dn_labelbook_size = dn_label_noise_ratio, dn_labelbook_size
        self.cls_out_channels = num_classes if self.loss_cls.use_sigmoid else num_classes + 1
        self.positional_encoding = POSITIONAL_ENCODING.build(positional_encoding)
        self.transformer = TRANSFORMER.build(transformer or dict(type="DINOTransformer"))
        self.embed_dims = self.transformer.embed_dims
        assert positional_encoding["num_feats"] * 2 == self.embed_dims
        self._init_layers()
        self.init_weights()

    def _init_layers(self):
        """dino_detr_head.py:215-282"""
        proj = []
        in_ch = self.backbone_channels[-1]
        for i in range(self.num_backbone_outs):
            in_ch = self.backbone_channels[i]
            proj.append(nn.Sequential(nn.Conv2d(in_ch, self.embed_dims, kernel_size=1),
                                      nn.GroupNorm(32, self.embed_dims)))
        for _ in range(self.num_feature_levels - self.num_backbone_outs):
            proj.append(nn.Sequential(nn.Conv2d(in_ch, self.embed_dims, kernel_size=3, stride=2, padding=1),
                                      nn.GroupNorm(32, self.embed_dims)))
            in_ch = self.embed_dims
        self.input_proj = nn.ModuleList(proj)
        cls_embed = nn.Linear(self.embed_dims, self.cls_out_channels)
        box_embed = MLP(self.embed_dims, self.embed_dims, 4, 3)
        cls_embed.bias.data = torch.ones(self.cls_out_channels) * (-math.log((1 - 0.01) / 0.01))
        nn.init.constant_(box_embed.layers[-1].weight.data, 0)
        nn.init.constant_(box_embed.layers[-1].bias.data, 0)
        n_dec = self.transformer.num_decoder_layers
        self.fc_reg = nn.ModuleList([box_embed for _ in range(n_dec)])      # shared across layers
        self.fc_cls = nn.ModuleList([cls_embed for _ in range(n_dec)])
        self.fc_enc_reg = copy.deepcopy(box_embed)
        self.fc_enc_cls = copy.deepcopy(cls_embed)
        self.label_enc = nn.Embedding(self.dn_labelbook_size + 1, self.embed_dims)

    def init_weights(self):
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)

    # ------------------------------------------------------------------------------------------------
    def _decode(self, mlvl_feats, img_metas, input_query_label=None, input_query_bbox=None, attn_mask=None):
        """Shared body of ``forward`` / ``forward_dummy``: -> hs (list of (bs, nq, C)), class logits and boxes of
        every decoder layer for ALL queries (denoising part included), encoder proposals' class / box."""
        bs = mlvl_feats[0].size(0)
        dev = mlvl_feats[0].device
        in_h, in_w = img_metas[0]["batch_input_shape"]
        shapes_key = (in_h, in_w, tuple(tuple(m["img_shape"][:2]) for m in img_metas))

        def build_img_masks():
            m = torch.ones((bs, in_h, in_w), device=dev)
            for i in range(bs):
                h, w, _ = img_metas[i]["img_shape"]
                m[i, :h, :w] = 0
            return m

        def mask_and_pos(size):
            # padding masks and sine embeddings depend only on the image / feature geometry: built once per
            # geometry (the reference rebuilds them every pass, dino_detr_head.py:321-347)
            size = tuple(size)
            mask = device_const(dev, "lvl_mask", (shapes_key, size), lambda: F.interpolate(
                device_const(dev, "img_masks", shapes_key, build_img_masks)[None], size=size).to(torch.bool).squeeze(0))
            pos = device_const(dev, "lvl_pos", (shapes_key, size, id(self.positional_encoding)),
                               lambda: self.positional_encoding(mask))
            return mask, pos

        srcs, masks, poss = [], [], []
        for lvl, feat in enumerate(mlvl_feats):
            mk, ps = mask_and_pos(feat.shape[-2:])
            masks.append(mk)
            poss.append(ps)
            srcs.append(self.input_proj[lvl](feat))
        for lvl in range(len(srcs), self.num_feature_levels):
            src = self.input_proj[lvl](mlvl_feats[-1] if lvl == len(mlvl_feats) else srcs[-1])
            srcs.append(src)
            mk, ps = mask_and_pos(src.shape[-2:])
            masks.append(mk)
            poss.append(ps)

        hs, reference, hs_enc, ref_enc, _ = self.transformer(
            srcs, masks, input_query_bbox, poss, input_query_label, attn_mask, fc_reg=self.fc_reg,
            fc_cls=self.fc_cls, fc_enc_reg=self.fc_enc_reg, fc_enc_cls=self.fc_enc_cls, geometry_key=shapes_key)
        hs[0] = hs[0] + self.label_enc.weight[0, 0] * 0.0      # keeps label_enc in the graph without a DN part

        hs_all = torch.stack(hs)                                # (n_dec, bs, nq, C); heads are shared across layers
        ref_all = torch.stack(reference[:-1])
        outputs_coord = (self.fc_reg[0](hs_all) + inverse_sigmoid(ref_all)).sigmoid()
        outputs_class = self.fc_cls[0](hs_all)
        interm_coord = ref_enc[-1]
        interm_class = self.fc_enc_cls(hs_enc[-1])
        return hs, outputs_class, outputs_coord, interm_class, interm_coord

    def forward(self, mlvl_feats, img_metas, input_query_label=None, input_query_bbox=None, attn_mask=None,
                dn_meta=None):
        """dino_detr_head.py:314-407 -> (outputs_class (n_dec, bs, Q, C), outputs_coord (n_dec, bs, Q, 4),
        interm_outputs_class (bs, Q, C), interm_outputs_coord (bs, Q, 4), dn_outputs_class, dn_outputs_coord)"""
        _, outputs_class, outputs_coord, interm_class, interm_coord = self._decode(
            mlvl_feats, img_metas, input_query_label, input_query_bbox, attn_mask)
        if self.dn_number > 0 and dn_meta is not None:
            outputs_class, outputs_coord, dn_class, dn_coord = dn_post_process(outputs_class, outputs_coord, dn_meta)
        else:
            dn_class, dn_coord = None, None
        return outputs_class, outputs_coord, interm_class, interm_coord, dn_class, dn_coord

    # ------------------------------------------------------------------------------------------------
    def _loss_sums(self, cls, box, gt_inds, prob_seg, targets, cls_weight=None):
        """Per-problem loss sums by the fused kernel (``fused_loss.detr_loss_sums``): targets gathered from the
        assignment, focal / L1 (+ xy, hw) / GIoU terms, one launch.  cls (P,Q,C), box (P,Q,4), gt_inds (P,Q), prob_seg
        (P,) int32 -> dict of (P,) sums (un-normalised, un-weighted)."""
        has_gt = targets.offsets_host[-1] > 0
        sums = fused_loss.detr_loss_sums(cls, box, gt_inds, prob_seg, targets.seg_offsets,
                                         targets.gt_bboxes if has_gt else None, targets.gt_labels if has_gt else None,
                                         targets.img_wh, cls_weight, self.loss_cls.alpha, self.loss_cls.gamma,
                                         self.loss_iou.eps)
        return {name: sums[:, i] for i, name in enumerate(fused_loss.LOSS_SUM_NAMES)}

    def _finish(self, sums, layers, bs, cls_avg, reg_avg):
        """(P,) per-problem sums -> per-layer losses with the reference's normalisers and weights."""
        out = {}
        for k, v in sums.items():
            v = v.view(layers, bs).sum(1)
            if k == "loss_cls":
                out[k] = v / cls_avg * self.loss_cls.loss_weight
            elif k == "loss_iou":
                out[k] = v / reg_avg * self.loss_iou.loss_weight
            else:
                out[k] = v / reg_avg * self.loss_bbox.loss_weight
        return out

    def loss(self, all_cls_scores, all_bbox_preds, enc_cls_scores, enc_bbox_preds, dn_cls_scores, dn_bbox_preds,
             gt_bboxes_list, gt_labels_list, gt_scores_list=None, img_metas=None, dn_metas=None,
             gt_bboxes_ignore=None, decouple=False, zero_weight_empty_dn=False):
        """Same inputs and the same 65 keys as dino_detr_head.py:506-632."""
        assert gt_bboxes_ignore is None and gt_scores_list is None
        all_cls_scores, all_bbox_preds = all_cls_scores.float(), all_bbox_preds.float()        # force_fp32
        L, bs, Q, C = all_cls_scores.shape
        dev = all_cls_scores.device
        counts = [int(b.shape[0]) for b in gt_bboxes_list]
        img_wh = [(m["img_shape"][1], m["img_shape"][0]) for m in img_metas]
        has_enc = enc_cls_scores is not None

        # --- matching part: L decoder layers (+ encoder proposals against class-0 labels, :574-577) -----------
        gtb = list(gt_bboxes_list) + (list(gt_bboxes_list) if has_enc else [])
        gtl = list(gt_labels_list) + ([torch.zeros_like(l) for l in gt_labels_list] if has_enc else [])
        targets = MatchTargets(gtb, gtl, img_wh * (2 if has_enc else 1), dev)
        cls_stack, box_stack = all_cls_scores, all_bbox_preds
        if has_enc:
            cls_stack = torch.cat([cls_stack, enc_cls_scores.float()[None]])
            box_stack = torch.cat([box_stack, enc_bbox_preds.float()[None]])
        layers = cls_stack.shape[0]
        P = layers * bs
        prob_seg = [(bs if (has_enc and l == L) else 0) + i for l in range(layers) for i in range(bs)]
        gt_inds, labels = self.assigner.assign_batch(box_stack.view(P, Q, 4), cls_stack.view(P, Q, C), targets,
                                                     prob_img=prob_seg)
        # targets (labels, normalised cxcywh boxes: dino_detr_head.py:969-976) are gathered inside the loss kernel
        prob_seg_d = device_const(dev, "prob_seg32", tuple(prob_seg), lambda: np.asarray(prob_seg, dtype=np.int32))
        sums = self._loss_sums(cls_stack.reshape(P, Q, C), box_stack.reshape(P, Q, 4), gt_inds, prob_seg_d, targets)
        num_pos = sum(min(c, Q) for c in counts)
        reg_avg = _clamp_min1(reduce_mean_scalar(num_pos, dev))
        main = self._finish(sums, layers, bs, max(num_pos * 1.0, 1), reg_avg)

        # --- denoising part: targets follow from the CDN layout, no matcher (:739-819) -------------------------
        if dn_cls_scores is not None and dn_bbox_preds is not None:
            dn = self._dn_terms(dn_cls_scores, dn_bbox_preds, gt_bboxes_list, gt_labels_list, img_metas, dn_metas,
                                zero_weight_empty_dn, targets=targets)
        else:
            dn = {k: torch.zeros(L, device=dev) for k in LOSS_PARTS}
        return self._assemble(main, dn, L, has_enc)

    def _dn_terms(self, dn_cls_scores, dn_bbox_preds, gt_bboxes_list, gt_labels_list, img_metas, dn_metas,
                  zero_weight_empty_dn=False, targets=None):
        """Per-layer denoising losses.  Slot of GT k of image b in group g is g*single_pad + k (positives first,
        negatives ``single_pad // 2`` later); everything else is background (dino_detr_head.py:739-819) -- i.e. the
        "assignment" is a constant of the batch geometry, so the same fused loss kernel serves.  With
        ``zero_weight_empty_dn`` an image without boxes contributes no classification loss
        (dino_detr_ssod_head.py:921-924).  ``targets``: a MatchTargets whose first ``bs`` segments are these images."""
        dev = dn_cls_scores.device
        Ld, bs, pad, C = dn_cls_scores.shape
        counts = [int(b.shape[0]) for b in gt_bboxes_list]
        groups = dn_metas["num_dn_group"]
        single_pad = pad // groups            # = 2 * max_gt: positives then negatives of one group
        total = sum(counts)
        if targets is None:
            img_wh = [(m["img_shape"][1], m["img_shape"][0]) for m in img_metas]
            targets = MatchTargets(list(gt_bboxes_list), list(gt_labels_list), img_wh, dev)

        def build_gt_inds():
            gi = np.zeros((bs, pad), dtype=np.int64)
            for b, c in enumerate(counts):
                for g in range(groups):
                    gi[b, g * single_pad:g * single_pad + c] = np.arange(1, c + 1)
            return np.tile(gi[None], (Ld, 1, 1)).reshape(Ld * bs, pad)
        gt_inds = device_const(dev, "dn_gt_inds", (tuple(counts), groups, single_pad, pad, Ld), build_gt_inds)
        Pd = Ld * bs
        prob_seg = device_const(dev, "dn_prob_seg32", (Ld, bs), lambda: np.tile(np.arange(bs, dtype=np.int32), Ld))
        cls_w = None
        if zero_weight_empty_dn and any(c == 0 for c in counts):
            cw = tuple(0.0 if c == 0 else 1.0 for c in counts) * Ld
            cls_w = device_const(dev, "dn_cls_w", cw, lambda: torch.tensor(cw, dtype=torch.float32))
        dsums = self._loss_sums(dn_cls_scores.float().reshape(Pd, pad, C), dn_bbox_preds.float().reshape(Pd, pad, 4),
                                gt_inds, prob_seg, targets, cls_w)
        dn_num_pos = total * groups
        return self._finish(dsums, Ld, bs, max(dn_num_pos * 1.0, 1), _clamp_min1(reduce_mean_scalar(dn_num_pos, dev, slot=1)))

    @staticmethod
    def _assemble(main, dn, L, has_enc):
        """Per-layer tensors -> the reference's loss dict (key order of dino_detr_head.py:584-632).  The dict also
        carries ``total`` -- the sum of all its entries, taken from the per-layer vectors with one concatenate + one
        reduction -- so ``_parse_losses`` need not add 65 scalars one kernel at a time (nor autograd undo them)."""
        loss_dict = LossDict()
        if has_enc:
            for k in LOSS_PARTS:
                loss_dict[f"enc_{k}"] = main[k][L]
        for k in LOSS_PARTS:
            loss_dict[k] = main[k][L - 1]
        for k in LOSS_PARTS:
            loss_dict[f"dn_{k}"] = dn[k][L - 1]
        for l in range(L - 1):
            for k in LOSS_PARTS:
                loss_dict[f"d{l}.{k}"] = main[k][l]
            for k in LOSS_PARTS:
                loss_dict[f"d{l}.dn_{k}"] = dn[k][l]
        loss_dict.total = torch.cat([main[k] for k in LOSS_PARTS] + [dn[k] for k in LOSS_PARTS]).sum()
        return loss_dict

    # ------------------------------------------------------------------------------------------------
    def forward_train(self, x, img_metas, gt_bboxes, gt_labels=None, gt_bboxes_ignore=None, proposal_cfg=None,
                      **kwargs):
        """dino_detr_head.py:983-1046"""
        assert proposal_cfg is None, '"proposal_cfg" must be None'
        counts = [int(b.shape[0]) for b in gt_bboxes]
        if self.dn_number > 0 and max(counts) > 0:
            boxes = []
            for meta, b in zip(img_metas, gt_bboxes):
                h, w, _ = meta["img_shape"]
                fac = device_const(b.device, "whwh", (w, h), lambda: torch.tensor([w, h, w, h], dtype=torch.float32))
                boxes.append(bbox_xyxy_to_cxcywh(b) / fac)
            q_label, q_bbox, attn_mask, dn_meta = prepare_for_cdn(
                dn_args=(dict(labels=gt_labels, boxes=boxes), self.dn_number, self.dn_label_noise_ratio,
                         self.dn_box_noise_scale),
                training=True, num_queries=self.num_query, num_classes=self.num_classes,
                hidden_dim=self.embed_dims, label_enc=self.label_enc)
        else:
            q_label = q_bbox = attn_mask = dn_meta = None
        outs = self(x, img_metas, q_label, q_bbox, attn_mask, dn_meta)
        return self.loss(*outs, gt_bboxes, gt_labels, img_metas=img_metas, dn_metas=dn_meta,
                         gt_bboxes_ignore=gt_bboxes_ignore)
