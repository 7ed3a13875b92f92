"""Host logic of the DINO path on the CPU (oracle ops injected): CDN layout, batched loss == per-problem
reference-style loss, loss-dict keys, gradients reach every trainable parameter."""
import copy

import numpy as np
import pytest
import torch

from oracle import hungarian_oracle as H
from oracle.cpu_path import reference_cpu_ops
from semi_detr_b200 import dino  # noqa: F401
from semi_detr_b200.dino.dn_components import prepare_for_cdn
from semi_detr_b200.dino.losses import FocalLoss, GIoULoss, L1Loss
from semi_detr_b200.matching.match_cost import bbox_cxcywh_to_xyxy, bbox_xyxy_to_cxcywh
from semi_detr_b200.registry import DETECTORS
from semi_detr_b200.synthetic import DINO_R50_4SCALE, coco_like_batch


def test_cdn_layout_matches_reference_rules():
    """SURVEY.md appendix A.5 / dn_components.py:21-112"""
    torch.manual_seed(0)
    labels = [torch.tensor([3, 7, 9]), torch.tensor([1])]
    boxes = [torch.rand(3, 4) * 0.4 + 0.3, torch.rand(1, 4) * 0.4 + 0.3]
    enc = torch.nn.Embedding(82, 16)
    ql, qb, mask, meta = prepare_for_cdn((dict(labels=labels, boxes=boxes), 100, 0.5, 0.4), True, 900, 80, 16, enc)
    groups = 200 // (3 * 2)
    assert meta == dict(pad_size=3 * 2 * groups, num_dn_group=groups)
    pad = meta["pad_size"]
    assert ql.shape == (2, pad, 16) and qb.shape == (2, pad, 4) and mask.shape == (pad + 900, pad + 900)
    # unused slots stay zero (image 1 has one GT: slots 1,2 of every repetition are empty)
    assert not qb[1, 1:3].any() and not ql[1, 1:3].any() and qb[1, 0].any()
    # positives (even repetitions) stay within half a box extent of the GT, negatives move at least that far
    gt = boxes[0][0]
    b = torch.sigmoid(qb[0, 0])                     # repetition 0, GT 0
    xyxy_gt = torch.cat([gt[:2] - gt[2:] / 2, gt[:2] + gt[2:] / 2])
    xyxy = torch.cat([b[:2] - b[2:] / 2, b[:2] + b[2:] / 2])
    assert ((xyxy - xyxy_gt).abs() <= torch.cat([gt[2:], gt[2:]]) / 2 * 0.4 + 1e-4).all()
    # mask: matching part never sees DN; groups are isolated; a group sees itself
    assert mask[pad:, :pad].all() and not mask[pad:, pad:].any()
    assert not mask[0:6, 0:6].any() and mask[0:6, 6:pad].all() and mask[6:12, 0:6].all()
    assert not mask[:pad, pad:].any()


def _reference_style_loss_single(cls, box, gtb, gtl, metas, num_classes=80):
    """dino_detr_head.py:634-736 written out per image with the reference's losses (oracle assigner)."""
    labels_l, box_t_l, w_l, fac_l = [], [], [], []
    for i in range(cls.shape[0]):
        h, w, _ = metas[i]["img_shape"]
        gi, lb = H.hungarian_assign(box[i].detach(), cls[i].detach(), gtb[i], gtl[i], h, w)
        pos = gi > 0
        labels = torch.full((cls.shape[1],), num_classes, dtype=torch.long)
        labels[pos] = gtl[i][gi[pos] - 1]
        bt = torch.zeros_like(box[i])
        bt[pos] = bbox_xyxy_to_cxcywh(gtb[i][gi[pos] - 1] / gtb[i].new_tensor([w, h, w, h]))
        bw = torch.zeros_like(box[i])
        bw[pos] = 1.0
        labels_l.append(labels); box_t_l.append(bt); w_l.append(bw)
        fac_l.append(box[i].new_tensor([w, h, w, h]).repeat(box.shape[1], 1))
    labels, bt, bw, fac = torch.cat(labels_l), torch.cat(box_t_l), torch.cat(w_l), torch.cat(fac_l)
    npos = float((bw.sum(-1) > 0).sum())
    cls_avg, reg_avg = max(npos, 1), max(npos, 1)
    lc = FocalLoss(loss_weight=2.0)(cls.reshape(-1, num_classes), labels, torch.ones(len(labels)), avg_factor=cls_avg)
    bp = box.reshape(-1, 4)
    li = GIoULoss(loss_weight=2.0)(bbox_cxcywh_to_xyxy(bp) * fac, bbox_cxcywh_to_xyxy(bt) * fac, bw, avg_factor=reg_avg)
    lb_ = L1Loss(loss_weight=5.0)(bp, bt, bw, avg_factor=reg_avg)
    return lc, lb_, li


def test_batched_loss_equals_per_layer_reference_loss():
    torch.manual_seed(1)
    cfg = copy.deepcopy(DINO_R50_4SCALE)
    head = DETECTORS.build(cfg).bbox_head
    L, bs, Q, C = 3, 2, 60, 80
    cls = torch.randn(L, bs, Q, C) - 2
    box = torch.rand(L, bs, Q, 4) * 0.5 + 0.2
    enc_cls, enc_box = torch.randn(bs, Q, C) - 2, torch.rand(bs, Q, 4) * 0.5 + 0.2
    data = coco_like_batch(bs, 300, 400, seed=3)
    with reference_cpu_ops():
        out = head.loss(cls, box, enc_cls, enc_box, None, None, data["gt_bboxes"], data["gt_labels"],
                        img_metas=data["img_metas"], dn_metas=None)
    assert len(out) == 5 + 5 + 5 + (L - 1) * 10
    for l in range(L):
        lc, lb_, li = _reference_style_loss_single(cls[l], box[l], data["gt_bboxes"], data["gt_labels"], data["img_metas"])
        pre = "" if l == L - 1 else f"d{l}."
        assert torch.allclose(out[pre + "loss_cls"], lc, rtol=1e-5)
        assert torch.allclose(out[pre + "loss_bbox"], lb_, rtol=1e-5)
        assert torch.allclose(out[pre + "loss_iou"], li, rtol=1e-5)
        assert torch.allclose(out[pre + "loss_bbox_xy"] + out[pre + "loss_bbox_hw"], lb_, rtol=1e-5)
    zl = [torch.zeros_like(l) for l in data["gt_labels"]]       # encoder aux loss: class-0 labels (:574-577)
    lc, lb_, li = _reference_style_loss_single(enc_cls, enc_box, data["gt_bboxes"], zl, data["img_metas"])
    assert torch.allclose(out["enc_loss_cls"], lc, rtol=1e-5) and torch.allclose(out["enc_loss_iou"], li, rtol=1e-5)
    assert float(out["dn_loss_cls"]) == 0.0


def test_train_step_on_cpu_oracle_path():
    torch.manual_seed(0)
    model = DETECTORS.build(copy.deepcopy(DINO_R50_4SCALE)).train()
    data = coco_like_batch(2, 256, 320, seed=0)
    with reference_cpu_ops():
        out = model.train_step(data)
        out["loss"].backward()
    keys = list(out["log_vars"])
    assert len(keys) == 66 and keys[-1] == "loss"          # 65 loss terms + total
    for part in ("loss_cls", "loss_bbox", "loss_iou", "loss_bbox_xy", "loss_bbox_hw"):
        for pre in ("", "enc_", "dn_", "d0.", "d4.dn_"):
            assert pre + part in out["log_vars"]
    total = sum(float(v) for k, v in out["log_vars"].items() if k != "loss")
    assert abs(total - float(out["loss"])) < 1e-3 * total
    assert np.isfinite(float(out["loss"]))
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert any("layer1" in n for n in frozen) and all("backbone" in n for n in frozen)


def test_five_scale_config_steps_on_cpu_oracle_path():
    """BASELINE config 4 plumbing: 5 feature levels (4 backbone stages + 1 extra stride-2 level)."""
    import torch
    from oracle.cpu_path import reference_cpu_ops
    from semi_detr_b200 import dino  # noqa: F401
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import coco_like_batch, dino_r50_5scale
    torch.manual_seed(0)
    model = DETECTORS.build(dino_r50_5scale()).train()
    assert model.bbox_head.transformer.num_feature_levels == 5 and len(model.bbox_head.input_proj) == 5
    assert model.bbox_head.transformer.encoder.layers[0].self_attn.n_levels == 5
    with reference_cpu_ops():
        out = model.train_step(coco_like_batch(1, 128, 160, seed=5))
    assert torch.isfinite(out["loss"]) and len(out["log_vars"]) == 66
    out["loss"].backward()
    assert model.bbox_head.transformer.level_embed.grad.shape == (5, 256)


def test_loss_dict_total_equals_entry_by_entry_sum():
    """The head's precomputed ``LossDict.total`` against mmdet's entry-by-entry ``_parse_losses`` sum
    (base.py:176-209): same value, same gradients, same log keys; editing the dict drops the shortcut."""
    import copy
    import torch
    from oracle.cpu_path import reference_cpu_ops
    from semi_detr_b200 import dino  # noqa: F401
    from semi_detr_b200.dino.head import LossDict
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import DINO_R50_4SCALE, coco_like_batch
    torch.manual_seed(0)
    model = DETECTORS.build(copy.deepcopy(DINO_R50_4SCALE)).train()
    data = coco_like_batch(1, 288, 352, seed=5)
    with reference_cpu_ops():
        losses = model(**data)
        assert isinstance(losses, LossDict) and losses.total is not None and len(losses) == 65
        fast, lv_fast = model._parse_losses(losses)
        slow, lv_slow = model._parse_losses(dict(losses))
        assert list(lv_fast) == list(lv_slow)
        assert abs(float(fast) - float(slow)) <= 1e-5 * abs(float(slow))
        for k in lv_slow:
            assert abs(float(lv_fast[k]) - float(lv_slow[k])) <= 1e-5 * abs(float(lv_slow[k])) + 1e-7
        w = model.bbox_head.fc_cls[0].weight
        g_fast = torch.autograd.grad(fast, w, retain_graph=True)[0]
        g_slow = torch.autograd.grad(slow, w)[0]
        assert (g_fast - g_slow).abs().max() <= 1e-5 * g_slow.abs().max()
    losses["extra_loss"] = torch.zeros(())
    assert losses.total is None


def test_decoder_self_attention_equals_multihead_attention():
    """The written-out decoder self-attention against nn.MultiheadAttention itself (same parameters, bool mask)."""
    import torch
    from semi_detr_b200.dino.transformer import DINOTransformerDecoderLayer
    torch.manual_seed(0)
    layer = DINOTransformerDecoderLayer(d_model=256, d_ffn=64, dropout=0.0).eval()
    T, N = 50, 3
    qk, v = torch.randn(T, N, 256), torch.randn(T, N, 256)
    mask = torch.rand(T, T) < 0.4
    mask.fill_diagonal_(False)                      # never a fully masked row (as in the CDN mask)
    want = layer.self_attn(qk, qk, v, attn_mask=mask, need_weights=False)[0]
    got = layer._self_attention(qk, v, mask)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(layer._self_attention(qk, v, None), layer.self_attn(qk, qk, v, need_weights=False)[0],
                          rtol=1e-5, atol=1e-6)
    qk.requires_grad_(True)
    g1 = torch.autograd.grad(layer._self_attention(qk, v, mask).square().sum(), qk)[0]
    g2 = torch.autograd.grad(layer.self_attn(qk, qk, v, attn_mask=mask, need_weights=False)[0].square().sum(), qk)[0]
    assert torch.allclose(g1, g2, rtol=1e-4, atol=1e-5)


def test_frozen_bn_fold_is_cached_and_tracks_writes():
    """dino/backbone.py ``folded_conv``: conv(x, w) + t equals BN_eval(conv(x)); with the cache enabled the fold is
    computed once per layer, reused (same tensor objects) and recomputed after any torch-side write to a BN tensor or
    a frozen convolution weight; trainable convolution weights are re-scaled every call and keep their gradient."""
    import torch.nn.functional as F
    from torch import nn
    from semi_detr_b200.dino.backbone import _FOLD_CACHE, enable_frozen_bn_fold_cache, folded_conv
    torch.manual_seed(0)
    conv = nn.Conv2d(5, 7, 3, padding=1, bias=False)
    bn = nn.BatchNorm2d(7).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    for p in bn.parameters():
        p.requires_grad = False
    x = torch.randn(2, 5, 9, 8)
    want = bn(conv(x))
    w, t = folded_conv(conv, bn)                       # cache off: plain computation
    assert torch.allclose(F.conv2d(x, w, t, padding=1), want, rtol=1e-5, atol=1e-6)
    assert bn not in _FOLD_CACHE
    enable_frozen_bn_fold_cache(nn.Sequential(conv, bn))
    w1, t1 = folded_conv(conv, bn)
    w2, t2 = folded_conv(conv, bn)
    assert t2 is t1 and w2 is not w1 and w2.requires_grad      # trainable weight: fresh product, gradient reaches it
    assert torch.equal(w1, w) and torch.equal(t1, t)
    F.conv2d(x, w2, t2, padding=1).sum().backward()
    ref = conv.weight.grad.clone()
    conv.weight.grad = None
    bn(conv(x)).sum().backward()
    assert torch.allclose(conv.weight.grad, ref, rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        conv.weight.mul_(1.5)                          # an optimizer step (no version bump needed: never cached)
    w3, _ = folded_conv(conv, bn)
    assert torch.allclose(w3, w * 1.5)
    conv.weight.requires_grad = False                  # frozen stage: the scaled weight is cached too
    _FOLD_CACHE.pop(bn)
    w4, t4 = folded_conv(conv, bn)
    w5, t5 = folded_conv(conv, bn)
    assert w5 is w4 and t5 is t4
    with torch.no_grad():
        conv.weight.add_(1.0)                          # torch-side write -> version bump -> recomputed
    w6, _ = folded_conv(conv, bn)
    assert w6 is not w4 and torch.allclose(w6, w4 + torch.rsqrt(bn.running_var + bn.eps).mul(bn.weight).view(-1, 1, 1, 1))
    with torch.no_grad():
        bn.running_var.mul_(2.0)                       # e.g. load_state_dict
    w7, t7 = folded_conv(conv, bn)
    assert t7 is not t4
    assert torch.allclose(F.conv2d(x, w7, t7, padding=1), bn(conv(x)), rtol=1e-4, atol=1e-4)
    bn.weight.requires_grad = True                     # trainable BN: never cached
    a = folded_conv(conv, bn)
    b = folded_conv(conv, bn)
    assert a[1] is not b[1] and a[1].requires_grad


def test_engines_enable_the_bn_fold_cache_for_the_student_only():
    from semi_detr_b200.dino.backbone import enable_frozen_bn_fold_cache
    from semi_detr_b200.engine import _cache_student_bn_folds
    from torch import nn

    class Wrapper(nn.Module):
        def __init__(self):
            super().__init__()
            self.student = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4))
            self.teacher = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4))
    from semi_detr_b200.dino.backbone import fold_cache_enabled
    import copy
    m = Wrapper()
    _cache_student_bn_folds(m)
    assert fold_cache_enabled(m.student[1]) and not fold_cache_enabled(m.teacher[1])
    plain = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4))
    _cache_student_bn_folds(plain)
    assert fold_cache_enabled(plain[1])
    assert not fold_cache_enabled(copy.deepcopy(plain)[1]), "a copy (e.g. a teacher made from the student) starts cold"
    enable_frozen_bn_fold_cache(plain, False)
    assert not fold_cache_enabled(plain[1])


def test_mask_geometry_cache_changes_nothing():
    """dino/transformer.py ``_mask_geometry``: valid ratios, encoder reference points and two-stage proposals depend on
    the padding masks alone; with the head's host geometry key they are built once and reused -- bit-identical losses
    with the cache cold, warm, and bypassed, and one entry per distinct geometry."""
    import copy
    from semi_detr_b200.dino import transformer as T
    from semi_detr_b200.synthetic import DINO_R50_4SCALE, coco_like_batch
    cfg = copy.deepcopy(DINO_R50_4SCALE)
    cfg["bbox_head"]["num_query"] = 60
    cfg["bbox_head"]["transformer"] = dict(type="DINOTransformer", num_queries=60, num_encoder_layers=1,
                                           num_decoder_layers=2, dim_feedforward=64)
    cfg["bbox_head"]["dn_number"] = 10
    torch.manual_seed(0)
    model = DETECTORS.build(cfg).train()

    def run(data, seed=3):
        torch.manual_seed(seed)
        with reference_cpu_ops():
            out = model(**data)
        return {k: v.detach().clone() for k, v in out.items() if torch.is_tensor(v)}

    data = coco_like_batch(2, 128, 160, seed=1)
    data["img_metas"][1]["img_shape"] = (100, 120, 3)            # a padded image: valid ratios < 1, masked proposals
    T._GEOMETRY.clear()
    cold = run(data)
    assert len(T._GEOMETRY) == 1
    entry = next(iter(T._GEOMETRY.values()))
    warm = run(data)
    assert next(iter(T._GEOMETRY.values())) is entry and len(T._GEOMETRY) == 1
    orig = T.DINOTransformer.forward
    try:
        T.DINOTransformer.forward = lambda self, *a, geometry_key=None, **k: orig(self, *a, geometry_key=None, **k)
        bypass = run(data)
    finally:
        T.DINOTransformer.forward = orig
    assert cold.keys() == warm.keys() == bypass.keys() and len(cold) > 10
    for k in cold:
        assert torch.equal(cold[k], warm[k]) and torch.equal(cold[k], bypass[k]), k
    other = coco_like_batch(2, 128, 160, seed=1)                  # same tensors, no padding: a second geometry
    run(other)
    assert len(T._GEOMETRY) == 2
    plain = run(other)
    assert any(not torch.equal(plain[k], cold[k]) for k in cold)


def test_sine_embedding_equals_the_per_coordinate_form():
    """gen_sineembed_for_position: one pass over all coordinates is bit-identical to the reference's per-coordinate
    sin / cos / stack / flatten (transformer.py:467-493)."""
    import math
    from semi_detr_b200.dino.transformer import gen_sineembed_for_position

    def reference(pos):
        dim_t = 10000 ** (2 * (torch.arange(128, dtype=torch.float32) // 2) / 128)

        def emb(c):
            e = (c * (2 * math.pi))[:, :, None] / dim_t
            return torch.stack((e[:, :, 0::2].sin(), e[:, :, 1::2].cos()), dim=3).flatten(2)
        px, py = emb(pos[:, :, 0]), emb(pos[:, :, 1])
        if pos.size(-1) == 2:
            return torch.cat((py, px), dim=2)
        return torch.cat((py, px, emb(pos[:, :, 2]), emb(pos[:, :, 3])), dim=2)

    g = torch.Generator().manual_seed(0)
    for n in (2, 4):
        pos = torch.rand(37, 3, n, generator=g)
        got, want = gen_sineembed_for_position(pos), reference(pos)
        assert got.shape == want.shape == (37, 3, 128 * n) and torch.equal(got, want)
        nc = pos.transpose(0, 1).contiguous().transpose(0, 1)          # non-contiguous input, as the decoder passes
        assert torch.equal(gen_sineembed_for_position(nc), want)
    with pytest.raises(ValueError):
        gen_sineembed_for_position(torch.rand(2, 2, 3))
