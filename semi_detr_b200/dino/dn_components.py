"""Contrastive denoising (CDN) query construction -- mirror of
detr_od/models/dense_heads/dn_components.py:6-125 (``prepare_for_cdn``) and :462-479 (``dn_post_process``).

Same layout and noise model as the reference: every GT is repeated 2*groups times,
repetition 2g = positives and 2g+1 = negatives of group g; label flipped w.p. label_noise_ratio*0.5; xyxy corners
jittered by sign * r * (w/2, h/2) * box_noise_scale with r~U[0,1) for positives and U[1,2) for negatives; slot of
GT k of image b in repetition i is (b, single_pad*i + k); block-structured (pad+Q)^2 bool mask.

What changed is only *how* it is built: all index arithmetic is done on the host from the (host-known) GT counts
and shipped in one small copy, the noise is drawn with fixed-shape device RNG calls (no ``nonzero`` /
``int(max(...))`` device syncs, no python loop over groups), and the attention mask is cached per
(single_pad, groups, num_queries).  The random stream therefore differs from the reference's, the distribution
does not.
"""
import numpy as np
import torch

from ..consts import device_const
from .transformer import inverse_sigmoid

_MASK_CACHE = {}


# RNG entry points (tests swap them for CPU-seeded draws to compare devices on identical noise)
def _rand(shape, device, generator=None):
    return torch.rand(shape, device=device, generator=generator)


def _randint(low, high, shape, device, generator=None):
    return torch.randint(low, high, shape, device=device, generator=generator)


def _attn_mask(single_pad, groups, num_queries, device):
    key = (single_pad, groups, num_queries, str(device))
    m = _MASK_CACHE.get(key)
    if m is None:
        pad = 2 * single_pad * groups
        size = pad + num_queries
        gid = torch.arange(size, device=device) // max(2 * single_pad, 1)
        is_dn = torch.arange(size, device=device) < pad
        # matching queries cannot see the denoising part; a denoising group sees only itself (+ matching part)
        m = (~is_dn[:, None] & is_dn[None, :]) | (is_dn[:, None] & is_dn[None, :] & (gid[:, None] != gid[None, :]))
        if len(_MASK_CACHE) > 64:
            _MASK_CACHE.clear()
        _MASK_CACHE[key] = m
    return m


def prepare_for_cdn(dn_args, training, num_queries, num_classes, hidden_dim, label_enc, generator=None,
                    fill_empty=False):
    """dn_args = (targets{'labels': [..], 'boxes': [normalised cxcywh ..]}, dn_number, label_noise_ratio,
    box_noise_scale) -> input_query_label (bs, pad, C), input_query_bbox (bs, pad, 4) [inverse-sigmoid],
    attn_mask (pad+Q, pad+Q) bool, dn_meta {'pad_size', 'num_dn_group'}"""
    if not training:
        return None, None, None, None
    targets, dn_number, label_noise_ratio, box_noise_scale = dn_args
    labels_list, boxes_list = list(targets["labels"]), list(targets["boxes"])
    empty = [int(t.shape[0]) == 0 for t in boxes_list]
    if fill_empty and any(empty):
        # prepare_for_cdn_plus (dn_components.py:137-160): an image without boxes gets one centred dummy box with a
        # random label so the batch layout stays rectangular; its targets stay empty (all background)
        dev0 = boxes_list[0].device
        dummy = device_const(dev0, "cdn_dummy_box", (), lambda: torch.tensor([[0.5, 0.5, 0.5, 0.5]]))
        for i, e in enumerate(empty):
            if e:
                boxes_list[i] = dummy
                labels_list[i] = _randint(0, num_classes, (1,), dev0, generator)
    bs = len(labels_list)
    counts = [int(t.shape[0]) for t in labels_list]           # host ints
    device = boxes_list[0].device
    max_gt = max(counts) if counts else 0
    dn_number = dn_number * 2
    if max_gt == 0:
        groups = 1
    elif dn_number >= 100:
        groups = dn_number // (max_gt * 2)
    else:
        groups = max(dn_number, 1)
    groups = max(groups, 1)
    single_pad = max_gt
    pad_size = single_pad * 2 * groups
    total = sum(counts)

    labels = torch.cat([t.reshape(-1) for t in labels_list]).long()
    boxes = torch.cat([t.reshape(-1, 4) for t in boxes_list])
    reps = 2 * groups
    known_labels = labels.repeat(reps)
    known_boxes = boxes.repeat(reps, 1)

    if label_noise_ratio > 0 and total > 0:
        p = _rand(known_labels.shape, device, generator)
        new_label = _randint(0, num_classes, known_labels.shape, device, generator)
        known_labels = torch.where(p < label_noise_ratio * 0.5, new_label, known_labels)

    if box_noise_scale > 0 and total > 0:
        xyxy = torch.cat([known_boxes[:, :2] - known_boxes[:, 2:] / 2, known_boxes[:, :2] + known_boxes[:, 2:] / 2], 1)
        diff = torch.cat([known_boxes[:, 2:] / 2, known_boxes[:, 2:] / 2], 1)
        sign = _randint(0, 2, known_boxes.shape, device, generator).float() * 2.0 - 1.0
        part = _rand(known_boxes.shape, device, generator)
        # negatives (odd repetitions) are pushed one half-extent further out
        neg = device_const(device, "cdn_neg", (reps, total),
                           lambda: (torch.arange(reps) % 2 == 1).repeat_interleave(total))
        part = (part + neg[:, None].float()) * sign
        xyxy = (xyxy + part * diff * box_noise_scale).clamp(min=0.0, max=1.0)
        known_boxes = torch.cat([(xyxy[:, :2] + xyxy[:, 2:]) / 2, xyxy[:, 2:] - xyxy[:, :2]], 1)

    input_label_embed = label_enc(known_labels)
    input_bbox_embed = inverse_sigmoid(known_boxes)
    input_query_label = torch.zeros(bs, pad_size, hidden_dim, device=device)
    input_query_bbox = torch.zeros(bs, pad_size, 4, device=device)
    if total > 0:
        def build_idx():
            bid = np.concatenate([np.full(c, i, dtype=np.int64) for i, c in enumerate(counts)])
            within = np.concatenate([np.arange(c, dtype=np.int64) for c in counts])
            return np.stack([np.tile(bid, reps), np.concatenate([within + single_pad * i for i in range(reps)])])
        idx = device_const(device, "cdn_idx", (tuple(counts), reps), build_idx)
        input_query_label[idx[0], idx[1]] = input_label_embed
        input_query_bbox[idx[0], idx[1]] = input_bbox_embed

    attn_mask = _attn_mask(single_pad, groups, num_queries, device)
    dn_meta = {"pad_size": pad_size, "num_dn_group": groups}
    if fill_empty:
        dn_meta["pad_mask"] = empty          # host flags: which images have no real box
    return input_query_label, input_query_bbox, attn_mask, dn_meta


def dn_post_process(outputs_class, outputs_coord, dn_meta):
    """Split (n_dec, bs, pad+Q, .) into the matching part and the denoising part (dn_components.py:462-479)."""
    if dn_meta and dn_meta["pad_size"] > 0:
        pad = dn_meta["pad_size"]
        return outputs_class[:, :, pad:], outputs_coord[:, :, pad:], outputs_class[:, :, :pad], outputs_coord[:, :, :pad]
    return outputs_class, outputs_coord, None, None
