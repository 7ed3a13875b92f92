"""Deterministic weights and inputs shared by the golden generator (reference side, make_golden.py) and the tests (our
side): both models are filled parameter-by-parameter from a seed derived from the parameter's NAME, so the fixture
needs no multi-megabyte state dict -- the reference's DINOTransformer and ours have the same 72 parameter names."""
import zlib

import torch

TRANSFORMER_KW = dict(d_model=256, nhead=8, num_queries=30, num_encoder_layers=1, num_decoder_layers=2,
                      dim_feedforward=64, num_feature_levels=4)
LEVELS = [(12, 16), (6, 8), (3, 4), (2, 2)]
NUM_CLASSES = 8
N_DN = 6


def fill_by_name(module, prefix=""):
    """Every parameter <- N(0, s) drawn from a generator seeded by crc32(name); s by kind (weights ~ xavier scale,
    LayerNorm gains around 1, sampling-offset biases a few pixels so taps sit off the pixel lattice)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            full = prefix + name
            g = torch.Generator().manual_seed(zlib.crc32(full.encode()))
            r = torch.randn(p.shape, generator=g, dtype=torch.float32)
            if "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * r)
            elif name.endswith("sampling_offsets.bias"):
                p.copy_(1.7 * r)
            elif name.endswith("bias"):
                p.copy_(0.1 * r)
            elif p.dim() >= 2:
                p.copy_(r * (2.0 / (p.shape[0] + p.shape[-1])) ** 0.5)
            else:
                p.copy_(0.5 * r)
    return module


def inputs(seed=11):
    """srcs / masks / pos_embeds per level (image 1 padded on the right and bottom), a denoising part of N_DN queries
    and its (T, T) attention mask."""
    g = torch.Generator().manual_seed(seed)
    bs, C = 2, TRANSFORMER_KW["d_model"]
    srcs, masks, poss = [], [], []
    for (h, w) in LEVELS:
        srcs.append(torch.randn(bs, C, h, w, generator=g))
        poss.append(torch.randn(bs, C, h, w, generator=g) * 0.5)
        m = torch.zeros(bs, h, w, dtype=torch.bool)
        m[1, max(1, (3 * h) // 4):, :] = True
        m[1, :, max(1, (2 * w) // 3):] = True
        masks.append(m)
    nq = TRANSFORMER_KW["num_queries"]
    refpoint = torch.randn(bs, N_DN, 4, generator=g)               # unsigmoided boxes of the denoising queries
    tgt = torch.randn(bs, N_DN, C, generator=g)
    T = N_DN + nq
    attn_mask = torch.zeros(T, T, dtype=torch.bool)
    attn_mask[N_DN:, :N_DN] = True                                 # matching queries do not see the denoising part
    attn_mask[:N_DN // 2, N_DN // 2:N_DN] = True                   # two denoising groups, isolated from each other
    attn_mask[N_DN // 2:N_DN, :N_DN // 2] = True
    return srcs, masks, poss, refpoint, tgt, attn_mask


FULL_GRADS = ["level_embed", "enc_output_norm.weight", "decoder.layers.1.cross_attn.sampling_offsets.bias",
              "encoder.layers.0.self_attn.attention_weights.bias", "decoder.ref_point_head.layers.1.bias",
              "decoder.layers.0.self_attn.in_proj_bias", "encoder.layers.0.norm2.bias"]


def scalar_loss(hs, references, hs_enc, ref_enc):
    """A fixed smooth scalar of every differentiable output of DINOTransformer.forward."""
    w = torch.linspace(0.5, 1.5, hs[0].shape[-1])
    loss = sum((h * w).sin().mean() for h in hs) + (hs_enc * w).cos().mean()
    loss = loss + sum((r * torch.tensor([1.0, -2.0, 0.5, 1.5])).sum(-1).square().mean() for r in references[1:])
    return loss + (ref_enc * torch.tensor([0.3, 0.7, -1.1, 0.9])).sum(-1).square().mean()


# ---- head loss fixture (reference: DINODETRHead.loss, dino_detr_head.py:506-980) ---------------------------------
LOSS_KW = dict(num_classes=9, num_query=40, n_dec=3, dn_groups=2)
LOSS_IMG_SHAPES = [(96, 128, 3), (80, 112, 3), (64, 64, 3)]
LOSS_GT_COUNTS = [3, 0, 5]            # one image without boxes


def loss_inputs(seed=23):
    """Decoder / encoder / denoising predictions, GT boxes (xyxy pixels) and labels, img_metas, dn_meta."""
    g = torch.Generator().manual_seed(seed)
    K, Q, L, G = LOSS_KW["num_classes"], LOSS_KW["num_query"], LOSS_KW["n_dec"], LOSS_KW["dn_groups"]
    bs = len(LOSS_IMG_SHAPES)
    gt_bboxes, gt_labels = [], []
    for (h, w, _), n in zip(LOSS_IMG_SHAPES, LOSS_GT_COUNTS):
        xy = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.6, h * 0.6])
        wh = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.35, h * 0.35]) + 4.0
        gt_bboxes.append(torch.cat([xy, xy + wh], 1))
        gt_labels.append(torch.randint(0, K, (n,), generator=g))
    pad = 2 * max(LOSS_GT_COUNTS) * G
    rnd = lambda *shape: torch.randn(*shape, generator=g)
    box = lambda *shape: torch.cat([torch.rand(*shape, 2, generator=g) * 0.8 + 0.1,
                                    torch.rand(*shape, 2, generator=g) * 0.3 + 0.03], -1)
    out = dict(all_cls_scores=rnd(L, bs, Q, K) - 2.0, all_bbox_preds=box(L, bs, Q),
               enc_cls_scores=rnd(bs, Q, K) - 2.0, enc_bbox_preds=box(bs, Q),
               dn_cls_scores=rnd(L, bs, pad, K) - 1.0, dn_bbox_preds=box(L, bs, pad),
               gt_bboxes=gt_bboxes, gt_labels=gt_labels,
               img_metas=[dict(img_shape=s, batch_input_shape=(96, 128)) for s in LOSS_IMG_SHAPES],
               dn_meta=dict(pad_size=pad, num_dn_group=G))
    return out


# ---- contrastive denoising fixture (reference: prepare_for_cdn, dn_components.py:6-125) ---------------------------
CDN_CASES = {
    # name: (GT counts per image, dn_number, label_noise_ratio, box_noise_scale, num_queries, num_classes)
    "config": ([3, 0, 5], 100, 0.5, 0.4, 30, 9),          # the shipped setting: dn_number 100 -> 20 groups
    "few_groups": ([2, 4], 2, 0.5, 1.0, 12, 9),            # dn_number < 100 is taken literally (4 groups)
    "no_noise": ([1, 2], 100, 0.0, 0.0, 10, 9),
}


def cdn_targets(counts, num_classes, seed=5):
    g = torch.Generator().manual_seed(seed + sum(counts))
    labels = [torch.randint(0, num_classes, (n,), generator=g) for n in counts]
    boxes = [torch.cat([torch.rand(n, 2, generator=g) * 0.6 + 0.2, torch.rand(n, 2, generator=g) * 0.3 + 0.05], 1)
             for n in counts]
    return dict(labels=labels, boxes=boxes)


def label_embedding(num_classes, hidden_dim=256):
    emb = torch.nn.Embedding(num_classes + 1, hidden_dim)
    return fill_by_name(emb, "label_enc.")


# ---- SSOD query construction fixture (reference: DinoDetrSSOD.prepare_unsup_cdn, dino_detr_ssod.py:484-760) -------
UNSUP_KW = dict(num_classes=80, num_queries=20, hidden_dim=256, dn_number=100, label_noise_ratio=0.5,
                box_noise_scale=0.4)
UNSUP_IMG_SHAPES = [(96, 128, 3), (80, 112, 3), (64, 64, 3)]
UNSUP_PSEUDO_COUNTS = [2, 0, 3]        # high-recall pseudo boxes per image (consistency queries); image 1 has none
UNSUP_RELIABLE_COUNTS = [1, 0, 2]      # reliable pseudo boxes per image (denoising part); image 1 gets the dummy box


def unsup_inputs(seed=31):
    g = torch.Generator().manual_seed(seed)
    K = UNSUP_KW["num_classes"]

    def boxes(n, h, w):
        xy = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.6, h * 0.6])
        return torch.cat([xy, xy + torch.rand(n, 2, generator=g) * torch.tensor([w * 0.3, h * 0.3]) + 3.0], 1)
    pseudo = [boxes(n, h, w) for n, (h, w, _) in zip(UNSUP_PSEUDO_COUNTS, UNSUP_IMG_SHAPES)]
    pseudo_labels = [torch.randint(0, K, (n,), generator=g) for n in UNSUP_PSEUDO_COUNTS]
    det = [boxes(n, h, w) for n, (h, w, _) in zip(UNSUP_PSEUDO_COUNTS, UNSUP_IMG_SHAPES)]
    rel = [boxes(n, h, w) for n, (h, w, _) in zip(UNSUP_RELIABLE_COUNTS, UNSUP_IMG_SHAPES)]
    rel_labels = [torch.randint(0, K, (n,), generator=g) for n in UNSUP_RELIABLE_COUNTS]
    rel_norm = []
    for b, (h, w, _) in zip(rel, UNSUP_IMG_SHAPES):
        cxcywh = torch.cat([(b[:, :2] + b[:, 2:]) / 2, b[:, 2:] - b[:, :2]], 1)
        rel_norm.append(cxcywh / torch.tensor([w, h, w, h], dtype=torch.float32))
    n_slots = 5 * max(max(UNSUP_PSEUDO_COUNTS), 1)
    bs = len(UNSUP_IMG_SHAPES)
    prior = dict(loss_weights=torch.rand(5 * sum(max(c, 1) for c in UNSUP_PSEUDO_COUNTS), 1, generator=g),
                 input_query_label_1=torch.randn(bs, n_slots, UNSUP_KW["hidden_dim"], generator=g))
    metas = [dict(img_shape=s) for s in UNSUP_IMG_SHAPES]
    return dict(pseudo=pseudo, pseudo_labels=pseudo_labels, det=det, det_labels=pseudo_labels, prior=prior, metas=metas,
                dn_targets=dict(labels=rel_labels, boxes=rel_norm), img=torch.zeros(bs, 3, 96, 128))


# ---- SSOD unsup_loss fixture (reference: DinoDetrSSOD.unsup_loss, dino_detr_ssod.py:204-482) -----------------------
# The heavy collaborators (query construction, the two decoder passes, the head loss) are replaced on BOTH sides by the
# same deterministic stand-ins below, so what is compared is the method's own logic: Hungarian costs of the student's
# predictions against the pseudo boxes, the GMM threshold, the double filter, and the cross-view consistency loss.
UNSUP_LOSS_KW = dict(num_query=40, num_classes=80, n_dec=3, score_thr=0.4)
UNSUP_LOSS_IMG_SHAPES = [(96, 128, 3), (80, 112, 3), (64, 64, 3)]
UNSUP_LOSS_COUNTS = [5, 0, 7]


def unsup_loss_inputs(seed=41):
    g = torch.Generator().manual_seed(seed)
    Q, K, bs = UNSUP_LOSS_KW["num_query"], UNSUP_LOSS_KW["num_classes"], len(UNSUP_LOSS_IMG_SHAPES)
    L = UNSUP_LOSS_KW["n_dec"]
    cls = torch.randn(L, bs, Q, K, generator=g) - 2.0
    box = torch.cat([torch.rand(L, bs, Q, 2, generator=g) * 0.8 + 0.1, torch.rand(L, bs, Q, 2, generator=g) * 0.3 + 0.03], -1)
    pseudo, labels, scores, det = [], [], [], []
    for n, (h, w, _) in zip(UNSUP_LOSS_COUNTS, UNSUP_LOSS_IMG_SHAPES):
        xy = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.6, h * 0.6])
        pseudo.append(torch.cat([xy, xy + torch.rand(n, 2, generator=g) * torch.tensor([w * 0.3, h * 0.3]) + 3.0], 1))
        labels.append(torch.randint(0, K, (n,), generator=g))
        scores.append(torch.rand(n, generator=g) * 0.7 + 0.05)            # some above, some below 0.4
        xy = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.6, h * 0.6])
        det.append(torch.cat([xy, xy + torch.rand(n, 2, generator=g) * torch.tensor([w * 0.3, h * 0.3]) + 3.0], 1))
    metas = [dict(img_shape=s) for s in UNSUP_LOSS_IMG_SHAPES]
    img = torch.zeros(bs, 3, 96, 128)
    student = dict(img=img, img_metas=metas, backbone_feature="student-feat",
                   outs=(cls, box, cls[-1], box[-1], None, None))
    teacher = dict(img=img, img_metas=metas, backbone_feature="teacher-feat", det_bboxes=det, det_labels=labels,
                   det_scores=scores)
    return student, teacher, pseudo, labels, scores


def fake_unsup_cdn(pseudo_bboxes, prior_info, hidden_dim=256, seed=0):
    """Stand-in for prepare_unsup_cdn: the real layout of part 1 (5 groups, dummy slot for an empty image), random
    content / weights (or the prior's), a 4-slot part 2."""
    counts = [max(int(b.shape[0]), 1) for b in pseudo_bboxes]
    bs, single = len(counts), max(counts)
    pad1 = 5 * single
    bid = torch.cat([torch.full((c,), i, dtype=torch.long) for i, c in enumerate(counts)]).repeat(5)
    slot = torch.cat([torch.cat([torch.arange(c) for c in counts]) + single * g for g in range(5)]).long()
    g = torch.Generator().manual_seed(1000 + seed + sum(counts))
    if prior_info is None:
        q1_label = torch.randn(bs, pad1, hidden_dim, generator=g)
        weights = (torch.rand(bid.numel(), 1, generator=g) > 0.2).float()
    else:
        q1_label, weights = prior_info["input_query_label_1"], prior_info["loss_weights"]
    meta = dict(pad_size_1=pad1, pad_size_2=4, num_dn_group_1=5, num_dn_group_2=1, known_bid_1=bid,
                map_known_indice_1=slot, loss_weights=weights)
    return (q1_label, torch.zeros(bs, pad1, 4), torch.zeros(bs, 4, hidden_dim), torch.zeros(bs, 4, 4),
            torch.zeros(pad1 + 4 + UNSUP_LOSS_KW["num_query"], pad1 + 4 + UNSUP_LOSS_KW["num_query"], dtype=torch.bool),
            meta)


def fake_forward_dummy(tag, q_label, hidden_dim=256):
    """Stand-in for bbox_head.forward_dummy: `n_dec` decoder states of the right shape, different per view."""
    bs, T = q_label.shape[0], q_label.shape[1] + UNSUP_LOSS_KW["num_query"]
    g = torch.Generator().manual_seed(77 if tag == "student" else 78)
    hs = [torch.randn(bs, T, hidden_dim, generator=g) for _ in range(UNSUP_LOSS_KW["n_dec"])]
    return (hs,) + (None,) * 8


# ---- teacher pseudo-label extraction fixture (reference: extract_teacher_info, dino_detr_ssod.py:893-951) ----------
def teacher_proposals(seed=53):
    """What ``simple_test_bboxes(..., for_pseudo_label=True)`` hands back per image: (n, 5) boxes with scores and (n,)
    labels -- incl. an image without detections, one with a single detection, and degenerate (zero-width) boxes."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in (40, 0, 1, 25):
        xy = torch.rand(n, 2, generator=g) * 200
        wh = torch.rand(n, 2, generator=g) * 80
        if n >= 25:
            wh[::6, 0] = 0.0                                   # degenerate boxes that must be dropped
        score = torch.rand(n, 1, generator=g) ** 2
        out.append((torch.cat([xy, xy + wh, score], 1), torch.randint(0, 80, (n,), generator=g)))
    return out


# ---- full head fixture (reference: DINODETRHead.__init__/_init_layers/forward + loss, dino_detr_head.py:74-632) ------
HEAD_CFG = dict(
    num_classes=9, in_channels=2048, num_query=30, num_feature_levels=4, num_backbone_outs=3,
    backbone_channels=[512, 1024, 2048], query_dim=4, dn_number=100, dn_box_noise_scale=0.4, dn_label_noise_ratio=0.5,
    dn_labelbook_size=9,
    transformer=dict(type="DINOTransformer", d_model=256, nhead=8, num_queries=30, num_encoder_layers=1,
                     num_decoder_layers=2, dim_feedforward=64, num_feature_levels=4),
    positional_encoding=dict(type="SinePositionalEncodingHW", num_feats=128, temperatureH=20, temperatureW=20,
                             normalize=True),
    loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
    loss_bbox=dict(type="L1Loss", loss_weight=5.0), loss_iou=dict(type="GIoULoss", loss_weight=2.0),
    train_cfg=dict(assigner=dict(type="HungarianAssigner", cls_cost=dict(type="FocalLossCost", weight=2.0),
                                 reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
                                 iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))),
    test_cfg=dict(max_per_img=300))


def head_inputs(seed=61):
    """Backbone features of a 96 x 128 padded batch (image 1 is 80 x 112), a 6-query denoising part, GT."""
    g = torch.Generator().manual_seed(seed)
    bs = 2
    feats = [torch.randn(bs, c, h, w, generator=g) * 0.5 for c, (h, w) in zip((512, 1024, 2048), LEVELS[:3])]
    metas = [dict(img_shape=(96, 128, 3), batch_input_shape=(96, 128)),
             dict(img_shape=(80, 112, 3), batch_input_shape=(96, 128))]
    q_label = torch.randn(bs, N_DN, 256, generator=g)
    q_bbox = torch.randn(bs, N_DN, 4, generator=g)
    T = N_DN + HEAD_CFG["num_query"]
    attn_mask = torch.zeros(T, T, dtype=torch.bool)
    attn_mask[N_DN:, :N_DN] = True
    dn_meta = dict(pad_size=N_DN, num_dn_group=1)
    gt_bboxes = [torch.tensor([[10., 12., 60., 70.], [40., 30., 120., 90.], [5., 50., 30., 80.]]),
                 torch.tensor([[20., 10., 100., 60.]])]
    gt_labels = [torch.tensor([1, 4, 7]), torch.tensor([3])]
    return dict(feats=feats, metas=metas, q_label=q_label, q_bbox=q_bbox, attn_mask=attn_mask, dn_meta=dn_meta,
                gt_bboxes=gt_bboxes, gt_labels=gt_labels)


# ---- backbone fixture (reference: mmdet ResNet-50 as the configs build it) ------------------------------------------
RESNET_KW = dict(depth=50, num_stages=4, out_indices=(1, 2, 3), frozen_stages=1,
                 norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="pytorch")


def fill_backbone(net):
    """By-name weights (He-like scale so activations neither vanish nor blow up over 50 layers) and non-trivial frozen
    BatchNorm statistics."""
    with torch.no_grad():
        for name, p in net.named_parameters():
            g = torch.Generator().manual_seed(zlib.crc32(("backbone." + name).encode()))
            r = torch.randn(p.shape, generator=g)
            if p.dim() == 4:
                p.copy_(r * (2.0 / (p.shape[1] * p.shape[2] * p.shape[3])) ** 0.5)
            elif name.endswith("weight"):
                p.copy_(1.0 + 0.1 * r)
            else:
                p.copy_(0.05 * r)
        for name, b in net.named_buffers():
            g = torch.Generator().manual_seed(zlib.crc32(("backbone." + name).encode()))
            if name.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif name.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
    return net


def backbone_input(seed=71):
    return torch.randn(2, 3, 96, 128, generator=torch.Generator().manual_seed(seed))


# second transformer case: 5 feature levels (BASELINE configs[3]), no denoising part, no padding, no attention mask
TRANSFORMER_KW5 = dict(d_model=256, nhead=8, num_queries=25, num_encoder_layers=2, num_decoder_layers=1,
                       dim_feedforward=96, num_feature_levels=5)
LEVELS5 = [(16, 20), (8, 10), (4, 5), (2, 3), (1, 2)]


def inputs5(seed=13):
    g = torch.Generator().manual_seed(seed)
    bs, C = 2, 256
    srcs = [torch.randn(bs, C, h, w, generator=g) for h, w in LEVELS5]
    poss = [torch.randn(bs, C, h, w, generator=g) * 0.5 for h, w in LEVELS5]
    masks = [torch.zeros(bs, h, w, dtype=torch.bool) for h, w in LEVELS5]
    return srcs, masks, poss


# ---- SSOD head fixture (reference: DINODETRSSODHead.__init__ / forward_dummy, dino_detr_ssod_head.py:77-505) --------
SSOD_HEAD_CFG = dict(
    num_classes=9, in_channels=2048, num_query=30, num_feature_levels=4, num_backbone_outs=3,
    backbone_channels=[512, 1024, 2048], query_dim=4, dn_number=100, dn_box_noise_scale=0.4, dn_label_noise_ratio=0.5,
    dn_labelbook_size=9, transformer=HEAD_CFG["transformer"], positional_encoding=HEAD_CFG["positional_encoding"],
    loss_cls1=dict(type="TaskAlignedFocalLoss", use_sigmoid=True, gamma=2.0, loss_weight=2.0),
    loss_cls2=HEAD_CFG["loss_cls"], loss_bbox=HEAD_CFG["loss_bbox"], loss_iou=HEAD_CFG["loss_iou"],
    train_cfg=dict(assigner1=dict(type="O2MAssigner"), assigner2=HEAD_CFG["train_cfg"]["assigner"], warm_up_step=50),
    test_cfg=dict(max_per_img=300, warm_up_step=50))


def ssod_head_inputs(seed=67):
    """head_inputs with the SSOD three-part query layout: 5 consistency slots | 4 denoising slots | matching."""
    x = head_inputs()
    g = torch.Generator().manual_seed(seed)
    p1, p2 = 5, 4
    x["q_label"] = torch.randn(2, p1 + p2, 256, generator=g)
    x["q_bbox"] = torch.randn(2, p1 + p2, 4, generator=g)
    T = p1 + p2 + SSOD_HEAD_CFG["num_query"]
    mask = torch.zeros(T, T, dtype=torch.bool)
    mask[p1 + p2:, :p1 + p2] = True
    mask[:p1, p1:p1 + p2] = True
    mask[p1:p1 + p2, :p1] = True
    x["attn_mask"] = mask
    x["dn_meta"] = dict(pad_size_1=p1, pad_size_2=p2, num_dn_group_1=5, num_dn_group_2=1)
    return x


# ---- SSOD wiring fixture (reference: foward_unsup_train / compute_pseudo_label_loss, dino_detr_ssod.py:154-201) ------
def unsup_wiring_inputs():
    """Three weak / strong pairs whose order differs between the two views (the method pairs them by filename), each
    with its own view matrices; teacher images are constant = their index so that the re-ordering is visible."""
    import numpy as np
    names = ["a.jpg", "b.jpg", "c.jpg"]
    order_t, order_s = [2, 0, 1], [0, 1, 2]
    flip = lambda w: np.array([[-1.0, 0.0, w], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    scale = lambda s: np.array([[s, 0.0, 0.0], [0.0, s, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    t_mats = {"a.jpg": scale(0.8), "b.jpg": flip(200.0), "c.jpg": np.eye(3, dtype=np.float32)}
    s_mats = {"a.jpg": flip(160.0) @ scale(0.8), "b.jpg": scale(1.25), "c.jpg": flip(128.0)}
    shapes = {"a.jpg": (120, 160, 3), "b.jpg": (150, 250, 3), "c.jpg": (96, 128, 3)}
    t_metas = [dict(filename=names[i], transform_matrix=t_mats[names[i]], img_shape=(100, 200, 3)) for i in order_t]
    s_metas = [dict(filename=names[i], transform_matrix=s_mats[names[i]], img_shape=shapes[names[i]]) for i in order_s]
    t_img = torch.stack([torch.full((3, 4, 4), float(i)) for i in order_t])
    s_img = torch.stack([torch.full((3, 4, 4), 10.0 + i) for i in order_s])
    teacher = dict(img=t_img, img_metas=t_metas)
    student = dict(img=s_img, img_metas=s_metas, gt_bboxes=[torch.zeros(0, 4)] * 3, gt_labels=[torch.zeros(0).long()] * 3)
    return teacher, student


def fake_teacher_detections(img_metas):
    """Stand-in for simple_test_bboxes(for_pseudo_label=True): detections that depend on the image's filename."""
    out = []
    for m in img_metas:
        g = torch.Generator().manual_seed(zlib.crc32(m["filename"].encode()))
        n = 12
        xy = torch.rand(n, 2, generator=g) * 150
        out.append((torch.cat([xy, xy + torch.rand(n, 2, generator=g) * 40 + 2, torch.rand(n, 1, generator=g) ** 2], 1),
                    torch.randint(0, 80, (n,), generator=g)))
    return out


def roi_inputs(seed=97, H=384, W=512, n=32):
    """Four feature levels (strides 8..64) of one 384 x 512 image pair and n RoIs (batch index, x1, y1, x2, y2) whose
    scales spread over the FPN level mapping (finest_scale 56: < 112 px -> level 0, < 224 -> 1, < 448 -> 2)."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(2, 256, -(-H // s), -(-W // s), generator=g) for s in (8, 16, 32, 64)]
    xy = torch.rand(n, 2, generator=g) * torch.tensor([W * 0.3, H * 0.3])
    wh = torch.exp(torch.rand(n, 1, generator=g) * 4.7 + 1.2) * torch.exp((torch.rand(n, 2, generator=g) - 0.5) * 0.6)   # ~3 .. 365 px
    boxes = torch.cat([xy, torch.minimum(xy + wh, torch.tensor([W - 1.0, H - 1.0]))], 1)
    rois = torch.cat([torch.randint(0, 2, (n, 1), generator=g).float(), boxes], 1)
    return feats, rois


def ssod_forward_train_inputs(seed=83):
    """An interleaved teacher-student batch as the SemiDataset collate would deliver it: 2 labelled images and 2 (weak,
    strong) pairs in mixed order; tiny images, per-sample ground truth lists."""
    g = torch.Generator().manual_seed(seed)
    tags = ["unsup_teacher", "sup", "unsup_student", "unsup_student", "sup", "unsup_teacher"]
    names = ["u1.jpg", "s0.jpg", "u0.jpg", "u1.jpg", "s1.jpg", "u0.jpg"]
    img = torch.randn(len(tags), 3, 8, 12, generator=g)
    metas = [dict(tag=t, filename=n, img_shape=(8, 12, 3), scale_factor=1.0) for t, n in zip(tags, names)]
    counts = [0, 3, 2, 1, 4, 2]
    gt_bboxes = [torch.rand(c, 4, generator=g) * 8 for c in counts]
    gt_labels = [torch.randint(0, 80, (c,), generator=g) for c in counts]
    return dict(img=img, img_metas=metas, gt_bboxes=gt_bboxes, gt_labels=gt_labels)
