"""Seeded synthetic inputs shared by the tests and bench.py."""
import torch

MICROBENCH_LEVELS = [(100, 134), (50, 67), (25, 34), (13, 17)]      # sum HW = 17821 (800x1066 image)
COCO_4SCALE_LEVELS = [(100, 167), (50, 84), (25, 42), (13, 21)]     # sum HW = 22223 (800x1333 image)


def level_tensors(levels, device):
    shapes = torch.as_tensor(levels, dtype=torch.long, device=device)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return shapes, start


def encoder_reference_points(levels, device):
    """Pixel centres of every level, (S, 2) as (x, y) in [0,1] (transformer.py:676-691 with valid_ratio 1)."""
    pts = []
    for (h, w) in levels:
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device) + 0.5,
                                torch.arange(w, dtype=torch.float32, device=device) + 0.5, indexing="ij")
        pts.append(torch.stack([xs.reshape(-1) / w, ys.reshape(-1) / h], -1))
    return torch.cat(pts)


def msda_inputs(levels, N, M=8, D=32, P=4, Lq=None, mode="encoder", seed=0, device="cuda", dtype=torch.float32):
    """value ~ N(0,1); loc: 'encoder' = own grid point + U(-4,4) px per level, 'uniform' = U(0,1),
    'wide' = U(-0.3,1.3) (exercises borders / out-of-range); attn = softmax(N(0,1)) over L*P; grad ~ N(0,1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = len(levels)
    S = sum(h * w for h, w in levels)
    shapes, start = level_tensors(levels, device)
    value = torch.randn(N, S, M, D, generator=g).to(device=device, dtype=dtype)
    if mode == "encoder":
        Lq = S
        ref = encoder_reference_points(levels, "cpu")                       # (S, 2)
        wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)  # (L, 2)
        off = (torch.rand(N, Lq, M, L, P, 2, generator=g) * 8 - 4) / wh[None, None, None, :, None, :]
        loc = ref[None, :, None, None, None, :] + off
    else:
        Lq = Lq or 1100
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
        if mode == "wide":
            loc = loc * 1.6 - 0.3
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value, shapes=shapes, start=start, loc=loc.to(device=device, dtype=dtype).contiguous(),
                attn=attn.to(device=device, dtype=dtype).contiguous(), gout=gout.to(device=device, dtype=dtype))


def msda_bytes(N, S, Lq, M=8, D=32, L=4, P=4, elt=4):
    """Algorithmic HBM bytes: fwd = value + loc + attn + out;
    bwd = grad_out + value + loc + attn (reads) + grad_value + grad_loc + grad_attn (writes)."""
    v, o = N * S * M * D, N * Lq * M * D
    loc, att = N * Lq * M * L * P * 2, N * Lq * M * L * P
    return elt * (v + loc + att + o), elt * (o + 2 * v + 2 * loc + 2 * att)


# ---- DINO train-step inputs ----------------------------------------

DINO_R50_4SCALE = dict(
    type="DINODETR",
    backbone=dict(type="ResNet", depth=50, num_stages=4, out_indices=(1, 2, 3), frozen_stages=1,
                  norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="pytorch"),
    bbox_head=dict(type="DINODETRHead", num_query=900, query_dim=4, random_refpoints_xy=False,
                   bbox_embed_diff_each_layer=False, num_classes=80, in_channels=2048,
                   transformer=dict(type="DINOTransformer"),
                   positional_encoding=dict(type="SinePositionalEncodingHW", temperatureH=20, temperatureW=20,
                                            num_feats=128, normalize=True),
                   loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
                   loss_bbox=dict(type="L1Loss", loss_weight=5.0), loss_iou=dict(type="GIoULoss", loss_weight=2.0)),
    train_cfg=dict(assigner=dict(type="HungarianAssigner", cls_cost=dict(type="FocalLossCost", weight=2.0),
                                 reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
                                 iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))),
    test_cfg=dict(max_per_img=300))
"""The ``model`` dict of configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:7-45 (init_cfg / pretrained weights dropped:
no network, random init)."""


def dino_r50_5scale():
    """the 5-scale variant -- all four ResNet stages feed the transformer plus one
    stride-2 extra level (head kwargs ``dino_detr_head.py:80-82``: num_feature_levels=5, num_backbone_outs=4,
    backbone_channels=[256, 512, 1024, 2048]; transformer num_feature_levels=5).  The reference ships no such config
    file; the head and transformer take these kwargs unchanged."""
    import copy
    cfg = copy.deepcopy(DINO_R50_4SCALE)
    cfg["backbone"]["out_indices"] = (0, 1, 2, 3)
    cfg["bbox_head"].update(num_feature_levels=5, num_backbone_outs=4, backbone_channels=[256, 512, 1024, 2048],
                            transformer=dict(type="DINOTransformer", num_feature_levels=5))
    return cfg


def coco_like_batch(batch_size=2, height=800, width=1333, seed=0, device="cpu", pin=False):
    """Synthetic COCO-shape batch: images ~ N(0,1); per image G ~ clamp(Poisson(7), 1, 100) boxes with
    cx,cy ~ U(.1,.9), w,h ~ U(.05,.5) clipped to the image, labels ~ U{0..79}."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch_size, 3, height, width, generator=g)
    gt_bboxes, gt_labels, metas = [], [], []
    for _ in range(batch_size):
        n = int(torch.poisson(torch.tensor(7.0), generator=g).clamp(1, 100))
        cxcy = torch.rand(n, 2, generator=g) * 0.8 + 0.1
        wh = torch.rand(n, 2, generator=g) * 0.45 + 0.05
        x1y1 = (cxcy - wh / 2).clamp(0, 1)
        x2y2 = (cxcy + wh / 2).clamp(0, 1)
        gt_bboxes.append(torch.cat([x1y1, x2y2], 1) * torch.tensor([width, height, width, height], dtype=torch.float32))
        gt_labels.append(torch.randint(0, 80, (n,), generator=g))
        metas.append(dict(img_shape=(height, width, 3), pad_shape=(height, width, 3), ori_shape=(height, width, 3),
                          batch_input_shape=(height, width), scale_factor=1.0))
    if pin:
        img = img.pin_memory()
        gt_bboxes = [b.pin_memory() for b in gt_bboxes]
        gt_labels = [l.pin_memory() for l in gt_labels]
    if device != "cpu":
        img = img.to(device, non_blocking=True)
        gt_bboxes = [b.to(device, non_blocking=True) for b in gt_bboxes]
        gt_labels = [l.to(device, non_blocking=True) for l in gt_labels]
    return dict(img=img, img_metas=metas, gt_bboxes=gt_bboxes, gt_labels=gt_labels)


# ---- Semi-DETR teacher-student step inputs -----------------------------

def ssod_model_cfg(warm_up_step=60000):
    """``semi_wrapper`` of configs/detr_ssod/detr_ssod_dino_detr_r50_coco_120k.py:27-39 around the model of
    configs/dino_detr/dino_detr_ssod_r50_coco_120k.py:7-53."""
    import copy
    model = copy.deepcopy(DINO_R50_4SCALE)
    head = model["bbox_head"]
    head["type"] = "DINODETRSSODHead"
    head["loss_cls1"] = dict(type="TaskAlignedFocalLoss", use_sigmoid=True, gamma=2.0, loss_weight=2.0)
    head["loss_cls2"] = head.pop("loss_cls")
    model["train_cfg"] = dict(assigner1=dict(type="O2MAssigner"), assigner2=model["train_cfg"]["assigner"],
                              warm_up_step=warm_up_step)
    model["test_cfg"] = dict(max_per_img=300, warm_up_step=warm_up_step)
    return dict(type="DinoDetrSSOD", model=model,
                train_cfg=dict(use_teacher_proposal=False, pseudo_label_initial_score_thr=0.4, min_pseduo_box_size=0,
                               unsup_weight=4.0, aug_query=False),
                test_cfg=dict(inference_on="student"))


def ssod_batch(n_sup=1, n_unsup=4, height=800, width=1333, seed=0, device="cpu"):
    """1 labelled + ``n_unsup`` unlabelled pairs per GPU (sample_ratio [1, 4], samples_per_gpu 5): the weak
    (teacher) and strong (student) views are the same synthetic image, the strong one horizontally flipped, with the
    3x3 ``transform_matrix`` each pipeline would have recorded."""
    import numpy as np
    sup = coco_like_batch(n_sup, height, width, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    weak = torch.randn(n_unsup, 3, height, width, generator=g)
    strong = weak.flip(-1)
    flip = np.array([[-1.0, 0.0, width], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    metas, imgs = [], []
    for m in sup["img_metas"]:
        m.update(tag="sup", filename=f"sup{len(metas)}", transform_matrix=np.eye(3, dtype=np.float32))
        metas.append(m)
    shape = dict(img_shape=(height, width, 3), pad_shape=(height, width, 3), ori_shape=(height, width, 3),
                 scale_factor=1.0)
    for i in range(n_unsup):
        metas.append(dict(shape, tag="unsup_teacher", filename=f"u{i}", transform_matrix=np.eye(3, dtype=np.float32)))
    for i in range(n_unsup):
        metas.append(dict(shape, tag="unsup_student", filename=f"u{i}", transform_matrix=flip))
    img = torch.cat([sup["img"], weak, strong])
    empty_b, empty_l = torch.zeros(0, 4), torch.zeros(0, dtype=torch.long)
    gt_bboxes = sup["gt_bboxes"] + [empty_b.clone() for _ in range(2 * n_unsup)]
    gt_labels = sup["gt_labels"] + [empty_l.clone() for _ in range(2 * n_unsup)]
    if device != "cpu":
        img = img.to(device)
        gt_bboxes = [b.to(device) for b in gt_bboxes]
        gt_labels = [l.to(device) for l in gt_labels]
    return dict(img=img, img_metas=metas, gt_bboxes=gt_bboxes, gt_labels=gt_labels)
