"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own python code.

Run in the authoring container only (needs read access to /root/reference):

    python tests/golden/make_golden.py

Outputs (committed, small):
  msda_golden.npz       inputs + outputs + gradients of the reference's
                        ms_deform_attn_core_pytorch (functions/ms_deform_attn_func.py:41-61),
                        fp64, incl. the ops/test.py case (seed 3; test.py:21-36)
  hungarian_golden.npz  inputs + (gt_inds, labels) of the real mmdet HungarianAssigner.assign
                        (hungarian_assigner.py:55-188) with the DINO config's costs, and its cost matrix
  lsap_golden.npz       float32 cost matrices + scipy.optimize.linear_sum_assignment answers
                        (scipy 1.18.1), incl. tie KATs
  ema_golden.npz        MeanTeacher.momentum_update / before_train_iter results
                        (detr_ssod/utils/hooks/mean_teacher.py:37-64)
  dino_transformer_golden.npz  outputs of the reference's own DINOTransformer.forward (detr_od/models/utils/
                        transformer.py:1047-1406, with its MSDeformAttn module on the pure-PyTorch op) on the
                        deterministic weights / inputs of dino_fixture.py
"""
import os
import sys

import numpy as np
import scipy
import scipy.optimize
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402


def _msda_case(rng_seed, N, M, D, Lq, L, P, shapes, value_scale=1.0, loc_mode="uniform", test_py=False):
    g = torch.Generator().manual_seed(rng_seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    S = int((shapes_t[:, 0] * shapes_t[:, 1]).sum())
    start = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    if test_py:
        # ops/test.py:21-36 draws with the global RNG after torch.manual_seed(3)
        torch.manual_seed(3)
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    else:
        value = torch.randn(N, S, M, D, generator=g) * value_scale
        if loc_mode == "uniform":
            loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
        elif loc_mode == "wide":        # exercises the out-of-range branch and all border cases
            loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.6 - 0.3
        else:
            raise ValueError(loc_mode)
        attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value, shapes=shapes_t, start=start, loc=loc, attn=attn, gout=gout)


def make_msda():
    ref = R.load_msda_python()
    cases = {
        "testpy": _msda_case(3, 1, 2, 2, 2, 2, 2, [(6, 4), (3, 2)], test_py=True),
        "small_d32": _msda_case(11, 2, 8, 32, 19, 4, 4, [(7, 9), (4, 5), (2, 3), (1, 2)]),
        "wide_d32": _msda_case(12, 1, 8, 32, 23, 4, 4, [(6, 7), (3, 4), (2, 2), (1, 1)], loc_mode="wide"),
        "odd_d30": _msda_case(13, 1, 2, 30, 9, 2, 3, [(5, 4), (3, 2)], loc_mode="wide"),
        "d64_l5": _msda_case(14, 1, 2, 64, 7, 5, 4, [(4, 5), (3, 3), (2, 3), (1, 2), (1, 1)], loc_mode="wide"),
        "one_point": _msda_case(15, 3, 1, 4, 5, 1, 1, [(4, 6)], loc_mode="wide"),
    }
    out = {}
    for name, c in cases.items():
        v = c["value"].double().requires_grad_(True)
        l = c["loc"].double().requires_grad_(True)
        a = c["attn"].double().requires_grad_(True)
        y = ref.ms_deform_attn_core_pytorch(v, c["shapes"], l, a)
        y.backward(c["gout"].double())
        for k in ("value", "loc", "attn", "gout"):
            out[f"{name}/{k}"] = c[k].numpy()
        out[f"{name}/shapes"] = c["shapes"].numpy()
        out[f"{name}/start"] = c["start"].numpy()
        out[f"{name}/out"] = y.detach().numpy()
        out[f"{name}/grad_value"] = v.grad.numpy()
        out[f"{name}/grad_loc"] = l.grad.numpy()
        out[f"{name}/grad_attn"] = a.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "msda_golden.npz"), **out)
    print("msda_golden.npz:", list(cases))


def _gt(g, G, img_h, img_w):
    cx = torch.rand(G, generator=g) * 0.8 + 0.1
    cy = torch.rand(G, generator=g) * 0.8 + 0.1
    w = torch.rand(G, generator=g) * 0.45 + 0.05
    h = torch.rand(G, generator=g) * 0.45 + 0.05
    x1 = (cx - w / 2).clamp(0, 1) * img_w
    x2 = (cx + w / 2).clamp(0, 1) * img_w
    y1 = (cy - h / 2).clamp(0, 1) * img_h
    y2 = (cy + h / 2).clamp(0, 1) * img_h
    return torch.stack([x1, y1, x2, y2], -1), torch.randint(0, 80, (G,), generator=g)


def make_hungarian():
    ha, mc, iou = R.load_hungarian()
    assigner = ha.HungarianAssigner(
        cls_cost=dict(type="FocalLossCost", weight=2.0),
        reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
        iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))
    out = {}
    specs = [("q900_g7", 900, 7, 800, 1333), ("q300_g1", 300, 1, 800, 1333), ("q400_g30", 400, 30, 750, 1333),
             ("q900_g100", 900, 100, 800, 1201), ("q100_g0", 100, 0, 800, 1333), ("q50_g60", 50, 60, 600, 800),
             ("q300_g13", 300, 13, 512, 640)]
    for i, (name, Q, G, ih, iw) in enumerate(specs):
        g = torch.Generator().manual_seed(100 + i)
        bbox_pred = torch.rand(Q, 4, generator=g) * torch.tensor([1.0, 1.0, 0.5, 0.5]) + torch.tensor([0, 0, 0.01, 0.01])
        cls_pred = torch.randn(Q, 80, generator=g) * 2 - 3
        gt_b, gt_l = _gt(g, G, ih, iw)
        meta = dict(img_shape=(ih, iw, 3))
        res = assigner.assign(bbox_pred, cls_pred, gt_b, gt_l, meta)
        out[f"{name}/bbox_pred"] = bbox_pred.numpy()
        out[f"{name}/cls_pred"] = cls_pred.numpy()
        out[f"{name}/gt_bboxes"] = gt_b.numpy()
        out[f"{name}/gt_labels"] = gt_l.numpy()
        out[f"{name}/img_hw"] = np.array([ih, iw])
        out[f"{name}/gt_inds"] = res.gt_inds.numpy()
        out[f"{name}/labels"] = res.labels.numpy()
        if G > 0:
            factor = gt_b.new_tensor([iw, ih, iw, ih]).unsqueeze(0)
            from mmdet.core.bbox.transforms import bbox_cxcywh_to_xyxy
            cost = (assigner.cls_cost(cls_pred, gt_l) + assigner.reg_cost(bbox_pred, gt_b / factor)
                    + assigner.iou_cost(bbox_cxcywh_to_xyxy(bbox_pred) * factor, gt_b))
            out[f"{name}/cost"] = cost.numpy()
    # the reference's only KAT on this path: IoUCost doctest, match_cost.py:155-162
    b = torch.FloatTensor([[1, 1, 2, 2], [2, 2, 3, 4]])
    gtb = torch.FloatTensor([[0, 0, 2, 4], [1, 2, 3, 4]])
    out["ioucost_doctest/out"] = mc.IoUCost()(b, gtb).numpy()
    np.savez_compressed(os.path.join(HERE, "hungarian_golden.npz"), **out)
    print("hungarian_golden.npz:", [s[0] for s in specs])


def make_lsap():
    rng = np.random.default_rng(7)
    out = {}
    mats = {
        "zeros_4x2": np.zeros((4, 2), np.float32),
        "tie_3x2": np.array([[1, 1], [1, 1], [0, 0]], np.float32),
        "n_900x7": rng.standard_normal((900, 7)).astype(np.float32),
        "n_900x30": rng.standard_normal((900, 30)).astype(np.float32),
        "n_900x100": rng.standard_normal((900, 100)).astype(np.float32),
        "n_40x70": rng.standard_normal((40, 70)).astype(np.float32),
        "int_60x25": rng.integers(0, 3, (60, 25)).astype(np.float32),
        "int_25x60": rng.integers(0, 2, (25, 60)).astype(np.float32),
        "sq_33": np.round(rng.standard_normal((33, 33)) * 2).astype(np.float32),
        "inf_some": np.where(rng.random((30, 9)) < 0.3, np.inf, rng.standard_normal((30, 9))).astype(np.float32),
        "one_1x1": np.array([[3.5]], np.float32),
    }
    for k, c in mats.items():
        r, cc = scipy.optimize.linear_sum_assignment(c)
        out[f"{k}/cost"] = c
        out[f"{k}/rows"] = r.astype(np.int64)
        out[f"{k}/cols"] = cc.astype(np.int64)
    out["scipy_version"] = np.array(scipy.__version__)
    np.savez_compressed(os.path.join(HERE, "lsap_golden.npz"), **out)
    print("lsap_golden.npz:", list(mats))


def make_ema():
    mt = R.load_mean_teacher()
    g = torch.Generator().manual_seed(5)
    shapes = [(16, 3, 7, 7), (64,), (48, 40), (1, 1), (300, 4), (17,), (1025,), (3, 5, 7)]

    class Net(torch.nn.Module):
        def __init__(self, scale):
            super().__init__()
            self.ps = torch.nn.ParameterList(
                [torch.nn.Parameter(torch.randn(*s, generator=g) * scale) for s in shapes])
            self.register_buffer("buf", torch.randn(9, generator=g))     # buffers are NOT blended

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.teacher = Net(1.0)
            self.student = Net(0.5)

    class LogBuf:
        output = {}

    class Runner:
        pass

    model = Model()
    hook = mt.MeanTeacher(momentum=0.999, interval=1, warm_up=0)
    runner = Runner()
    runner.model = model
    runner.log_buffer = LogBuf()
    out = {}
    for i, p in enumerate(model.student.ps):
        out[f"student/{i}"] = p.detach().numpy().copy()
    for i, p in enumerate(model.teacher.ps):
        out[f"teacher0/{i}"] = p.detach().numpy().copy()
    moms = []
    for it in (1, 2, 7, 999, 5000):           # iter 0 is a plain copy (momentum 0)
        runner.iter = it
        hook.before_train_iter(runner)
        moms.append(runner.log_buffer.output["ema_momentum"])
        for i, p in enumerate(model.teacher.ps):
            out[f"teacher_after_{it}/{i}"] = p.detach().numpy().copy()
    out["iters"] = np.array([1, 2, 7, 999, 5000])
    out["momenta"] = np.array(moms, dtype=np.float64)
    out["buf_teacher"] = model.teacher.buf.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "ema_golden.npz"), **out)
    print("ema_golden.npz: momenta", moms)


def make_dino_transformer():
    import dino_fixture as F
    T, _ = R.load_dino_transformer()
    torch.manual_seed(0)
    model = F.fill_by_name(T.DINOTransformer(**F.TRANSFORMER_KW)).eval()
    C, K = F.TRANSFORMER_KW["d_model"], F.NUM_CLASSES
    n_dec = F.TRANSFORMER_KW["num_decoder_layers"]
    heads = torch.nn.ModuleDict(dict(
        fc_reg=torch.nn.ModuleList([T.MLP(C, C, 4, 3) for _ in range(n_dec)]),
        fc_cls=torch.nn.ModuleList([torch.nn.Linear(C, K) for _ in range(n_dec)]),
        fc_enc_reg=T.MLP(C, C, 4, 3), fc_enc_cls=torch.nn.Linear(C, K)))
    F.fill_by_name(heads, "heads.")
    srcs, masks, poss, refpoint, tgt, attn_mask = F.inputs()
    srcs = [s_.requires_grad_(True) for s_ in srcs]
    hs, references, hs_enc, ref_enc, init_box = model(srcs, masks, refpoint, poss, tgt, attn_mask,
                                                      fc_reg=heads["fc_reg"], fc_cls=heads["fc_cls"],
                                                      fc_enc_reg=heads["fc_enc_reg"], fc_enc_cls=heads["fc_enc_cls"])
    out = {}
    out["hs"] = torch.stack(list(hs)).detach().numpy()
    out["references"] = torch.stack(list(references)).detach().numpy()
    out["hs_enc"], out["ref_enc"] = hs_enc.detach().numpy(), ref_enc.detach().numpy()
    out["init_box_proposal"] = init_box.detach().numpy()
    # backward of a fixed scalar of every differentiable output (dino_fixture.scalar_loss): gradient norm of every
    # parameter, full gradients of a few small ones and of the finest input level
    F.scalar_loss(hs, references, hs_enc, ref_enc).backward()
    names = sorted(n for n, _ in model.named_parameters())
    params = dict(model.named_parameters())
    out["grad_norms"] = np.array([float(params[n].grad.norm()) if params[n].grad is not None else -1.0 for n in names])
    for n in F.FULL_GRADS:
        out["grad/" + n] = params[n].grad.numpy()
    out["grad_src3"] = srcs[3].grad.numpy()
    out["grad_head_norms"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                       for _, p in sorted(heads.named_parameters())])
    out["head_param_names"] = np.array(sorted(n for n, _ in heads.named_parameters()))
    out["param_names"] = np.array(sorted(n for n, _ in model.named_parameters()))
    # second case: 5 levels, 2 encoder layers, inference-style call (no denoising part, no mask)
    model5 = F.fill_by_name(T.DINOTransformer(**F.TRANSFORMER_KW5), "five.").eval()
    heads5 = torch.nn.ModuleDict(dict(
        fc_reg=torch.nn.ModuleList([T.MLP(C, C, 4, 3)]), fc_cls=torch.nn.ModuleList([torch.nn.Linear(C, K)]),
        fc_enc_reg=T.MLP(C, C, 4, 3), fc_enc_cls=torch.nn.Linear(C, K)))
    F.fill_by_name(heads5, "heads5.")
    s5, m5, p5 = F.inputs5()
    with torch.no_grad():
        hs5, ref5, hs_enc5, ref_enc5, init5 = model5(s5, m5, None, p5, None, None, fc_reg=heads5["fc_reg"],
                                                     fc_cls=heads5["fc_cls"], fc_enc_reg=heads5["fc_enc_reg"],
                                                     fc_enc_cls=heads5["fc_enc_cls"])
    out["five/hs"], out["five/references"] = torch.stack(list(hs5)).numpy(), torch.stack(list(ref5)).numpy()
    out["five/hs_enc"], out["five/ref_enc"], out["five/init_box_proposal"] = hs_enc5.numpy(), ref_enc5.numpy(), init5.numpy()
    np.savez_compressed(os.path.join(HERE, "dino_transformer_golden.npz"), **out)
    print("dino_transformer_golden.npz:", {k: v.shape for k, v in out.items()})


def make_dino_head_loss():
    """The reference's own DINODETRHead.loss on CPU tensors (its `.to('cuda')` placeholders are not reached: the
    fixture has a denoising part), with the DINO config's assigner / sampler / losses
    (configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:30-60)."""
    import dino_fixture as F
    m = R.load_dino_head()
    H = m["head"].DINODETRHead
    head = H.__new__(H)
    torch.nn.Module.__init__(head)
    K = F.LOSS_KW["num_classes"]
    head.num_classes, head.cls_out_channels, head.num_query = K, K, F.LOSS_KW["num_query"]
    head.bg_cls_weight, head.sync_cls_avg_factor = 0.0, False
    mc = sys.modules["mmdet.core.bbox.match_costs.match_cost"]
    head.assigner = m["assigner"].HungarianAssigner(
        cls_cost=dict(type="FocalLossCost", weight=2.0), reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
        iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))
    head.sampler = m["sampler"].PseudoSampler()
    head.loss_cls = m["focal"].FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0)
    head.loss_bbox = m["l1"].L1Loss(loss_weight=5.0)
    head.loss_iou = m["iou"].GIoULoss(loss_weight=2.0)
    x = F.loss_inputs()
    # the denoising targets place their index tensors with `.cuda()` (dino_detr_head.py:774-790): device placement
    # only, made a no-op here so the reference code runs on the CPU
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            losses = head.loss(x["all_cls_scores"], x["all_bbox_preds"], x["enc_cls_scores"], x["enc_bbox_preds"],
                               x["dn_cls_scores"], x["dn_bbox_preds"], x["gt_bboxes"], x["gt_labels"],
                               img_metas=x["img_metas"], dn_metas=x["dn_meta"])
    finally:
        torch.Tensor.cuda = saved
    out = {"keys": np.array(list(losses.keys())), "values": np.array([float(v) for v in losses.values()], np.float64)}
    np.savez_compressed(os.path.join(HERE, "dino_head_loss_golden.npz"), **out)
    print("dino_head_loss_golden.npz:", len(losses), "losses;", dict(list(losses.items())[:6]))


def make_dino_cdn():
    """The reference's own prepare_for_cdn on the CPU (`.cuda()` / `.to('cuda')` made no-ops), with every random draw
    recorded so that the test can replay the same noise through our fixed-shape RNG entry points."""
    import dino_fixture as F
    m = R.load_dino_head()
    dn = m["dn"]
    out = {}
    saved = torch.Tensor.cuda, torch.Tensor.to, torch.rand_like, torch.randint_like
    draws = []

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return self
        return saved[1](self, *a, **k)

    def rec(fn):
        def wrapped(*a, **k):
            r = fn(*a, **k)
            draws.append(r.clone())
            return r
        return wrapped
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    torch.rand_like, torch.randint_like = rec(saved[2]), rec(saved[3])
    try:
        for name, (counts, dn_number, lnr, bns, nq, K) in F.CDN_CASES.items():
            torch.manual_seed(100 + len(name))
            del draws[:]
            tg = F.cdn_targets(counts, K)
            emb = F.label_embedding(K)
            with torch.no_grad():
                ql, qb, mask, meta = dn.prepare_for_cdn((tg, dn_number, lnr, bns), True, nq, K, 256, emb)
            out[name + "/query_label"], out[name + "/query_bbox"] = ql.numpy(), qb.numpy()
            out[name + "/attn_mask"] = mask.numpy()
            out[name + "/meta"] = np.array([meta["pad_size"], meta["num_dn_group"]])
            for i, d in enumerate(draws):
                out[f"{name}/draw{i}"] = d.numpy()
            out[name + "/n_draws"] = np.array(len(draws))
            print(name, "pad", meta, "draws", [tuple(d.shape) for d in draws])
    finally:
        torch.Tensor.cuda, torch.Tensor.to, torch.rand_like, torch.randint_like = saved
    np.savez_compressed(os.path.join(HERE, "dino_cdn_golden.npz"), **out)


def make_ssod_pieces():
    """Small reference-pinned pieces of rows a10 / a12 / a14: SinePositionalEncodingHW (positional_encoding.py),
    Transform2D.transform_bboxes (bbox_utils.py:167-192) and O2MAssigner.assign (o2m_assigner.py:50-170)."""
    out = {}
    g = torch.Generator().manual_seed(77)
    # a14: the DINO setting (num_feats 128, temperature 20/20, normalize) on a padded batch
    pe = R.load_positional_encoding().SinePositionalEncodingHW(128, temperatureH=20, temperatureW=20, normalize=True)
    mask = torch.zeros(2, 7, 9, dtype=torch.bool)
    mask[1, 5:, :] = True
    mask[1, :, 6:] = True
    out["pe/mask"], out["pe/out"] = mask.numpy(), pe(mask).numpy()
    # a12: weak -> strong view warp (flip + scale + translate homographies, boxes with scores, clamp at the border)
    bu = R.load_bbox_utils()
    boxes = [torch.cat([torch.rand(6, 2, generator=g) * 200, torch.rand(6, 2, generator=g) * 200 + 210,
                        torch.rand(6, 1, generator=g)], 1), torch.zeros(0, 5),
             torch.cat([torch.rand(3, 2, generator=g) * 100, torch.rand(3, 2, generator=g) * 100 + 120], 1)]
    Ms = [torch.tensor([[-1.3, 0.0, 400.0], [0.0, 1.3, -20.0], [0.0, 0.0, 1.0]]),
          torch.eye(3), torch.tensor([[0.8, 0.1, 5.0], [-0.05, 0.9, 12.0], [0.0, 0.0, 1.0]])]
    shapes = [(300, 380, 3), (200, 200, 3), (256, 256, 3)]
    warped = bu.Transform2D.transform_bboxes(boxes, Ms, shapes)
    for i in range(3):
        out[f"warp/box{i}"], out[f"warp/M{i}"], out[f"warp/out{i}"] = boxes[i].numpy(), Ms[i].numpy(), warped[i].numpy()
    out["warp/shapes"] = np.array(shapes)
    # a10: O2M assignment (alpha 1, beta 6, top-13 candidates), predictions already sigmoid-ed
    o2m = R.load_o2m_assigner().O2MAssigner(candidate_topk=13)
    for ci, (Q, G, w, h) in enumerate([(200, 7, 640.0, 480.0), (60, 1, 320.0, 200.0), (150, 20, 500.0, 500.0)]):
        bbox = torch.rand(Q, 4, generator=g) * torch.tensor([1, 1, 0.5, 0.5]) + 0.01
        scores = torch.rand(Q, 80, generator=g)
        xy = torch.rand(G, 2, generator=g) * 0.5
        gtb = torch.cat([xy, xy + torch.rand(G, 2, generator=g) * 0.4 + 0.05], 1) * torch.tensor([w, h, w, h])
        gtl = torch.randint(0, 80, (G,), generator=g)
        res = o2m.assign(bbox, scores, gtb, gtl, dict(img_shape=(int(h), int(w), 3)))
        out[f"o2m{ci}/bbox"], out[f"o2m{ci}/scores"] = bbox.numpy(), scores.numpy()
        out[f"o2m{ci}/gtb"], out[f"o2m{ci}/gtl"], out[f"o2m{ci}/wh"] = gtb.numpy(), gtl.numpy(), np.array([w, h])
        out[f"o2m{ci}/gt_inds"], out[f"o2m{ci}/labels"] = res.gt_inds.numpy(), res.labels.numpy()
        out[f"o2m{ci}/max_overlaps"], out[f"o2m{ci}/assign_metrics"] = res.max_overlaps.numpy(), res.assign_metrics.numpy()
    # a3: the reference MSDeformAttn's deterministic initialisation (ms_deform_attn.py:67-76): the sampling-offset
    # bias grid (8 directions x point index), zero offset / attention weights and biases
    _, msda_mod = R.load_dino_transformer()
    for tag, (L, P) in (("l4p4", (4, 4)), ("l5p4", (5, 4))):
        torch.manual_seed(0)
        mref = msda_mod.MSDeformAttn(d_model=256, n_levels=L, n_heads=8, n_points=P)
        out[f"msda_init/{tag}/sampling_offsets.bias"] = mref.sampling_offsets.bias.detach().numpy()
        for n in ("sampling_offsets.weight", "attention_weights.weight", "attention_weights.bias", "value_proj.bias",
                  "output_proj.bias"):
            assert float(dict(mref.named_parameters())[n].abs().max()) == 0.0
    np.savez_compressed(os.path.join(HERE, "ssod_pieces_golden.npz"), **out)
    print("ssod_pieces_golden.npz:", len(out), "arrays; O2M positives",
          [int((out[f"o2m{c}/gt_inds"] > 0).sum()) for c in range(3)])


def make_dino_ssod_head_loss():
    """The reference's own DINODETRSSODHead.loss (dino_detr_ssod_head.py:508-1205) in its three regimes: warm-up on
    labelled data (O2M assignment + TaskAlignedFocalLoss, denoising part on), warm-up on pseudo labels (no denoising
    loss), and the Hungarian phase on pseudo labels (soft scores passed, second denoising block of the SSOD layout)."""
    import dino_fixture as F
    m = R.load_dino_ssod_head()
    H = m["ssod_head"].DINODETRSSODHead
    head = H.__new__(H)
    torch.nn.Module.__init__(head)
    K = F.LOSS_KW["num_classes"]
    head.num_classes, head.cls_out_channels, head.num_query = K, K, F.LOSS_KW["num_query"]
    head.bg_cls_weight, head.sync_cls_avg_factor = 0.0, False
    head.assigner1 = m["o2m"].O2MAssigner(candidate_topk=13)
    head.assigner2 = m["assigner"].HungarianAssigner(
        cls_cost=dict(type="FocalLossCost", weight=2.0), reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
        iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))
    head.sampler = m["sampler"].PseudoSampler()
    head.loss_cls1 = m["tal"].TaskAlignedFocalLoss(use_sigmoid=True, gamma=2.0, loss_weight=2.0)
    head.loss_cls2 = m["focal"].FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0)
    head.loss_bbox = m["l1"].L1Loss(loss_weight=5.0)
    head.loss_iou = m["iou"].GIoULoss(loss_weight=2.0)
    x = F.loss_inputs()
    g = torch.Generator().manual_seed(9)
    scores = [torch.rand(n, generator=g) * 0.5 + 0.4 for n in F.LOSS_GT_COUNTS]
    ssod_meta = dict(pad_size_2=x["dn_meta"]["pad_size"], num_dn_group_2=x["dn_meta"]["num_dn_group"],
                     pad_size=x["dn_meta"]["pad_size"], num_dn_group=x["dn_meta"]["num_dn_group"])
    saved = torch.Tensor.cuda, torch.Tensor.to

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return self
        return saved[1](self, *a, **k)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    out = {}
    try:
        for name, warm, pseudo, gts in (("warmup_sup", True, False, None), ("warmup_pseudo", True, True, scores),
                                        ("hungarian_pseudo", False, True, scores)):
            head.in_warm_up = warm
            with torch.no_grad():
                losses = head.loss(x["all_cls_scores"], x["all_bbox_preds"], x["enc_cls_scores"], x["enc_bbox_preds"],
                                   x["dn_cls_scores"], x["dn_bbox_preds"], x["gt_bboxes"], x["gt_labels"],
                                   gt_scores_list=gts, img_metas=x["img_metas"], dn_metas=ssod_meta,
                                   is_pseudo_label=pseudo)
            out[name + "/keys"] = np.array(list(losses.keys()))
            out[name + "/values"] = np.array([float(v) for v in losses.values()], np.float64)
            print(name, len(losses), [round(float(v), 4) for v in list(losses.values())[:7]])
    finally:
        torch.Tensor.cuda, torch.Tensor.to = saved
    for i, s_ in enumerate(scores):
        out[f"gt_scores{i}"] = s_.numpy()
    np.savez_compressed(os.path.join(HERE, "dino_ssod_head_loss_golden.npz"), **out)


def make_ssod_gmm():
    """The reference's own DinoDetrSSOD._fit_gmm (dino_detr_ssod.py:832-890; the method body compiled from the file,
    sklearn underneath) on float32 cost pools, as the wrapper calls it.  A pool is flagged `tie` when the two best
    log-likelihoods inside the chosen component differ by less than 1e-5: the reference then picks by float32 rounding
    noise (e.g. a component holding two points symmetric about its mean)."""
    import types
    import sklearn.mixture as skm
    fn = R.load_methods(R.REF + "/detr_ssod/models/dino_detr_ssod.py", "DinoDetrSSOD", ["_fit_gmm"],
                        dict(np=np, torch=torch, skm=skm))["_fit_gmm"]
    me = types.SimpleNamespace(covariance_type="diag")
    rng = np.random.default_rng(0)
    pools, thr, tie = [], [], []
    for t in range(160):
        k = int(rng.integers(1, 200)) if t > 5 else t            # sizes 0..5 first
        if t % 3 == 0:
            x = rng.normal(0, 1, k)
        elif t % 3 == 1:
            x = np.concatenate([rng.normal(-2, 0.5, k // 2 + 1), rng.normal(1.5, 0.8, k - k // 2)])[:max(k, 0)]
        else:
            x = np.concatenate([rng.normal(-1, 0.3, k // 3 + 1), rng.normal(3.0, 1.5, k)])
        x = x.astype(np.float32)
        r = fn(me, torch.from_numpy(x), device="cpu")
        r = float(np.asarray(r).reshape(-1)[0])
        is_tie = False
        if x.size >= 2:
            xs = np.sort(x).reshape(-1, 1)
            gm = skm.GaussianMixture(2, weights_init=np.array([.5, .5]), means_init=np.array([xs.min(), xs.max()]).reshape(2, 1),
                                     precisions_init=np.ones((2, 1)), covariance_type="diag", reg_covar=1e-5).fit(xs)
            a, sc = gm.predict(xs), gm.score_samples(xs)
            comp = 0 if (a == 0).any() else 1
            top = np.sort(sc[a == comp])[::-1]
            is_tie = top.size > 1 and (top[0] - top[1]) < 1e-5
        pools.append(x)
        thr.append(r)
        tie.append(is_tie)
    out = {f"pool{i}": p for i, p in enumerate(pools)}
    out["thresholds"], out["tie"] = np.array(thr, np.float64), np.array(tie)
    np.savez_compressed(os.path.join(HERE, "ssod_gmm_golden.npz"), **out)
    print("ssod_gmm_golden.npz:", len(pools), "pools,", int(np.sum(tie)), "ties")


def make_ssod_unsup_cdn():
    """The reference's own DinoDetrSSOD.prepare_unsup_cdn (dino_detr_ssod.py:484-760; method body compiled from the
    file) in its ``prior_info`` form -- the consistency content is given (the teacher pass re-uses the student's), so
    neither RoIAlign nor the projector runs -- with every random draw of the denoising part recorded."""
    import types
    import dino_fixture as F
    m = R.load_dino_head()
    T = sys.modules["detr_od_ref.models.utils.transformer"]
    tr = sys.modules["mmdet.core.bbox.transforms"]
    fn = R.load_methods(R.REF + "/detr_ssod/models/dino_detr_ssod.py", "DinoDetrSSOD", ["prepare_unsup_cdn"],
                        dict(torch=torch, np=np, inverse_sigmoid=T.inverse_sigmoid,
                             bbox_xyxy_to_cxcywh=tr.bbox_xyxy_to_cxcywh))["prepare_unsup_cdn"]
    kw = F.UNSUP_KW
    emb = F.label_embedding(kw["num_classes"])
    me = types.SimpleNamespace(curr_step=0, student=types.SimpleNamespace(
        bbox_head=types.SimpleNamespace(label_enc=emb, warm_up_step=10)))
    x = F.unsup_inputs()
    info = dict(img_metas=x["metas"], img=x["img"])
    saved = torch.Tensor.cuda, torch.Tensor.to, torch.rand_like, torch.randint_like, torch.randint
    draws = []

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return self
        return saved[1](self, *a, **k)

    def rec(fn_):
        def wrapped(*a, **k):
            r = fn_(*a, **k)
            draws.append(r.clone())
            return r
        return wrapped
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    torch.rand_like, torch.randint_like, torch.randint = rec(saved[2]), rec(saved[3]), rec(saved[4])
    try:
        torch.manual_seed(4)
        with torch.no_grad():
            q1l, q1b, q2l, q2b, mask, meta = fn(
                me, info, info, x["pseudo"], x["pseudo_labels"], x["det"], x["det_labels"],
                dn_args=(x["dn_targets"], kw["dn_number"], kw["label_noise_ratio"], kw["box_noise_scale"]),
                hidden_dim=kw["hidden_dim"], num_queries=kw["num_queries"], num_classes=kw["num_classes"],
                prior_info=x["prior"])
    finally:
        torch.Tensor.cuda, torch.Tensor.to, torch.rand_like, torch.randint_like, torch.randint = saved
    out = dict(q1_label=q1l.numpy(), q1_bbox=q1b.numpy(), q2_label=q2l.numpy(), q2_bbox=q2b.numpy(),
               attn_mask=mask.numpy(),
               meta=np.array([meta["pad_size_1"], meta["pad_size_2"], meta["num_dn_group_1"], meta["num_dn_group_2"]]),
               known_bid_1=meta["known_bid_1"].long().numpy(), map_known_indice_1=meta["map_known_indice_1"].numpy(),
               loss_weights=meta["loss_weights"].numpy(), n_draws=np.array(len(draws)))
    for i, d in enumerate(draws):
        out[f"draw{i}"] = d.numpy()
    np.savez_compressed(os.path.join(HERE, "ssod_unsup_cdn_golden.npz"), **out)
    print("ssod_unsup_cdn_golden.npz: meta", out["meta"], "draws", [tuple(d.shape) for d in draws])


def make_ssod_unsup_loss():
    """The reference's own DinoDetrSSOD.unsup_loss (dino_detr_ssod.py:204-482; method body compiled from the file) with
    its own _fit_gmm, the real mmdet match costs and scipy's solver; query construction, decoder passes and head loss
    are the deterministic stand-ins of dino_fixture.py (the same ones the test gives our method)."""
    import types
    import scipy.optimize
    import sklearn.mixture as skm
    import torch.nn.functional as TF
    import dino_fixture as F
    m = R.load_dino_head()
    tr = sys.modules["mmdet.core.bbox.transforms"]
    ns = dict(torch=torch, np=np, F=TF, skm=skm, bbox_cxcywh_to_xyxy=tr.bbox_cxcywh_to_xyxy,
              bbox_xyxy_to_cxcywh=tr.bbox_xyxy_to_cxcywh, linear_sum_assignment=scipy.optimize.linear_sum_assignment,
              get_dist_info=lambda: (0, 1), concat_all_gather=lambda t: t)
    fns = R.load_methods(R.REF + "/detr_ssod/models/dino_detr_ssod.py", "DinoDetrSSOD", ["unsup_loss", "_fit_gmm"], ns)
    assigner = m["assigner"].HungarianAssigner(
        cls_cost=dict(type="FocalLossCost", weight=2.0), reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
        iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))
    out = {}
    for phase, curr_step in (("warmup", 0), ("after", 100)):
        rec = {}

        def cdn(teacher_info, student_info, pb, pl, db, dl, dn_args=None, prior_info=None, **kw):
            key = "cdn2" if prior_info is not None else "cdn1"
            rec[key] = dict(pseudo=[b.clone() for b in pb], labels=[l.clone() for l in pl], det=[b.clone() for b in db],
                            det_labels=[l.clone() for l in dl], dn_boxes=[b.clone() for b in dn_args[0]["boxes"]],
                            dn_labels=[l.clone() for l in dn_args[0]["labels"]])
            return F.fake_unsup_cdn(pb, prior_info)

        def loss(*a, **k):
            rec["loss"] = dict(boxes=[b.clone() for b in k["gt_bboxes_list"]], labels=[l.clone() for l in k["gt_labels_list"]],
                               scores=[s_.clone() for s_ in k["gt_scores_list"]], pseudo=k["is_pseudo_label"])
            return dict(loss_cls=torch.tensor(1.0))
        head_s = types.SimpleNamespace(assigner2=assigner, warm_up_step=50, in_warm_up=None, dn_number=100,
                                       dn_label_noise_ratio=0.5, dn_box_noise_scale=0.4, loss=loss,
                                       forward_dummy=lambda feat, metas, ql, qb, mask, meta: F.fake_forward_dummy("student", ql))
        head_t = types.SimpleNamespace(forward_dummy=lambda feat, metas, ql, qb, mask, meta: F.fake_forward_dummy("teacher", ql))
        me = types.SimpleNamespace(curr_step=curr_step, covariance_type="diag",
                                   train_cfg=types.SimpleNamespace(pseudo_label_initial_score_thr=F.UNSUP_LOSS_KW["score_thr"]),
                                   student=types.SimpleNamespace(bbox_head=head_s),
                                   teacher=types.SimpleNamespace(bbox_head=head_t, extract_feat=lambda img: "teacher-feat"),
                                   prepare_unsup_cdn=cdn)
        me._fit_gmm = types.MethodType(fns["_fit_gmm"], me)
        student, teacher, pseudo, labels, scores = F.unsup_loss_inputs()
        losses = fns["unsup_loss"](me, student, teacher, pseudo, labels, scores)
        out[phase + "/keys"] = np.array(list(losses.keys()))
        out[phase + "/values"] = np.array([float(v) for v in losses.values()], np.float64)
        out[phase + "/in_warm_up"] = np.array(bool(head_s.in_warm_up))
        for key in ("cdn1", "cdn2"):
            for field, lst in rec[key].items():
                for i, t in enumerate(lst):
                    out[f"{phase}/{key}/{field}{i}"] = t.numpy()
        for field in ("boxes", "labels", "scores"):
            for i, t in enumerate(rec["loss"][field]):
                out[f"{phase}/loss/{field}{i}"] = t.numpy()
        print(phase, dict(zip(losses.keys(), [round(float(v), 6) for v in losses.values()])),
              "reliable", [len(b) for b in rec["loss"]["boxes"]], "high-recall", [len(b) for b in rec["cdn1"]["pseudo"]])
    np.savez_compressed(os.path.join(HERE, "ssod_unsup_loss_golden.npz"), **out)


def make_ssod_teacher_info():
    """The reference's own extract_teacher_info (dino_detr_ssod.py:893-951; method body compiled from the file) on fixed
    teacher detections: the per-image mean + std score filter and the removal of degenerate boxes."""
    import types
    import dino_fixture as F
    fn = R.load_methods(R.REF + "/detr_ssod/models/dino_detr_ssod.py", "DinoDetrSSOD", ["extract_teacher_info"],
                        dict(torch=torch, np=np))["extract_teacher_info"]
    props = F.teacher_proposals()
    feat = (torch.zeros(1, 1),)
    head = types.SimpleNamespace(simple_test_bboxes=lambda *a, **k: props)
    me = types.SimpleNamespace(curr_step=0, teacher=types.SimpleNamespace(extract_feat=lambda img: feat, bbox_head=head))
    metas = [dict(transform_matrix=np.eye(3, dtype=np.float32) * (i + 1)) for i in range(len(props))]
    info = fn(me, torch.zeros(len(props), 3, 8, 8), metas)
    out = {}
    for i in range(len(props)):
        out[f"det_bboxes{i}"], out[f"det_labels{i}"] = info["det_bboxes"][i].numpy(), info["det_labels"][i].numpy()
        out[f"det_scores{i}"], out[f"transform_matrix{i}"] = info["det_scores"][i].numpy(), info["transform_matrix"][i].numpy()
    np.savez_compressed(os.path.join(HERE, "ssod_teacher_info_golden.npz"), **out)
    print("ssod_teacher_info_golden.npz: kept", [len(b) for b in info["det_bboxes"]])


def make_dino_head_forward():
    """The reference's own DINODETRHead, constructed from a config like the shipped one (small transformer), by-name
    weights: forward (masks, positional encodings, input projections incl. the extra stride-2 level, transformer,
    shared heads, box refinement, denoising split) and the loss of its own outputs."""
    import dino_fixture as F
    m = R.load_dino_head_buildable()
    torch.manual_seed(0)
    head = m["head"].DINODETRHead(**F.HEAD_CFG)
    F.fill_by_name(head, "head.")
    head.eval()
    x = F.head_inputs()
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            outs = head(x["feats"], x["metas"], x["q_label"], x["q_bbox"], x["attn_mask"], x["dn_meta"])
            losses = head.loss(*outs, x["gt_bboxes"], x["gt_labels"], img_metas=x["metas"], dn_metas=x["dn_meta"])
    finally:
        torch.Tensor.cuda = saved
    out = {f"out{i}": o.numpy() for i, o in enumerate(outs)}
    out["loss_keys"] = np.array(list(losses.keys()))
    out["loss_values"] = np.array([float(v) for v in losses.values()], np.float64)
    out["param_names"] = np.array(sorted(n for n, _ in head.named_parameters()))
    # forward_train (:983-1046): GT normalisation -> prepare_for_cdn (noise recorded) -> forward -> loss, end to end
    saved = torch.Tensor.cuda, torch.Tensor.to, torch.rand_like, torch.randint_like
    draws = []

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return self
        return saved[1](self, *a, **k)

    def rec(fn_):
        def wrapped(*a, **k):
            r = fn_(*a, **k)
            draws.append(r.clone())
            return r
        return wrapped
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    torch.rand_like, torch.randint_like = rec(saved[2]), rec(saved[3])
    try:
        torch.manual_seed(8)
        with torch.no_grad():
            tl = head.forward_train(x["feats"], x["metas"], x["gt_bboxes"], x["gt_labels"])
    finally:
        torch.Tensor.cuda, torch.Tensor.to, torch.rand_like, torch.randint_like = saved
    out["train_keys"] = np.array(list(tl.keys()))
    out["train_values"] = np.array([float(v) for v in tl.values()], np.float64)
    for i, d in enumerate(draws):
        out[f"train_draw{i}"] = d.numpy()
    out["train_n_draws"] = np.array(len(draws))
    np.savez_compressed(os.path.join(HERE, "dino_head_forward_golden.npz"), **out)
    print("forward_train:", len(tl), "losses, draws", [tuple(d.shape) for d in draws])
    print("dino_head_forward_golden.npz:", [tuple(o.shape) for o in outs], len(losses), "losses")


def make_backbone():
    """mmdet's own ResNet-50 (backbones/resnet.py) as the DINO configs build it, in train() mode (frozen stem + layer1,
    BatchNorm in eval): the three output levels on by-name weights, and which parameters receive a gradient."""
    import dino_fixture as F
    rn = R.load_mmdet_resnet()
    net = F.fill_backbone(rn.ResNet(**F.RESNET_KW))
    net.train()                                   # mmdet's train() returns None
    x = F.backbone_input()
    outs = net(x)
    sum(o.square().mean() for o in outs).backward()
    out = {f"level{i}": o.detach()[:, ::8, ::2, ::2].numpy() for i, o in enumerate(outs)}
    out["level2_full"] = outs[2].detach().numpy()
    out["shapes"] = np.array([o.shape for o in outs])
    names = [n for n, _ in net.named_parameters()]
    params = dict(net.named_parameters())
    out["param_names"] = np.array(names)
    out["grad_norms"] = np.array([float(params[n].grad.norm()) if params[n].grad is not None else -1.0 for n in names])
    out["bn_training"] = np.array([m.training for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)])
    np.savez_compressed(os.path.join(HERE, "backbone_golden.npz"), **out)
    print("backbone_golden.npz:", [tuple(o.shape) for o in outs], "params with grad", int((out["grad_norms"] >= 0).sum()))


def make_dino_ssod_head_forward():
    """The reference's own DINODETRSSODHead built by its __init__ (both assigners, both classification losses), and its
    forward_dummy on the three-part query layout of the unsupervised pass."""
    import dino_fixture as F
    m = R.load_dino_ssod_head_buildable()
    torch.manual_seed(0)
    head = m["ssod_head"].DINODETRSSODHead(**F.SSOD_HEAD_CFG)
    F.fill_by_name(head, "head.")
    head.eval()
    x = F.ssod_head_inputs()
    with torch.no_grad():
        outs = head.forward_dummy(x["feats"], x["metas"], x["q_label"], x["q_bbox"], x["attn_mask"], x["dn_meta"])
    out = {"hs": torch.stack(list(outs[0])).numpy(), "param_names": np.array(sorted(n for n, _ in head.named_parameters())),
           "warm_up_step": np.array(head.warm_up_step), "in_warm_up": np.array(head.in_warm_up)}
    for i, o in enumerate(outs[1:], 1):
        out[f"out{i}"] = o.numpy()
    np.savez_compressed(os.path.join(HERE, "dino_ssod_head_forward_golden.npz"), **out)
    print("dino_ssod_head_forward_golden.npz:", [tuple(o.shape) for o in outs[1:]])


def make_ssod_wiring():
    """The reference's own foward_unsup_train -> compute_pseudo_label_loss chain (dino_detr_ssod.py:154-201) with its
    extract_teacher_info / extract_student_info / _get_trans_mat / _transform_bbox and the real Transform2D: pairing of
    the two views by filename, the teacher->student view matrix, the warped pseudo boxes handed to unsup_loss."""
    import types
    import dino_fixture as F
    bu = R.load_bbox_utils()
    names = ["foward_unsup_train", "compute_pseudo_label_loss", "extract_teacher_info", "extract_student_info",
             "_get_trans_mat", "_transform_bbox"]
    fns = R.load_methods(R.REF + "/detr_ssod/models/dino_detr_ssod.py", "DinoDetrSSOD", names,
                         dict(torch=torch, np=np, Transform2D=bu.Transform2D))
    rec = {}

    def unsup_loss(student_info, teacher_info, pseudo_bboxes, pseudo_labels, pseudo_scores):
        rec.update(student=student_info, teacher=teacher_info, boxes=pseudo_bboxes, labels=pseudo_labels,
                   scores=pseudo_scores)
        return dict(loss_x=torch.tensor(2.0))
    t_head = types.SimpleNamespace(simple_test_bboxes=lambda feat, metas, **k: F.fake_teacher_detections(metas))
    s_head = types.SimpleNamespace(forward=lambda feat, metas: ("outs", [m["filename"] for m in metas]))
    me = types.SimpleNamespace(curr_step=3, unsup_loss=unsup_loss,
                               teacher=types.SimpleNamespace(extract_feat=lambda img: (img,), bbox_head=t_head),
                               student=types.SimpleNamespace(extract_feat=lambda img: (img,), bbox_head=s_head))
    for n in names:
        setattr(me, n, types.MethodType(fns[n], me))
    teacher, student = F.unsup_wiring_inputs()
    loss = me.foward_unsup_train(teacher, student)
    out = {"loss_keys": np.array(list(loss.keys())), "teacher_img": rec["teacher"]["img"].numpy(),
           "teacher_names": np.array([m["filename"] for m in rec["teacher"]["img_metas"]]),
           "student_outs": np.array(rec["student"]["outs"][1])}
    for i in range(3):
        out[f"boxes{i}"], out[f"labels{i}"], out[f"scores{i}"] = (rec["boxes"][i].numpy(), rec["labels"][i].numpy(),
                                                                  rec["scores"][i].numpy())
        out[f"det{i}"] = rec["teacher"]["det_bboxes"][i].numpy()
        out[f"t_mat{i}"], out[f"s_mat{i}"] = (rec["teacher"]["transform_matrix"][i].numpy(),
                                              rec["student"]["transform_matrix"][i].numpy())
    np.savez_compressed(os.path.join(HERE, "ssod_wiring_golden.npz"), **out)
    print("ssod_wiring_golden.npz: teacher order", list(out["teacher_names"]), "kept", [len(b) for b in rec["boxes"]])


def make_ssod_decode():
    """The reference's own teacher decode for pseudo labels: ``DINODETRSSODHead._get_bboxes_single`` with
    ``for_pseudo_label=True`` (dino_detr_ssod_head.py:1331-1395; method body compiled from the file) calling the real
    mmdet ``multiclass_nms`` (thirdparty/mmdetection/mmdet/core/post_processing/bbox_nms.py:8-95, loaded from the file)
    and the real ``bbox_cxcywh_to_xyxy``.  The one piece that is not under /root/reference is mmcv-full 1.3.16's compiled
    ``mmcv.ops.nms`` (README.md:30): ``batched_nms`` is restated below from mmcv's published python (ops/nms.py:
    class-aware offset trick, then -- because the reference passes split_thr=-1 -- the per-class loop with the
    descending-score re-sort) on ``torchvision.ops.nms``, which implements the same greedy IoU > threshold suppression
    with offset 0.  Inputs: peaked / overlapping predictions so that NMS, the 0.01 threshold and max_per_img all bite."""
    import types
    import torchvision
    R._install_mmcv_stub()
    for pk in ("mmcv.ops", "mmdet", "mmdet.core", "mmdet.core.bbox", "mmdet.core.bbox.iou_calculators",
               "mmdet.core.post_processing"):
        R._pkg(pk)

    def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):        # mmcv/ops/nms.py (1.3.16), restated
        nms_cfg_ = dict(nms_cfg)
        assert not class_agnostic and nms_cfg_.pop("type", "nms") == "nms"
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        boxes_for_nms = boxes + offsets[:, None]
        split_thr = nms_cfg_.pop("split_thr", 10000)
        thr = nms_cfg_["iou_threshold"]
        if boxes_for_nms.shape[0] < split_thr:
            keep = torchvision.ops.nms(boxes_for_nms, scores, thr)
            boxes, scores = boxes[keep], scores[keep]
        else:
            total_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
            for cid in torch.unique(idxs):
                mask = (idxs == cid).nonzero(as_tuple=False).view(-1)
                k = torchvision.ops.nms(boxes_for_nms[mask], scores[mask], thr)
                total_mask[mask[k]] = True
            keep = total_mask.nonzero(as_tuple=False).view(-1)
            keep = keep[scores[keep].argsort(descending=True)]
            boxes, scores = boxes[keep], scores[keep]
        return torch.cat([boxes, scores[:, None]], -1), keep

    nms_mod = types.ModuleType("mmcv.ops.nms")
    nms_mod.batched_nms = batched_nms
    sys.modules["mmcv.ops.nms"] = nms_mod
    sys.modules["mmcv.ops"].nms = nms_mod
    R.load_hungarian()                      # iou2d_calculator (bbox_nms.py imports bbox_overlaps) and transforms, as for the assigner
    tr = sys.modules["mmdet.core.bbox.transforms"]
    bn = R._load("mmdet.core.post_processing.bbox_nms", R.MMDET + "/core/post_processing/bbox_nms.py")
    fn = R.load_methods(R.REF + "/detr_od/models/dense_heads/dino_detr_ssod_head.py", "DINODETRSSODHead",
                        ["_get_bboxes_single"],
                        dict(torch=torch, multiclass_nms=bn.multiclass_nms,
                             bbox_cxcywh_to_xyxy=tr.bbox_cxcywh_to_xyxy))["_get_bboxes_single"]
    g = torch.Generator().manual_seed(11)
    out = {}
    cases = [(900, 80, 300, (800, 1333, 3), 40), (300, 80, 100, (640, 480, 3), 12), (50, 80, 300, (200, 300, 3), 0)]
    for ci, (Q, C, max_per_img, img_shape, n_obj) in enumerate(cases):
        me = types.SimpleNamespace(test_cfg=dict(max_per_img=max_per_img), num_query=Q, num_classes=C, in_warm_up=False,
                                   loss_cls2=types.SimpleNamespace(use_sigmoid=True))
        logits = torch.randn(Q, C, generator=g) * 1.2 - 6.5                  # most (query, class) pairs below 0.01
        boxes = torch.rand(Q, 4, generator=g) * torch.tensor([0.8, 0.8, 0.3, 0.3]) + torch.tensor([0.1, 0.1, 0.02, 0.02])
        for o in range(n_obj):                                                # clusters of near-duplicate detections
            ctr = torch.rand(4, generator=g) * torch.tensor([0.7, 0.7, 0.3, 0.3]) + torch.tensor([0.15, 0.15, 0.05, 0.05])
            members = torch.randperm(Q, generator=g)[:int(torch.randint(2, 9, (1,), generator=g))]
            cls = int(torch.randint(0, C, (1,), generator=g))
            boxes[members] = ctr + torch.randn(len(members), 4, generator=g) * 0.01
            logits[members, cls] = torch.randn(len(members), generator=g) * 1.5 + 1.0
            if o % 5 == 0:                                                    # a second class on the same boxes
                logits[members, (cls + 3) % C] = torch.randn(len(members), generator=g) - 1.0
        boxes = boxes.clamp(0.001, 0.999)
        det, lab = fn(me, logits.clone(), boxes.clone(), img_shape, None, rescale=False, for_pseudo_label=True)
        out[f"c{ci}/logits"], out[f"c{ci}/boxes"] = logits.numpy(), boxes.numpy()
        out[f"c{ci}/img_shape"], out[f"c{ci}/max_per_img"] = np.asarray(img_shape), np.asarray(max_per_img)
        out[f"c{ci}/det_bboxes"], out[f"c{ci}/det_labels"] = det.numpy(), lab.numpy()
        print(f"ssod_decode case {ci}: {int((logits.sigmoid() > 0.01).sum())} candidates -> {len(lab)} detections")
    np.savez_compressed(os.path.join(HERE, "ssod_decode_golden.npz"), **out)


def make_ssod_roi_projector():
    """Content of the consistency queries (dino_detr_ssod.py:592-607): the real mmdet ``SingleRoIExtractor``
    (roi_heads/roi_extractors/single_level_roi_extractor.py + base_roi_extractor.py, loaded from the files, built with
    the reference's config dino_detr_ssod.py:97-100) followed by the reference's own ``Projector`` class
    (dino_detr_ssod.py:33-72, class body compiled from the file), by-name deterministic weights, BatchNorm in training
    mode as in the train step.  Outside /root/reference: mmcv-full 1.3.16's compiled ``mmcv.ops.RoIAlign`` -- restated as
    a module over ``torchvision.ops.roi_align`` with mmcv's defaults for this call (pool_mode 'avg', aligned=True)."""
    import ast
    import types
    import torchvision
    import dino_fixture as F
    R._install_mmcv_stub()
    for pk in ("mmcv.ops", "mmdet", "mmdet.models", "mmdet.models.roi_heads", "mmdet.models.roi_heads.roi_extractors"):
        R._pkg(pk)

    class RoIAlign(torch.nn.Module):                  # mmcv/ops/roi_align.py (1.3.16) restated on torchvision's op
        def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True):
            super().__init__()
            assert pool_mode == "avg"
            self.output_size = torch.nn.modules.utils._pair(output_size)
            self.spatial_scale, self.sampling_ratio, self.aligned = float(spatial_scale), int(sampling_ratio), aligned

        def forward(self, feat, rois):
            return torchvision.ops.roi_align(feat, rois, self.output_size, self.spatial_scale, self.sampling_ratio,
                                             self.aligned)
    sys.modules["mmcv.ops"].RoIAlign = RoIAlign
    sys.modules["mmcv"].ops = sys.modules["mmcv.ops"]
    runner = sys.modules["mmcv.runner"]

    class BaseModule(torch.nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
    runner.BaseModule = BaseModule
    runner.force_fp32 = lambda *a, **k: (lambda f: f)
    builder = types.ModuleType("mmdet.models.builder")
    builder.ROI_EXTRACTORS = R._Registry("roi_extractor")
    sys.modules["mmdet.models.builder"] = builder
    sys.modules["mmdet.models"].builder = builder
    d = R.MMDET + "/models/roi_heads/roi_extractors"
    R._load("mmdet.models.roi_heads.roi_extractors.base_roi_extractor", d + "/base_roi_extractor.py")
    sl = R._load("mmdet.models.roi_heads.roi_extractors.single_level_roi_extractor", d + "/single_level_roi_extractor.py")
    extractor = sl.SingleRoIExtractor(roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=0), out_channels=256,
                                      featmap_strides=[8, 16, 32, 64])
    path = R.REF + "/detr_ssod/models/dino_detr_ssod.py"
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Projector")
    ns = dict(nn=torch.nn, torch=torch)
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)
    proj = ns["Projector"]()
    F.fill_by_name(proj)
    proj.train()
    feats, rois = F.roi_inputs()               # seeded; the test regenerates them instead of shipping 6 MB of noise
    pooled = extractor(feats, rois)
    lvls = extractor.map_roi_levels(rois, 4)
    emb = proj(pooled)
    out = dict(rois=rois.numpy(), pooled_every_8th_channel=pooled.detach().numpy()[:, ::8], levels=lvls.numpy(),
               embed=emb.detach().numpy(), feat_checksum=np.array([float(f.double().sum()) for f in feats]),
               names=np.array([k for k, _ in proj.state_dict().items()]))
    np.savez_compressed(os.path.join(HERE, "ssod_roi_projector_golden.npz"), **out)
    print("ssod_roi_projector_golden.npz: levels", np.bincount(lvls.numpy(), minlength=4).tolist(),
          "embed mean", float(emb.detach().mean()))


def make_ssod_forward_train():
    """The reference's own ``DinoDetrSSOD.forward_train`` (dino_detr_ssod.py:112-152; method body compiled from the file,
    minus its first statement -- ``super().forward_train(...)``, a call into MultiSteamDetector that needs the class
    cell) with the real ``dict_split`` / ``dict_select`` / ``weighted_loss`` of detr_ssod/utils/structure_utils.py
    (loaded from the file; ``collections.Mapping`` aliased for python >= 3.10, mmdet's BitmapMasks stubbed).  The student
    and ``foward_unsup_train`` are recorders: what each receives from an interleaved batch, and the loss dict that
    comes back (prefixes, the 4.0 weight on keys containing 'loss')."""
    import ast
    import collections
    import collections.abc
    import types
    collections.Mapping, collections.Sequence = collections.abc.Mapping, collections.abc.Sequence
    for pk in ("mmdet", "mmdet.core", "mmdet.core.mask"):
        R._pkg(pk)
    st = types.ModuleType("mmdet.core.mask.structures")
    st.BitmapMasks = type("BitmapMasks", (), {})
    sys.modules["mmdet.core.mask.structures"] = st
    su = R._load("detr_ssod_ref.utils.structure_utils", R.REF + "/detr_ssod/utils/structure_utils.py") \
        if "detr_ssod_ref.utils" in sys.modules else None
    if su is None:
        for pk in ("detr_ssod_ref", "detr_ssod_ref.utils"):
            R._pkg(pk)
        su = R._load("detr_ssod_ref.utils.structure_utils", R.REF + "/detr_ssod/utils/structure_utils.py")
    path = R.REF + "/detr_ssod/models/dino_detr_ssod.py"
    tree = ast.parse(open(path).read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DinoDetrSSOD")
    node = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward_train")
    assert isinstance(node.body[0], ast.Expr) and "super" in ast.unparse(node.body[0])
    node.body = node.body[1:]
    ns = dict(torch=torch, dict_split=su.dict_split, weighted_loss=su.weighted_loss, log_every_n=lambda *a, **k: None)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    import dino_fixture as F
    data = F.ssod_forward_train_inputs()
    rec = {}

    def student_forward_train(img, img_metas, gt_bboxes, gt_labels, curr_step=None, **kw):
        rec["sup"] = dict(img=img, names=[m["filename"] for m in img_metas], gt_bboxes=gt_bboxes, gt_labels=gt_labels,
                          curr_step=curr_step, extra=sorted(kw))
        return dict(loss_cls=torch.tensor(1.5), loss_bbox=torch.tensor(0.25), pos_num=torch.tensor(3.0))

    def foward_unsup_train(teacher_data, student_data):
        rec["teacher"], rec["student"] = teacher_data, student_data
        return dict(loss_cls=torch.tensor(2.0), **{"d0.loss_iou": torch.tensor(0.5)}, consis_count=torch.tensor(7.0))
    me = types.SimpleNamespace(student=types.SimpleNamespace(forward_train=student_forward_train),
                               foward_unsup_train=foward_unsup_train, curr_step=1234, unsup_weight=4.0)
    loss = ns["forward_train"](me, data["img"], data["img_metas"], gt_bboxes=data["gt_bboxes"],
                               gt_labels=data["gt_labels"])
    out = {"loss_keys": np.array(sorted(loss)), "loss_vals": np.array([float(loss[k]) for k in sorted(loss)]),
           "sup/img": rec["sup"]["img"].numpy(), "sup/names": np.array(rec["sup"]["names"]),
           "sup/curr_step": np.asarray(rec["sup"]["curr_step"]), "sup/extra": np.array(rec["sup"]["extra"], dtype=str)}
    for i, (b, l) in enumerate(zip(rec["sup"]["gt_bboxes"], rec["sup"]["gt_labels"])):
        out[f"sup/gt_bboxes{i}"], out[f"sup/gt_labels{i}"] = b.numpy(), l.numpy()
    for side in ("teacher", "student"):
        d = rec[side]
        out[f"{side}/keys"] = np.array(sorted(d))
        out[f"{side}/img"] = d["img"].numpy()
        out[f"{side}/names"] = np.array([m["filename"] for m in d["img_metas"]])
        out[f"{side}/tags"] = np.array([m["tag"] for m in d["img_metas"]])
        for i, (b, l) in enumerate(zip(d["gt_bboxes"], d["gt_labels"])):
            out[f"{side}/gt_bboxes{i}"], out[f"{side}/gt_labels{i}"] = b.numpy(), l.numpy()
    np.savez_compressed(os.path.join(HERE, "ssod_forward_train_golden.npz"), **out)
    print("ssod_forward_train_golden.npz:", dict(zip(out["loss_keys"].tolist(), out["loss_vals"].tolist())),
          "sup", out["sup/names"].tolist(), "teacher", out["teacher/names"].tolist(), "student", out["student/names"].tolist())


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_msda()
    make_hungarian()
    make_lsap()
    make_ema()
    make_dino_transformer()
    make_dino_head_loss()
    make_dino_cdn()
    make_ssod_pieces()
    make_dino_ssod_head_loss()
    make_ssod_gmm()
    make_ssod_unsup_cdn()
    make_ssod_unsup_loss()
    make_ssod_teacher_info()
    make_dino_head_forward()
    make_backbone()
    make_dino_ssod_head_forward()
    make_ssod_wiring()
    make_ssod_decode()
    make_ssod_roi_projector()
    make_ssod_forward_train()
