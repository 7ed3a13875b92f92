"""Host logic of the teacher-student step on the CPU: O2M assigner vs a loop restatement of the reference, GMM
threshold vs sklearn, box warp, SSOD attention-mask layout, and a full DinoDetrSSOD step in both phases."""
import numpy as np
import pytest
import torch

from oracle.cpu_path import reference_cpu_ops
from semi_detr_b200 import dino, ssod  # noqa: F401
from semi_detr_b200.registry import DETECTORS
from semi_detr_b200.ssod.bbox_utils import Transform2D
from semi_detr_b200.ssod.dino_detr_ssod import DinoDetrSSOD, weighted_loss
from oracle.gmm_oracle import fit_gmm_threshold
from semi_detr_b200.ssod.o2m_assigner import INF, O2MAssigner, normalized_alignment_metrics, pairwise_iou
from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg


def _o2m_reference_loop(bbox_pred, scores, gt_bboxes, gt_labels, w, h, topk=13, alpha=1, beta=6):
    """o2m_assigner.py:95-170 written with the reference's per-GT python loops."""
    cx, cy, bw, bh = bbox_pred.unbind(-1)
    pred = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], -1) * torch.tensor([w, h, w, h])
    overlaps = pairwise_iou(pred, gt_bboxes)
    metrics = scores[:, gt_labels] ** alpha * overlaps ** beta
    Q, G = overlaps.shape
    _, cand = metrics.topk(topk, dim=0)
    cand_metrics = metrics[cand, torch.arange(G)]
    is_pos = cand_metrics > 0
    for g in range(G):
        cand[:, g] += g * Q
    ov_inf = torch.full_like(overlaps, -INF).t().contiguous().view(-1)
    index = cand.view(-1)[is_pos.view(-1)]
    ov_inf[index] = overlaps.t().contiguous().view(-1)[index]
    ov_inf = ov_inf.view(G, -1).t()
    mx, arg = ov_inf.max(dim=1)
    gt_inds = torch.zeros(Q, dtype=torch.long)
    gt_inds[mx != -INF] = arg[mx != -INF] + 1
    am = torch.zeros(Q)
    am[mx != -INF] = metrics[mx != -INF, arg[mx != -INF]]
    # head-side normalisation, dino_detr_ssod_head.py:1146-1157
    ious = mx.clone()
    ious[ious == -INF] = 0
    norm = torch.zeros(Q)
    pos_inds = torch.nonzero(gt_inds > 0).reshape(-1)
    pag = gt_inds[pos_inds] - 1
    for g in torch.unique(pag):
        sel = pos_inds[pag == g]
        norm[sel] = am[sel] / (am[sel].max() + 10e-8) * ious[sel].max()
    return gt_inds, mx, am, norm


def test_o2m_assigner_matches_loop_restatement():
    g = torch.Generator().manual_seed(0)
    for trial in range(5):
        Q, G, w, h = 200, 1 + 3 * trial, 640.0, 480.0
        bbox = torch.rand(Q, 4, generator=g) * torch.tensor([1, 1, 0.5, 0.5]) + 0.01
        scores = torch.rand(Q, 80, generator=g)
        xy = torch.rand(G, 2, generator=g) * 0.5
        gtb = torch.cat([xy, xy + torch.rand(G, 2, generator=g) * 0.4 + 0.05], 1) * torch.tensor([w, h, w, h])
        gtl = torch.randint(0, 80, (G,), generator=g)
        res = O2MAssigner().assign(bbox, scores, gtb, gtl, dict(img_shape=(int(h), int(w), 3)))
        gi, mx, am, norm = _o2m_reference_loop(bbox, scores, gtb, gtl, w, h)
        assert torch.equal(res.gt_inds, gi)
        assert torch.allclose(res.max_overlaps, mx) and torch.allclose(res.assign_metrics, am)
        assert torch.allclose(normalized_alignment_metrics(res), norm, atol=1e-7)
        assert torch.equal(res.labels[gi > 0], gtl[gi[gi > 0] - 1]) and (res.labels[gi == 0] == -1).all()
    empty = O2MAssigner().assign(bbox, scores, gtb[:0], gtl[:0], dict(img_shape=(480, 640, 3)))
    assert (empty.gt_inds == 0).all() and empty.num_gts == 0


def test_gmm_threshold_against_sklearn():
    skm = pytest.importorskip("sklearn.mixture")
    rng = np.random.default_rng(0)
    agree, n = 0, 120
    for t in range(n):
        k = int(rng.integers(2, 150))
        x = np.concatenate([rng.normal(-2, 0.5, k // 2 + 1), rng.normal(1.5, 0.8, k - k // 2)]) if t % 3 else rng.normal(0, 1, k)
        xs = np.sort(x).reshape(-1, 1)                 # float64 in, so both sides run the same precision
        gm = skm.GaussianMixture(2, weights_init=np.array([.5, .5]), means_init=np.array([xs.min(), xs.max()]).reshape(2, 1),
                                 precisions_init=np.array([1., 1.]).reshape(2, 1), covariance_type="diag", reg_covar=1e-5)
        gm.fit(xs)
        a, s = gm.predict(xs), gm.score_samples(xs)
        m = a == 0
        want = xs[m][s[m].argmax()][0] if m.any() else xs[a == 1][s[a == 1].argmax()][0]
        agree += abs(fit_gmm_threshold(x) - want) < 1e-9
    assert agree == n
    assert fit_gmm_threshold([]) == 0.0 and fit_gmm_threshold([0.7]) == pytest.approx(0.7)


def test_transform_bboxes_flip_and_clamp():
    M = torch.tensor([[-1.0, 0.0, 100.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    b = torch.tensor([[10.0, 20.0, 30.0, 40.0, 0.9], [-5.0, 0.0, 200.0, 90.0, 0.5]])
    out = Transform2D.transform_bboxes(b, M, (80, 100))
    assert torch.allclose(out[0], torch.tensor([70.0, 20.0, 90.0, 40.0, 0.9]))
    assert torch.allclose(out[1], torch.tensor([0.0, 0.0, 100.0, 80.0, 0.5]))
    assert Transform2D.transform_bboxes([b[:0]], [M], [(80, 100)])[0].shape == (0, 5)


def test_ssod_mask_layout():
    """dino_detr_ssod.py:723-744: [consistency | denoising | matching]; every dn group sees itself + matching."""
    s1, g1, sp2, g2, Q = 3, 5, 2, 4, 7
    pad1, pad2 = s1 * g1, 2 * sp2 * g2
    m = DinoDetrSSOD._ssod_mask(s1, g1, pad2, g2, Q, "cpu")
    ref = torch.zeros(pad1 + pad2 + Q, pad1 + pad2 + Q, dtype=torch.bool)
    ref[pad1 + pad2:, :pad1 + pad2] = True
    for i in range(g1):
        ref[s1 * i:s1 * (i + 1), s1 * (i + 1):pad1 + pad2] = True
        if i:
            ref[s1 * i:s1 * (i + 1), :s1 * i] = True
    for j in range(g2):
        r = slice(pad1 + sp2 * 2 * j, pad1 + sp2 * 2 * (j + 1))
        if j == g2 - 1:
            ref[r, :pad1 + sp2 * j * 2] = True
        else:
            ref[r, pad1 + sp2 * 2 * (j + 1):pad1 + pad2] = True
            ref[r, :pad1 + sp2 * 2 * j] = True
    assert torch.equal(m, ref)


def test_weighted_loss_only_scales_loss_keys():
    out = weighted_loss({"loss_cls": torch.tensor(1.0), "acc": torch.tensor(1.0), "consis_loss.d0": torch.tensor(2.0)}, 4.0)
    assert float(out["loss_cls"]) == 4.0 and float(out["acc"]) == 1.0 and float(out["consis_loss.d0"]) == 8.0


def _small_ssod():
    cfg = ssod_model_cfg()
    cfg["model"]["bbox_head"]["num_query"] = 100
    cfg["model"]["bbox_head"]["transformer"] = dict(type="DINOTransformer", num_queries=100, num_encoder_layers=1,
                                                    num_decoder_layers=2, dim_feedforward=64)
    return DETECTORS.build(cfg).train()


@pytest.mark.parametrize("curr_step", [0, 70000])
def test_teacher_student_step_on_cpu_oracle_path(curr_step):
    torch.manual_seed(0)
    model = _small_ssod()
    model.curr_step = curr_step
    assert not any(p.requires_grad for p in model.teacher.parameters()) and not model.teacher.training
    data = ssod_batch(1, 2, 192, 256, seed=0)
    with reference_cpu_ops():
        losses = model(**data)
        loss, log_vars = model._parse_losses(losses)
        loss.backward()
    assert np.isfinite(float(loss))
    n_dec = 2
    assert {f"unsup_consis_loss.d{l}" for l in range(n_dec)} <= set(log_vars)
    assert "sup_loss_cls" in log_vars and "unsup_loss_cls" in log_vars and "sup_enc_loss_iou" in log_vars
    if curr_step >= 60000:     # consistency weights are zeroed after warm-up (:469-470)
        assert all(float(log_vars[f"unsup_consis_loss.d{l}"]) == 0.0 for l in range(n_dec))
    else:                      # DN losses are skipped for pseudo labels during warm-up (:548-553)
        assert float(log_vars["unsup_dn_loss_cls"]) == 0.0
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing
    assert all(p.grad is None for p in model.teacher.parameters())


def test_mean_teacher_hook_drives_the_wrapper():
    """MeanTeacher + StepRecord on the wrapper with the CPU EMA injected (mean_teacher.py:26-64, step_record.py)."""
    from semi_detr_b200.teacher import MeanTeacher, StepRecord
    torch.manual_seed(0)
    model = _small_ssod()

    class Runner:
        iter = 0
        log_buffer = type("B", (), {"output": {}})()
    r = Runner()
    r.model = model
    with torch.no_grad():
        for p in model.student.parameters():
            p.add_(0.01)
    hook, rec = MeanTeacher(momentum=0.999, interval=1, warm_up=0), StepRecord(normalize=False)
    with reference_cpu_ops():
        hook.before_run(r)
        for (n, t), (_, s) in zip(model.teacher.named_parameters(), model.student.named_parameters()):
            assert torch.equal(t, s), n
        r.iter = 3
        rec.before_train_iter(r)
        hook.before_train_iter(r)
    assert model.curr_step == 3 and r.log_buffer.output["ema_momentum"] == 0.75


