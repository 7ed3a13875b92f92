"""Device kernels of the pseudo-label side path (csrc/ssod.cu) against the oracle (oracle/ssod_oracle.py: torchvision
batched NMS + the reference's mean / std filter; oracle/gmm_oracle.py: float64 EM pinned to sklearn and to the reference's
own ``_fit_gmm`` goldens).  Integer / index outputs (labels, counts, which boxes survive, which cost sample is the
threshold) must be identical; scores and boxes are copies of inputs, hence bit-equal."""
import os

import numpy as np
import pytest
import torch

from oracle import ssod_oracle
from oracle.gmm_oracle import fit_gmm_threshold

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _detections(B, Q, C, seed, spread, peaked):
    """Random teacher outputs: `peaked` puts a few confident classes per query (a trained teacher), otherwise every
    (query, class) score sits near 0.5 (an untrained one: 72 000 candidates per image)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, Q, C, generator=g) * (2.5 if peaked else 0.05) - (4.0 if peaked else 0.0)
    cxcy = torch.rand(B, Q, 2, generator=g)
    wh = torch.rand(B, Q, 2, generator=g) * spread + 0.01
    if peaked:                       # clusters of near-duplicate boxes so that NMS has work to do
        cxcy = (cxcy * 12).floor() / 12 + torch.randn(B, Q, 2, generator=g) * 0.004
        wh = (wh * 4).floor() / 4 * spread + 0.05
    whwh = torch.tensor([1333.0, 800.0, 1333.0, 800.0])
    b = torch.cat([cxcy - wh / 2, cxcy + wh / 2], -1) * whwh
    b = torch.minimum(b.clamp(min=0), whwh)
    return logits.sigmoid(), b


@pytest.mark.parametrize("peaked,filt", [(True, True), (True, False), (False, True), (False, False)])
def test_pseudo_label_nms_matches_torchvision_batched_nms(peaked, filt):
    from semi_detr_b200.ssod import device_ops
    scores, boxes = _detections(3, 900, 80, seed=int(peaked) * 2 + int(filt), spread=0.4, peaked=peaked)
    want = ssod_oracle.pseudo_label_nms(scores, boxes, 0.01, 0.6, 300, filt)
    got = device_ops.pseudo_label_nms(scores.cuda(), boxes.cuda(), 0.01, 0.6, 300, filt)
    assert got[3].cpu().tolist() == want[3].tolist(), "survivor counts"
    assert got[4].cpu().tolist() == want[4].tolist(), "NMS counts"
    for b in range(3):
        n = int(want[3][b])
        gs, ws = got[1][b, :n].cpu(), want[1][b, :n]
        assert torch.equal(gs, ws), "scores in descending order"
        # equal scores may come out in either order (torch.sort is not stable on either side): compare as sets of rows
        rows_g = sorted(map(tuple, torch.cat([got[0][b, :n].cpu(), gs[:, None], got[2][b, :n].cpu()[:, None].float()], 1).tolist()))
        rows_w = sorted(map(tuple, torch.cat([want[0][b, :n], ws[:, None], want[2][b, :n][:, None].float()], 1).tolist()))
        assert rows_g == rows_w
        assert float(got[0][b, n:].abs().sum()) == 0.0 and float(got[1][b, n:].abs().sum()) == 0.0


def test_pseudo_label_nms_edge_cases():
    """No candidate above the threshold; a single detection (std is NaN -> nothing survives the filter, like the
    reference); degenerate boxes; fewer than max_per_img survivors."""
    from semi_detr_b200.ssod import device_ops
    scores = torch.full((4, 20, 5), 0.001)
    boxes = torch.rand(4, 20, 4) * 100
    boxes[..., 2:] += boxes[..., :2]
    scores[1, 3, 2] = 0.9                                     # image 1: exactly one detection
    scores[2, :6, 1] = torch.tensor([0.9, 0.8, 0.7, 0.2, 0.15, 0.1])
    boxes[2, 0] = torch.tensor([10., 10., 10., 50.])          # the best one is degenerate (zero width)
    scores[3, :, 0] = torch.linspace(0.02, 0.4, 20)
    for filt in (True, False):
        want = ssod_oracle.pseudo_label_nms(scores, boxes, 0.01, 0.6, 300, filt)
        got = device_ops.pseudo_label_nms(scores.cuda(), boxes.cuda(), 0.01, 0.6, 300, filt)
        assert got[3].cpu().tolist() == want[3].tolist() and got[4].cpu().tolist() == want[4].tolist()
        for b in range(4):
            n = int(want[3][b])
            assert torch.equal(got[0][b, :n].cpu(), want[0][b, :n]) and torch.equal(got[2][b, :n].cpu(), want[2][b, :n])
    assert device_ops.pseudo_label_nms(scores.cuda(), boxes.cuda(), 0.01, 0.6, 300, True)[3].cpu().tolist()[:2] == [0, 0]


def test_gmm_threshold_matches_the_reference_goldens():
    """The 160 pools of tests/golden/ssod_gmm_golden.npz (thresholds of the REFERENCE's own ``_fit_gmm``): the device EM
    picks the same cost sample; where the reference's pick is a float32 rounding tie between two equally likely samples
    (flagged at generation) either sample of the pool is accepted."""
    from semi_detr_b200.ssod import device_ops
    z = np.load(os.path.join(HERE, "golden", "ssod_gmm_golden.npz"), allow_pickle=False)
    thr, tie = z["thresholds"], z["tie"]
    for i, (want, is_tie) in enumerate(zip(thr, tie)):
        pool = z[f"pool{i}"].astype(np.float32)
        got = device_ops.gmm_threshold(torch.from_numpy(pool).cuda()).cpu()
        assert int(got[1]) == pool.size
        if is_tie:
            assert pool.size == 0 or np.isclose(pool, float(got[0]), rtol=0, atol=1e-6).any(), i
            continue
        assert abs(float(got[0]) - want) <= 1e-6 * max(1.0, abs(want)), (i, pool.size, float(got[0]), want)


def test_gmm_threshold_segments_and_sizes():
    from semi_detr_b200.ssod import device_ops
    g = torch.Generator().manual_seed(3)
    for n in (0, 1, 2, 3, 17, 300, 4096):
        x = torch.cat([torch.randn(n // 2, generator=g) * 0.3 - 2, torch.randn(n - n // 2, generator=g) * 0.5 + 1])
        got = device_ops.gmm_threshold(x.cuda()).cpu()
        assert int(got[1]) == n
        assert abs(float(got[0]) - fit_gmm_threshold(x.numpy())) <= 1e-6 * max(1.0, abs(float(got[0])))
    # the padded all-gather layout of two ranks (37 and 5 costs in segments of 65 floats, count in slot 0 ignored here)
    a, b = torch.randn(37, generator=g) - 2, torch.randn(5, generator=g) + 1.5
    buf = torch.zeros(2 * 65)
    buf[1:38], buf[66:71] = a, b
    got = device_ops.gmm_threshold(buf[1:].cuda(), torch.tensor([37, 5], dtype=torch.int32).cuda(), 65).cpu()
    assert int(got[1]) == 42
    assert abs(float(got[0]) - fit_gmm_threshold(torch.cat([a, b]).numpy())) <= 1e-6


def test_teacher_student_step_reads_back_only_what_decides_shapes():
    """One unsupervised pass at a small size: the pseudo-label path (decode + NMS + filter, matching, GMM threshold,
    double filter) must not call nonzero / boolean indexing -- counted as aten::nonzero dispatches."""
    import copy
    from torch.utils._python_dispatch import TorchDispatchMode
    from semi_detr_b200 import dino, ssod  # noqa: F401
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg

    class Count(TorchDispatchMode):
        def __init__(self):
            super().__init__()
            self.n = {}

        def __torch_dispatch__(self, func, types, args=(), kwargs=None):
            name = func.overloadpacket.__name__
            self.n[name] = self.n.get(name, 0) + 1
            return func(*args, **(kwargs or {}))

    torch.manual_seed(0)
    cfg = ssod_model_cfg()
    cfg["model"]["bbox_head"]["transformer"] = dict(type="DINOTransformer", num_encoder_layers=1, num_decoder_layers=2)
    model = DETECTORS.build(cfg).cuda().train()
    model.curr_step = 70000
    data = ssod_batch(1, 2, 256, 320, seed=3, device="cuda")
    model(**copy.deepcopy(data))                       # warm the per-geometry caches
    with Count() as c:
        losses = model(**data)
    assert torch.isfinite(sum(v for k, v in losses.items() if "loss" in k))
    assert c.n.get("nonzero", 0) == 0, c.n.get("nonzero")
    assert c.n.get("_local_scalar_dense", 0) + c.n.get("item", 0) <= 8, {k: v for k, v in c.n.items() if "scalar" in k or k == "item"}


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_teacher_decode_on_the_device_matches_the_reference_head(ci):
    """The device decode (box arithmetic of `_decode_last_layer` + `sdb_pseudo_label_nms_f32`) against detections
    produced by the reference's OWN `_get_bboxes_single` + mmdet `multiclass_nms` (tests/golden/ssod_decode_golden.npz):
    labels and order identical, scores to 2 ulp (the sigmoid runs on the device), boxes to 1e-4 px."""
    import sys
    sys.path.insert(0, HERE)
    from test_dino_reference_golden import _decode_case
    from semi_detr_b200.ssod import device_ops
    scores, xyxy, max_per_img, want_det, want_lab = _decode_case(ci, "cuda")
    ob, os_, ol, cnt, _ = device_ops.pseudo_label_nms(scores, xyxy, 0.01, 0.6, max_per_img, False)
    n = int(cnt[0])
    assert n == len(want_lab)
    assert np.array_equal(ol[0, :n].cpu().numpy(), want_lab)
    np.testing.assert_allclose(os_[0, :n].cpu().numpy(), want_det[:, 4], rtol=0, atol=2.4e-7)   # sigmoid on the device: <= 2 ulp
    np.testing.assert_allclose(ob[0, :n].cpu().numpy(), want_det[:, :4], rtol=1e-6, atol=1e-4)
