"""The teacher-student step on the device against the same host code on the CPU oracle path (identical weights,
identical CDN noise, identical teacher detections): losses within 1e-3 relative in both phases."""
import copy

import pytest
import torch

from oracle.cpu_path import reference_cpu_ops

pytestmark = pytest.mark.gpu


@pytest.fixture
def cpu_noise(monkeypatch):
    from semi_detr_b200.dino import dn_components as dn
    state = {"g": None}

    def reseed():
        state["g"] = torch.Generator().manual_seed(4321)
    monkeypatch.setattr(dn, "_rand", lambda shape, device, generator=None: torch.rand(shape, generator=state["g"]).to(device))
    monkeypatch.setattr(dn, "_randint", lambda lo, hi, shape, device, generator=None:
                        torch.randint(lo, hi, shape, generator=state["g"]).to(device))
    return reseed


def _to(obj, dev):
    if torch.is_tensor(obj):
        return obj.to(dev)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to(o, dev) for o in obj)
    if isinstance(obj, dict):
        return {k: _to(v, dev) for k, v in obj.items()}
    return obj


@pytest.mark.parametrize("curr_step", [0, 70000])
def test_unsup_branch_matches_cpu_reference_path(cpu_noise, curr_step):
    from semi_detr_b200 import dino, ssod  # noqa: F401
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cfg = ssod_model_cfg()
    cfg["model"]["bbox_head"]["transformer"] = dict(type="DINOTransformer", num_encoder_layers=2, num_decoder_layers=2)
    cpu_model = DETECTORS.build(cfg).train()
    with torch.no_grad():     # off the pixel-centre lattice (see test_dino_gpu.py) and spread the teacher's scores
        for name, p in cpu_model.named_parameters():
            if name.endswith("sampling_offsets.bias"):
                p.add_(torch.randn_like(p) * 0.37)
            if name.endswith("fc_cls.0.bias") or name.endswith("fc_cls.0.weight"):
                p.add_(torch.randn_like(p) * (0.8 if p.dim() == 1 else 0.05))
    cpu_model.curr_step = curr_step
    gpu_model = copy.deepcopy(cpu_model).cuda().train()
    data = ssod_batch(1, 2, 320, 384, seed=2)
    metas = data["img_metas"]
    for m in metas:
        m["batch_input_shape"] = (320, 384)
    t_idx = [i for i, m in enumerate(metas) if m["tag"] == "unsup_teacher"]
    s_idx = [i for i, m in enumerate(metas) if m["tag"] == "unsup_student"]
    t_metas, s_metas = [metas[i] for i in t_idx], [metas[i] for i in s_idx]

    # teacher detections once, on the device; the CPU path consumes the very same pseudo boxes
    with torch.no_grad():
        t_gpu = gpu_model.extract_teacher_info(data["img"][t_idx].cuda(), [dict(m) for m in t_metas])
    assert sum(b.shape[0] for b in t_gpu["det_bboxes"]) > 0, "the synthetic teacher should produce pseudo boxes"

    def run(model, dev, teacher_info):
        cpu_noise()
        s_info = model.extract_student_info(data["img"][s_idx].to(dev), [dict(m) for m in s_metas])
        from semi_detr_b200.ssod.bbox_utils import Transform2D
        M = [b @ a.inverse() for b, a in zip(s_info["transform_matrix"], teacher_info["transform_matrix"])]
        pb = Transform2D.transform_bboxes(teacher_info["det_bboxes"], M, [m["img_shape"] for m in s_metas])
        return model.unsup_loss(s_info, teacher_info, pb, teacher_info["det_labels"], teacher_info["det_scores"])

    out = run(gpu_model, "cuda", t_gpu)
    t_cpu = dict(t_gpu)
    t_cpu.update(img=t_gpu["img"].cpu(), det_bboxes=_to(t_gpu["det_bboxes"], "cpu"), det_labels=_to(t_gpu["det_labels"], "cpu"),
                 det_scores=_to(t_gpu["det_scores"], "cpu"), transform_matrix=_to(t_gpu["transform_matrix"], "cpu"))
    with reference_cpu_ops():
        with torch.no_grad():
            t_cpu["backbone_feature"] = cpu_model.teacher.extract_feat(t_cpu["img"])
        ref = run(cpu_model, "cpu", t_cpu)
    assert list(out) == list(ref)
    for k in ref:
        a, b = float(out[k]), float(ref[k])
        assert abs(a - b) <= 2e-3 * abs(b) + 2e-5, (k, a, b)
    total = sum(v for k, v in out.items() if "loss" in k)
    total.backward()
    assert all(p.grad is not None for p in gpu_model.student.parameters() if p.requires_grad)


def test_full_size_ssod_step_runs():
    """BASELINE.json configs[2] shape on one GPU: 1 labelled + 4 unlabelled pairs at 800x1333, Hungarian phase, EMA."""
    from semi_detr_b200 import _lib, dino, ssod  # noqa: F401
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg
    from semi_detr_b200.teacher import MeanTeacher
    torch.manual_seed(0)
    model = DETECTORS.build(ssod_model_cfg()).cuda().train()
    model.curr_step = 60000
    data = ssod_batch(1, 4, 800, 1333, seed=0, device="cuda")
    runner = type("R", (), dict(model=model, iter=0, log_buffer=type("B", (), {"output": {}})()))()
    hook = MeanTeacher(momentum=0.999, interval=1, warm_up=0)
    before = dict(_lib.LAUNCHES)
    hook.before_run(runner)
    losses = model(**data)
    loss, _ = model._parse_losses(losses)
    loss.backward()
    assert torch.isfinite(loss)
    n = {k: _lib.LAUNCHES[k] - before[k] for k in before}
    assert n["ema_update"] == 1 and n["msda_forward"] + n["msda_fused_forward"] >= 5 * 12
    assert n["msda_backward"] + n["msda_fused_backward"] == 2 * 12


def test_fused_ssod_engine_keeps_the_hook_semantics():
    """FusedSSODTrainStep (student clip+AdamW and teacher EMA in one pass, frozen parameters through the EMA kernel)
    against the reference's ordering: teacher <- student at start; after every step
    teacher == m * teacher_prev + (1 - m) * student_new with the hook's momentum schedule, for every parameter
    (trainable, frozen, and nothing for the projector)."""
    from semi_detr_b200 import dino, ssod  # noqa: F401
    from semi_detr_b200.engine import FusedSSODTrainStep
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg
    torch.manual_seed(0)
    cfg = ssod_model_cfg()
    cfg["model"]["bbox_head"]["transformer"] = dict(type="DINOTransformer", num_encoder_layers=1, num_decoder_layers=2)
    model = DETECTORS.build(cfg).cuda().train()
    with torch.no_grad():                      # make the teacher differ from the student before the engine copies
        for p in model.teacher.parameters():
            p.add_(1.0)
    step = FusedSSODTrainStep(model, momentum=0.999, warm_up=0, start_iter=0, lr=1e-3)   # iteration 0: before_run copies
    for (n, s), (_, t) in zip(model.student.named_parameters(), model.teacher.named_parameters()):
        assert torch.equal(s, t), n
    step.iter = 70000                           # continue in the Hungarian phase
    data = ssod_batch(1, 2, 256, 320, seed=3, device="cuda")
    for it in range(2):
        t_prev = {n: p.detach().clone() for n, p in model.teacher.named_parameters()}
        s_prev = {n: p.detach().clone() for n, p in model.student.named_parameters()}
        loss, log_vars = step(data)
        assert torch.isfinite(loss) and model.curr_step == 70000 + it
        m = min(0.999, 1 - 1 / (70000 + it + 2))
        assert log_vars["ema_momentum"] == m
        moved = 0
        for (n, s), (_, t) in zip(model.student.named_parameters(), model.teacher.named_parameters()):
            want = t_prev[n] * m + s.detach() * (1 - m)
            assert torch.allclose(t, want, rtol=1e-6, atol=1e-7), n
            if s.requires_grad:
                moved += int(not torch.equal(s, s_prev[n]))
            else:
                assert torch.equal(s, s_prev[n]), n
        assert moved > 100
    assert all(not p.requires_grad for p in model.teacher.parameters())


def test_graphed_inference_sections_give_the_eager_pseudo_labels(monkeypatch):
    """The teacher pass (backbone + transformer + decode + NMS / filter) and the student's no-grad head pass replay from
    CUDA graphs after two eager calls (engine.GraphedNoGrad).  Same kernels on the same numbers: from the third call on,
    the graphed teacher must hand out the detections the eager teacher produces, also after its weights were changed in
    place between two replays (what the EMA update does every step)."""
    from semi_detr_b200 import dino, ssod  # noqa: F401
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        cfg = ssod_model_cfg()
        cfg["model"]["bbox_head"]["transformer"] = dict(type="DINOTransformer", num_encoder_layers=1, num_decoder_layers=2)
        model = DETECTORS.build(cfg).cuda().train()
        model.curr_step = 70000
        data = ssod_batch(1, 2, 256, 320, seed=3, device="cuda")
        tags = [m["tag"] for m in data["img_metas"]]
        idx = [i for i, t in enumerate(tags) if t == "unsup_teacher"]
        img = data["img"][idx]
        metas = [dict(data["img_metas"][i], batch_input_shape=tuple(img.shape[-2:])) for i in idx]

        def detections():
            info = model.extract_teacher_info(img, [dict(m) for m in metas])
            return [torch.cat([b, s[:, None], l[:, None].float()], 1).clone()
                    for b, s, l in zip(info["det_bboxes"], info["det_scores"], info["det_labels"])]

        def check(tag):
            monkeypatch.setenv("SDB_SSOD_GRAPHS", "1")
            got = detections()
            monkeypatch.setenv("SDB_SSOD_GRAPHS", "0")
            want = detections()
            assert [g.shape for g in got] == [w.shape for w in want], tag
            for g, w in zip(got, want):
                assert torch.equal(g[:, 5], w[:, 5]) and torch.allclose(g[:, :5], w[:, :5], rtol=1e-5, atol=1e-5), tag
        monkeypatch.setenv("SDB_SSOD_GRAPHS", "1")
        detections()
        detections()                                   # two eager warm-up calls
        check("first replay")
        assert any("graph" in st for st in model._teacher_graphs.cache.values()), "the teacher pass was not captured"
        with torch.no_grad():                          # in-place weight change, as the EMA blend does
            for p in model.teacher.bbox_head.fc_cls.parameters():
                p.add_(0.05 * torch.randn_like(p))
        check("after an in-place weight update")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
