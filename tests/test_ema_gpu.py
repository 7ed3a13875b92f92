"""Fused EMA kernel against goldens from the real MeanTeacher hook and the oracle loop."""
import numpy as np
import pytest
import torch

from oracle.ema_oracle import ema_momentum, ema_update

pytestmark = pytest.mark.gpu


class _LogBuf:
    def __init__(self):
        self.output = {}


class _Runner:
    def __init__(self, model):
        self.model = model
        self.iter = 0
        self.log_buffer = _LogBuf()


class _Net(torch.nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(t.clone()) for t in tensors])
        self.register_buffer("buf", torch.arange(9.0))


class _Model(torch.nn.Module):
    def __init__(self, teacher, student):
        super().__init__()
        self.teacher, self.student = _Net(teacher), _Net(student)


def test_matches_reference_hook(ema_golden):
    from semi_detr_b200.teacher import MeanTeacher
    g = ema_golden
    n = len(g["student"])
    model = _Model([torch.from_numpy(g["teacher0"][str(i)]) for i in range(n)],
                   [torch.from_numpy(g["student"][str(i)]) for i in range(n)]).cuda()
    hook = MeanTeacher(momentum=0.999, interval=1, warm_up=0)
    runner = _Runner(model)
    for it, m in zip(g["iters"], g["momenta"]):
        runner.iter = int(it)
        hook.before_train_iter(runner)
        assert runner.log_buffer.output["ema_momentum"] == m
        for i in range(n):
            got = model.teacher.ps[i].detach().cpu().numpy()
            want = g[f"teacher_after_{int(it)}"][str(i)]
            # same two roundings as mul_ + add_(alpha): allow 1 ulp for hosts whose add_ is not fused
            np.testing.assert_allclose(got, want, rtol=2e-7, atol=1e-9)
    assert torch.equal(model.teacher.buf.cpu(), torch.arange(9.0))     # buffers untouched


def test_before_run_copies_student():
    from semi_detr_b200.teacher import MeanTeacher
    torch.manual_seed(0)
    shapes = [(3, 5), (1,), (40000,), (257, 129)]
    model = _Model([torch.randn(*s) for s in shapes], [torch.randn(*s) for s in shapes]).cuda()
    hook = MeanTeacher(momentum=0.999, interval=1, warm_up=0)
    hook.before_run(_Runner(model))
    for t, s in zip(model.teacher.ps, model.student.ps):
        assert torch.equal(t, s)


def test_large_random_vs_oracle_and_unaligned():
    from semi_detr_b200.teacher import EmaPlan
    torch.manual_seed(1)
    flat_t = torch.randn(3_000_001, device="cuda")
    flat_s = torch.randn(3_000_001, device="cuda")
    # views at odd offsets -> unaligned chunks take the scalar path
    cuts = [0, 1, 1025, 70001, 2_000_000, 3_000_001]
    tp = [flat_t[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    sp = [flat_s[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    want = [t.clone().cpu() for t in tp]
    plan = EmaPlan(tp, sp)
    for it in (1, 5, 2000):
        m = ema_momentum(it)
        plan.step(m)
        ema_update(want, [s.cpu() for s in sp], m)
        for a, b in zip(tp, want):
            np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=2e-7, atol=1e-9)


def test_interval_and_schedule():
    from semi_detr_b200.teacher import MeanTeacher
    model = _Model([torch.zeros(4)], [torch.ones(4)]).cuda()
    hook = MeanTeacher(momentum=0.9, interval=2, warm_up=0)
    r = _Runner(model)
    r.iter = 1
    hook.before_train_iter(r)                      # skipped: 1 % 2 != 0
    assert model.teacher.ps[0].abs().sum() == 0
    r.iter = 2
    hook.before_train_iter(r)                      # m = min(0.9, 1 - 1/3)
    assert torch.allclose(model.teacher.ps[0], torch.full((4,), 1 - (1 - 1 / 3), device="cuda"))
