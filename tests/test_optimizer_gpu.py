"""Fused clip + AdamW (+ EMA) kernel against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW (+ the EMA loop)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 5, 1))
        self.head = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Linear(19, 3))
        self.frozen = torch.nn.Parameter(torch.randn(7), requires_grad=False)


@pytest.mark.parametrize("with_teacher", [False, True])
def test_fused_adamw_matches_torch(with_teacher):
    from semi_detr_b200.engine import FusedAdamW, build_optimizer
    torch.manual_seed(0)
    ref = _Net().cuda()
    ref.backbone.to(memory_format=torch.channels_last)
    mine = copy.deepcopy(ref)
    teacher_ref = copy.deepcopy(ref) if with_teacher else None
    teacher_mine = copy.deepcopy(ref) if with_teacher else None
    opt_ref = build_optimizer(ref, lr=1e-3, weight_decay=1e-2, backbone_lr_mult=0.1, fused=False)
    opt = FusedAdamW(mine, lr=1e-3, weight_decay=1e-2, backbone_lr_mult=0.1,
                     teacher_params=list(teacher_mine.named_parameters()) if with_teacher else None)
    params_ref = [p for g in opt_ref.param_groups for p in g["params"]]
    for it in range(4):
        g = torch.Generator(device="cuda").manual_seed(it)
        opt.zero_grad()
        for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
            if p.requires_grad:
                gr = torch.randn(p.shape, device="cuda", generator=g) * (5.0 if it % 2 else 0.01)
                p.grad = gr.clone()
                q.grad.copy_(gr)
        torch.nn.utils.clip_grad_norm_(params_ref, 0.1)
        opt_ref.step()
        m = min(0.999, 1 - 1 / (it + 2))
        if with_teacher:
            for (_, t), (_, s) in zip(teacher_ref.named_parameters(), ref.named_parameters()):
                if s.requires_grad:
                    t.data.mul_(m).add_(s.data, alpha=1 - m)
        opt.step(0.1, ema_momentum=m if with_teacher else None)
        for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-7), (it, n)
            assert q.stride() == p.stride()
        if with_teacher:
            for (n, t), (_, u) in zip(teacher_ref.named_parameters(), teacher_mine.named_parameters()):
                assert torch.allclose(t, u, rtol=2e-5, atol=2e-7), (it, n)
    assert float(opt.step_count) == 4


def test_gathered_gradients_equal_accumulated_ones():
    """FusedSupervisedTrainStep: autograd.grad + multi-tensor pack into the flat buffer gives the gradients that
    backward() accumulates into the zeroed .grad views (unused parameters -> zeros; shared modules summed)."""
    from semi_detr_b200.engine import FusedSupervisedTrainStep

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = torch.nn.Conv2d(3, 8, 3).to(memory_format=torch.channels_last)
            self.shared = torch.nn.Linear(8, 8)
            self.unused = torch.nn.Linear(4, 4)

        def forward(self, img):
            f = self.backbone(img).mean((2, 3))
            return {"loss_a": self.shared(self.shared(f)).pow(2).mean(), "loss_b": f.abs().mean()}

        @staticmethod
        def _parse_losses(losses):
            return sum(losses.values()), dict(losses)

    torch.manual_seed(0)
    a = Toy().cuda()
    b = copy.deepcopy(a)
    sa = FusedSupervisedTrainStep(a, gather_grads=True, lr=0.0)
    sb = FusedSupervisedTrainStep(b, gather_grads=False, lr=0.0)
    img = torch.randn(2, 3, 16, 16, device="cuda")
    for st in (sa, sb):           # stale gradients must not leak into either mode
        for p in st.opt.params:
            p.grad.fill_(7.0)
    la, _ = sa(dict(img=img))
    lb, _ = sb(dict(img=img))
    assert torch.allclose(la, lb)
    assert torch.allclose(sa.opt.flat_g, sb.opt.flat_g, rtol=1e-6, atol=1e-8)
    assert sa.opt.flat_g.abs().sum() > 0


def test_fused_adamw_follows_param_groups_and_checkpoints():
    """The schedule is read from ``param_groups`` on every step (a scheduler writing ``group['lr']`` takes effect, also
    under CUDA-graph replay through the device-side hyper-parameter buffer), world-size averaging rides in the clip
    coefficient, and state_dict()/load_state_dict() resume the moments and the step counter."""
    from semi_detr_b200.engine import FusedAdamW, build_optimizer
    torch.manual_seed(1)
    ref = _Net().cuda()
    mine = copy.deepcopy(ref)
    opt_ref = build_optimizer(ref, lr=1e-3, weight_decay=1e-2, backbone_lr_mult=0.1, fused=False)
    opt = FusedAdamW(mine, lr=1e-3, weight_decay=1e-2, backbone_lr_mult=0.1)
    params_ref = [p for g in opt_ref.param_groups for p in g["params"]]
    world = 4

    def one_step(it, o, o_ref):
        g = torch.Generator(device="cuda").manual_seed(100 + it)
        for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
            if p.requires_grad:
                gr = torch.randn(p.shape, device="cuda", generator=g) * 0.05
                p.grad = gr.clone()
                q.grad.copy_(gr * world)             # what a summing all-reduce over `world` equal ranks leaves
        torch.nn.utils.clip_grad_norm_(params_ref, 0.1)
        o_ref.step()
        o.step(0.1, grad_scale=1.0 / world)

    for it in range(3):
        if it == 2:                                  # step decay, written the way a scheduler / LrUpdaterHook does
            for grp in opt_ref.param_groups:
                grp["lr"] *= 0.1
            for grp in opt.param_groups:
                grp["lr"] *= 0.1
        one_step(it, opt, opt_ref)
    for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-7), n

    # checkpoint -> fresh optimizer over a copy of the weights -> same continuation
    state = opt.state_dict()
    mine2 = copy.deepcopy(mine)
    opt2 = FusedAdamW(mine2, lr=1e-3, weight_decay=1e-2, backbone_lr_mult=0.1)
    opt2.load_state_dict(state)
    assert float(opt2.step_count) == 3 and opt2.param_groups[0]["lr"] == pytest.approx(1e-4)
    g = torch.Generator(device="cuda").manual_seed(7)
    for (n, q), (_, q2) in zip(mine.named_parameters(), mine2.named_parameters()):
        if q.requires_grad:
            gr = torch.randn(q.shape, device="cuda", generator=g) * 0.05
            q.grad.copy_(gr)
            q2.grad.copy_(gr)
    opt.step(0.1)
    opt2.step(0.1)
    for (n, q), (_, q2) in zip(mine.named_parameters(), mine2.named_parameters()):
        assert torch.equal(q, q2), n

    # a captured step follows set_lr() between replays
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt2.step(0.1)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        opt2.step(0.1)
    opt2.set_lr([0.0, 0.0])
    frozen = mine2.head[0].weight.detach().clone()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(mine2.head[0].weight, frozen)          # lr 0 -> no movement (decay factor 1 - lr*wd = 1)


def test_ssod_step_keeps_a_loaded_teacher_on_resume():
    """MeanTeacher.before_run clones the student only at iteration 0 (mean_teacher.py:26-35)."""
    from semi_detr_b200.engine import FusedSSODTrainStep

    class Wrapper(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.student = _Net()
            self.teacher = _Net()

    torch.manual_seed(2)
    w = Wrapper().cuda()
    teacher_before = [p.detach().clone() for p in w.teacher.parameters()]
    FusedSSODTrainStep(w, start_iter=500)
    for a, b in zip(teacher_before, w.teacher.parameters()):
        assert torch.equal(a, b)
    w0 = Wrapper().cuda()
    FusedSSODTrainStep(w0, start_iter=0)
    for s, t in zip(w0.student.parameters(), w0.teacher.parameters()):
        assert torch.equal(s, t)
