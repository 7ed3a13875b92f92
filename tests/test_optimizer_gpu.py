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
