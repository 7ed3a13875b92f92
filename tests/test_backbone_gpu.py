"""ResNet-50 with the fused cuDNN convolution-bias-ReLU calls against the same network as separate convolution,
bias and ReLU kernels: identical features and gradients (fp32 convolutions)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_conv_bias_relu_matches_unfused(monkeypatch):
    from semi_detr_b200.dino.backbone import ResNet
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        net = ResNet(depth=50, out_indices=(0, 1, 2, 3), frozen_stages=1, norm_eval=True)
        with torch.no_grad():                     # non-trivial frozen statistics
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.normal_(0, 0.1)
                    m.running_var.uniform_(0.5, 1.5)
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.normal_(0, 0.1)
        net = net.cuda().train().to(memory_format=torch.channels_last)
        ref = copy.deepcopy(net)
        x = torch.randn(2, 3, 160, 192, device="cuda").contiguous(memory_format=torch.channels_last)
        monkeypatch.setenv("SDB_FUSED_CONV", "1")
        outs = net(x)
        sum(o.square().mean() for o in outs).backward()
        monkeypatch.setenv("SDB_FUSED_CONV", "0")
        outs_ref = ref(x)
        sum(o.square().mean() for o in outs_ref).backward()
        for a, b in zip(outs, outs_ref):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))
        checked = 0
        for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            if q.grad is None:
                assert p.grad is None, n
                continue
            assert torch.allclose(p.grad, q.grad, rtol=1e-3, atol=1e-4 * float(q.grad.abs().max())), n
            checked += 1
        assert checked >= 40
    finally:
        torch.backends.cudnn.allow_tf32 = prev
