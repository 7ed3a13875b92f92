"""Host logic of the linear-layer dispatch (no GPU): which products go to the tcgen05 kernel under each policy, that
the kernel is never chosen when torch's TF32 switch is off, and that CPU tensors take the plain torch path."""
import pytest
import torch


@pytest.fixture
def tf32_on():
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    yield
    torch.backends.cuda.matmul.allow_tf32 = prev


def test_product_plan_policies(monkeypatch, tf32_on):
    from semi_detr_b200.layers.linear import product_plan
    monkeypatch.setenv("SDB_LINEAR", "auto")
    assert product_plan(256, 256, False) == (False, False, True)      # split-K grad-weight (+ bias gradient) only
    assert product_plan(256, 256, True) == (True, False, True)        # value_proj: the mask rides in the epilogue
    assert product_plan(256, 2048, True, has_mask=False) == (False, False, False)   # FFN linear1: ReLU is a library epilogue too;
    #                                                     the layer keeps its fused ReLU-backward + bias-gradient pass
    assert product_plan(2048, 256, False) is None                     # FFN linear2: library
    monkeypatch.setenv("SDB_LINEAR", "tcgen05")
    assert product_plan(256, 384, False) == (True, True, True)
    assert product_plan(256, 2048, True) is None
    monkeypatch.setenv("SDB_LINEAR", "tcgen05_all")
    assert product_plan(2048, 256, False) == (True, True, True)
    monkeypatch.setenv("SDB_LINEAR", "cublas")
    assert product_plan(256, 256, True) is None
    monkeypatch.setenv("SDB_LINEAR", "bogus")
    with pytest.raises(RuntimeError, match="SDB_LINEAR"):
        product_plan(256, 256, True)


def test_kernel_follows_torch_tf32_switch(monkeypatch):
    from semi_detr_b200.layers.linear import product_plan
    monkeypatch.setenv("SDB_LINEAR", "tcgen05_all")
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        assert product_plan(256, 256, True) is None
        torch.backends.cuda.matmul.allow_tf32 = True
        assert product_plan(256, 256, True) == (True, True, True)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_cpu_tensors_take_the_torch_path(tf32_on):
    """The Linear subclass is an nn.Linear on the CPU (the oracle path of the host tests), epilogue flags included."""
    from semi_detr_b200.layers.linear import Linear
    torch.manual_seed(0)
    lin = Linear(256, 512)
    x = torch.randn(3, 1200, 256)
    mask = torch.rand(3, 1200) < 0.3
    y = lin(x, relu=True, row_mask=mask)
    want = torch.relu(torch.nn.functional.linear(x, lin.weight, lin.bias)).masked_fill(mask[..., None], 0.0)
    assert torch.equal(y, want)
    assert set(lin.state_dict()) == {"weight", "bias"}


def test_loss_dict_drops_total_when_edited():
    from semi_detr_b200.dino.head import LossDict
    d = LossDict()
    d["loss_a"] = torch.tensor(1.0)
    d.total = torch.tensor(1.0)
    d.update(loss_b=torch.tensor(2.0))
    assert d.total is None
    d.total = torch.tensor(3.0)
    d["loss_c"] = torch.tensor(0.5)
    assert d.total is None and list(d) == ["loss_a", "loss_b", "loss_c"]


def test_ssod_engine_momentum_schedule():
    """FusedSSODTrainStep.momentum_at is the MeanTeacher hook's schedule (mean_teacher.py:37-50)."""
    from semi_detr_b200.engine import FusedSSODTrainStep
    eng = FusedSSODTrainStep.__new__(FusedSSODTrainStep)
    eng.momentum, eng.warm_up = 0.999, 0
    assert eng.momentum_at(0) == 0.0 and eng.momentum_at(1) == 0.5 and eng.momentum_at(10 ** 6) == 0.999
    eng.warm_up = 100
    assert abs(eng.momentum_at(0) - (1 - 101 / 101)) < 1e-12 and abs(eng.momentum_at(50) - (1 - 101 / 151)) < 1e-12
