"""The fused decoder self-attention kernels (csrc/attention.cu, `sdb_mha_forward/backward_f32`) against the written-out
softmax attention in float64 -- the computation of nn.MultiheadAttention's core as the DINO decoder layer uses it
(detr_od/models/utils/transformer.py:795-812) with the denoising mask of dn_components.py:97-113.

The kernels round the operands of their four products to TF32 (10 mantissa bits, round-to-nearest) and accumulate in
fp32.  Tolerances, relative to the largest magnitude of each compared tensor: 2e-3 for the output, 4e-3 for the
gradients (measured values are printed); the log-sum-exp to 2e-3 absolute."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dn_mask(T, pad, groups, device):
    """The contrastive-denoising attention mask (True = blocked): matching queries cannot see the denoising part, each
    denoising group sees only itself."""
    m = torch.zeros(T, T, dtype=torch.bool, device=device)
    m[pad:, :pad] = True
    size = pad // max(groups, 1)
    for i in range(groups):
        m[size * i:size * (i + 1), :size * i] = True
        m[size * i:size * (i + 1), size * (i + 1):pad] = True
    return m


def reference(qk, v, mask_add, H):
    T, N, C2 = qk.shape
    C = C2 // 2
    d = C // H
    q, k = qk.double().split(C, -1)
    q = (q * d ** -0.5).reshape(T, N * H, d).transpose(0, 1)
    k = k.reshape(T, N * H, d).transpose(0, 1)
    vv = v.double().reshape(T, N * H, d).transpose(0, 1)
    s = q @ k.transpose(1, 2)
    if mask_add is not None:
        s = s + mask_add.double()
    p = s.softmax(-1)
    return (p @ vv).transpose(0, 1).reshape(T, N, C), torch.logsumexp(s, -1)


@pytest.mark.parametrize("T,N,pad,groups", [(1100, 2, 200, 5), (64, 1, 0, 0), (333, 3, 120, 3), (900, 2, 0, 0), (17, 2, 8, 2)])
def test_fused_self_attention_matches_float64_softmax_attention(T, N, pad, groups):
    from semi_detr_b200 import _lib
    from semi_detr_b200.layers.attention import _SelfAttentionFn
    H, C = 8, 256
    g = torch.Generator(device="cuda").manual_seed(T)
    qk = (torch.randn(T, N, 2 * C, device="cuda", generator=g) * 1.5).requires_grad_(True)
    v = torch.randn(T, N, C, device="cuda", generator=g).requires_grad_(True)
    gout = torch.randn(T, N, C, device="cuda", generator=g)
    mask_add = mask_add_t = None
    if pad:
        mask_add = torch.zeros(T, T, device="cuda").masked_fill_(dn_mask(T, pad, groups, "cuda"), float("-inf"))
        mask_add_t = mask_add.t().contiguous()
    before = _lib.LAUNCHES["mha_forward"], _lib.LAUNCHES["mha_backward"]
    out = _SelfAttentionFn.apply(qk, v, mask_add, mask_add_t, H)
    out.backward(gout)
    assert (_lib.LAUNCHES["mha_forward"], _lib.LAUNCHES["mha_backward"]) == (before[0] + 1, before[1] + 2)
    got = out.detach(), qk.grad.clone(), v.grad.clone()
    qk2, v2 = qk.detach().clone().requires_grad_(True), v.detach().clone().requires_grad_(True)
    want, lse = reference(qk2, v2, mask_add, H)
    want.backward(gout.double())
    worst = {}
    for name, a, b, tol in (("out", got[0], want.detach(), 2e-3), ("grad_qk", got[1], qk2.grad, 4e-3),
                            ("grad_v", got[2], v2.grad, 4e-3)):
        assert torch.isfinite(a).all(), name
        err = float((a.double() - b.double()).abs().max() / b.abs().max())
        worst[name] = err
        assert err < tol, (name, err)
    print(f"[attention T={T} N={N}] " + " ".join(f"{k} {e:.1e}" for k, e in worst.items()))


def test_decoder_layer_uses_the_fused_attention_and_matches_the_library_path():
    """DINOTransformerDecoderLayer._self_attention: fused kernels (TF32 switch on) against the written-out library path
    with full-fp32 products, forward and parameter gradients."""
    from semi_detr_b200 import _lib
    from semi_detr_b200.dino.transformer import DINOTransformerDecoderLayer
    torch.manual_seed(0)
    layer = DINOTransformerDecoderLayer(256, 2048, 0.0, "relu", 4, 8, 4).cuda().train()
    T, N = 420, 2
    x = torch.randn(T, N, 256, device="cuda")
    pos = torch.randn(T, N, 256, device="cuda")
    mask = torch.zeros(T, T, device="cuda").masked_fill_(dn_mask(T, 120, 4, "cuda"), float("-inf"))
    prev = torch.backends.cuda.matmul.allow_tf32
    res = {}
    try:
        for mode in (True, False):
            torch.backends.cuda.matmul.allow_tf32 = mode
            layer.zero_grad()
            before = _lib.LAUNCHES["mha_forward"]
            y = layer._self_attention(x + pos, x, mask)
            y.square().sum().backward()
            assert (_lib.LAUNCHES["mha_forward"] > before) == mode
            res[mode] = (y.detach(), layer.self_attn.in_proj_weight.grad.clone(), layer.self_attn.out_proj.weight.grad.clone())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    for a, b in zip(res[True], res[False]):
        assert float((a - b).abs().max() / b.abs().max()) < 5e-3
