"""Seeded inputs of the fused-prologue MSDA tests (shared by test_msda_fused_gpu.py-style tests)."""
import torch


def fused_inputs(levels, N, Lq, ref_dim, seed):
    from semi_detr_b200.synthetic import level_tensors
    g = torch.Generator().manual_seed(seed)
    L, M, D, P = len(levels), 8, 32, 4
    S = sum(h * w for h, w in levels)
    shapes, start = level_tensors(levels, "cuda")
    value = torch.randn(N, S, M, D, generator=g).cuda()
    if ref_dim == 2:
        ref = torch.rand(N, Lq, L, 2, generator=g).cuda()
        off = (torch.randn(N, Lq, M, L, P, 2, generator=g) * 3).cuda()
    else:
        ref = torch.cat([torch.rand(N, Lq, L, 2, generator=g), torch.rand(N, Lq, L, 2, generator=g) * 0.4 + 0.02], -1).cuda()
        off = (torch.randn(N, Lq, M, L, P, 2, generator=g) * 2).cuda()
    logits = (torch.randn(N, Lq, M, L * P, generator=g) * 2).cuda()
    gout = torch.randn(N, Lq, M * D, generator=g).cuda()
    return value, shapes, start, ref, off, logits, gout
