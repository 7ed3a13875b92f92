"""Launches of the FFN grad-input product for an `ncu --set full` capture at the encoder shape (44 446 x 256 -> 2048):
two with the ReLU-backward + column-sum epilogue (sdb_gemm_tf32_relu_grad, round_mode 2), two plain (sdb_gemm_tf32)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200.layers import gemm as G  # noqa: E402

torch.manual_seed(0)
g = torch.randn(44446, 256, device="cuda")
w = torch.randn(256, 2048, device="cuda") * 0.05
h = torch.relu(torch.randn(44446, 2048, device="cuda"))
for _ in range(2):
    y, s = G.linear_grad_input_relu(g, w, h, round_mode=2)
for _ in range(2):
    z = G.gemm_tf32(g, 0, w, 1, 44446, 2048, 256, round_mode=2)
torch.cuda.synchronize()
print("done", float(y.sum()), float(z.sum()))
