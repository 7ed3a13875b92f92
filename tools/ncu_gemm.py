"""Two launches of the tcgen05 GEMM for an `ncu --set full` capture: the projection shape (m=44446, n=256, k=256,
B-resident variant) and the FFN-1 shape (n=2048, streaming variant)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200.layers import gemm as G  # noqa: E402

torch.manual_seed(0)
x = torch.randn(44446, 256, device="cuda")
for n in (256, 2048):
    w = torch.randn(n, 256, device="cuda") * 0.05
    b = torch.randn(n, device="cuda")
    for _ in range(2):
        y = G.linear_forward(x, w, b, relu=True)
torch.cuda.synchronize()
print("done", float(y.sum()))
