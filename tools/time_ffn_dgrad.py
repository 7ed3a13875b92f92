"""FFN linear2 grad-input + ReLU backward + linear1 bias gradient (44 446 x 256 -> 2048 and the decoder's 2 184 rows):
this library's tcgen05 GEMM with the fused epilogue (sdb_gemm_tf32_relu_grad) against the layer-by-layer route (library
GEMM, then sdb_relu_backward_colsum_f32) and against the bare products.  L2 flushed between launches, median of 12."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200.layers import gemm as G  # noqa: E402
from semi_detr_b200.layers.linear import relu_backward_colsum  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
res = {}
for rows in (44446, 2184):
    g = [torch.randn(rows, 256, device=dev) for _ in range(4)]
    w = torch.randn(256, 2048, device=dev) * 0.05
    h = torch.relu(torch.randn(rows, 2048, device=dev))
    flush = torch.empty(64 * 1024 * 1024, device=dev)

    def timed(fn, n=12):
        for i in range(3):
            fn(g[i % 4])
        ts = []
        for i in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(g[i % 4])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return round(ts[len(ts) // 2], 1)
    res[f"rows{rows}"] = dict(
        ours_fused_round3=timed(lambda t: G.linear_grad_input_relu(t, w, h, round_mode=3)),
        ours_fused_round2=timed(lambda t: G.linear_grad_input_relu(t, w, h, round_mode=2)),
        ours_plain_round3=timed(lambda t: G.gemm_tf32(t, 0, w, 1, rows, 2048, 256, round_mode=3)),
        ours_plain_round2=timed(lambda t: G.gemm_tf32(t, 0, w, 1, rows, 2048, 256, round_mode=2)),
        library_mm=timed(lambda t: t @ w),
        library_mm_then_relu_backward_colsum=timed(lambda t: relu_backward_colsum(t @ w, h)))
print(json.dumps(res))
