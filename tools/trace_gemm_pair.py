"""Timeline of the first tile of the cta_group::2 GEMM (SDB_GEMM_2CTA=1): when each of the first 8 k-blocks was issued by
the producer, landed locally (seen by the rounding warps), and became ready for the leader's MMA thread."""
import os
import sys

import torch

os.environ["SDB_GEMM_2CTA"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200 import _lib  # noqa: E402
from semi_detr_b200.layers import gemm as G  # noqa: E402

# the traced GEMM is the debug library's build of the same source (-DSDB_GEMM_TRACE=1): route this tool's launches to it
_lib.lib().sdb_gemm_tf32 = _lib.debug_lib().sdb_gemm_tf32

m, n, k = 44446, 256, 2048
x = torch.randn(m, k, device="cuda")
w = torch.randn(n, k, device="cuda") * 0.05
for rm in (3, 0):
    for _ in range(3):
        G.gemm_tf32(x, 0, w, 0, m, n, k, round_mode=rm)
    torch.cuda.synchronize()
    trace = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
    _lib.debug_lib().sdb_gemm_tf32_set_trace(trace.data_ptr())
    G.gemm_tf32(x, 0, w, 0, m, n, k, round_mode=rm)
    torch.cuda.synchronize()
    _lib.debug_lib().sdb_gemm_tf32_set_trace(None)
    tr = trace.view(148, 64).cpu()
    t0 = int(tr[0, 0])
    us = lambda cta, slot: (int(tr[cta, slot]) - t0) / 1e3 if int(tr[cta, slot]) else float("nan")
    print(f"== pair kernel m={m} n={n} k={k} round_mode={rm} (us since CTA 0 started)")
    print("   setup done: leader %.2f  peer %.2f" % (us(0, 1), us(1, 1)))
    for kb in range(8):
        print(f"   k-block {kb}: issued leader {us(0, 3 + kb):6.2f} peer {us(1, 3 + kb):6.2f} | landed leader "
              f"{us(0, 40 + kb):6.2f} peer {us(1, 40 + kb):6.2f} | ready for MMA {us(0, 17 + kb):6.2f}")
    print("   tiles committed:", " ".join(f"{us(0, 26 + t):.2f}" for t in range(4)))
