"""TMA feed rate per SM on B200 (debug microbenchmark sdb_debug_tma_rate): how fast K-major fp32 tiles arrive in shared
memory for different box heights, boxes per stage and ring depths, with nothing consuming them -- the question the GEMM
traces raise (docs/ROUND2_NOTES.md).  Results: profiles/tma_rate_r2.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200 import _lib  # noqa: E402

lib = _lib.debug_lib()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in (256, 2048):
    rows = 44446 if k == 256 else 22223
    x = torch.randn(rows, k, device="cuda")
    nbytes = x.numel() * 4
    for box_rows, boxes, kpb in ((128, 1, 1), (64, 2, 1), (32, 4, 1), (128, 2, 1), (256, 1, 1), (128, 1, 2), (128, 1, 4),
                                 (64, 1, 8), (32, 1, 8), (16, 1, 8)):
        for stages in (2, 4, 8):
            if boxes * box_rows * 128 * kpb * stages > 200 * 1024:
                continue
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.sdb_debug_tma_rate(None, x.data_ptr(), rows, k, box_rows, boxes, kpb, stages, 148, out.data_ptr())
                e1.record()
                _lib.check(rc, "tma_rate")
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            us = sorted(ts)[len(ts) // 2]
            print(f"k={k:5d} box {box_rows:3d} rows x {kpb} k-blocks, {boxes} per stage, {stages:2d} stages: {us:8.1f} us  "
                  f"{nbytes / us / 1e3:7.1f} GB/s total  {nbytes / us / 1e3 / 148:6.1f} GB/s per SM")
