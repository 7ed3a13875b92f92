"""tcgen05.mma kind::tf32 issue rate on B200 (debug microbenchmark sdb_debug_umma_rate)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200 import _lib  # noqa: E402

out = torch.zeros(1, dtype=torch.int64, device="cuda")
lib = _lib.debug_lib()
for grid in (1,):
    for n in (128, 256):
        for mode in (0, 1):
            if n == 256 and mode & 7:
                continue
            for iters in (2048,):
                rc = lib.sdb_debug_umma_rate(None, n, mode, iters, grid, out.data_ptr())
                _lib.check(rc, "umma_rate")
                torch.cuda.synchronize()
                cyc = int(out)
                print(f"grid={grid:3d} N={n} A={'tmem' if mode & 1 else 'smem'} ring={'8x16KB' if mode & 4 else 'none  '} commit/4={'y' if mode & 8 else 'n'} iters={iters:5d}: "
                      f"{cyc / iters:7.1f} cycles per 128x{n}x8 MMA  ({2 * 128 * n * 8 * iters / cyc:7.0f} flop/clk/SM)")

# cta_group::2: the leader of a CTA pair issues 256 x N x 8 (each SM: 128 rows x N columns, half of B from the peer)
for grid in (2, 148):
    for n in (128, 256):
        rc = lib.sdb_debug_umma_rate(None, n, 16, 2048, grid, out.data_ptr())
        _lib.check(rc, "umma_rate")
        torch.cuda.synchronize()
        cyc = int(out)
        print(f"grid={grid:3d} cta_group::2 N={n}: {cyc / 2048:7.1f} cycles per 256x{n}x8 MMA  "
              f"({2 * 128 * n * 8 * 2048 / cyc:7.0f} flop/clk/SM)")
