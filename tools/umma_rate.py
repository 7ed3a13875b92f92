"""tcgen05.mma kind::tf32 issue rate on B200 (debug microbenchmark sdb_debug_umma_rate)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200 import _lib  # noqa: E402

out = torch.zeros(1, dtype=torch.int64, device="cuda")
lib = _lib.lib()
for grid in (1, 148):
    for n in (128, 256):
        for mode in (0, 1, 4, 5, 8, 9, 12, 13):
            if n == 256 and mode & 7:
                continue
            for iters in (2048,):
                rc = lib.sdb_debug_umma_rate(None, n, mode, iters, grid, out.data_ptr())
                _lib.check(rc, "umma_rate")
                torch.cuda.synchronize()
                cyc = int(out)
                print(f"grid={grid:3d} N={n} A={'tmem' if mode & 1 else 'smem'} ring={'8x16KB' if mode & 4 else 'none  '} commit/4={'y' if mode & 8 else 'n'} iters={iters:5d}: "
                      f"{cyc / iters:7.1f} cycles per 128x{n}x8 MMA  ({2 * 128 * n * 8 * iters / cyc:7.0f} flop/clk/SM)")
