"""Timeline of one sdb_gemm_tf32 launch from the %globaltimer stamps its warp roles leave (debug hook
sdb_gemm_tf32_set_trace): when did the producer issue, the MMA warp get its first stage, the epilogue its accumulator."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200 import _lib  # noqa: E402
from semi_detr_b200.layers import gemm as G  # noqa: E402

# the traced GEMM is the debug library's build of the same source (-DSDB_GEMM_TRACE=1): route this tool's launches to it
_lib.lib().sdb_gemm_tf32 = _lib.debug_lib().sdb_gemm_tf32

NAMES = {0: "start", 1: "setup done", 2: "producer: B issued", 16: "mma: B ready", 48: "xform: B landed",
         49: "B rounded (smem variant) / weights in TMEM (wres)", 62: "epilogue: stores drained", 63: "end"}
for t in range(6):
    NAMES[3 + 2 * t] = f"producer: tile {t} first A issue"
    NAMES[4 + 2 * t] = f"producer: tile {t} last A issue"
    NAMES[17 + 2 * t] = f"mma: tile {t} first stage ready"
    NAMES[18 + 2 * t] = f"mma: tile {t} committed"
    NAMES[32 + 2 * t] = f"epilogue: tile {t} accumulator ready"
    NAMES[33 + 2 * t] = f"epilogue: tile {t} done"


def run(m, n, k, round_mode):
    x = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") * 0.05
    b = torch.randn(n, device="cuda")
    for _ in range(3):
        G.gemm_tf32(x, 0, w, 0, m, n, k, bias=b, round_mode=round_mode)
    torch.cuda.synchronize()
    trace = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
    _lib.debug_lib().sdb_gemm_tf32_set_trace(trace.data_ptr())
    G.gemm_tf32(x, 0, w, 0, m, n, k, bias=b, round_mode=round_mode)
    torch.cuda.synchronize()
    _lib.debug_lib().sdb_gemm_tf32_set_trace(None)
    tr = trace.view(148, 64).cpu()
    t0 = int(tr[:, 0][tr[:, 0] > 0].min())
    print(f"== m={m} n={n} k={k} round_mode={round_mode}: CTA 0 and CTA 147 (us since the first CTA started)")
    for cta in (0, 147):
        row = tr[cta]
        ev = sorted((int(row[s]) - t0, NAMES[s]) for s in NAMES if int(row[s]) > 0)
        print(f"-- CTA {cta}")
        for ts, name in ev:
            print(f"   {ts / 1e3:8.2f}  {name}")
    ends = tr[:, 63][tr[:, 63] > 0]
    print(f"   all CTAs end between {(int(ends.min()) - t0) / 1e3:.2f} and {(int(ends.max()) - t0) / 1e3:.2f} us")


if __name__ == "__main__":
    run(44446, 256, 256, 3)
    run(44446, 256, 256, 0)
    run(44446, 2048, 256, 3)
