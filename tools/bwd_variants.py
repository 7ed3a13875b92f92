"""Fused MSDA backward and forward at the train-step encoder / decoder shapes for the register-budget variants
(sdb_msda_set_variant): median of 30 L2-flushed launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler  # noqa: E402  (nvidia-smi clocks / throttle reasons sampled while the timings run)
from semi_detr_b200 import _lib  # noqa: E402
from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA  # noqa: E402

_clocks = ClockSampler(0)
levels = [(100, 167), (50, 84), (25, 42), (13, 21)]
S = sum(h * w for h, w in levels)
shapes = torch.tensor(levels, dtype=torch.int64, device="cuda")
start = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, Lq, refdim in (("enc", S, 2), ("dec", 1092, 4)):
    N = 2
    value = torch.randn(N, S, 8, 32, device="cuda", generator=g)
    if refdim == 2:   # pixel centres, like the encoder
        ref = torch.cat([torch.stack(torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device="cuda") / h,
                                                    torch.linspace(0.5, w - 0.5, w, device="cuda") / w, indexing="ij"), -1)
                         .flip(-1).reshape(-1, 2) for h, w in levels])[None, :, None, :].expand(N, S, 4, 2).contiguous()
    else:
        ref = torch.rand(N, Lq, 4, 4, device="cuda", generator=g) * 0.5 + 0.25
    off = torch.randn(N, Lq, 8, 4, 4, 2, device="cuda", generator=g) * 2.0
    logits = torch.randn(N, Lq, 8, 16, device="cuda", generator=g)
    gout = torch.randn(N, Lq, 256, device="cuda", generator=g)
    # 7 / 8 = experimental 4-lane x 8-channel mapping (msda_backward_x8.cu), unrolled / rolled batch loop; 10 / 11 / 12 = 2 / 2 / 4 points of corner
    # loads in flight per warp at 4 / 3 / 3 CTAs per SM
    # 15 / 16 = 8 x 8 pixel tiles with 256 threads, 1 / 2 points in flight
    # 0 / 20 = default: tile-combining kernel for the encoder shape (msda_backward_tile.cu), 8-lane kernel otherwise
    variants = (0, 5, 6, 3, 2, 7, 8, 10, 11, 12, 15, 16) if "--all" in sys.argv else \
        ((0, 20, 24, 25) if "--tile" in sys.argv else (0, 5, 7, 12))
    for variant in variants:
        _lib.lib().sdb_msda_set_variant(0, variant)
        ts = []
        for _ in range(33):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            MSDA.ms_deform_attn_fused_backward(value, shapes, start, ref, off, logits, gout)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[3:])
        print(f"{name} Lq={Lq} backward variant {variant}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f}")
    for variant in (0, 3, 5, 6, 8):
        if name == "dec" and variant in (5, 8):
            continue
        _lib.lib().sdb_msda_set_variant(variant, 0)
        ts = []
        for _ in range(33):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            MSDA.ms_deform_attn_fused_forward(value, shapes, start, ref, off, logits)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[3:])
        print(f"{name} Lq={Lq} forward variant {variant}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f}")
_lib.lib().sdb_msda_set_variant(0, 0)
import json  # noqa: E402
print("clocks:", json.dumps(_clocks.stop()))
