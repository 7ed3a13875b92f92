"""bf16-storage MSDA kernels next to the fp32 ones at the train-step shapes (4-scale and 5-scale, encoder and decoder):
median of 30 L2-flushed launches, algorithmic GB/s against the measured HBM peak.  NOT YET RUN ON HARDWARE."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA  # noqa: E402
from semi_detr_b200.synthetic import COCO_4SCALE_LEVELS, msda_inputs  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def median_us(fn, iters=30):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def algorithmic_bytes(N, S, Lq, L, velt):
    """fwd / bwd bytes with `velt`-byte value / out / grad_out and fp32 everything else (grad_value fp32 either way)."""
    v, o = N * S * 256, N * Lq * 256
    loc, att = N * Lq * 8 * L * 4 * 2, N * Lq * 8 * L * 4
    return velt * (v + o) + 4 * (loc + att), velt * (o + v) + 4 * (v + 2 * loc + 2 * att)


for name, levels in (("4-scale", COCO_4SCALE_LEVELS), ("5-scale", [(200, 334)] + COCO_4SCALE_LEVELS)):
    S = sum(h * w for h, w in levels)
    for shape, mode, Lq in (("enc", "encoder", None), ("dec", "uniform", 1100)):
        x = msda_inputs(levels, N=2, Lq=Lq, mode=mode, seed=0)
        q = x["loc"].shape[1]
        vb, gb = x["value"].to(torch.bfloat16), x["gout"].to(torch.bfloat16)
        for dt, v, g, velt in (("fp32", x["value"], x["gout"], 4), ("bf16", vb, gb, 2)):
            fb, bb = algorithmic_bytes(2, S, q, len(levels), velt)
            tf = median_us(lambda: MSDA.ms_deform_attn_forward(v, x["shapes"], x["start"], x["loc"], x["attn"], 64))
            tb = median_us(lambda: MSDA.ms_deform_attn_backward(v, x["shapes"], x["start"], x["loc"], x["attn"], g, 64))
            print(json.dumps(dict(config=name, shape=shape, S=S, Lq=q, storage=dt, fwd_us=round(tf, 1),
                                  fwd_gbs=round(fb / tf / 1e3, 1), fwd_frac=round(fb / tf / 1e3 / PEAK, 4),
                                  bwd_us=round(tb, 1), bwd_gbs=round(bb / tb / 1e3, 1),
                                  bwd_frac=round(bb / tb / 1e3 / PEAK, 4))), flush=True)
        del x, vb, gb
