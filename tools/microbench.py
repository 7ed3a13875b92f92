"""Kernel-level timing sweep on one GPU (CUDA events, L2 flushed between runs).

    python tools/microbench.py [--n 2] [--iters 30] [--variants] [--ref]

Prints one JSON line per measurement: MSDA fwd/bwd per kernel variant (ours) and for the reference's own CUDA
op compiled for sm_100a (oracle/_ref), Hungarian solves, EMA.  Numbers here steer tuning; bench.py is the
contract."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from semi_detr_b200 import _lib  # noqa: E402
from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA  # noqa: E402
from semi_detr_b200.synthetic import COCO_4SCALE_LEVELS, MICROBENCH_LEVELS, msda_bytes, msda_inputs  # noqa: E402

PEAK = 6556.5
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--variants", action="store_true")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    lib = _lib.lib()
    cases = [("micro_enc", MICROBENCH_LEVELS, "encoder", None), ("micro_uniform", MICROBENCH_LEVELS, "uniform", 17821),
             ("coco_enc", COCO_4SCALE_LEVELS, "encoder", None), ("coco_dec", COCO_4SCALE_LEVELS, "uniform", 1100)]
    fvars = [1, 2, 3, 4, 5, 6, 7, 8, 9] if args.variants else [0]
    bvars = [1, 2, 3, 4, 5, 6, 9] if args.variants else [0]
    for name, levels, mode, Lq in cases:
        if args.only and args.only not in name:
            continue
        x = msda_inputs(levels, N=args.n, mode=mode, Lq=Lq, seed=0)
        S = x["value"].shape[1]
        q = x["loc"].shape[1]
        fb, bb = msda_bytes(args.n, S, q)
        a = (x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
        MSDA.USE_TMA = True
        if q == S:
            med, best = timeit(lambda: MSDA.ms_deform_attn_forward(*a, 64), args.iters)
            emit(op="msda_fwd_tma", case=name, n=args.n, us=round(med, 2), best_us=round(best, 2),
                 gbs=round(fb / med / 1e3, 1), frac=round(fb / med / 1e3 / PEAK, 4))
        MSDA.USE_TMA = False
        for v in fvars:
            lib.sdb_msda_set_variant(v, 0)
            med, best = timeit(lambda: MSDA.ms_deform_attn_forward(*a, 64), args.iters)
            emit(op="msda_fwd", case=name, n=args.n, variant=v, us=round(med, 2), best_us=round(best, 2),
                 gbs=round(fb / med / 1e3, 1), frac=round(fb / med / 1e3 / PEAK, 4))
        for v in bvars:
            lib.sdb_msda_set_variant(0, v)
            med, best = timeit(lambda: MSDA.ms_deform_attn_backward(*a, x["gout"], 64), args.iters)
            emit(op="msda_bwd", case=name, n=args.n, variant=v, us=round(med, 2), best_us=round(best, 2),
                 gbs=round(bb / med / 1e3, 1), frac=round(bb / med / 1e3 / PEAK, 4))
        lib.sdb_msda_set_variant(0, 0)
        if args.ref:
            import ref_cuda
            if ref_cuda.available():
                med, best = timeit(lambda: ref_cuda.forward(*a), args.iters)
                emit(op="ref_msda_fwd", case=name, n=args.n, us=round(med, 2), best_us=round(best, 2),
                     gbs=round(fb / med / 1e3, 1), frac=round(fb / med / 1e3 / PEAK, 4))
                med, best = timeit(lambda: ref_cuda.backward(*a, x["gout"]), args.iters)
                emit(op="ref_msda_bwd", case=name, n=args.n, us=round(med, 2), best_us=round(best, 2),
                     gbs=round(bb / med / 1e3, 1), frac=round(bb / med / 1e3 / PEAK, 4))
        del x, a
    if args.only and "hung" not in args.only and "ema" not in args.only:
        return
    # Hungarian: 7 layers x 2 images, G in {7, 30, 100}
    from semi_detr_b200.matching import HungarianAssigner, MatchTargets
    assigner = HungarianAssigner(cls_cost=dict(type="FocalLossCost", weight=2.0),
                                 reg_cost=dict(type="BBoxL1Cost", weight=5.0, box_format="xywh"),
                                 iou_cost=dict(type="IoUCost", iou_mode="giou", weight=2.0))
    g = torch.Generator().manual_seed(0)
    for G in (7, 30, 100):
        gtb, gtl = [], []
        for _ in range(2):
            xy = torch.rand(G, 2, generator=g) * 0.6
            wh = torch.rand(G, 2, generator=g) * 0.35 + 0.03
            gtb.append(torch.cat([xy, xy + wh], 1) * torch.tensor([1333., 800., 1333., 800.]))
            gtl.append(torch.randint(0, 80, (G,), generator=g))
        bbox = (torch.rand(14, 900, 4, generator=g) * torch.tensor([1, 1, 0.5, 0.5]) + 0.01).cuda()
        cls = (torch.randn(14, 900, 80, generator=g) * 2 - 3).cuda()
        t = MatchTargets(gtb, gtl, [(1333, 800)] * 2, "cuda")
        med, best = timeit(lambda: assigner.assign_batch(bbox, cls, t), args.iters)
        emit(op="hungarian_14x900", G=G, us=round(med, 2), best_us=round(best, 2), solves_per_s=round(14 / med * 1e6))
    # EMA: 47M parameters in 430 tensors
    from semi_detr_b200.teacher import EmaPlan
    sizes = [47_000_000 // 430] * 430
    tp = [torch.randn(s, device="cuda") for s in sizes]
    sp = [torch.randn(s, device="cuda") for s in sizes]
    plan = EmaPlan(tp, sp)
    med, best = timeit(lambda: plan.step(0.999), args.iters)
    nb = 12 * sum(sizes)
    emit(op="ema_430tensors", params=sum(sizes), us=round(med, 2), best_us=round(best, 2), gbs=round(nb / med / 1e3, 1),
         frac=round(nb / med / 1e3 / PEAK, 4))

    def loop():
        for t_, s_ in zip(tp, sp):
            t_.mul_(0.999).add_(s_, alpha=0.001)
    med, best = timeit(loop, 5)
    emit(op="ema_reference_python_loop_gpu", us=round(med, 2), best_us=round(best, 2))


if __name__ == "__main__":
    main()
