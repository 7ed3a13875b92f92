"""Timing of the semi-supervised teacher-student step (BASELINE.json configs[2] shape) on one GPU, eager:
1 labelled + 4 unlabelled (weak, strong) pairs at 800x1333 per GPU, EMA hook + forward + backward + clip + AdamW.
Prints one JSON line; images/s counts source images (5 per step per GPU)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200 import _lib, dino, ssod  # noqa: E402,F401
from semi_detr_b200.engine import FlatGrads, build_optimizer  # noqa: E402
from semi_detr_b200.registry import DETECTORS  # noqa: E402
from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg  # noqa: E402
from semi_detr_b200.teacher import MeanTeacher  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--curr-step", type=int, default=60000)
ap.add_argument("--engine", default="fused", choices=["fused", "hook"])
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = DETECTORS.build(ssod_model_cfg()).cuda().train()
model.curr_step = args.curr_step
data = ssod_batch(1, 4, 800, 1333, seed=0, device="cuda")
if args.engine == "fused":
    # clip + AdamW + EMA in one pass, gradients gathered with autograd.grad (engine.FusedSSODTrainStep)
    from semi_detr_b200.engine import FusedSSODTrainStep  # noqa: E402
    fused = FusedSSODTrainStep(model, momentum=0.999, warm_up=0, start_iter=args.curr_step)

    def step(i):
        fused.iter = args.curr_step          # stay in the requested phase
        return fused(data)[0]
else:
    # the reference's structure: MeanTeacher hook, backward() into flat .grad views, clip, torch AdamW
    opt = build_optimizer(model)
    grads = FlatGrads([p for g in opt.param_groups for p in g["params"]])
    runner = type("R", (), dict(model=model, iter=0, log_buffer=type("B", (), {"output": {}})()))()
    hook = MeanTeacher(momentum=0.999, interval=1, warm_up=0)
    hook.before_run(runner)

    def step(i):
        runner.iter = i
        hook.before_train_iter(runner)            # fused EMA
        grads.zero()
        losses = model(**data)
        loss, _ = model._parse_losses(losses)
        loss.backward()
        grads.clip_(0.1)
        opt.step()
        return loss


for i in range(args.warmup):
    step(i)
torch.cuda.synchronize()
l0 = dict(_lib.LAUNCHES)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.steps):
    loss = step(args.warmup + i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(json.dumps(dict(workload="configs[2] shape: Semi-DETR teacher-student step, 1 sup + 4 unsup pairs 800x1333, 1 GPU, eager",
                      phase="warm-up (O2M)" if args.curr_step < 60000 else "Hungarian", engine=args.engine, ms_per_step=ms,
                      images_per_s=5 / (ms / 1e3), loss=float(loss),
                      launches_per_step={k: (_lib.LAUNCHES[k] - l0[k]) / args.steps for k in l0})))
