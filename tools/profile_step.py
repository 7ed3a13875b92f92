"""Where does a train step go?  torch.profiler over a few steps: GPU time by kernel, CPU wall, launch counts."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200 import dino  # noqa: E402,F401
from semi_detr_b200.engine import FusedSupervisedTrainStep  # noqa: E402
from semi_detr_b200.registry import DETECTORS  # noqa: E402
from semi_detr_b200.synthetic import DINO_R50_4SCALE, coco_like_batch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = DETECTORS.build(copy.deepcopy(DINO_R50_4SCALE)).cuda().train()
step = FusedSupervisedTrainStep(model)
data = coco_like_batch(2, 800, 1333, seed=0, device="cuda")
for _ in range(5):
    step(data)
torch.cuda.synchronize()
N = 5
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step(data)
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = [(e.key, e.device_time_total / N / 1e3, e.count / N) for e in ev if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"GPU kernel time per step: {tot:.2f} ms in {sum(r[2] for r in rows):.0f} launches")
for k, ms, c in rows[:45]:
    print(f"{ms:8.3f} ms  {c:7.1f}x  {k[:110]}")
