"""Which tensor shapes do the small / elementwise kernels of a train step work on?  torch.profiler with
record_shapes: device time grouped by (aten op, input shapes), aten ops only (our own kernels show up under the
op that wraps them)."""
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
for _ in range(4):
    step(data)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
N = 2
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(N):
        step(data)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.self_device_time_total > 0 and e.key.startswith("aten::"):
        rows.append((e.self_device_time_total / N / 1e3, e.count / N, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"aten self device time per step: {tot:.2f} ms")
skip = ("aten::mm", "aten::addmm", "aten::convolution", "aten::cudnn", "aten::bmm", "aten::_scaled_dot")
for ms, c, k, sh in rows[:90]:
    if k.startswith(skip):
        continue
    print(f"{ms:7.3f} ms {c:6.1f}x  {k:34s} {sh}")
