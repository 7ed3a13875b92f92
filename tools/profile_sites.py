"""Which host code issues the small device ops of the train step?  torch.profiler with stacks on ONE eager step: every
aten op that launched kernels is attributed to the innermost frame inside semi_detr_b200/ (forward ops) or, for ops
issued by the autograd engine, to "<autograd> op [shapes]".  Prints launches and device time per site.

    python tools/profile_sites.py [--top 60]
"""
import collections
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step(data)
    torch.cuda.synchronize()
LIB = ("aten::mm", "aten::addmm", "aten::convolution", "aten::cudnn", "aten::bmm", "aten::_addmm_activation",
       "aten::convolution_backward")
sites = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
for e in prof.events():
    if not e.kernels or not e.name.startswith("aten::") or e.name.startswith(LIB):
        continue
    site = None
    for fr in e.stack or []:
        if "semi_detr_b200/" in fr and "engine.py" not in fr:
            site = fr.split("semi_detr_b200/")[-1].strip()
            break
    if site is None:
        site = f"<autograd> {e.name} {str(e.input_shapes)[:70]}"
    s = sites[site]
    s[0] += len(e.kernels)
    s[1] += sum(k.duration for k in e.kernels)
    s[2][e.name.replace("aten::", "")] += 1
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 60
tot_l = sum(s[0] for s in sites.values())
tot_t = sum(s[1] for s in sites.values())
print(f"non-library aten ops of one eager step: {tot_l} launches, {tot_t / 1e3:.3f} ms of device time")
for site, s in sorted(sites.items(), key=lambda kv: -kv[1][1])[:top]:
    ops = ", ".join(f"{k} x{v}" for k, v in s[2].most_common(4))
    print(f"{s[1] / 1e3:7.3f} ms {s[0]:5d}  {site[:95]:95s} {ops}")
