"""Can the teacher-student step be captured into a CUDA graph?  Builds the configs[2] per-GPU batch, runs the Hungarian
phase eagerly and as a replayed graph, prints both step times and the loss of each (same inputs, same teacher)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200 import dino, ssod  # noqa: E402,F401
from semi_detr_b200.engine import FusedSSODTrainStep, GraphedTrainStep  # noqa: E402
from semi_detr_b200.registry import DETECTORS  # noqa: E402
from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = DETECTORS.build(ssod_model_cfg()).to(dev).train()
fused = FusedSSODTrainStep(model, momentum=0.999, warm_up=0, world_size=1)
host = ssod_batch(1, 4, 800, 1333, seed=0)
data = dict(img=host["img"].to(dev), img_metas=[dict(m) for m in host["img_metas"]],
            gt_bboxes=[x.to(dev) for x in host["gt_bboxes"]], gt_labels=[x.to(dev) for x in host["gt_labels"]])
IT0 = int(os.environ.get("SSOD_ITER", "60000"))


def step(batch):
    fused.iter = IT0
    return fused(batch)


def timed(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(out[0])


res = {}
for _ in range(5):
    step(data)
res["eager_ms"], res["eager_loss"] = timed(lambda: step(data))
try:
    g = GraphedTrainStep(step, data, warmup=2)
    res["graph_ms"], res["graph_loss"] = timed(lambda: g())
except Exception as e:
    res["graph_error"] = f"{type(e).__name__}: {str(e)[:600]}"
print(json.dumps(res))
