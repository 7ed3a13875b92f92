import copy, os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200 import dino  # noqa
from semi_detr_b200.engine import GraphedTrainStep, SupervisedTrainStep, build_optimizer
from semi_detr_b200.registry import DETECTORS
from semi_detr_b200.synthetic import DINO_R50_4SCALE, coco_like_batch
torch.manual_seed(0)
model = DETECTORS.build(copy.deepcopy(DINO_R50_4SCALE)).cuda().train()
step = SupervisedTrainStep(model, build_optimizer(model, capturable=True))
data = coco_like_batch(2, 512, 640, seed=0, device="cuda")
for _ in range(3):
    step(data)
try:
    g = GraphedTrainStep(step, data, warmup=2)
    for _ in range(3):
        l, _ = g()
    torch.cuda.synchronize()
    print("graph ok", float(l))
except Exception:
    traceback.print_exc()
