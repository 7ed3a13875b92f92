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

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if WORLD > 1:   # python -m torch.distributed.run --nproc-per-node 2 ... tools/profile_step.py: rank 0 reports
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
FIVE = os.environ.get("SDB_PROFILE_WORKLOAD", "sup") == "sup5"     # configs[3]: 5-scale model under bf16 autocast
if FIVE:
    from semi_detr_b200.synthetic import dino_r50_5scale
    model = DETECTORS.build(dino_r50_5scale()).cuda().train()
else:
    model = DETECTORS.build(copy.deepcopy(DINO_R50_4SCALE)).cuda().train()
step = FusedSupervisedTrainStep(model, world_size=WORLD, autocast=torch.bfloat16 if FIVE else None)
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
if WORLD > 1 and int(os.environ["RANK"]) != 0:
    dist.barrier()
    os._exit(0)
ev = prof.key_averages()
rows = [(e.key, e.device_time_total / N / 1e3, e.count / N) for e in ev if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"GPU kernel time per step: {tot:.2f} ms in {sum(r[2] for r in rows):.0f} launches")
for k, ms, c in rows[:45]:
    print(f"{ms:8.3f} ms  {c:7.1f}x  {k[:110]}")

if WORLD > 1:
    # wall time of the profiled steps on the device vs the serialised kernel time: how much of the exchange is exposed
    evs = [e for e in prof.events() if e.device_type.name == "CUDA"]
    t0 = min(e.time_range.start for e in evs)
    t1 = max(e.time_range.end for e in evs)
    nccl = [e for e in evs if "nccl" in e.name.lower()]
    print(f"device span per step {(t1 - t0) / N / 1e3:.2f} ms; nccl kernels per step: "
          + ", ".join(f"{(e.time_range.end - e.time_range.start) / 1e3:.3f}" for e in nccl[:len(nccl) // N]))
    dist.barrier()
    os._exit(0)
