import os, sys, time, copy
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from semi_detr_b200 import _lib, dino, ssod  # noqa
from semi_detr_b200.registry import DETECTORS
from semi_detr_b200.synthetic import ssod_batch, ssod_model_cfg
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0)
model = DETECTORS.build(ssod_model_cfg()).cuda().train()
model.curr_step = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
data = ssod_batch(1, 4, 800, 1333, seed=int(os.environ.get("SSOD_SEED", "0")), device="cuda")
for _ in range(2):
    losses = model(**data); loss, _ = model._parse_losses(losses); loss.backward()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t = time.time()
    losses = model(**data); loss, _ = model._parse_losses(losses); loss.backward()
    torch.cuda.synchronize()
    print("wall ms", (time.time() - t) * 1e3)
ev = prof.key_averages()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in ev if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
print("GPU ms", sum(r[1] for r in rows), "launches", sum(r[2] for r in rows))
for k, ms, c in rows[:14]:
    print(f"{ms:9.3f} ms {c:6d}x  {k[:100]}")
cpu = [(e.key, e.self_cpu_time_total / 1e3, e.count) for e in ev]
cpu.sort(key=lambda r: -r[1])
for k, ms, c in cpu[:10]:
    print(f"cpu {ms:9.3f} ms {c:6d}x  {k[:100]}")
