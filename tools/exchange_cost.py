"""What does the data-parallel exchange cost per step?  Two (or more) ranks, the graph-replayed supervised step, timed
with the collectives switched off one group at a time (numbers are then WRONG on purpose -- this is a timing aid):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 \
        tools/exchange_cost.py

  full          everything on (what bench.py times)
  no_grad       the gradient all-reduces skipped (every all_reduce over more than 1e5 elements)
  no_scalar     the two 1-element normaliser all-reduces skipped
  none          no collective at all = the single-GPU step inside a multi-process run
and, separately, the bare all-reduce of the flat gradient buffer (whole, and in the three buckets) replayed from a graph.
"""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200 import dino  # noqa: E402,F401
from semi_detr_b200.engine import FusedSupervisedTrainStep, GraphedTrainStep  # noqa: E402
from semi_detr_b200.registry import DETECTORS  # noqa: E402
from semi_detr_b200.synthetic import DINO_R50_4SCALE, coco_like_batch  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True

real_all_reduce = dist.all_reduce
mode = {"skip_big": False, "skip_small": False}


class _Done:
    def wait(self):
        return True


def patched(t, *a, **k):
    big = t.numel() > 100000
    if (big and mode["skip_big"]) or (not big and mode["skip_small"]):
        return _Done() if k.get("async_op") else None
    return real_all_reduce(t, *a, **k)


dist.all_reduce = patched


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    real_all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    real_all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {}
torch.manual_seed(0)
base = DETECTORS.build(copy.deepcopy(DINO_R50_4SCALE)).to(dev).train()
data = coco_like_batch(2, 800, 1333, seed=rank, device=dev)
for name, big, small, overlap in (("full", False, False, True), ("full_single_bucket", False, False, False),
                                  ("no_grad", True, False, True), ("no_scalar", False, True, True),
                                  ("none", True, True, True)):
    mode["skip_big"], mode["skip_small"] = big, small
    model = copy.deepcopy(base)
    step = FusedSupervisedTrainStep(model, world_size=world, overlap=overlap)
    g = GraphedTrainStep(step, data, warmup=3)
    res[name] = timed(lambda: g())
    del g, step, model
    torch.cuda.empty_cache()

mode["skip_big"] = mode["skip_small"] = False
n = 47_000_000
buf = torch.randn(n, device=dev)
for label, parts in (("allreduce_188MB", [(0, n)]), ("allreduce_3_buckets", [(0, n // 2), (n // 2, n // 2 + 15_000_000), (n // 2 + 15_000_000, n)])):
    for _ in range(3):
        for a, b in parts:
            real_all_reduce(buf[a:b])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for a, b in parts:
            real_all_reduce(buf[a:b])
    res[label] = timed(lambda: graph.replay())
    buf.normal_()
if rank == 0:
    print(json.dumps({k: round(v, 3) for k, v in res.items()}), flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
