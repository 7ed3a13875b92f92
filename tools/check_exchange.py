"""Two-rank check of the exchange fused with the optimizer (csrc/exchange.cu, engine.FusedAdamW.step_exchange):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_exchange.py

Both ranks hold the same small two-group model (a 'backbone' and a head group, sizes that do not divide evenly into
shards) with different gradients.  Reference: NCCL all-reduce of the flat gradient, then the single-GPU fused
clip + AdamW kernel (`FusedAdamW.step`, itself held to torch.optim.AdamW by tests/test_optimizer_gpu.py).  Compared
after each of 3 steps: parameters, both moments (on every rank: the moments of foreign shards stay untouched by design,
so they are compared shard-wise), the small all-reduce, and that every rank ends with bit-identical parameters.
Also replays the three launches from a CUDA graph.  Prints one JSON line on rank 0; exit code 1 on mismatch."""
import json
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semi_detr_b200.engine import FusedAdamW  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = nn.ParameterList([nn.Parameter(torch.randn(1031, 77)), nn.Parameter(torch.randn(5))])
        self.head = nn.ParameterList([nn.Parameter(torch.randn(300, 257)), nn.Parameter(torch.randn(123457))])


def make():
    torch.manual_seed(7)
    return Net().to(dev)


ref_model, our_model = make(), make()
ref, ours = FusedAdamW(ref_model, lr=1e-3, weight_decay=1e-2), FusedAdamW(our_model, lr=1e-3, weight_decay=1e-2)
res = {"multicast": bool(ours.enable_peer_exchange())}
ok = res["multicast"]
if ok:
    total = ours.flat_p.numel()
    per = ((total // 4 + world - 1) // world) * 4
    lo, hi = min(rank * per, total), min(rank * per + per, total)
    worst = dict(param=0.0, exp_avg=0.0, exp_avg_sq=0.0)
    for it in range(3):
        g = torch.Generator(device=dev).manual_seed(100 * it + rank)
        grad = torch.randn(total, device=dev, generator=g) * (0.05 if it else 5.0)      # first step clips, later ones do not
        ref.flat_g.copy_(grad)
        ours.flat_g.copy_(grad)
        dist.all_reduce(ref.flat_g)
        ref.step(0.1 if it < 2 else None, grad_scale=1.0 / world)
        ours.step_exchange(0.1 if it < 2 else None)
        torch.cuda.synchronize()
        for name, a, b in (("param", ours.flat_p, ref.flat_p), ("exp_avg", ours.flat_m[lo:hi], ref.flat_m[lo:hi]),
                           ("exp_avg_sq", ours.flat_v[lo:hi], ref.flat_v[lo:hi])):
            worst[name] = max(worst[name], float((a - b).abs().max() / b.abs().max()))
    res.update(worst)
    ok = ok and worst["param"] < 1e-6 and worst["exp_avg"] < 1e-5 and worst["exp_avg_sq"] < 1e-5
    # module parameters are views of the symmetric buffer
    res["views"] = bool(our_model.head[1].data_ptr() >= ours.flat_p.data_ptr())
    # every rank holds the same bits
    chk = ours.flat_p.view(torch.int32).sum(dtype=torch.int64).reshape(1)
    alls = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(alls, chk)
    res["identical"] = bool(all(int(a) == int(alls[0]) for a in alls))
    # small all-reduce, a few rounds through both parities
    small = []
    for k in range(5):
        t = torch.tensor([rank + 1.0 + k], device=dev)
        ours.small_allreduce(t, slot=k % 2)
        small.append(float(t))
    want = [sum(r + 1.0 + k for r in range(world)) for k in range(5)]
    res["small"] = small
    ok = ok and small == want and res["identical"]
    # graph replay of the fused exchange
    grad = torch.randn(total, device=dev) * 0.01
    ours.flat_g.copy_(grad)
    ours.step_exchange(0.1)                      # warm-up outside capture
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ours.step_exchange(0.1)
    before = ours.flat_p.clone()
    for _ in range(3):
        ours.flat_g.copy_(grad)
        graph.replay()
    torch.cuda.synchronize()
    res["graph_moved_params"] = bool((ours.flat_p - before).abs().max() > 0)
    res["error_word"] = bool(ours.peer_error())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(20):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    res["us_per_exchange_small_model"] = round(e0.elapsed_time(e1) * 1e3 / 20, 1)
    ok = ok and res["graph_moved_params"] and not res["error_word"]
res["ok"] = bool(ok)
if rank == 0:
    print(json.dumps(res), flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if ok else 1)
