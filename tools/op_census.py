"""Which host code issues the small ops of the train step?  Runs one supervised step on the CPU oracle path (no GPU
needed) under a TorchDispatchMode that attributes every aten op -- forward ops directly, backward ops through the
forward frame that created their autograd node is not attempted: backward ops are listed under "<autograd>" -- to the
innermost frame inside semi_detr_b200/.  Op counts on this path approximate the device launch counts of the same host
code (our own kernels replace some of them on the device).

    python tools/op_census.py [--top 40]
"""
import collections
import copy
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.cpu_path import reference_cpu_ops  # noqa: E402
from semi_detr_b200 import dino  # noqa: E402,F401
from semi_detr_b200.registry import DETECTORS  # noqa: E402
from semi_detr_b200.synthetic import DINO_R50_4SCALE, coco_like_batch  # noqa: E402

SKIP = {"detach", "alias", "view", "_unsafe_view", "expand", "t", "transpose", "permute", "unsqueeze", "squeeze", "slice",
        "select", "as_strided", "reshape", "split", "unbind", "split_with_sizes", "lift_fresh", "_reshape_alias",
        "empty", "empty_like", "empty_strided", "new_empty", "size", "stride", "sym_size", "is_same_size", "unflatten"}


class Census(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.by_site = collections.Counter()
        self.by_op = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = func.overloadpacket.__name__
        if name not in SKIP:
            site = "<autograd>"
            for fr in reversed(traceback.extract_stack(limit=40)):
                if "/semi_detr_b200/" in fr.filename:
                    site = f"{os.path.relpath(fr.filename, ROOT)}:{fr.name}"
                    break
            self.by_site[site] += 1
            self.by_op[(site, name)] += 1
        return func(*args, **(kwargs or {}))


def main():
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    torch.manual_seed(0)
    cfg = copy.deepcopy(DINO_R50_4SCALE)
    model = DETECTORS.build(cfg).train()
    data = coco_like_batch(2, 256, 320, seed=5)
    with reference_cpu_ops():
        model.train_step(data)["loss"].backward()          # warm the per-geometry caches
        with Census() as c:
            model.train_step(data)["loss"].backward()
    total = sum(c.by_site.values())
    print(f"{total} aten ops in one step (views / allocations excluded)")
    for site, n in c.by_site.most_common(top):
        ops = collections.Counter({op: k for (s, op), k in c.by_op.items() if s == site}).most_common(6)
        print(f"{n:6d}  {site:60s} " + ", ".join(f"{op} x{k}" for op, k in ops))


if __name__ == "__main__":
    main()
