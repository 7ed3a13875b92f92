"""Train-step plumbing around the detector.

What mmcv's ``EpochBasedRunner`` + ``OptimizerHook`` + ``MMDistributedDataParallel`` do per iteration for
configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:122-128 (AdamW lr 1e-4, wd 1e-4, backbone lr x0.1, grad_clip
max_norm 0.1; DDP wrap at detr_ssod/apis/train.py:84-93), laid out for one process per B200:

 * all trainable gradients live in ONE flat fp32 buffer (``p.grad`` are views), so the data-parallel exchange is a
   single NCCL all-reduce over NVLink (~188 MB, well under a millisecond on NVSwitch -- <2% of the step, so it is
   not split into overlap buckets), the grad-norm clip is one reduction + one scale over the flat buffer, and
   zeroing is one memset;
 * the whole step (forward, loss with device-side Hungarian matching, backward, all-reduce, clip, AdamW) has no
   host synchronisation and no per-step host->device copies (``consts.device_const``), so it can be captured in a
   CUDA graph and replayed: ``GraphedTrainStep``.  A graph is specific to the batch geometry (image sizes and GT
   counts); training keeps one per geometry bucket, the eager ``SupervisedTrainStep`` handles the rest.

torch's fused AdamW is library plumbing here; fusing clip+AdamW+EMA over the flat buffers is a "next" row
(SURVEY.md section 8f, rank 3).
"""
import torch
import torch.distributed as dist


def build_optimizer(model, lr=1e-4, weight_decay=1e-4, backbone_lr_mult=0.1, fused=None, capturable=False):
    backbone, rest = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (backbone if "backbone" in name else rest).append(p)
    groups = [dict(params=rest, lr=lr, weight_decay=weight_decay),
              dict(params=backbone, lr=lr * backbone_lr_mult, weight_decay=weight_decay)]
    if fused is None:
        fused = all(p.is_cuda for p in rest + backbone)
    return torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, fused=fused,
                             capturable=capturable and fused)


class FlatGrads:
    """One contiguous gradient buffer; every trainable parameter's ``.grad`` is a view into it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, world_size):
        if world_size > 1:
            dist.all_reduce(self.flat)
            self.flat.div_(world_size)

    def clip_(self, max_norm):
        """torch.nn.utils.clip_grad_norm_(norm_type=2) on the flat buffer: no host sync."""
        total_norm = torch.linalg.vector_norm(self.flat, 2)
        coef = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
        self.flat.mul_(coef)
        return total_norm


class SupervisedTrainStep:
    """One data-parallel rank's step: zero grads -> forward/loss -> backward -> all-reduce -> clip -> AdamW."""

    def __init__(self, model, optimizer, max_grad_norm=0.1, world_size=None):
        self.model, self.optimizer, self.max_grad_norm = model, optimizer, max_grad_norm
        if world_size is None:
            world_size = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.world_size = world_size
        self.grads = FlatGrads([p for g in optimizer.param_groups for p in g["params"]])

    def __call__(self, data):
        self.grads.zero()
        losses = self.model(**data)
        loss, log_vars = self.model._parse_losses(losses)
        loss.backward()
        self.grads.all_reduce_mean(self.world_size)
        if self.max_grad_norm is not None:
            self.grads.clip_(self.max_grad_norm)
        self.optimizer.step()
        return loss.detach(), log_vars


class GraphedTrainStep:
    """The same step captured once into a CUDA graph over static input buffers and replayed.

    ``static_data`` is a device-resident batch whose tensors become the graph's inputs; ``load`` refills them from
    another batch of the SAME geometry (image size, GT counts) -- e.g. from pinned host memory each step."""

    def __init__(self, step, static_data, warmup=3):
        self.step, self.data = step, static_data
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):              # warms cuDNN autotune, kernel attributes, device_const caches
                step(static_data)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.log_vars = step(static_data)

    def load(self, batch):
        d = self.data
        d["img"].copy_(batch["img"], non_blocking=True)
        for dst, src in zip(d["gt_bboxes"], batch["gt_bboxes"]):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(d["gt_labels"], batch["gt_labels"]):
            dst.copy_(src, non_blocking=True)

    def __call__(self, batch=None):
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        return self.loss, self.log_vars
