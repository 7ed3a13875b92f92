"""Train-step plumbing around the detector.

What mmcv's ``EpochBasedRunner`` + ``OptimizerHook`` + ``MMDistributedDataParallel`` do per iteration for
configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:122-128 (AdamW lr 1e-4, wd 1e-4, backbone lr x0.1, grad_clip
max_norm 0.1; DDP wrap at detr_ssod/apis/train.py:84-93), laid out for one process per B200:

 * all trainable gradients live in ONE flat fp32 buffer (``p.grad`` are views), so the data-parallel exchange is a
   single NCCL all-reduce (sum) over NVLink (~188 MB; the 1/world of the mean is folded into the clip coefficient),
   the grad-norm clip is one reduction + one scale over the flat buffer, and
   zeroing is one memset;
 * the whole step (forward, loss with device-side Hungarian matching, backward, all-reduce, clip, AdamW) has no
   host synchronisation and no per-step host->device copies (``consts.device_const``), so it can be captured in a
   CUDA graph and replayed: ``GraphedTrainStep``.  A graph is specific to the batch geometry (image sizes and GT
   counts); training keeps one per geometry bucket, the eager ``SupervisedTrainStep`` handles the rest.

torch's fused AdamW is library plumbing here; fusing clip+AdamW+EMA over the flat buffers is a "next" row
(SURVEY.md section 8f, rank 3).
"""
import os
import sys

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


def train_step_options(cfg):
    """What the train-step engines need from a loaded config (``semi_detr_b200.config.Config``): the optimizer /
    grad-clip hook settings (dino_detr_r50_8x2_12e_coco.py:122-128) and, if present, the MeanTeacher hook's
    (detr_ssod_dino_detr_r50_coco_120k.py:41-45).  ``FusedSupervisedTrainStep(model, **opts)`` /
    ``FusedSSODTrainStep(model, **opts)`` take the result."""
    opt = cfg.optimizer
    if opt["type"] != "AdamW":
        raise ValueError(f"optimizer type {opt['type']!r}: the fused step implements AdamW")
    keys = (opt.get("paramwise_cfg") or {}).get("custom_keys") or {}
    extra = set(keys) - {"backbone"}
    if extra or any(v.get("decay_mult", 1.0) != 1.0 for v in keys.values()):
        raise ValueError(f"paramwise_cfg beyond a backbone lr_mult is not supported: {dict(keys)}")
    out = dict(lr=opt["lr"], weight_decay=opt.get("weight_decay", 0.0),
               backbone_lr_mult=keys.get("backbone", {}).get("lr_mult", 1.0))
    clip = (cfg.get("optimizer_config") or {}).get("grad_clip")
    if clip is not None and clip.get("norm_type", 2) != 2:
        raise ValueError("only 2-norm gradient clipping is implemented")
    out["max_grad_norm"] = clip["max_norm"] if clip else None
    for hook in cfg.get("custom_hooks") or []:
        if hook.get("type") == "MeanTeacher":
            if hook.get("interval", 1) != 1:
                raise ValueError("MeanTeacher interval != 1 is not supported by the fused step")
            out.update(momentum=hook.get("momentum", 0.999), warm_up=hook.get("warm_up", 100))
    return out


def _cache_student_bn_folds(model):
    """The engines own the parameter updates of the model they train: frozen BatchNorm tensors are never written per
    step, so their folded affine maps are computed once (dino/backbone.py ``folded_conv``).  For the teacher-student
    wrapper only the student opts in -- the EMA teacher is rewritten through raw pointers every step."""
    from .dino.backbone import enable_frozen_bn_fold_cache
    enable_frozen_bn_fold_cache(model.student if hasattr(model, "student") and hasattr(model, "teacher") else model)


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
            p.grad = self.flat[off:off + n].as_strided(p.shape, p.stride())   # same memory format as the parameter
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
        _cache_student_bn_folds(model)

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


class GraphedNoGrad:
    """Replay a no-grad, static-shape section of an otherwise eager step from a CUDA graph.

    The teacher-student step cannot be captured as a whole (the number of pseudo boxes decides tensor shapes), but its
    two inference passes can: the teacher's backbone + transformer + decode + NMS / filter on the weak views, and the
    student's no-grad head pass on the strong views are ~3 000 launches with shapes fixed by the batch geometry.  After
    ``warmup`` eager calls per key (cuDNN autotuning, per-geometry constant caches) the section is captured once over
    copies of its tensor inputs; later calls copy the inputs in, replay, and hand out the same output tensors -- valid
    until the next replay, i.e. for the rest of the step.  Weights are read in place (the EMA teacher is updated between
    replays on the same storage).  Any failure to capture switches the key back to eager execution for good.
    ``SDB_SSOD_GRAPHS=0`` disables it."""

    def __init__(self, fn, warmup=2):
        self.fn, self.warmup, self.cache = fn, warmup, {}

    def __call__(self, key, tensors, *args):
        on = (os.environ.get("SDB_SSOD_GRAPHS", "1") != "0" and all(t.is_cuda for t in tensors)
              and not torch.is_grad_enabled() and not torch.cuda.is_current_stream_capturing())
        if not on:
            return self.fn(*tensors, *args)
        key = (key, tuple((tuple(t.shape), t.dtype, t.stride()) for t in tensors))
        st = self.cache.setdefault(key, {"calls": 0})
        if st.get("eager") or st["calls"] < self.warmup:
            st["calls"] += 1
            return self.fn(*tensors, *args)
        if "graph" not in st:
            try:
                static_in = [t.detach().clone() for t in tensors]
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self.fn(*static_in, *args)
                st.update(graph=graph, inputs=static_in, out=out)      # capturing does not execute: replayed below
            except Exception as e:   # a synchronising op inside the section: stay eager, loudly
                print(f"GraphedNoGrad: capture failed ({type(e).__name__}: {str(e)[:160]}); section stays eager",
                      file=sys.stderr)
                torch.cuda.synchronize()
                st["eager"] = True
                return self.fn(*tensors, *args)
        for dst, src in zip(st["inputs"], tensors):
            dst.copy_(src)
        st["graph"].replay()
        return st["out"]


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
        # The captured kernels read the per-geometry constants by address; the caches that own them are bounded (LRU),
        # so the graph keeps its own references -- an eviction can then never free memory a replay still reads.
        from . import consts
        from .dino import backbone, transformer
        self._keepalive = (list(consts._CACHE.values()), list(transformer._GEOMETRY.values()),
                           list(backbone._FOLD_CACHE.values()))

    def load(self, batch):
        d = self.data
        d["img"].copy_(batch["img"], non_blocking=True)
        for dst, src in zip(d["gt_bboxes"], batch["gt_bboxes"]):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(d["gt_labels"], batch["gt_labels"]):
            dst.copy_(src, non_blocking=True)

    def prefetch(self, batch):
        """Start the host->device copy of the NEXT batch (pinned host memory) on a side stream into staging buffers,
        so it overlaps the replay of the current step; ``__call__(prefetched=True)`` then moves it into the graph's
        input buffers with device-to-device copies (a few microseconds)."""
        if getattr(self, "_stage", None) is None:
            d = self.data
            self._stage = dict(img=torch.empty_like(d["img"]), gt_bboxes=[torch.empty_like(t) for t in d["gt_bboxes"]],
                               gt_labels=[torch.empty_like(t) for t in d["gt_labels"]])
            self._copy_stream = torch.cuda.Stream()
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        st = self._stage
        self._copy_stream.wait_event(self._consumed)      # the previous staged batch has been moved out
        with torch.cuda.stream(self._copy_stream):
            st["img"].copy_(batch["img"], non_blocking=True)
            for dst, src in zip(st["gt_bboxes"], batch["gt_bboxes"]):
                dst.copy_(src, non_blocking=True)
            for dst, src in zip(st["gt_labels"], batch["gt_labels"]):
                dst.copy_(src, non_blocking=True)
            self._staged.record()

    def __call__(self, batch=None, prefetched=False):
        if prefetched:
            torch.cuda.current_stream().wait_event(self._staged)
            self.load(self._stage)
            self._consumed.record()
        elif batch is not None:
            self.load(batch)
        self.graph.replay()
        return self.loss, self.log_vars


class FusedAdamW:
    """clip-by-global-norm + AdamW (+ optional mean-teacher EMA of the same parameters) as ONE kernel over flat
    buffers (``sdb_adamw_ema_step_f32``).  Parameters are re-pointed into a flat fp32 buffer (layout preserved with
    ``as_strided``), their ``.grad`` into a second one; the two moments are flat as well.  Param groups follow
    ``build_optimizer`` (head/transformer at ``lr``, backbone at ``lr * backbone_lr_mult``).

    Semantics match ``torch.nn.utils.clip_grad_norm_(max_norm)`` followed by ``torch.optim.AdamW.step()``; with
    ``teacher_params`` the teacher is blended right after the update, which is what the reference's MeanTeacher hook
    does at the start of the next iteration (mean_teacher.py:37-64)."""

    def __init__(self, model, lr=1e-4, weight_decay=1e-4, backbone_lr_mult=0.1, betas=(0.9, 0.999), eps=1e-8,
                 teacher_params=None):
        from . import _lib
        self._lib = _lib
        groups = [[], []]
        names = [[], []]
        for name, p in model.named_parameters():
            if p.requires_grad:
                gi = 1 if "backbone" in name else 0
                groups[gi].append(p)
                names[gi].append(name)
        self.params = groups[0] + groups[1]
        self.param_names = names[0] + names[1]
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW needs CUDA parameters (no CPU path)")
        sizes = [sum(p.numel() for p in g) for g in groups]
        pad = [(-s) % 4 for s in sizes]
        bounds, off = [], 0
        for s, pd in zip(sizes, pad):
            bounds.append((off, off + s + pd))
            off += s + pd
        total = off
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_t = torch.zeros(total, dtype=torch.float32, device=dev) if teacher_params is not None else None
        # parameters without a teacher twin (e.g. the SSOD projector) blend into a scratch slot nobody reads
        teacher_by_name = dict(teacher_params) if teacher_params is not None else {}
        for gi, g in enumerate(groups):
            o = bounds[gi][0]
            for name, p in zip(names[gi], g):
                n = p.numel()
                view = self.flat_p[o:o + n].as_strided(p.shape, p.stride())
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[o:o + n].as_strided(p.shape, p.stride())
                if self.flat_t is not None and name in teacher_by_name:
                    t = teacher_by_name[name]
                    if t.shape != p.shape or t.stride() != p.stride():
                        raise RuntimeError(f"FusedAdamW: teacher parameter {name} does not share the student's layout")
                    tv = self.flat_t[o:o + n].as_strided(t.shape, t.stride())
                    tv.copy_(t.data)
                    t.data = tv
                o += n
        import ctypes
        self._bounds = (ctypes.c_int64 * 4)(bounds[0][0], bounds[0][1], bounds[1][0], bounds[1][1])
        self.betas, self.eps = betas, eps
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        # the schedule lives in ``param_groups`` like a torch optimizer's (an lr scheduler / mmcv LrUpdaterHook writes
        # ``group["lr"]``); the kernel reads (lr, weight_decay) per group from a small DEVICE buffer that ``step``
        # refreshes whenever the host values changed, so a captured graph follows the schedule too (``sync_hparams``)
        self.param_groups = [dict(params=groups[0], lr=lr, initial_lr=lr, weight_decay=weight_decay),
                             dict(params=groups[1], lr=lr * backbone_lr_mult, initial_lr=lr * backbone_lr_mult,
                                  weight_decay=weight_decay)]
        self._hparams_dev = torch.zeros(4, dtype=torch.float32, device=dev)
        self._hparams_host = None
        self.sync_hparams()

    def _host_hparams(self):
        return tuple(float(g[k]) for g in self.param_groups for k in ("lr", "weight_decay"))

    def sync_hparams(self):
        """Copy the current ``param_groups`` lr / weight_decay to the device buffer the kernel reads.  ``step`` calls
        it; with a captured graph call it between replays after changing ``param_groups`` (outside capture)."""
        h = self._host_hparams()
        if h != self._hparams_host:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdamW: param_groups changed during CUDA graph capture; call sync_hparams() "
                                   "before capturing / between replays")
            self._hparams_dev.copy_(torch.tensor(h, dtype=torch.float32), non_blocking=False)
            self._hparams_host = h

    def set_lr(self, lrs):
        """lrs: one learning rate per param group (head/transformer, backbone)"""
        for g, lr in zip(self.param_groups, lrs):
            g["lr"] = float(lr)
        self.sync_hparams()

    def state_dict(self):
        """Moments, step counter and the schedule state -- what torch.optim.AdamW.state_dict() carries, flat."""
        return dict(exp_avg=self.flat_m.clone(), exp_avg_sq=self.flat_v.clone(), step=self.step_count.clone(),
                    param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                    param_names=list(self.param_names), betas=tuple(self.betas), eps=self.eps)

    def load_state_dict(self, state):
        if list(state["param_names"]) != list(self.param_names):
            raise ValueError("FusedAdamW.load_state_dict: parameter list differs from the checkpoint's")
        if state["exp_avg"].numel() != self.flat_m.numel():
            raise ValueError("FusedAdamW.load_state_dict: flat buffer size differs from the checkpoint's")
        self.flat_m.copy_(state["exp_avg"])
        self.flat_v.copy_(state["exp_avg_sq"])
        self.step_count.copy_(state["step"])
        for g, sg in zip(self.param_groups, state["param_groups"]):
            g.update(sg)
        self.betas, self.eps = tuple(state["betas"]), state["eps"]
        self.sync_hparams()

    def zero_grad(self):
        self.flat_g.zero_()

    def all_reduce_sum(self, world_size):
        """One NCCL all-reduce (sum) of the flat gradient; the 1/world of the mean rides in the clip coefficient of
        ``step`` (no separate pass over the 188 MB buffer)."""
        if world_size > 1:
            dist.all_reduce(self.flat_g)

    # ---- gradient exchange fused with the update, over NVLink peer memory (csrc/exchange.cu) ---------------------------
    def enable_peer_exchange(self, group=None):
        """Move the flat parameter / gradient buffers into SYMMETRIC memory (torch's symmetric-memory allocator: the same
        allocation on every rank of ``group``, mapped peer-to-peer and bound to an NVLS multicast object) so that
        ``step_exchange`` can sum the gradients and broadcast the parameters through the NVSwitch.  Returns False --
        leaving everything as it was -- when the node has no multicast support; raises when the ranks disagree."""
        import ctypes

        import torch.distributed._symmetric_memory as symm
        if self.flat_t is not None:
            raise RuntimeError("FusedAdamW.enable_peer_exchange: not combined with the EMA teacher blend")
        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = self.flat_p.device
        lib = self._lib.lib()
        total = self.flat_p.numel()
        new_p = symm.empty(total, dtype=torch.float32, device=dev)
        new_g = symm.empty(total, dtype=torch.float32, device=dev)
        ctrl = symm.empty(lib.sdb_dp_ctrl_bytes() // 4, dtype=torch.int32, device=dev)
        ctrl.zero_()
        handles = [symm.rendezvous(t, group.group_name) for t in (new_p, new_g, ctrl)]
        ok = torch.tensor([1 if all(h.multicast_ptr != 0 for h in handles[:2]) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            return False
        new_p.copy_(self.flat_p)
        new_g.copy_(self.flat_g)
        off = 0
        for gi, g in enumerate(self.param_groups):
            off = int(self._bounds[2 * gi])
            for p in g["params"]:
                n = p.numel()
                p.data = new_p[off:off + n].as_strided(p.shape, p.stride())
                p.grad = new_g[off:off + n].as_strided(p.shape, p.stride())
                off += n
        self.flat_p, self.flat_g = new_p, new_g

        def local_offset(t, h):
            return t.data_ptr() - int(h.buffer_ptrs[h.rank])
        hp, hg, hc = handles
        self._peer = dict(
            rank=rank, world=world, group=group, handles=handles, ctrl=ctrl,
            p_mc=int(hp.multicast_ptr) + local_offset(new_p, hp), g_mc=int(hg.multicast_ptr) + local_offset(new_g, hg),
            ctrl_ptrs=(ctypes.c_void_p * world)(*[int(ptr) + local_offset(ctrl, hc) for ptr in hc.buffer_ptrs]))
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        return True

    @property
    def peer_exchange(self):
        return getattr(self, "_peer", None) is not None

    def step_exchange(self, max_grad_norm=None):
        """Gradient mean over the ranks + clip + AdamW + parameter broadcast as one fused exchange
        (``sdb_dp_adamw_exchange_f32``): no NCCL call, no separate norm pass, optimizer state touched once per node."""
        x = self._peer
        self.sync_hparams()
        dev = self.flat_p.device
        with torch.cuda.device(dev):
            rc = self._lib.lib().sdb_dp_adamw_exchange_f32(
                self._lib.current_stream(dev), x["rank"], x["world"], x["ctrl_ptrs"], self.flat_g.data_ptr(), x["g_mc"],
                self.flat_p.data_ptr(), x["p_mc"], self.flat_m.data_ptr(), self.flat_v.data_ptr(),
                self.step_count.data_ptr(), self._bounds, self._hparams_dev.data_ptr(), 2, self.betas[0], self.betas[1],
                self.eps, float(max_grad_norm) if max_grad_norm is not None else 0.0, 1.0 / x["world"],
                self.flat_p.numel())
        self._lib.check(rc, "dp_adamw_exchange")
        self._lib.LAUNCHES["dp_adamw_exchange"] += 3
        self.step_count.add_(1.0)

    def peer_error(self):
        """True if a bounded wait of the exchange kernels ran out (a peer never arrived); reads the control block."""
        x = self._peer
        word = self._lib.lib().sdb_dp_error_word_offset() // 4
        return bool(int(x["ctrl"][word]) != 0)

    def small_allreduce(self, values, slot=0):
        """In-place sum over the ranks of a 1- or 2-element fp32 device tensor through the peer control blocks."""
        x = self._peer
        dev = values.device
        with torch.cuda.device(dev):
            rc = self._lib.lib().sdb_dp_small_allreduce_f32(self._lib.current_stream(dev), x["rank"], x["world"],
                                                            x["ctrl_ptrs"], slot, values.data_ptr(), values.numel())
        self._lib.check(rc, "dp_small_allreduce")
        self._lib.LAUNCHES["dp_small_allreduce"] += 1
        return values

    def step(self, max_grad_norm=None, ema_momentum=None, grad_scale=1.0):
        """``grad_scale``: factor applied to the gradient buffer before clipping (1/world after a summing all-reduce)."""
        self.sync_hparams()
        coef = None
        if max_grad_norm is not None:
            total_norm = torch.linalg.vector_norm(self.flat_g, 2)
            if grad_scale != 1.0:
                total_norm = total_norm * grad_scale
            coef = torch.clamp(max_grad_norm / (total_norm + 1e-6), max=1.0).reshape(1)
            if grad_scale != 1.0:
                coef = coef * grad_scale
        elif grad_scale != 1.0:
            coef = torch.full((1,), grad_scale, dtype=torch.float32, device=self.flat_p.device)
        dev = self.flat_p.device
        teacher = self.flat_t if (self.flat_t is not None and ema_momentum is not None) else None
        with torch.cuda.device(dev):
            rc = self._lib.lib().sdb_adamw_ema_step_sched_f32(
                self._lib.current_stream(dev), self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                self.flat_v.data_ptr(), self._lib.ptr(teacher), self._lib.ptr(coef), self.step_count.data_ptr(),
                self._bounds, self._hparams_dev.data_ptr(), 2, self.betas[0], self.betas[1], self.eps,
                float(ema_momentum) if ema_momentum is not None else 0.0)
        self._lib.check(rc, "adamw_ema_step")
        self._lib.LAUNCHES["adamw_ema_step"] += 1
        self.step_count.add_(1.0)


class FusedSupervisedTrainStep:
    """``SupervisedTrainStep`` with the fused optimizer kernel (same step, fewer passes over the parameters).

    ``gather_grads`` (default): gradients are taken with ``torch.autograd.grad`` and packed into the flat buffer by
    multi-tensor copies, instead of ``backward()`` accumulating into pre-zeroed ``.grad`` views -- that costs one
    read-modify-write kernel per parameter (~430 launches and a 188 MB memset per step) for sums that have a single
    term.  Same numbers either way (``tests/test_optimizer_gpu.py``)."""

    def __init__(self, model, max_grad_norm=0.1, world_size=None, gather_grads=True, overlap=True, autocast=None,
                 exchange=None, **opt_kw):
        """``autocast``: None (fp32 storage, TF32 products) or ``torch.bfloat16`` -- forward and loss run under
        ``torch.autocast``: library bf16 GEMMs / convolutions, bf16-storage MSDA kernels with fp32 sampling arithmetic,
        fp32 normalisations, matching, losses, master weights and optimizer (BASELINE.json configs[3])."""
        self.model, self.max_grad_norm, self.overlap, self.autocast = model, max_grad_norm, overlap, autocast
        if world_size is None:
            world_size = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.world_size = world_size
        self.opt = FusedAdamW(model, **opt_kw)
        self.gather_grads = gather_grads
        # how the ranks exchange gradients: "peer" = fused with the optimizer over NVLink multicast (csrc/exchange.cu,
        # the default when the node supports it), "nccl" = bucketed all-reduce overlapped with the backward
        exchange = exchange or os.environ.get("SDB_EXCHANGE", "auto")
        self.exchange = "nccl"
        if (world_size > 1 and gather_grads and exchange in ("auto", "peer") and dist.is_available()
                and dist.is_initialized() and dist.get_world_size() == world_size):
            try:
                if self.opt.enable_peer_exchange():
                    self.exchange = "peer"
            except Exception as e:   # no symmetric memory on this box: keep NCCL, loudly if peer was demanded
                if exchange == "peer":
                    raise
                print(f"FusedSupervisedTrainStep: peer exchange unavailable ({type(e).__name__}: {e}); using NCCL",
                      file=sys.stderr)
        if exchange == "peer" and self.exchange != "peer" and world_size > 1:
            raise RuntimeError("FusedSupervisedTrainStep: exchange='peer' needs NVLink multicast (NVLS) on this node")
        _cache_student_bn_folds(model)

    def check(self):
        """Off-the-hot-path health check (synchronises; call it at the logging interval): raises what the reference's
        host-side matcher would have raised inside the step -- scipy's ValueError on NaN / infeasible cost matrices
        (hungarian_assigner.py:136; an out-of-range label also shows up as an invalid entry) -- and reports a peer
        that never arrived at the fused gradient exchange."""
        assigner = getattr(getattr(self.model, "bbox_head", None), "assigner", None)
        if assigner is not None and hasattr(assigner, "check_status"):
            assigner.check_status()
        if self.opt.peer_exchange and self.opt.peer_error():
            raise RuntimeError("fused gradient exchange: a rank did not arrive within the wait bound")

    def _pack(self, grads, params=None):
        """grads (one per parameter, None for unused ones) -> the flat gradient buffer"""
        views = [p.grad for p in (self.opt.params if params is None else params)]
        fast_dst, fast_src = [], []
        for v, g in zip(views, grads):
            if g is None:
                v.zero_()
            elif g.stride() == v.stride() and g.dtype == v.dtype:
                fast_dst.append(v)
                fast_src.append(g)
            else:
                v.copy_(g)
        if fast_dst:
            torch._foreach_copy_(fast_dst, fast_src)

    def _overlapped_backward(self, data):
        """Backward in three segments so that the gradient exchange overlaps it (world_size > 1, DINODETR).

        The flat gradient buffer is laid out [head + transformer | backbone layer2, layer3, layer4] (``FusedAdamW``
        keeps ``named_parameters`` order inside each group).  Segment 1 differentiates the loss down to the backbone's
        output features and packs the head / transformer gradients, whose all-reduce (half of the 188 MB) is issued
        asynchronously on NCCL's stream; segment 2 is layer4's backward (two thirds of the backbone's parameters), its
        bucket is exchanged while segment 3 -- layer3 and layer2, most of the backbone's backward time -- computes.
        Only the last, smallest bucket (layer2 + layer3, 34 MB) is left exposed.  Same sums as one ``autograd.grad``
        over all parameters."""
        model, opt = self.model, self.opt
        n_head = len(opt.param_groups[0]["params"])
        head_params, bb_params = opt.params[:n_head], opt.params[n_head:]
        batch_input_shape = tuple(data["img"].shape[-2:])
        for m in data["img_metas"]:
            m["batch_input_shape"] = batch_input_shape
        stage_cut = []
        feats = model.extract_feat(data["img"], cut_before_last_stage=stage_cut)
        # the head sees detached copies of the backbone features: segment 1 then stops exactly at the cut (the feature
        # levels are nested outputs of one convolution chain, so a cut on the live tensors would run -- and free --
        # part of the backbone graph already in segment 1)
        cut_in = [f.detach().requires_grad_(f.requires_grad) for f in feats]
        rest = {k: v for k, v in data.items() if k not in ("img", "img_metas", "gt_bboxes", "gt_labels")}
        losses = model.bbox_head.forward_train(tuple(cut_in), data["img_metas"], data["gt_bboxes"], data["gt_labels"],
                                               **rest)
        loss, log_vars = model._parse_losses(losses)
        live = [i for i, f in enumerate(feats) if f.requires_grad]
        g = torch.autograd.grad(loss, head_params + [cut_in[i] for i in live], allow_unused=True)
        self._pack(g[:n_head], head_params)
        b = opt._bounds
        works = [dist.all_reduce(opt.flat_g[b[0]:b[1]], async_op=True)]
        if bb_params:
            g_feat = {i: (gc if gc is not None else torch.zeros_like(feats[i])) for i, gc in zip(live, g[n_head:])}
            split = self._last_stage_split(bb_params) if stage_cut else None
            if split is not None and (len(feats) - 1) in g_feat:
                # segment 2: layer4 (its output is the last feature level) down to the detached copy of layer3's output
                n_early, off_late = split
                late = bb_params[n_early:]
                (x3, leaf3), last = stage_cut[0], len(feats) - 1
                g2 = torch.autograd.grad([feats[last]], late + [leaf3], grad_outputs=[g_feat.pop(last)],
                                         allow_unused=True)
                self._pack(g2[:-1], late)
                works.append(dist.all_reduce(opt.flat_g[off_late:b[3]], async_op=True))
                # segment 3: the remaining feature levels plus layer4's input gradient, through layer3 / layer2
                outs = [feats[i] for i in g_feat] + [x3]
                gouts = [g_feat[i] for i in g_feat] + [g2[-1] if g2[-1] is not None else torch.zeros_like(x3)]
                early = bb_params[:n_early]
                self._pack(torch.autograd.grad(outs, early, grad_outputs=gouts, allow_unused=True), early)
                works.append(dist.all_reduce(opt.flat_g[b[2]:off_late], async_op=True))
            else:
                outs = [feats[i] for i in g_feat]
                self._pack(torch.autograd.grad(outs, bb_params, grad_outputs=[g_feat[i] for i in g_feat],
                                               allow_unused=True), bb_params)
                works.append(dist.all_reduce(opt.flat_g[b[2]:b[3]], async_op=True))
        for w in reversed(works):
            w.wait()
        return loss, log_vars

    def _last_stage_split(self, bb_params):
        """-> (number of backbone parameters before layer4, offset of layer4's first gradient in the flat buffer), or
        None when the backbone group is not laid out [..., layer4] with a 4-aligned boundary."""
        if getattr(self, "_stage_split", "unset") != "unset":
            return self._stage_split
        opt = self.opt
        names = opt.param_names[len(opt.params) - len(bb_params):]
        late = [".layer4." in n or n.startswith("backbone.layer4.") for n in names]
        n_early = late.index(True) if True in late else len(late)
        ok = 0 < n_early < len(late) and all(late[n_early:]) and not any(late[:n_early])
        off = int(opt._bounds[2]) + sum(p.numel() for p in bb_params[:n_early])
        self._stage_split = (n_early, off) if ok and off % 4 == 0 else None
        return self._stage_split

    def __call__(self, data):
        if self.exchange == "peer":
            from .dino import head as _head
            # the two loss normalisers go through the peer control blocks too (one 32-thread kernel each, no NCCL launch)
            _head.set_scalar_allreduce(lambda t, slot: self.opt.small_allreduce(t, slot))
            try:
                with torch.autocast("cuda", dtype=self.autocast, enabled=self.autocast is not None):
                    losses = self.model(**data)
                    loss, log_vars = self.model._parse_losses(losses)
            finally:
                _head.set_scalar_allreduce(None)
            self._pack(torch.autograd.grad(loss, self.opt.params, allow_unused=True))
            self.opt.step_exchange(self.max_grad_norm)
            return loss.detach(), log_vars
        if self.world_size > 1 and self.gather_grads and self.overlap and self.autocast is None \
                and hasattr(self.model, "extract_feat") \
                and hasattr(self.model, "bbox_head"):
            loss, log_vars = self._overlapped_backward(data)
            self.opt.step(self.max_grad_norm, grad_scale=1.0 / self.world_size)
            return loss.detach(), log_vars
        if not self.gather_grads:
            self.opt.zero_grad()
        with torch.autocast("cuda", dtype=self.autocast, enabled=self.autocast is not None):
            losses = self.model(**data)
            loss, log_vars = self.model._parse_losses(losses)
        if self.gather_grads:
            self._pack(torch.autograd.grad(loss, self.opt.params, allow_unused=True))
        else:
            loss.backward()
        self.opt.all_reduce_sum(self.world_size)
        self.opt.step(self.max_grad_norm, grad_scale=1.0 / self.world_size)
        return loss.detach(), log_vars


class FusedSSODTrainStep:
    """The teacher-student iteration of configs/detr_ssod (``DinoDetrSSOD`` wrapper + ``MeanTeacher`` hook +
    OptimizerHook) with the student's clip + AdamW and the teacher's EMA blend in one pass over flat buffers.

    The reference blends at the START of iteration i with ``m_i = min(momentum, 1 - (1 + warm_up) / (i + 1 +
    warm_up))`` (mean_teacher.py:37-50); here the blend with ``m_{i+1}`` rides on the optimizer kernel at the END of
    iteration i -- the same teacher for every forward pass.  Student parameters that are frozen (stem, layer1, BN
    affine) are still blended (the reference has no ``requires_grad`` filter, :60-64) by a second, small
    ``sdb_ema_update_f32`` launch.  ``before_run``'s momentum-0 copy happens in the constructor when ``start_iter == 0``
    (a run resumed at ``start_iter > 0`` keeps the EMA teacher it loaded)."""

    def __init__(self, model, momentum=0.999, warm_up=0, max_grad_norm=0.1, world_size=None, start_iter=0, **opt_kw):
        from .teacher.mean_teacher import EmaPlan
        self.model, self.max_grad_norm = model, max_grad_norm
        self.momentum, self.warm_up = momentum, warm_up
        if world_size is None:
            world_size = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.world_size = world_size
        student = dict(model.student.named_parameters())
        teacher = dict(model.teacher.named_parameters())
        if start_iter == 0:                         # before_run clones only at runner.iter == 0
            with torch.no_grad():                   # (mean_teacher.py:26-35); a resumed run keeps the loaded teacher
                for n, t in teacher.items():
                    t.copy_(student[n])
        self.opt = FusedAdamW(model, teacher_params={"student." + n: t for n, t in teacher.items()}, **opt_kw)
        frozen = [n for n, p in student.items() if not p.requires_grad]
        self.frozen_plan = EmaPlan([teacher[n].data for n in frozen], [student[n].data for n in frozen])
        self.iter = start_iter
        _cache_student_bn_folds(model)

    def momentum_at(self, it):
        return min(self.momentum, 1 - (1 + self.warm_up) / (it + 1 + self.warm_up))

    _pack = FusedSupervisedTrainStep._pack

    def __call__(self, data):
        self.model.curr_step = self.iter            # StepRecord hook (step_record.py:7-27)
        losses = self.model(**data)
        loss, log_vars = self.model._parse_losses(losses)
        self._pack(torch.autograd.grad(loss, self.opt.params, allow_unused=True))
        self.opt.all_reduce_sum(self.world_size)
        m = self.momentum_at(self.iter + 1)
        self.opt.step(self.max_grad_norm, ema_momentum=m, grad_scale=1.0 / self.world_size)
        self.frozen_plan.step(m)
        self.iter += 1
        log_vars["ema_momentum"] = m
        return loss.detach(), log_vars
