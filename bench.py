"""bench.py -- the driver's measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): one supervised DINO-4scale R50 train step -- forward, 13-way Hungarian-matched
loss with contrastive denoising, backward, grad-clip 0.1, AdamW -- on a synthetic COCO-shape batch of 2 images
800x1333 per GPU, random-init weights (no network for checkpoints / datasets).  Data parallel over N GPUs of one
node: one process per GPU, one NCCL all-reduce of the flat gradient buffer over NVLink per step, weak scaling.

Prints ONE JSON line on rank 0 (see README / DESIGN.md for the keys):
  value     images/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       images/s with the batch copied from pinned host memory every step and the loss read back
  roofline  achieved algorithmic GB/s of the dominant kernel of this library (MSDA backward, encoder shape),
            timed live with CUDA events around each launch inside the timed region
  cpu_baseline  (N=1 only) the reference's CPU path (oracle: python ms_deform_attn fallback + host Hungarian +
            torch-cpu) timed on this box's host cores on a bounded sample
`--impl reference` times that CPU path alone (rank 0 only) and prints the same line with "impl": "reference".
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (train step) DINO-4scale R50"
UNIT = "images/s"
IMG_H, IMG_W, PER_GPU_BATCH = 800, 1333, 2
WORKLOAD = ("configs[1]: DINO-4scale R50 supervised train step (fwd + CDN + 13x Hungarian-matched loss + bwd + "
            "clip 0.1 + AdamW), bs=2/GPU, synthetic COCO-shape 800x1333")


def measured_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


def build_model(device):
    import torch
    from semi_detr_b200 import dino  # noqa: F401
    from semi_detr_b200.registry import DETECTORS
    from semi_detr_b200.synthetic import DINO_R50_4SCALE
    torch.manual_seed(0)
    model = DETECTORS.build(copy.deepcopy(DINO_R50_4SCALE))
    return model.to(device).train()


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path (oracle restatement) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step_time(n_images, height, width, steps, warmup, threads):
    import torch
    from oracle.cpu_path import reference_cpu_ops
    from semi_detr_b200.engine import SupervisedTrainStep, build_optimizer
    from semi_detr_b200.synthetic import coco_like_batch
    torch.set_num_threads(threads)
    model = build_model("cpu")
    step = SupervisedTrainStep(model, build_optimizer(model, fused=False))
    data = coco_like_batch(n_images, height, width, seed=0)
    times = []
    with reference_cpu_ops():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            loss, _ = step(data)
            float(loss)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: ONE full-size image per step unless K+W steps of that would not end within a few minutes,
    # then a half-resolution image counted in full-size-image equivalents (pixel ratio)
    total_steps = args.steps + args.warmup
    h, w, scale_note = IMG_H, IMG_W, ""
    if total_steps > 16:
        h, w = IMG_H // 2, (IMG_W + 1) // 2
        scale_note = " (half resolution; images/s in 800x1333-pixel equivalents)"
    times = cpu_reference_step_time(1, h, w, args.steps, args.warmup, threads)
    sec = sum(times) / len(times)
    equiv = (h * w) / float(IMG_H * IMG_W)
    value = equiv / sec
    sample = f"{args.steps} steps of 1 synthetic image {h}x{w} per step{scale_note}, {threads} host threads"
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, sample=sample, parallelism="cpu"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from semi_detr_b200 import _lib
    from semi_detr_b200.engine import FusedSupervisedTrainStep, GraphedTrainStep, SupervisedTrainStep, build_optimizer
    from semi_detr_b200.msda import MultiScaleDeformableAttention as MSDA
    from semi_detr_b200.synthetic import coco_like_batch, msda_bytes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a CUDA device: semi_detr_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner to file descriptor 1 when the first communicator is created: point fd 1 at
        # stderr until then, so stdout carries exactly ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    _lib.lib()  # fail loudly if the extension is missing
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True

    model = build_model(device)
    if world > 1:   # identical seeds give identical weights; broadcast anyway so the ranks cannot drift
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    if args.torch_optimizer:
        step = SupervisedTrainStep(model, build_optimizer(model, capturable=True), world_size=world)
    else:   # clip + AdamW as one kernel over flat buffers (sdb_adamw_ema_step_f32)
        step = FusedSupervisedTrainStep(model, world_size=world)
    host = coco_like_batch(PER_GPU_BATCH, IMG_H, IMG_W, seed=rank, pin=True)

    def to_device(b):
        return dict(img=b["img"].to(device, non_blocking=True), img_metas=[dict(m) for m in b["img_metas"]],
                    gt_bboxes=[x.to(device, non_blocking=True) for x in b["gt_bboxes"]],
                    gt_labels=[x.to(device, non_blocking=True) for x in b["gt_labels"]])

    resident = to_device(host)
    h2d_bytes = host["img"].numel() * 4 + sum(x.numel() * 4 for x in host["gt_bboxes"]) + \
        sum(x.numel() * 8 for x in host["gt_labels"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    if args.ncu:
        step(resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps(dict(note="--ncu mode: no measurement", launches=dict(_lib.LAUNCHES))))
        return
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(resident)
    barrier()
    # kernels of this library per step (python-side counters tick when the launchers run, i.e. in eager mode)
    launches0 = sum(_lib.LAUNCHES.values())
    step(resident)
    launches_per_step = sum(_lib.LAUNCHES.values()) - launches0

    graphed, graph_note = None, "eager"
    if not args.no_graph:
        try:
            graphed = GraphedTrainStep(step, resident, warmup=2)
            graph_note = "whole step captured in one CUDA graph, replayed"
        except Exception as e:  # pragma: no cover - fall back loudly, never silently
            graphed, graph_note = None, f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            print("WARNING: CUDA graph capture failed, running eagerly:", e, file=sys.stderr)
            torch.cuda.synchronize()
    run = (lambda: graphed()) if graphed else (lambda: step(resident))
    for _ in range(warm):
        run()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- timed region 1: inputs resident in HBM --------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = launches_per_step * args.steps

    # ---- timed region 2: end to end (pinned host -> device every step, loss read back) -------------------
    # Every step's inputs cross PCIe inside the timed region; with the graph engine the copy of step i+1 is issued
    # (side stream, staging buffers) right after step i is enqueued, so it overlaps step i's compute like a
    # prefetching data loader would -- K copies for K steps, the first one exposed.
    run_e2e = (lambda: graphed(host)) if graphed else (lambda: step(to_device(host)))
    for _ in range(2):
        float(run_e2e()[0])
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = 0.0
    if graphed:
        graphed.prefetch(host)
        for i in range(args.steps):
            loss, _ = graphed(prefetched=True)
            if i + 1 < args.steps:
                graphed.prefetch(host)
            last = float(loss)                    # device -> host read of the step's result
    else:
        for _ in range(args.steps):
            loss, _ = run_e2e()
            last = float(loss)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel timing for the roofline: CUDA events around every MSDA launch.  Events recorded inside a
    # captured graph cannot be queried, so the same step runs eagerly for this part -----------------------
    MSDA.EVENT_LOG = []
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_roof = min(args.steps, 5)
    barrier()
    e4.record()
    for _ in range(n_roof):
        step(resident)
    e5.record()
    barrier()
    ms_roof = e4.elapsed_time(e5)
    log, MSDA.EVENT_LOG = MSDA.EVENT_LOG, None

    if rank != 0:
        _leave(world)
        return

    imgs = PER_GPU_BATCH * world * args.steps
    value = imgs / (ms_total / 1e3)
    e2e = imgs / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: MSDA backward at the encoder shape (Lq == S) ---------------------
    peak, peak_src = measured_peak()
    by = {}
    for kind, b, s, q, a, z in log:
        by.setdefault((kind, "enc" if q == s else "dec", b, s, q), []).append(a.elapsed_time(z) * 1e-3)
    kernels = {}
    for (kind, shape, b, s, q), ts in by.items():
        fb, bb = msda_bytes(b, s, q)
        nbytes = fb if kind == "fwd" else bb
        avg = sum(ts) / len(ts)
        kernels[f"msda_{kind}_{shape}"] = dict(launches=len(ts), avg_us=avg * 1e6, bytes=nbytes,
                                               gbs=nbytes / avg / 1e9, total_ms=sum(ts) * 1e3,
                                               ms_per_step=sum(ts) * 1e3 / n_roof)
    dom_name = max(kernels, key=lambda k: kernels[k]["total_ms"]) if kernels else None
    roofline = None
    if dom_name:
        d = kernels[dom_name]
        # DRAM bytes per launch of the same kernel at the same shape, from the committed `ncu --set full` capture of
        # `bench.py --ncu` (profiles/ncu_msda_step_r1.txt); ncu cannot run inside the timed region.
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_msda_traffic_r1.json")) as f:
                tj = json.load(f)
            traffic = tj["kernels"][dom_name]["dram_bytes_per_launch"]
            traffic_src = tj["source"]
        except (OSError, KeyError, ValueError):
            pass
        roofline = dict(bound="hbm", kernel=dom_name, achieved=d["gbs"], peak=peak, unit="GB/s",
                        frac=d["gbs"] / peak, traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                        avg_launch_us=d["avg_us"],
                        algorithmic_bytes_per_launch=d["bytes"], launches_timed=d["launches"],
                        share_of_step=d["ms_per_step"] / (ms_total / args.steps),
                        timed_in="eager replica of the timed step (events inside a captured graph cannot be read); "
                                 f"{n_roof} steps, {ms_roof / n_roof:.2f} ms/step eager",
                        all_msda={k: dict(avg_us=round(v["avg_us"], 2), gbs=round(v["gbs"], 1),
                                          frac=round(v["gbs"] / peak, 4), launches=v["launches"],
                                          share_of_step=round(v["ms_per_step"] / (ms_total / args.steps), 4))
                                  for k, v in kernels.items()})

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times = cpu_reference_step_time(PER_GPU_BATCH, IMG_H, IMG_W, 1, 0, threads)
        cpu_baseline = dict(value=PER_GPU_BATCH / times[0], unit=UNIT, cores=threads, kind="port",
                            sample=f"1 full train step, {PER_GPU_BATCH} images {IMG_H}x{IMG_W}, reference CPU path "
                                   f"(python ms_deform_attn fallback + host LSAP + torch-cpu), {threads} threads")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warm,
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 (tf32 tensor-core matmul/conv; MSDA, matching and losses in f32)", data="synthetic",
                config=dict(workload=WORKLOAD, global_batch=PER_GPU_BATCH * world, parallelism=f"dp{world}",
                            execution=graph_note,
                            l2="per-step working set (activations + 25.6 MB input) exceeds the 126 MB L2"),
                clocks=clocks,
                e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=4,
                         ms_per_step=ms_e2e / args.steps, last_loss=last),
                gpu_launches=launches, gpu_launches_per_step=launches_per_step,
                roofline=roofline, cpu_baseline=cpu_baseline)
    print(json.dumps(line), flush=True)
    _leave(world)


def _leave(world):
    """End a multi-rank run without NCCL teardown: destroying a communicator that a live CUDA graph still references
    was observed to block (N=2, round 1), so every rank drains its device, meets the others at a host-side barrier
    and exits immediately."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--torch-optimizer", action="store_true",
                    help="use torch.optim.AdamW(fused) + flat-buffer clip instead of the fused clip+AdamW kernel")
    ap.add_argument("--ncu", action="store_true",
                    help="profiling aid (numbers printed in this mode are NOT bench values): 1 eager warm-up step + "
                         "1 eager step, nothing else, so an `ncu --metrics gpu__time_duration.sum` launch list stays short")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
