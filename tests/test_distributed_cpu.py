"""World-size-2 gloo checks of the data-parallel host logic (no GPU): the cross-rank loss normaliser
(dino_detr_head.py:698-699, 720-723), log-var reduction (base.py:202-207) and DDP gradient averaging with the
reference CPU ops injected."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _small_cfg():
    from semi_detr_b200.synthetic import DINO_R50_4SCALE
    cfg = copy.deepcopy(DINO_R50_4SCALE)
    cfg["bbox_head"]["num_query"] = 60
    cfg["bbox_head"]["transformer"] = dict(type="DINOTransformer", num_queries=60, num_encoder_layers=1,
                                           num_decoder_layers=2, dim_feedforward=64)
    cfg["bbox_head"]["dn_number"] = 10
    return cfg


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle.cpu_path import reference_cpu_ops
        from semi_detr_b200 import dino  # noqa: F401
        from semi_detr_b200.dino.head import reduce_mean_scalar
        from semi_detr_b200.registry import DETECTORS
        from semi_detr_b200.synthetic import coco_like_batch
        # 1. normaliser: mean over ranks of a host scalar, returned as a tensor
        m = reduce_mean_scalar(float(3 + 4 * rank), "cpu")
        assert torch.is_tensor(m) and abs(float(m) - 5.0) < 1e-6
        # 2. DDP step on different data per rank
        torch.manual_seed(0)
        model = DETECTORS.build(_small_cfg()).train()
        from semi_detr_b200.engine import FlatGrads
        grads = FlatGrads(list(model.parameters()))
        data = coco_like_batch(1, 128, 160, seed=10 + rank)
        with reference_cpu_ops():
            losses = model(**data)
            loss, log_vars = model._parse_losses(losses)
            loss.backward()
            _, reduced = model._parse_losses(losses, reduce_log_vars=True)
        local = grads.flat.clone()
        grads.all_reduce_mean(world)                       # the step's one gradient exchange
        both = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        assert torch.allclose(grads.flat, (both[0] + both[1]) / 2, rtol=1e-6, atol=1e-9)
        norm_before = float(grads.flat.norm())
        returned = float(grads.clip_(0.1))
        assert abs(returned - norm_before) < 1e-5 * norm_before and float(grads.flat.norm()) <= 0.1 + 1e-5
        g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        gathered = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(gathered, g)
        assert torch.equal(gathered[0], gathered[1]), "the exchange must leave identical (averaged) gradients on all ranks"
        losses_all = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(losses_all, loss.detach().reshape(1))
        mean_loss = sum(float(x) for x in losses_all) / world
        assert abs(reduced["loss"] - mean_loss) < 1e-4 * abs(mean_loss)
        assert isinstance(reduced["loss_cls"], float) and len(reduced) == len(log_vars)
        if rank == 0:
            out.put(("ok", float(loss)))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            out.put(("fail", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(280)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    status, val = q.get(timeout=5)
    assert status == "ok", val


def test_bench_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without work."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _ssod_worker(rank, world, port, out):
    """Cross-rank couplings of the teacher-student step: the pooled matched costs behind the GMM
    threshold (dino_detr_ssod.py:303, dist_utils.py:5-30) -- ragged lengths, an empty rank, overflow of the buffer
    size -- must give every rank the same pool, hence the same threshold."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from semi_detr_b200.ssod.dino_detr_ssod import concat_all_gather_1d
        from oracle.gmm_oracle import fit_gmm_threshold
        g = torch.Generator().manual_seed(7)
        pools = [torch.randn(37, generator=g) - 2.0, torch.randn(5, generator=g) + 1.5]   # rank 0 / rank 1 costs
        got = concat_all_gather_1d(pools[rank])
        assert torch.equal(got, torch.cat(pools)), "rank order and values of the pooled costs"
        thr = fit_gmm_threshold(got.numpy())
        both = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(both, torch.tensor([thr], dtype=torch.float64))
        assert float(both[0]) == float(both[1]) == fit_gmm_threshold(torch.cat(pools).numpy())
        # one rank without a single pseudo box: the other rank's costs alone decide
        mine = pools[0] if rank == 0 else torch.zeros(0)
        got = concat_all_gather_1d(mine)
        assert torch.equal(got, pools[0])
        # nothing anywhere: empty pool, threshold 0 (gmm.py), no hang
        got = concat_all_gather_1d(torch.zeros(0))
        assert got.numel() == 0 and fit_gmm_threshold(got.numpy()) == 0.0
        # longer than the fixed buffer: an error on every rank (it used to truncate silently), before any collective
        long = torch.arange(10, dtype=torch.float32) + 100 * rank
        try:
            concat_all_gather_1d(long, max_len=4)
            raise AssertionError("expected ValueError")
        except ValueError:
            pass
        # the segment form the GMM kernel consumes: padded layout + per-rank counts, nothing read back
        from semi_detr_b200.ssod.dino_detr_ssod import pooled_cost_segments
        costs, seg_counts, stride = pooled_cost_segments(pools[rank], max_len=64)
        assert stride == 65 and seg_counts.tolist() == [37, 5] and seg_counts.dtype == torch.int32
        assert torch.equal(costs[:37], pools[0]) and torch.equal(costs[stride:stride + 5], pools[1])
        assert np.isfinite(thr)
        if rank == 0:
            out.put(("ok", thr))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            out.put(("fail", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_ssod_cost_pool_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ssod_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(280)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    status, val = q.get(timeout=5)
    assert status == "ok", val
