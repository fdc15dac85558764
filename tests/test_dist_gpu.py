"""Data-parallel training on 2 GPUs (NCCL): the gradient averaged over two ranks, each holding half of the rays, must equal the
single-GPU gradient of the whole batch -- including the per-sample means (eikonal term) whose sample counts differ per rank
 -- with the bucketed all-reduce that overlaps the colour hash table's reduction with the SDF backward
(Trainer.enable_data_parallel). Skipped on boxes with one GPU."""
import os
import socket

import numpy as np
import pytest
import torch

import scenes

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_trainer(hp_over=None):
    from permuto_sdf import OccupancyGrid, RaySampler, RaySamplesPacked, VolumeRendering
    from permuto_sdf_b200.permuto_sdf import _Pcg32Host
    from permuto_sdf_b200.train import HyperParams, Trainer
    OccupancyGrid.m_rng, RaySampler.m_rng, VolumeRendering.m_rng = _Pcg32Host(), _Pcg32Host(), _Pcg32Host()
    RaySamplesPacked.static_capacity = False
    hp = HyperParams()
    hp.max_nr_samples_per_ray = 32
    hp.nr_samples_imp_sampling = 8
    hp.min_dist_between_samples = 1e-3
    hp.offsurface_weight = 0.0          # its random points come from torch's generator (differ per process)
    hp.jitter_samples = False           # jitter streams are indexed by the LOCAL ray number
    tr = Trainer(hp, nr_levels=8, capacity=2 ** 14, sdf_hidden=64, occupancy_resolution=128, nr_images=4, seed=7, optimizer="fused")
    tr.set_analytic_scene()
    tr.iter_nr = 60000                  # past the curvature phase (random directions from torch's generator)
    return tr


def _batch(n):
    rgb, mask, K, tf = scenes.synthetic_reel(nimg=4, H=60, W=80)

    class Reel:
        pass
    reel = Reel()
    reel.rgb_reel, reel.mask_reel, reel.K_reel, reel.tf_world_cam_reel = [torch.from_numpy(a).cuda() for a in (rgb, mask, K, tf)]
    gen = torch.Generator().manual_seed(21)
    pix = torch.randint(0, 60 * 80, (n,), generator=gen, dtype=torch.int32)
    img = torch.randint(0, 4, (n,), generator=gen, dtype=torch.int32)
    return reel, pix, img


def _worker(rank, world, port, q):
    import sys
    import traceback
    try:
        _worker_body(rank, world, port, q)
    except Exception:
        traceback.print_exc()
        sys.stderr.flush()
        raise


def _worker_body(rank, world, port, q):
    import torch.distributed as dist
    from permuto_sdf import PermutoSDF
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        reel, pix, img = _batch(512)
        lo, hi = rank * 256, (rank + 1) * 256
        tr = _make_trainer()
        tr.enable_data_parallel(world, overlap=True)
        with torch.no_grad():
            o, d, gt, gm, idx = PermutoSDF.rays_from_reel_indices(reel, pix[lo:hi].cuda(), img[lo:hi].cuda())
        tr.step(o, d, gt, gm, idx, update_occupancy=False, optimizer_step=False)
        n_local = int(tr.last["nr_samples"])
        tr.dp_reduce_gradients()
        torch.cuda.synchronize()
        if rank == 0:
            g_dp = tr.optimizer.flat_grad.detach().clone()
            # the same batch on one GPU, no data parallelism
            tr1 = _make_trainer()
            with torch.no_grad():
                o, d, gt, gm, idx = PermutoSDF.rays_from_reel_indices(reel, pix.cuda(), img.cuda())
            tr1.step(o, d, gt, gm, idx, update_occupancy=False, optimizer_step=False)
            g_1 = tr1.optimizer.flat_grad.detach()
            q.put(("result", float((g_dp - g_1).norm() / g_1.norm()), float(g_1.norm()), n_local, int(tr1.last["nr_samples"])))
        else:
            q.put(("count", n_local))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_equals_single_gpu_gradient(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = []
    import queue as _queue
    import time
    deadline = time.time() + 240
    while len(got) < 2:
        try:
            got.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() > deadline:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                pytest.fail("data-parallel worker failed (exit codes %s)" % [p.exitcode for p in procs])
    for p in procs:
        p.join(timeout=60)
    res = next(g for g in got if g[0] == "result")
    cnt = next(g for g in got if g[0] == "count")
    rel, norm, n0, n_all = res[1], res[2], res[3], res[4]
    assert n0 + cnt[1] == n_all and n0 != cnt[1], "the two ranks must hold different sample counts for the weighting to matter"
    assert norm > 0 and rel < 1e-3, "rank-averaged gradient differs from the global-batch gradient: %g" % rel


def _peer_worker(rank, world, port, q):
    import sys
    import traceback
    try:
        import torch.distributed as dist
        from permuto_sdf import PermutoSDF
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        reel, pix, img = _batch(512)
        lo, hi = rank * 256, (rank + 1) * 256
        tr = _make_trainer()
        tr.enable_data_parallel(world, mode="peer")
        opt = tr.optimizer
        with torch.no_grad():
            o, d, gt, gm, idx = PermutoSDF.rays_from_reel_indices(reel, pix[lo:hi].cuda(), img[lo:hi].cuda())
        before = opt.flat_param.detach().clone()
        tr.step(o, d, gt, gm, idx, update_occupancy=False, optimizer_step=False)
        g_mean = opt.flat_grad.detach().clone()
        dist.all_reduce(g_mean)
        g_mean /= world
        tr.optimizer_step()                       # barrier, fused reduce + AdamW + broadcast over peer memory, barrier
        torch.cuda.synchronize()
        after = opt.flat_param.detach().clone()
        # the owner's first moment of its shard = (1 - beta1) * mean gradient
        ok_m = True
        for g, (slo, cnt) in zip(opt.param_groups, opt._peer["shards"]):
            a = g["_off"] + slo
            want = 0.1 * g_mean[a:a + cnt]
            got = opt.exp_avg[a:a + cnt]
            if cnt and float((got - want).abs().max()) > 1e-6 * max(float(want.abs().max()), 1e-20) + 1e-12:
                ok_m = False
        other = [torch.empty_like(after) for _ in range(world)]
        dist.all_gather(other, after)
        same = all(torch.equal(other[0], t) for t in other)
        q.put((rank, bool(ok_m), bool(same), float((after - before).abs().max()), float(opt.flat_grad.abs().max())))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        traceback.print_exc()
        sys.stderr.flush()
        raise


def test_peer_memory_optimizer_step(cuda):
    """Trainer.enable_data_parallel(mode='peer'): gradient reduction + AdamW + parameter broadcast in one kernel per group over NVLink
    peer memory (csrc/optim.cu k_adamw_dp): every rank ends with bit-identical parameters, the owner's moments hold the rank-mean
    gradient, gradients are zeroed for the next iteration"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import queue as _queue
    import time
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, deadline = [], time.time() + 240
    while len(got) < 2:
        try:
            got.append(q.get(timeout=2))
        except _queue.Empty:
            if [p.exitcode for p in procs if p.exitcode not in (None, 0)] or time.time() > deadline:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                pytest.fail("peer-step worker failed (exit codes %s)" % [p.exitcode for p in procs])
    for p in procs:
        p.join(timeout=60)
    for rank, ok_m, same, moved, gmax in got:
        assert ok_m, "rank %d: moments of the owned shard do not hold the mean gradient" % rank
        assert same, "parameters differ between the ranks after the peer step"
        assert moved > 0 and gmax == 0.0
