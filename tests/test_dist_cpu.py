"""Host-side logic of the data-parallel path on CPU with gloo, world size 2: ray sharding, identical replicas,
one all-reduce of the flat gradient buffer giving the gradient of the global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from permuto_sdf_b200.dist import FlatGrads, broadcast_parameters, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(123)                      # identical parameter init on every rank
        model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 3))
        table = torch.nn.Parameter(torch.randn(64, 2))
        params = list(model.parameters()) + [table]
        if rank == 1:                               # perturb, then restore through the broadcast helper
            with torch.no_grad():
                params[0].add_(1.0)
        broadcast_parameters(params, src=0)
        flat = FlatGrads(params)
        g = torch.Generator().manual_seed(7)
        rays = torch.randn(10, 6, generator=g)      # global batch of 10 rays (same on every rank)
        idx = torch.randint(0, 64, (10,), generator=g)
        lo, hi = shard_range(10, rank, world)
        x, ix = rays[lo:hi], idx[lo:hi]
        # loss = mean over the GLOBAL batch -> local sum / global count, then SUM all-reduce == mean of rank means * world...
        out = model(x) + table[ix].sum(1, keepdim=True)
        loss = (out ** 2).sum() / 10.0
        flat.zero()
        loss.backward()
        assert params[0].grad.data_ptr() == flat.flat.data_ptr(), "gradients must accumulate inside the flat buffer"
        dist.all_reduce(flat.flat, op=dist.ReduceOp.SUM)
        # plain Python data through the queue: tensors would travel as shared-memory handles that die with this process
        q.put((rank, flat.flat.tolist(), lo, hi))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    res = [(r, torch.tensor(v), lo, hi) for r, v, lo, hi in res]
    assert torch.equal(res[0][1], res[1][1]), "ranks disagree after the all-reduce"
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 5, 5, 10)
    # single-process reference on the whole batch
    torch.manual_seed(123)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.GELU(), torch.nn.Linear(16, 3))
    table = torch.nn.Parameter(torch.randn(64, 2))
    params = list(model.parameters()) + [table]
    g = torch.Generator().manual_seed(7)
    rays = torch.randn(10, 6, generator=g)
    idx = torch.randint(0, 64, (10,), generator=g)
    loss = ((model(rays) + table[idx].sum(1, keepdim=True)) ** 2).sum() / 10.0
    grads = torch.autograd.grad(loss, params)
    ref = torch.cat([x.reshape(-1) for x in grads])
    assert torch.allclose(res[0][1], ref, atol=1e-6, rtol=1e-5)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 512, 65536):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
