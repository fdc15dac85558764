"""Timing of the data-parallel optimizer step in isolation (torchrun, one rank per GPU): peer-memory fused step vs NCCL all-reduce +
AdamW, plus the cost of the two cross-rank barriers and of the gradient zero-fill. usage (under gpurun --gpus N):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dp_step_timing.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "compat")]


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / n * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    from permuto_sdf_b200.optim import FusedAdamW
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    sizes = [2 ** 23 + 20000, 2 ** 23, 60000, 24]          # the bench's groups: sdf table + mlp, rgb table, rgb mlp, colorcal
    def make():
        ps = [torch.nn.Parameter(torch.randn(n, device="cuda") * 0.01) for n in sizes]
        return FusedAdamW([{"params": [p], "weight_decay": 0.0, "lr": 1e-3, "name": "g%d" % i} for i, p in enumerate(ps)], betas=(0.9, 0.99), eps=1e-15)
    out = {"world": world}
    o1 = make()
    g = torch.randn_like(o1.flat_grad) * 1e-3
    def nccl_step():
        o1.flat_grad.copy_(g)
        dist.all_reduce(o1.flat_grad)
        o1.step(grad_scale=1.0 / world)
    def fill_only():
        o1.flat_grad.copy_(g)
    out["fill_us"] = timeit(fill_only)
    out["nccl_allreduce_plus_adamw_us"] = timeit(nccl_step) - out["fill_us"]
    def ar_only():
        dist.all_reduce(o1.flat_grad)
    out["nccl_allreduce_us"] = timeit(ar_only)
    def adamw_only():
        o1.flat_grad.copy_(g); o1.step()
    out["adamw_full_us"] = timeit(adamw_only) - out["fill_us"]
    o2 = make()
    o2.enable_peer_step()
    def peer_step():
        o2.flat_grad.copy_(g)
        o2.step(grad_scale=1.0 / world)
    out["peer_step_us"] = timeit(peer_step) - out["fill_us"]
    pe = o2._peer
    out["multicast"] = bool(pe["multicast"])
    if pe["multicast"]:
        pe["multicast"] = False
        out["peer_step_unicast_us"] = timeit(peer_step) - out["fill_us"]
        pe["multicast"] = True
    out["barrier_us"] = timeit(lambda: pe["hg"].barrier(channel=0))
    out["zero_us"] = timeit(lambda: o2.flat_grad.zero_())
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
