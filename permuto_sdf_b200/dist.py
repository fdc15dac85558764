"""Data-parallel plumbing of the hot path: rays shard by rank, parameters are replicated, one all-reduce of a flat
gradient buffer per step (NCCL over NVLink on the GPUs; the same code runs on gloo/CPU tensors in the tests).
The reference has no distributed code; this is the single collective of SURVEY.md 8(e)."""
import torch


class FlatGrads:
    """All parameter gradients as views into one flat buffer, so that autograd accumulates straight into the
    communication buffer and the step needs exactly one all-reduce."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise RuntimeError("no trainable parameters")
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=self.params[0].dtype)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self, world=None, group=None):
        import torch.distributed as dist
        world = world or dist.get_world_size(group)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


def shard_range(n_items, rank, world):
    """contiguous shard [lo, hi) of n_items for `rank` (remainder spread over the first ranks)"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(params, src=0, group=None):
    import torch.distributed as dist
    for p in params:
        dist.broadcast(p.data, src=src, group=group)
