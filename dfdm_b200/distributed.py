"""
Data-parallel sharding of a batch of independent designs: one process per GPU, the plan and goal
program replicated, contiguous slices of the batch per rank, and exactly one collective per batched
evaluation — an all-gather of the per-design losses and gradients, or an all-reduce
of their sums in ensemble-mean mode.  There is no exchange inside an evaluation: designs do not couple.

The reference evaluates one design at a time in one process (src/dfdm/optimization.py:89-97); this
module is what lets an ensemble / parameter sweep use 8 x B200.  Works with the ``nccl`` backend on
GPUs and with ``gloo`` on CPU tensors (tests).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard", "gather_designs", "ensemble_mean", "ShardedBatch"]


def shard_bounds(batch, world_size, rank):
    """Contiguous slice [lo, hi) of ``batch`` designs owned by ``rank``; the first ``batch % world_size`` ranks take one more."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("rank %r outside world of size %r" % (rank, world_size))
    base, extra = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(array, world_size, rank):
    lo, hi = shard_bounds(array.shape[0], world_size, rank)
    return array[lo:hi]


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_designs(local, batch, group=None):
    """
    All-gather per-design rows (``local`` is this rank's [b_local, ...] slice) into [batch, ...] on every
    rank, in design order.  Ragged shards are padded to the largest shard for the collective.
    """
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    padded = local
    if local.shape[0] != most:
        pad = torch.zeros((most - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * most,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(hi - lo == most for lo, hi in sizes):
        return out
    return torch.cat([out[r * most: r * most + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def ensemble_mean(loss_local, dq_local, batch, group=None):
    """Mean loss and mean gradient over the whole ensemble: one all-reduce of ``[E + 1]`` doubles."""
    world, _ = _world(group)
    packed = torch.cat([dq_local.sum(dim=0), loss_local.sum().reshape(1)])
    if world > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    packed = packed / float(batch)
    return packed[-1], packed[:-1]


class ShardedBatch:
    """
    Evaluate a global batch with this rank's evaluator and gather the results.

    ``evaluate(q_local) -> (loss[b_local], dq[b_local, E])`` is the per-rank compute (on the GPU:
    ``Evaluator.loss_and_grad``); the class owns only the partition and the collective.
    """

    def __init__(self, evaluate, group=None):
        self.evaluate = evaluate
        self.group = group
        self.world_size, self.rank = _world(group)

    def local_slice(self, q):
        return shard(q, self.world_size, self.rank)

    def loss_and_grad(self, q, gather=True):
        batch = q.shape[0]
        loss, dq = self.evaluate(self.local_slice(q))
        if not gather:
            return loss, dq
        return gather_designs(loss, batch, self.group), gather_designs(dq, batch, self.group)

    def ensemble(self, q):
        batch = q.shape[0]
        loss, dq = self.evaluate(self.local_slice(q))
        return ensemble_mean(loss, dq, batch, self.group)
