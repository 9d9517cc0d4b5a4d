"""
Worker of tests/test_gpu_multi.py, one process per GPU under torchrun: the sharded evaluation of a batch against the
same batch evaluated whole on this rank's own GPU - row offsets of the peer-store gather, chunked pushes, the
clone-by-default lifetime of the results, every exchange mode, close().
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dfdm_b200 import workloads as W                                   # noqa: E402
from dfdm_b200.engine import Plan, GoalProgram, Evaluator             # noqa: E402
from dfdm_b200.distributed import ShardedBatch, PeerGather, shard_bounds   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    edges, fixed, xyz, loads, q0 = W.grid_arrays(24)
    plan = Plan(len(fixed), edges, fixed)
    ev = Evaluator(plan, loads, xyz[plan.fixed], device=device, solver="pcg")
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    batch = 16 * world
    q = torch.from_numpy(W.sample_q(q0, batch, seed=4, spread=0.3, relative=True)).to(device)
    whole = ev.loss_and_grad(prog, q)
    loss_ref, dq_ref = whole["loss"].clone(), whole["dq"].clone()

    def evaluate(q_local):
        out = ev.loss_and_grad(prog, q_local)
        return out["loss"], out["dq"]

    assert PeerGather.available(batch)
    for chunks in (1, 2, 4):
        sb = ShardedBatch(evaluate, chunks=chunks)
        loss, dq = sb.loss_and_grad(q)
        assert "peer stores" in sb.collective, sb.collective
        # per-design results do not depend on what else is in the batch beyond rounding of nothing: designs never mix
        assert torch.allclose(loss, loss_ref, rtol=1e-10, atol=0), (rank, chunks)
        assert float((dq - dq_ref).abs().max() / dq_ref.abs().max()) < 1e-9, (rank, chunks)
        # a caller may keep a result across later evaluations (the IPC buffers behind it are reused two calls later)
        kept_loss, kept_dq = loss, dq
        snapshot = dq.clone()
        for scale in (0.9, 1.1, 0.8):
            sb.loss_and_grad(q * scale)
        assert torch.equal(kept_dq, snapshot) and torch.equal(kept_loss, loss)
        sb.close()
    zc = ShardedBatch(evaluate, zero_copy=True)
    loss, dq = zc.loss_and_grad(q)
    assert float((dq - dq_ref).abs().max() / dq_ref.abs().max()) < 1e-9
    zc.close()
    # NCCL path and the other exchanges
    nccl = ShardedBatch(evaluate, peer=False)
    loss, dq = nccl.loss_and_grad(q)
    assert "nccl" in nccl.collective and float((dq - dq_ref).abs().max() / dq_ref.abs().max()) < 1e-9
    lo, hi = shard_bounds(batch, world, rank)
    root_loss, root_dq = ShardedBatch(evaluate, exchange="root").loss_and_grad(q)
    if rank == 0:
        assert float((root_dq - dq_ref).abs().max() / dq_ref.abs().max()) < 1e-9 and torch.allclose(root_loss, loss_ref, rtol=1e-10, atol=0)
    else:
        assert root_loss is None and root_dq is None
    all_loss, own_dq = ShardedBatch(evaluate, exchange="losses").loss_and_grad(q)
    assert torch.allclose(all_loss, loss_ref, rtol=1e-10, atol=0) and float((own_dq - dq_ref[lo:hi]).abs().max() / dq_ref.abs().max()) < 1e-9
    mean_loss, mean_dq = ShardedBatch(evaluate, exchange="mean").loss_and_grad(q)
    assert torch.allclose(mean_loss, loss_ref.mean(), rtol=1e-10, atol=0)
    assert float((mean_dq - dq_ref.mean(dim=0)).abs().max() / dq_ref.mean(dim=0).abs().max()) < 1e-9
    dist.barrier()
    if rank == 0:
        print("multi_gpu_worker: ok on %d ranks" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
