"""
configs[3] of BASELINE.json in ensemble form: a batch of dome designs improved in lock step, sharded over the
GPUs of one node (one process per GPU).  Designs are independent, so every rank optimises its own contiguous
slice with its optimiser state on its own GPU (device-resident lock-step L-BFGS): no design data crosses PCIe or NVLink
during the optimisation - one flag per iteration keeps the ranks in step - and ONE all-gather at the end collects the
optimal force densities and losses.

    python examples/ensemble_sharded.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29500 examples/ensemble_sharded.py                  # 8 x B200
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from dfdm_b200 import workloads
from dfdm_b200.distributed import gather_designs, shard_bounds
from dfdm_b200.equilibrium import EquilibriumModel
from dfdm_b200.goals import NodeLineGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import BatchedLBFGS


def main(batch_per_gpu=512, maxiter=30, verbose=True):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    batch = batch_per_gpu * world

    network = workloads.dome()
    model = EquilibriumModel(network, device=f"cuda:{local_rank}")
    goals = []
    for node in network.nodes_free():
        xyz = network.node_coordinates(node)
        goals.append(NodeLineGoal(node, target=(xyz, [xyz[0], xyz[1], xyz[2] + 1.0])))
    loss = Loss(SquaredError(goals))
    q0 = np.asarray(network.edges_forcedensities())
    q_all = workloads.sample_q(q0, batch, seed=5, spread=0.2, relative=True)     # the same ensemble on every rank
    lo, hi = shard_bounds(batch, world, rank)

    device = torch.device("cuda", local_rank)
    q_local = torch.as_tensor(np.ascontiguousarray(q_all[lo:hi]), device=device)        # this rank's slice: on its GPU from here on
    start, _ = loss.value_and_grad_device(q_local, model)
    start = start.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    optimizer = BatchedLBFGS(maxiter=maxiter)                  # state (q, gradients, curvature pairs) lives with the designs
    q_opt, value_opt = optimizer.minimize(model, loss, q_local, bounds=(None, -1e-3))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0

    # the ONLY exchange of the run: the optimal force densities and losses of every rank, once, at the end
    q_opt_all = gather_designs(q_opt.contiguous(), batch)
    value_all = gather_designs(value_opt.contiguous(), batch)
    start_all = gather_designs(start, batch)
    if verbose and rank == 0:
        print(f"{batch} dome designs on {world} GPU(s): mean loss {float(start_all.mean()):.5f} -> {float(value_all.mean()):.5f} "
              f"in {optimizer.iterations} lock-step L-BFGS iterations, {optimizer.evaluations} batched evaluations, {dt:.2f} s "
              f"({batch * optimizer.evaluations / dt:.0f} design evaluations/s)")
    if world > 1:
        dist.barrier()
    return start_all.cpu().numpy(), value_all.cpu().numpy(), q_opt_all.cpu().numpy()


if __name__ == "__main__":
    main()
    if dist.is_initialized():
        dist.destroy_process_group()
