"""
Optimiser iterations per second of the device-resident lock-step L-BFGS on BASELINE configs[3] in ensemble form: dome
network, 4 096 designs in all, sharded over the GPUs of the node (one process per GPU; run under torchrun for N > 1).
Prints one JSON line on rank 0.   python tools/bench_optimizer.py [--designs 4096] [--maxiter 30]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from dfdm_b200 import workloads                                # noqa: E402
from dfdm_b200.distributed import shard_bounds                 # noqa: E402
from dfdm_b200.equilibrium import EquilibriumModel            # noqa: E402
from dfdm_b200.goals import NodeLineGoal                      # noqa: E402
from dfdm_b200.losses import Loss, SquaredError              # noqa: E402
from dfdm_b200.optimization import BatchedLBFGS              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--designs", type=int, default=4096)
    ap.add_argument("--maxiter", type=int, default=30)
    ap.add_argument("--rings", type=int, default=6)
    ap.add_argument("--sides", type=int, default=16)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    network = workloads.dome(num_sides=args.sides, num_rings=args.rings)
    model = EquilibriumModel(network, device=f"cuda:{local}")
    goals = []
    for node in network.nodes_free():
        xyz = network.node_coordinates(node)
        goals.append(NodeLineGoal(node, target=(xyz, [xyz[0], xyz[1], xyz[2] + 1.0])))
    loss = Loss(SquaredError(goals))
    q0 = np.asarray(network.edges_forcedensities())
    q_all = workloads.sample_q(q0, args.designs, seed=5, spread=0.2, relative=True)
    lo, hi = shard_bounds(args.designs, world, rank)
    q_local = torch.as_tensor(np.ascontiguousarray(q_all[lo:hi]), device=device)
    BatchedLBFGS(maxiter=2).minimize(model, loss, q_local, bounds=(None, -1e-3))      # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    opt = BatchedLBFGS(maxiter=args.maxiter)
    q_opt, value = opt.minimize(model, loss, q_local, bounds=(None, -1e-3))
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    stats = torch.stack([value.sum(), torch.tensor(float(hi - lo), dtype=torch.float64, device=device)])
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats)
    if rank == 0:
        secs = float(dt[0])
        print(json.dumps({"metric": "lock-step L-BFGS iterations/s over the ensemble", "value": opt.iterations / secs, "unit": "iterations/s",
                          "n_gpus": world, "designs": args.designs, "iterations": opt.iterations, "batched_evaluations": opt.evaluations,
                          "design_evaluations_per_s": args.designs * opt.evaluations / secs, "seconds": secs,
                          "mean_loss_end": float(stats[0] / stats[1]), "n_free": model.plan.n_free, "n_edges": model.plan.n_edges,
                          "state": "device resident (q, gradients, curvature pairs per rank); inter-GPU traffic: one flag per iteration"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
