"""
What the reference cannot do: a whole ensemble of designs in one call.  4096 dome designs (random
force densities around the nominal ones) are evaluated, differentiated and improved in lock step.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from dfdm_b200 import workloads
from dfdm_b200.equilibrium import EquilibriumModel
from dfdm_b200.goals import NodeLineGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import BatchedLBFGS


def main(batch=4096, maxiter=30, verbose=True):
    network = workloads.dome()
    model = EquilibriumModel(network)
    goals = []
    for node in network.nodes_free():
        xyz = network.node_coordinates(node)
        goals.append(NodeLineGoal(node, target=(xyz, [xyz[0], xyz[1], xyz[2] + 1.0])))
    loss = Loss(SquaredError(goals))
    q0 = np.asarray(network.edges_forcedensities())
    q = workloads.sample_q(q0, batch, seed=5, spread=0.2, relative=True)
    loss(q[:2], model)                                   # warm-up
    t0 = time.perf_counter()
    value, grad = loss.value_and_grad(q, model)
    dt = time.perf_counter() - t0
    optimizer = BatchedLBFGS(maxiter=maxiter)
    q_opt, value_opt = optimizer.minimize(model, loss, q, bounds=(None, -1e-3))
    if verbose:
        print(f"{batch} designs: loss + gradient in {1e3 * dt:.2f} ms through the Python API ({batch / dt:.0f} evals/s incl. copies)")
        print(f"mean loss {value.mean():.5f} -> {value_opt.mean():.5f} after {optimizer.iterations} lock-step L-BFGS iterations "
              f"({optimizer.evaluations} batched evaluations)")
    return value, value_opt


if __name__ == "__main__":
    main()
