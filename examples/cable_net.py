"""
Large quad cable net (BASELINE.json configs[4] at a chosen size): the network is too big for the
shared-memory factorisation, so the evaluation runs multigrid-preconditioned CG on the GPU.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from dfdm_b200 import _lib, workloads
from dfdm_b200.engine import Evaluator, GoalProgram, Plan


def main(nfree=256, batch=64, verbose=True):
    edges, fixed, xyz, loads, q0 = workloads.grid_arrays(nfree)
    plan = Plan(len(fixed), edges, fixed)
    ev = Evaluator(plan, loads, xyz[plan.fixed])
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    prog = GoalProgram(plan, [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))], [("SquaredError", 1.0)], 0.1)
    q = workloads.sample_q(q0, batch, seed=6, spread=0.2, relative=True)
    ev.loss_and_grad(prog, q[:2])
    t0 = time.perf_counter()
    out = ev.loss_and_grad(prog, q, outputs=("xyz",))
    loss = out["loss"].cpu().numpy()
    dt = time.perf_counter() - t0
    if verbose:
        print(f"{plan.n_free} free nodes, {plan.n_edges} edges, solver {plan.auto_solver}, {plan.info['mg_levels'] + 1} multigrid levels")
        print(f"{batch} designs: {1e3 * dt:.1f} ms, PCG iterations (forward, adjoint) = {_lib.last_pcg_iterations()}")
        print(f"loss min / mean / max: {loss.min():.6f} / {loss.mean():.6f} / {loss.max():.6f}")
        print(f"lowest point of design 0: z = {out['xyz'].cpu().numpy()[0, :, 2].min():.6f}")
    return loss


if __name__ == "__main__":
    main()
