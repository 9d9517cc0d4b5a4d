"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): PCG with the multigrid hierarchy
(cluster kernel included), its Jacobi variant, the matrix-free product, and the fused direct kernel."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfdm_b200 import workloads as W, _lib
from dfdm_b200.engine import Plan, GoalProgram, Evaluator

edges, fixed, xyz, loads, q0 = W.grid_arrays(48)
plan = Plan(len(fixed), edges, fixed)
lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
for B in (6, 5):
    q = W.sample_q(q0, B, seed=1, spread=0.3, relative=True)
    for kw in (dict(), dict(pcg_precond="jacobi", pcg_maxiter=40), dict(flags=_lib.FLAG_SPMV_MATRIX_FREE), dict(pcg_check_every=3)):
        ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", **kw)
        out = ev.loss_and_grad(prog, q, outputs=("xyz", "residuals", "vectors", "lengths", "forces"))
        print("pcg", B, kw, "levels", plan.info["mg_levels"], "iters", _lib.last_pcg_iterations(), "status", out["status"].cpu().numpy().tolist(), flush=True)
edges, fixed, xyz, loads, q0 = W.grid_arrays(9)
plan = Plan(len(fixed), edges, fixed)
lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
prog = GoalProgram(plan, [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))], [("SquaredError", 1.0)], 0.0)
ev = Evaluator(plan, loads, xyz[plan.fixed], solver="direct")
out = ev.loss_and_grad(prog, W.sample_q(q0, 7, seed=2, spread=0.3, relative=True), outputs=("xyz",))
print("direct status", out["status"].cpu().numpy().tolist(), float(out["loss"].sum()))
