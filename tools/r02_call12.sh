#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "beyond_shared" > gpurun_out/r02_pytest12a.log 2>&1; echo "pytest new rc $?"
tail -25 gpurun_out/r02_pytest12a.log | cut -c1-400
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest12.log 2>&1; echo "pytest all rc $?"
tail -8 gpurun_out/r02_pytest12.log | cut -c1-300
python - <<'PY'
# how slow is the global-memory factorisation? grid 128 (nf = 16384, w ~ 128) and grid256, a few designs
import time, numpy as np, torch
from dfdm_b200 import workloads as W
from dfdm_b200.engine import Plan, GoalProgram, Evaluator
for nfree, B in ((128, 4), (256, 2)):
    edges, fixed, xyz, loads, q0 = W.grid_arrays(nfree)
    plan = Plan(len(fixed), edges, fixed)
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    prog = GoalProgram(plan, [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))], [("SquaredError", 1.0)], 0.1)
    q = W.sample_q(q0, B, seed=1, spread=0.3, relative=True)
    d = Evaluator(plan, loads, xyz[plan.fixed], solver="direct")
    p = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg")
    od = d.loss_and_grad(prog, q); torch.cuda.synchronize()
    t0 = time.perf_counter(); od = d.loss_and_grad(prog, q); torch.cuda.synchronize(); td = time.perf_counter() - t0
    op = p.loss_and_grad(prog, q); torch.cuda.synchronize()
    t0 = time.perf_counter(); op = p.loss_and_grad(prog, q); torch.cuda.synchronize(); tp = time.perf_counter() - t0
    rel = float((od["dq"] - op["dq"]).abs().max() / op["dq"].abs().max())
    print("grid%d B=%d w=%d direct-global %.3f s, pcg %.4f s, status %s, dq rel diff %.2e" % (nfree, B, plan.info["bandwidth_solver"], td, tp, od["status"].cpu().tolist(), rel))
PY
