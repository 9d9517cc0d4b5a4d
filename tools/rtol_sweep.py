"""
Calibration of the PCG stop tests at the headline configuration (grid256): error of coordinates, loss and gradient against
the sparse oracle (SuperLU on the same K) as a function of pcg_rtol (forward solve) and pcg_rtol_adjoint, with the
iteration counts.  Run on a GPU box:  python tools/rtol_sweep.py [nfree] > gpurun_out/rtol_sweep.json
The defaults in include/dfdm_b200.h are the loosest values that keep a 10x margin to the 1e-9 / 1e-7 contract.
"""
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfdm_b200 import workloads as W, _lib           # noqa: E402
from dfdm_b200.engine import Plan, GoalProgram, Evaluator   # noqa: E402
from oracle import fdm_oracle as O                    # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    nfree = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    edges, fixed, xyz, loads, q0 = W.grid_arrays(nfree)
    plan = Plan(len(fixed), edges, fixed)
    B = 8
    q = W.sample_q(q0, B, seed=6, spread=0.2, relative=True)
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    designs = (0, 3, 7)
    ref = {b: sm.loss_and_grad_edge_length(q[b], lengths0, 1.0, 0.1) for b in designs}
    rows = []
    pairs = [(1e-12, 1e-12), (1e-11, 1e-11), (1e-10, 1e-10), (1e-9, 1e-9), (1e-8, 1e-8), (1e-7, 1e-7),
             (1e-11, 1e-9), (1e-11, 1e-8), (1e-11, 1e-7), (1e-10, 1e-8), (1e-10, 1e-7), (1e-10, 1e-6), (1e-12, 1e-6),
             (2e-11, 1e-7), (3e-11, 1e-7), (5e-11, 1e-7), (3e-11, 3e-7), (3e-11, 1e-6)]
    for rf, ra in pairs:
        ev = Evaluator(plan, loads, xyz[plan.fixed], pcg_rtol=rf, pcg_rtol_adjoint=ra)
        out = ev.loss_and_grad(prog, q, outputs=("xyz",))
        it = _lib.last_pcg_iterations()
        X, dq, loss = out["xyz"].cpu().numpy(), out["dq"].cpu().numpy(), out["loss"].cpu().numpy()
        ex = max(rel(X[b], ref[b][2].xyz) for b in designs)
        eg = max(rel(dq[b], ref[b][1]) for b in designs)
        el = max(abs(loss[b] - ref[b][0]) / abs(ref[b][0]) for b in designs)
        rows.append({"pcg_rtol": rf, "pcg_rtol_adjoint": ra, "iterations": it, "xyz_rel_err": ex, "loss_rel_err": el, "dq_rel_err": eg})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
