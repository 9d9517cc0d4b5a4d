import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_hierarchy_cpu import stiffness, hierarchy, cycle
from dfdm_b200 import workloads as W
from dfdm_b200.engine import Plan
import scipy.sparse.linalg as spl

def pcg(A, b, M, rtol, p32, maxit=100):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy()
    if p32: p = p.astype(np.float32).astype(np.float64)
    rz = np.sum(r * z, axis=0); bb = np.sum(b * b, axis=0)
    hist = []
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / np.sum(p * Ap, axis=0)
        x += alpha * p; r -= alpha * Ap
        hist.append(np.sqrt(np.max(np.sum(r * r, axis=0) / bb)))
        if np.all(np.sum(r * r, axis=0) <= rtol ** 2 * bb):
            return it, x, hist
        z = M(r); rz_new = np.sum(r * z, axis=0)
        p = z + (rz_new / rz) * p; rz = rz_new
        if p32: p = p.astype(np.float32).astype(np.float64)
    return maxit, x, hist

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
edges, fixed, xyz, loads, q0 = W.grid_arrays(N)
plan = Plan(len(fixed), edges, fixed)
rng = np.random.default_rng(0)
for trial in range(2):
    q = q0 * (1.0 + 0.3 * (rng.random(len(edges)) - 0.5))
    K = stiffness(plan, edges, fixed, q)
    perm, levels, inv = hierarchy(plan, K)
    A0 = levels[0][0]
    # realistic right-hand side: the loads / boundary term, or random
    b = rng.standard_normal((A0.shape[0], 3)) if trial else np.tile((-1.0 / N**2) * np.ones((A0.shape[0], 1)), (1, 3)) + 1e-3 * rng.standard_normal((A0.shape[0], 3))
    w_levels = min(3, len(levels) - 1)
    M = lambda r: cycle(levels, inv, 0, r.astype(np.float32).astype(np.float64), w_levels).astype(np.float32).astype(np.float64)
    xs = spl.spsolve(A0.tocsc(), b)
    for rtol in (1e-12, 5e-9):
        for p32 in (False, True):
            it, x, hist = pcg(A0, b, M, rtol, p32)
            err = np.max(np.abs(x - xs)) / np.max(np.abs(xs))
            true_res = np.sqrt(np.max(np.sum((b - A0 @ x) ** 2, axis=0) / np.sum(b * b, axis=0)))
            print("N %d trial %d rtol %.0e p32 %d: iters %d, x err %.2e, true residual %.2e (recurrence %.2e)" % (N, trial, rtol, p32, it, err, true_res, hist[-1]))
