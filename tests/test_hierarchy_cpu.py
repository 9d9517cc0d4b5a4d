"""The multigrid hierarchy the plan compiles, judged on the CPU: a NumPy / SciPy restatement of the preconditioned solve
(csrc/pcg.cu: damped-Jacobi smoothing, piecewise-constant transfers, Galerkin operators, W-cycle on the first coarse
levels, over-corrected coarse corrections, exact coarsest solve) runs on the aggregates the plan exports and must
converge in the number of iterations the GPU path takes.  No GPU involved: this pins the integer work of plan.cpp."""
import numpy as np
import pytest
import scipy.sparse as sp

from dfdm_b200 import _lib
from dfdm_b200 import workloads as W
from dfdm_b200.engine import Plan


def stiffness(plan, edges, fixed, q):
    """K = C_f^T Q C_f in the plan's free order (reference: model.py:53)."""
    n, E = len(fixed), len(edges)
    C = sp.csr_matrix((np.r_[-np.ones(E), np.ones(E)], (np.r_[np.arange(E), np.arange(E)], np.r_[edges[:, 0], edges[:, 1]])), shape=(E, n))
    Cf = C[:, plan.free]
    return (Cf.T @ sp.diags(q) @ Cf).tocsr()


def hierarchy(plan, K):
    perm = plan.array(_lib.ARR_PCG_PERM)
    A = K[perm][:, perm].tocsr()
    levels = []
    for l in range(plan.info["mg_levels"]):
        agg = plan.array(_lib.ARR_MG_AGG + l)
        assert len(agg) == A.shape[0]
        P = sp.csr_matrix((np.ones(len(agg)), (np.arange(len(agg)), agg)), shape=(len(agg), int(agg.max()) + 1))
        levels.append((A, A.diagonal(), P))
        A = (P.T @ A @ P).tocsr()
    return perm, levels, np.linalg.inv(A.toarray())


def cycle(levels, inv, l, b, w_levels, omega=0.8, over=1.3, over_w=1.7):
    if l == len(levels):
        return inv @ b
    A, d, P = levels[l]
    twice = l + 1 <= w_levels and l + 1 < len(levels) + 0          # the coarse problem (level l+1) is a W level
    s = over_w if twice else over
    x = omega * b / d[:, None]
    for _ in range(2 if twice else 1):
        r = b - A @ x
        x = x + s * (P @ cycle(levels, inv, l + 1, P.T @ r, w_levels, omega, over, over_w))
    return x + omega * (b - A @ x) / d[:, None]


def pcg_iterations(A, b, M, rtol=1e-12, maxit=300):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy()
    rz = np.sum(r * z, axis=0); bb = np.sum(b * b, axis=0)
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / np.sum(p * Ap, axis=0)
        x += alpha * p; r -= alpha * Ap
        if np.all(np.sum(r * r, axis=0) <= rtol ** 2 * bb):
            return it, x
        z = M(r); rz_new = np.sum(r * z, axis=0)
        p = z + (rz_new / rz) * p; rz = rz_new
    return maxit, x


def delaunay(n, seed):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    tri = Delaunay(pts)
    pairs = np.vstack([tri.simplices[:, [0, 1]], tri.simplices[:, [1, 2]], tri.simplices[:, [2, 0]]])
    edges = np.unique(np.sort(pairs, axis=1), axis=0).astype(np.int32)
    fixed = np.zeros(n, dtype=np.uint8)
    fixed[np.unique(tri.convex_hull)] = 1
    return edges, fixed


@pytest.mark.parametrize("mesh, bound", [("grid96", 22), ("delaunay6000", 28)])
def test_plan_hierarchy_converges_like_the_gpu_path(mesh, bound):
    if mesh == "grid96":
        edges, fixed, _, _, _ = W.grid_arrays(96)
    else:
        edges, fixed = delaunay(6000, 21)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.info["mg_levels"] >= 3
    rng = np.random.default_rng(0)
    q = 1.0 + 0.2 * (rng.random(len(edges)) - 0.5)
    K = stiffness(plan, edges, fixed, q)
    perm, levels, inv = hierarchy(plan, K)
    # aggregates: at most 4 rows each, consecutive in the hierarchical order (k_mg_down and k_mg_bottom rely on it)
    for _, _, P in levels:
        agg = P.indices
        assert np.all(np.diff(agg) >= 0) and np.bincount(agg).max() <= 4
    A0 = levels[0][0]
    b = rng.standard_normal((A0.shape[0], 3))
    w_levels = min(3, len(levels) - 1)
    M = lambda r: cycle(levels, inv, 0, r.astype(np.float32).astype(np.float64), w_levels)
    iters, x = pcg_iterations(A0, b, M)
    assert iters <= bound, iters
    assert np.max(np.abs(A0 @ x - b)) <= 1e-10 * np.max(np.abs(b))
    # the plain V-cycle needs visibly more iterations on the same hierarchy: the W levels are doing their job
    iters_v, _ = pcg_iterations(A0, b, lambda r: cycle(levels, inv, 0, r, 0, over=1.7, over_w=1.7))
    assert iters_v > iters


def pcg_as_the_kernels_store_it(A, b, M, rtol=1e-12, maxit=300, window=8):
    """The recurrence of csrc/pcg.cu with its storage choices: the search direction rounded to single precision every time it is
    stored, the multigrid cycle fed a single-precision copy of r, and x updated once per window of iterations from the kept
    directions and step lengths; K p, the dot products, x and r in double precision."""
    f32 = lambda v: v.astype(np.float32).astype(np.float64)
    x = np.zeros_like(b); r = b.copy(); z = M(f32(r)); p = f32(z)
    rz = np.sum(r * z, axis=0); bb = np.sum(b * b, axis=0)
    kept = []
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / np.sum(p * Ap, axis=0)
        kept.append((alpha, p))
        r -= alpha * Ap
        done = np.all(np.sum(r * r, axis=0) <= rtol ** 2 * bb)
        if len(kept) == window or done:
            for a_k, p_k in kept:
                x += a_k * p_k
            kept = []
        if done:
            return it, x
        z = M(f32(r)); rz_new = np.sum(r * z, axis=0)
        p = f32(z + (rz_new / rz) * p); rz = rz_new
    return maxit, x


def test_single_precision_directions_and_windowed_x_updates_change_nothing():
    """What DESIGN.md section 4.2 claims for the storage choices of the PCG kernels, on the CPU: the same iteration count as
    the all-double recurrence and the same solution to rounding, because x and r move by the same stored vector."""
    edges, fixed, _, _, _ = W.grid_arrays(64)
    plan = Plan(len(fixed), edges, fixed)
    rng = np.random.default_rng(4)
    q = 1.0 + 0.3 * (rng.random(len(edges)) - 0.5)
    K = stiffness(plan, edges, fixed, q)
    perm, levels, inv = hierarchy(plan, K)
    A0 = levels[0][0]
    b = rng.standard_normal((A0.shape[0], 3))
    w_levels = min(3, len(levels) - 1)
    M = lambda r: cycle(levels, inv, 0, r.astype(np.float32).astype(np.float64), w_levels).astype(np.float32).astype(np.float64)
    it_ref, x_ref = pcg_iterations(A0, b, M)
    it_new, x_new = pcg_as_the_kernels_store_it(A0, b, M)
    assert it_new == it_ref and it_ref > 8                      # more than one window
    assert np.max(np.abs(x_new - x_ref)) <= 1e-11 * np.max(np.abs(x_ref))
    assert np.max(np.abs(A0 @ x_new - b)) <= 1e-10 * np.max(np.abs(b))
