"""Prototype: smoothed aggregation (prolongator smoothing) versus the plain-aggregation cycle, iteration counts only."""
import sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from mg_proto import grid_K, pcg

def block_agg(side, bs):
    i, j = np.divmod(np.arange(side * side), side)
    nc = (side + bs - 1) // bs
    return (i // bs) * nc + (j // bs), nc

class Level: pass

def build(K, N, bs, smooth_P, omega_p=2.0 / 3.0, coarsest=96, bs_coarse=None):
    levels = []; A = K.tocsr(); side = N
    first = True
    while A.shape[0] > coarsest:
        L = Level(); L.A = A; L.d = A.diagonal()
        b_ = bs if first or bs_coarse is None else bs_coarse
        agg, nc = block_agg(side, b_)
        Pt = sp.csr_matrix((np.ones(len(agg)), (np.arange(len(agg)), agg)), shape=(len(agg), nc * nc))
        if smooth_P:
            Dinv = sp.diags(1.0 / L.d)
            L.P = (Pt - omega_p * (Dinv @ (A @ Pt))).tocsr()
        else:
            L.P = Pt
        levels.append(L)
        A = (L.P.T @ A @ L.P).tocsr(); side = nc; first = False
    C = Level(); C.A = A; C.inv = np.linalg.inv(A.toarray()); levels.append(C)
    return levels

def cycle(levels, l, b, opt):
    if l == len(levels) - 1: return levels[l].inv @ b
    L = levels[l]; om = opt["omega"]
    x = om * b / L.d[:, None]
    for _ in range(opt.get("nu", 1) - 1): x = x + om * (b - L.A @ x) / L.d[:, None]
    visits = 2 if (1 <= l + 1 <= opt["w_levels"]) else 1
    for v in range(visits):
        r = b - L.A @ x
        x = x + opt["over"] * (L.P @ cycle(levels, l + 1, L.P.T @ r, opt))
    for _ in range(opt.get("nu", 1)): x = x + om * (b - L.A @ x) / L.d[:, None]
    return x

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    K = grid_K(N)
    b = np.random.default_rng(0).standard_normal((N * N, 3))
    tests = [
        ("plain 2x2, W2 over 1.3 (current)", dict(bs=2, smooth_P=False), dict(omega=0.8, over=1.3, w_levels=2)),
        ("SA 2x2, V", dict(bs=2, smooth_P=True), dict(omega=0.8, over=1.0, w_levels=0)),
        ("SA 3x3, V", dict(bs=3, smooth_P=True), dict(omega=0.8, over=1.0, w_levels=0)),
        ("SA 3x3, V, nu=2", dict(bs=3, smooth_P=True), dict(omega=0.8, over=1.0, w_levels=0, nu=2)),
        ("SA 3x3, W2", dict(bs=3, smooth_P=True), dict(omega=0.8, over=1.0, w_levels=2)),
        ("SA 4x4, V", dict(bs=4, smooth_P=True), dict(omega=0.8, over=1.0, w_levels=0)),
        ("SA first level only (2x2), plain below, W2 1.3", "hybrid", dict(omega=0.8, over=1.3, w_levels=2)),
    ]
    for name, bopt, copt in tests:
        if bopt == "hybrid":
            levels = build(K, N, 2, False)
            L = levels[0]
            Dinv = sp.diags(1.0 / L.d)
            L.P = (L.P - (2.0 / 3.0) * (Dinv @ (L.A @ L.P))).tocsr()
            # rebuild below with the smoothed first-level Galerkin operator, plain aggregation
            A = (L.P.T @ L.A @ L.P).tocsr()
            side = (N + 1) // 2
            rest = build(A, side, 2, False)
            levels = [L] + rest
            over_first = 1.0
        else:
            levels = build(K, N, **bopt)
        nnz = [l.A.nnz for l in levels]
        def M(r, levels=levels, copt=copt, hybrid=(bopt == "hybrid")):
            if not hybrid: return cycle(levels, 0, r, copt)
            # first level with over 1.0 (smoothed P), the rest as configured
            L = levels[0]; om = copt["omega"]
            x = om * r / L.d[:, None]
            for v in range(2):
                rr = r - L.A @ x
                x = x + 1.0 * (L.P @ cycle(levels, 1, L.P.T @ rr, dict(copt, w_levels=copt["w_levels"] + 0)))
            return x + om * (r - L.A @ x) / L.d[:, None]
        x, it = pcg(K, b, M)
        print("%-50s iterations %3d  rows %s  nnz %s  op.complexity %.2f" % (name, it, [l.A.shape[0] for l in levels], nnz, sum(nnz) / nnz[0]))
