"""CPU prototype of the preconditioner cycle (NumPy / SciPy): iteration counts of PCG for cycle variants.
Not part of the product; used to decide what is worth writing as kernels."""
import sys
import numpy as np
import scipy.sparse as sp

def grid_K(N, seed=6):
    # N x N free nodes, boundary fixed, 4-neighbour edges, q = 1 + 0.2 (u - 0.5)
    rng = np.random.default_rng(seed)
    idx = -np.ones((N + 2, N + 2), dtype=np.int64)
    idx[1:-1, 1:-1] = np.arange(N * N).reshape(N, N)
    rows, cols, vals = [], [], []
    diag = np.zeros(N * N)
    for (di, dj) in ((0, 1), (1, 0)):
        a = idx[:N + 2 - di, :N + 2 - dj].ravel()
        b = idx[di:, dj:].ravel()
        q = 1.0 + 0.2 * (rng.random(len(a)) - 0.5)
        for u, v in ((a, b), (b, a)):
            m = u >= 0
            np.add.at(diag, u[m], q[m])
        m = (a >= 0) & (b >= 0)
        rows += [a[m], b[m]]; cols += [b[m], a[m]]; vals += [-q[m], -q[m]]
    rows.append(np.arange(N * N)); cols.append(np.arange(N * N)); vals.append(diag)
    K = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N * N, N * N))
    return K

def block_agg(n_side):
    # 2 x 2 blocks of an n_side x n_side grid
    i, j = np.divmod(np.arange(n_side * n_side), n_side)
    nc = (n_side + 1) // 2
    return (i // 2) * nc + (j // 2), nc

class Level: pass

def build(K, N, coarsest=96):
    levels = []
    A = K; side = N
    while A.shape[0] > coarsest:
        L = Level(); L.A = A.tocsr(); L.d = A.diagonal(); L.side = side
        agg, nc = block_agg(side)
        L.P = sp.csr_matrix((np.ones(len(agg)), (np.arange(len(agg)), agg)), shape=(len(agg), nc * nc))
        # colours (red / black) of the grid for Gauss-Seidel
        i, j = np.divmod(np.arange(side * side), side)
        L.red = ((i + j) % 2 == 0)
        levels.append(L)
        A = (L.P.T @ A @ L.P).tocsr(); side = nc
    C = Level(); C.A = A; C.inv = np.linalg.inv(A.toarray()); levels.append(C)
    return levels

def smooth_jacobi(L, b, x, omega):
    if x is None: return omega * b / L.d[:, None]
    return x + omega * (b - L.A @ x) / L.d[:, None]

def smooth_gs(L, b, x, order):
    # coloured Gauss-Seidel sweep; order = (first mask, second mask)
    if x is None: x = np.zeros_like(b)
    for mask in order:
        r = b - L.A @ x
        x = x.copy(); x[mask] += r[mask] / L.d[mask, None]
    return x

def cycle(levels, l, b, opt):
    if l == len(levels) - 1:
        return levels[l].inv @ b
    L = levels[l]
    nu = opt.get("nu", 1) if l >= opt.get("nu_from", 0) else 1
    x = None
    for _ in range(nu):
        if opt["smoother"] == "jacobi": x = smooth_jacobi(L, b, x, opt["omega"])
        else: x = smooth_gs(L, b, x, (L.red, ~L.red))
    visits = 2 if (1 <= l + 1 <= opt["w_levels"]) else 1     # the coarse problem of level l is level l+1
    for v in range(visits):
        r = b - L.A @ x
        e = cycle(levels, l + 1, L.P.T @ r, opt)
        x = x + opt["over"] * (L.P @ e)
    for _ in range(nu):
        if opt["smoother"] == "jacobi": x = smooth_jacobi(L, b, x, opt["omega"])
        else: x = smooth_gs(L, b, x, (~L.red, L.red))
    return x

def pcg(K, b, M, rtol=1e-12, maxit=500):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy()
    rz = np.sum(r * z, axis=0); b2 = np.sum(b * b, axis=0)
    for it in range(1, maxit + 1):
        Ap = K @ p
        alpha = rz / np.sum(p * Ap, axis=0)
        x += alpha * p; r -= alpha * Ap
        if np.all(np.sum(r * r, axis=0) <= rtol ** 2 * b2): return x, it
        z = M(r); rz_new = np.sum(r * z, axis=0)
        p = z + (rz_new / rz) * p; rz = rz_new
    return x, maxit

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    K = grid_K(N)
    levels = build(K, N)
    print("levels", [l.A.shape[0] for l in levels])
    rng = np.random.default_rng(0)
    b = rng.standard_normal((N * N, 3))
    variants = {
        "jacobi V": dict(smoother="jacobi", omega=0.8, over=1.7, w_levels=0),
        "jacobi W2 (current)": dict(smoother="jacobi", omega=0.8, over=1.3, w_levels=2),
        "jacobi W2 nu=2": dict(smoother="jacobi", omega=0.8, over=1.3, w_levels=2, nu=2),
        "jacobi W2 nu=2 on coarse only": dict(smoother="jacobi", omega=0.8, over=1.3, w_levels=2, nu=2, nu_from=1),
        "rbgs V over 1.0": dict(smoother="gs", over=1.0, w_levels=0),
        "rbgs V over 1.5": dict(smoother="gs", over=1.5, w_levels=0),
        "rbgs W2 over 1.0": dict(smoother="gs", over=1.0, w_levels=2),
        "rbgs W2 over 1.3": dict(smoother="gs", over=1.3, w_levels=2),
        "rbgs W2 over 1.5": dict(smoother="gs", over=1.5, w_levels=2),
    }
    for name, opt in variants.items():
        M = lambda r, opt=opt: cycle(levels, 0, r.astype(np.float32).astype(np.float64), opt)
        x, it = pcg(K, b, M)
        err = np.max(np.abs(K @ x - b)) / np.max(np.abs(b))
        print("%-34s iterations %3d   residual %.1e" % (name, it, err))
