"""Iterations of the multigrid-preconditioned solve on different networks for several over-correction factors of the
single-visit levels (mg_over) - how much margin the default has before the cycle loses definiteness."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfdm_b200 import workloads as W, _lib
from dfdm_b200.engine import Plan, Evaluator

def delaunay(n, seed):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2)); tri = Delaunay(pts)
    pairs = np.vstack([tri.simplices[:, [0, 1]], tri.simplices[:, [1, 2]], tri.simplices[:, [2, 0]]])
    edges = np.unique(np.sort(pairs, axis=1), axis=0).astype(np.int32)
    fixed = np.zeros(n, dtype=np.uint8); fixed[np.unique(tri.convex_hull)] = 1
    xyz = np.column_stack([pts, 0.2 * np.sin(3 * pts[:, 0])]); loads = np.zeros_like(xyz); loads[:, 2] = -1.0 / n
    return edges, fixed, xyz, loads, np.ones(len(edges))

def clustered(k, m, seed):
    # k x k blobs of m x m grid nodes, blobs joined by single weak edges: near-disconnected components
    side = k * m
    edges, fixed, xyz, loads, q0 = W.grid_arrays(side)
    idx = lambda e: None
    n_side = side + 2
    q = np.ones(len(edges))
    for e, (u, v) in enumerate(edges):
        iu, ju = divmod(int(u), n_side); iv, jv = divmod(int(v), n_side)
        bu = ((iu - 1) // m, (ju - 1) // m); bv = ((iv - 1) // m, (jv - 1) // m)
        if bu != bv and 1 <= iu <= side and 1 <= iv <= side and 1 <= ju <= side and 1 <= jv <= side: q[e] = 1e-3
    return edges, fixed, xyz, loads, q

cases = {"grid 128": W.grid_arrays(128), "grid 300": W.grid_arrays(300), "delaunay 20k": delaunay(20000, 3),
         "clustered 8x8 blobs of 16x16": clustered(8, 16, 0)}
for name, (edges, fixed, xyz, loads, q0) in cases.items():
    plan = Plan(len(fixed), edges, fixed)
    rng = np.random.default_rng(1)
    for spread in (0.2, 1.6):
        q = q0[None, :] * (1.0 + spread * (rng.random((8, len(edges))) - 0.5))
        row = []
        for over in (1.2, 1.3, 1.4, 1.5):
            ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", mg_over=over)
            out = ev.forward(q, outputs=("xyz",))
            row.append("%s:%d%s" % (over, _lib.last_pcg_iterations()[0], "" if np.all(out["status"].cpu().numpy() == 0) else "!"))
        print("%-30s levels %d  q spread %.1f   " % (name, plan.info["mg_levels"], spread) + "  ".join(row), flush=True)
