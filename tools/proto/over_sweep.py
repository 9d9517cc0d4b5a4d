"""Prototype sweep: per-level over-correction factors and smoother damping of the plain-aggregation W-cycle."""
import sys, itertools
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from mg_proto import grid_K, build, pcg

def cycle(levels, l, b, opt):
    if l == len(levels) - 1: return levels[l].inv @ b
    L = levels[l]; om = opt["omega"][min(l, len(opt["omega"]) - 1)]
    x = om * b / L.d[:, None]
    visits = 2 if (1 <= l + 1 <= opt["w_levels"]) else 1
    ov = opt["over"][min(l, len(opt["over"]) - 1)]
    for v in range(visits):
        r = b - L.A @ x
        x = x + ov * (L.P @ cycle(levels, l + 1, L.P.T @ r, opt))
    return x + om * (b - L.A @ x) / L.d[:, None]

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = grid_K(N); levels = build(K, N)
b = np.random.default_rng(0).standard_normal((N * N, 3))
for w in (2, 3):
    for o0, o1, o2 in itertools.product((1.3, 1.5, 1.7), (1.2, 1.3, 1.4), (1.3, 1.7)):
        opt = dict(omega=[0.8], over=[o0, o1, o2], w_levels=w)
        x, it = pcg(K, b, lambda r: cycle(levels, 0, r, opt))
        print("w_levels %d over %.1f %.1f %.1f  iterations %d" % (w, o0, o1, o2, it), flush=True)
for om0, om1 in itertools.product((0.7, 0.8, 0.9), (0.7, 0.8, 0.9)):
    opt = dict(omega=[om0, om1], over=[1.3], w_levels=2)
    x, it = pcg(K, b, lambda r: cycle(levels, 0, r, opt))
    print("omega %.1f %.1f  iterations %d" % (om0, om1, it), flush=True)
