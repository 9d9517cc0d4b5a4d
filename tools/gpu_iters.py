import sys, numpy as np
sys.path.insert(0,'.')
from dfdm_b200 import workloads as W, _lib
from dfdm_b200.engine import Plan, Evaluator
for N in (64, 256):
    edges,fixed,xyz,loads,q0=W.grid_arrays(N)
    plan=Plan(len(fixed),edges,fixed)
    q=W.sample_q(q0,256,seed=6,spread=0.2,relative=True)
    for B in (1,2,256):
        ev=Evaluator(plan,loads,xyz[plan.fixed],solver="pcg",pcg_check_every=1)
        out=ev.forward(q[:B],outputs=("xyz",))
        print("N",N,"B",B,"iters",_lib.last_pcg_iterations(), "levels", plan.info["mg_levels"], plan.info["mg_n_coarse"], plan.info["mg_n_coarsest"])
