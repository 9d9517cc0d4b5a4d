"""
Arches whose support reactions point along prescribed directions (reference: examples/arches.py).
One SLSQP run per target direction, bounds (-inf, 0], tol 1e-9.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from dfdm_b200 import workloads
from dfdm_b200.equilibrium import constrained_fdm
from dfdm_b200.goals import NodeResidualDirectionGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import SLSQP


def main(num_segments=10, vertical_comps=(0.1, 0.2, 0.4, 0.8, 1.6, 3.2), verbose=True):
    network = workloads.arch(num_segments, length=5.0, q_init=-1.0, pz=-0.1)
    heights = []
    for comp in vertical_comps:
        goals = [NodeResidualDirectionGoal(0, target=[-1.0, 0.0, -comp]),
                 NodeResidualDirectionGoal(num_segments, target=[1.0, 0.0, -comp])]
        arch = constrained_fdm(network, optimizer=SLSQP(disp=False), loss=Loss(SquaredError(goals)), bounds=(-np.inf, 0.0),
                               maxiter=200, tol=1e-9)
        res = np.asarray(arch.node_residual(0))
        height = max(arch.node_attribute(k, "z") for k in arch.nodes())
        heights.append(height)
        if verbose:
            print(f"vertical component {comp}: rise {height:.4f}, reaction direction {res / np.linalg.norm(res)}")
    return heights


if __name__ == "__main__":
    main()
