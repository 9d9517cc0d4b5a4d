"""
Monkey saddle (reference: examples/monkey_saddle_constraints.py): prescribed reaction-force magnitudes
on the supported boundary, edge-length constraints, trust-constr.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from _report import report
from dfdm_b200 import workloads
from dfdm_b200.constraints import NetworkEdgesLengthConstraint
from dfdm_b200.equilibrium import constrained_fdm, fdm
from dfdm_b200.goals import NodeResidualForceGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import OptimizationRecorder, TrustRegionConstrained


def main(maxiter=60, tol=1e-2, verbose=True):
    network0, targets = workloads.monkey_saddle(n=3, q0=-1.0, load=(0.0, 0.0, -1.0))
    network0 = fdm(network0)
    goals = [NodeResidualForceGoal(key, targets[key]) for key in network0.nodes_supports()]
    loss = Loss(SquaredError(goals, name="ReactionForceGoal"))
    constraints = [NetworkEdgesLengthConstraint(bound_low=0.5, bound_up=4.5)]
    recorder = OptimizationRecorder()
    network = constrained_fdm(network0, optimizer=TrustRegionConstrained(disp=False), loss=loss, bounds=(-20.0, -0.01),
                              constraints=constraints, maxiter=maxiter, tol=tol, callback=recorder)
    reactions = {key: float(np.linalg.norm(network.node_residual(key))) for key in network.nodes_supports()}
    error = max(abs(reactions[key] - targets[key]) for key in reactions)
    if verbose:
        report("monkey_saddle start", network0)
        report("monkey_saddle optimised", network)
        print(f"largest deviation of a reaction magnitude from its target: {error:.4f}")
    return network, reactions, targets


if __name__ == "__main__":
    main()
