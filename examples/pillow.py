"""
Pillow: a square quad-grid shell on boundary supports (reference: examples/pillow.py).
Horizontal-projection goals on every free node; length, force and curvature constraints; SLSQP.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from _report import report
from dfdm_b200 import workloads
from dfdm_b200.constraints import NetworkEdgesForceConstraint, NetworkEdgesLengthConstraint, NodeCurvatureConstraint
from dfdm_b200.equilibrium import constrained_fdm, fdm
from dfdm_b200.goals import NodeLineGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import SLSQP


def main(divisions=6, maxiter=100, verbose=True, seed=2):
    network = workloads.pillow(divisions, q0=-2.0, pz=-100.0)
    rng = np.random.default_rng(seed)
    for edge in network.edges():
        network.edge_forcedensity(edge, -2.0 + 0.1 * (rng.random() - 0.5))
    networks = {"loaded": fdm(network)}

    goals = []
    for node in network.nodes_free():
        xyz = network.node_coordinates(node)
        goals.append(NodeLineGoal(node, target=(xyz, [xyz[0], xyz[1], xyz[2] + 1.0]), weight=1.0))
    loss = Loss(SquaredError(goals=goals))

    average_length = np.mean([network.edge_length(*edge) for edge in network.edges()])
    constraints = [NetworkEdgesLengthConstraint(bound_low=0.5 * average_length, bound_up=3.0 * average_length),
                   NetworkEdgesForceConstraint(bound_low=-100.0, bound_up=-1.0)]
    mid = divisions // 2
    width = divisions + 1
    for i in range(1, divisions):                       # the middle polyedge, as in pillow.py:244-266
        key = mid * width + i
        polygon = [key + 1, key + width, key - 1, key - width]
        constraints.append(NodeCurvatureConstraint(key, polygon, bound_low=-100.0, bound_up=-0.3))

    networks["uncstr_opt"] = constrained_fdm(network, optimizer=SLSQP(disp=False), bounds=(None, None), loss=loss, maxiter=maxiter)
    networks["cstr_opt"] = constrained_fdm(network, optimizer=SLSQP(disp=False), bounds=(None, None), loss=loss,
                                           constraints=constraints, maxiter=maxiter)
    stats = {}
    for name, net in networks.items():
        stats[name] = report(name, net) if verbose else {"loadpath": net.loadpath()}
    return networks, stats


if __name__ == "__main__":
    main()
