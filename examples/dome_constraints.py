"""
Dome on a ring of supports (reference: examples/dome_constraints.py): horizontal-projection goals,
edge-angle constraints on the radial edges between consecutive rings, SLSQP.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from _report import report
from dfdm_b200 import workloads
from dfdm_b200.constraints import EdgeVectorAngleConstraint
from dfdm_b200.equilibrium import EquilibriumModel, constrained_fdm, fdm
from dfdm_b200.goals import NodeLineGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import SLSQP


def main(maxiter=100, angle_min=10.0, angle_max=30.0, verbose=True):
    network = workloads.dome(num_sides=16, num_rings=6, q0_ring=-2.0, q0_cross=-0.5, pz=-0.1)
    goals = []
    for node in network.nodes_free():
        xyz = network.node_coordinates(node)
        goals.append(NodeLineGoal(node, target=(xyz, [xyz[0], xyz[1], xyz[2] + 1.0])))
    loss = Loss(SquaredError(goals=goals))
    constraint_angles = [EdgeVectorAngleConstraint(tuple(edge), vector=[0.0, 0.0, 1.0], bound_low=angle_min, bound_up=angle_max)
                         for edge in network.attributes["edges_cross"]]

    networks = {"eq": fdm(network)}
    networks["eq_g"] = constrained_fdm(network, optimizer=SLSQP(disp=False), bounds=(None, None), loss=loss, maxiter=maxiter)
    networks["eq_g_c"] = constrained_fdm(network, optimizer=SLSQP(disp=False), bounds=(None, None), loss=loss,
                                         constraints=constraint_angles, maxiter=maxiter)
    angles = {}
    for name, net in networks.items():
        model = EquilibriumModel(net)
        eqstate = model(np.array(net.edges_forcedensities()))
        angles[name] = [c.constraint(eqstate, model) for c in constraint_angles]
        if verbose:
            report(name, net)
            print(f"Angles\t\tMin: {round(min(angles[name]), 3)}\tMax: {round(max(angles[name]), 3)}")
    return networks, angles


if __name__ == "__main__":
    main()
