"""
Vault on two rows of supports ("monkey_vault"; reference: examples/pringle.py): reaction-force goals on
the supports, plane goals on the free nodes, length goals on the cross edges; SLSQP with bounds; the
optimisation history is recorded, written to JSON and replayed (reference: examples/history.py).
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from _report import report
from dfdm_b200 import workloads
from dfdm_b200.equilibrium import EquilibriumModel, constrained_fdm
from dfdm_b200.goals import EdgeLengthGoal, NodePlaneGoal, NodeResidualForceGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import SLSQP, OptimizationRecorder


def main(maxiter=200, verbose=True):
    network = workloads.pringle(num_u=10, num_v=9, q_init=-0.25, pz=-0.1)
    arches = network.attributes["arches"]
    rz_min, rz_max = 0.45, 2.0
    num_steps = (len(arches) - 1) / 2.0
    rzs = [rz_min + i * (rz_max - rz_min) / num_steps for i in range(int(num_steps) + 1)]
    rzs = rzs + rzs[0:-1][::-1]

    goals = []
    for rz, arch in zip(rzs, arches):
        goals.append(NodeResidualForceGoal(arch[0], target=rz, weight=100.0))
        goals.append(NodeResidualForceGoal(arch[-1], target=rz, weight=100.0))
    for node in network.nodes_free():
        goals.append(NodePlaneGoal(node, target=(network.node_coordinates(node), [1.0, 0.0, 0.0]), weight=10.0))
    for edge in network.attributes["cross_edges"]:
        goals.append(EdgeLengthGoal(tuple(edge), target=network.edge_length(*edge), weight=1.0))
    loss = Loss(SquaredError(goals))

    recorder = OptimizationRecorder()
    c_network = constrained_fdm(network, optimizer=SLSQP(disp=False), loss=loss, bounds=(-5.0, -0.1), maxiter=maxiter, tol=1e-9,
                                callback=recorder)
    # history round trip + replay of the loss along the optimisation path
    path = os.path.join(tempfile.gettempdir(), "monkey_vault_history.json")
    recorder.to_json(path)
    history = OptimizationRecorder.from_json(path).history
    model = EquilibriumModel(network)
    losses = loss(np.asarray(history), model)               # one batched evaluation of every recorded iterate
    if verbose:
        report("monkey_vault", c_network)
        print(f"iterations recorded: {len(history)}; loss {losses[0]:.6f} -> {losses[-1]:.6f}")
        reactions = [np.linalg.norm(c_network.node_residual(arch[0])) for arch in arches]
        print("support reactions:", np.round(reactions, 3), "targets:", np.round(rzs, 3))
    return c_network, losses


if __name__ == "__main__":
    main()
