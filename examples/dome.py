"""
Dome on a ring of supports, unconstrained (reference: examples/dome.py): 8 sides x 40 rings (320 free nodes), length and
direction goals on the radial edges, BFGS; the loss terms are then evaluated along the recorded optimisation history
(reference lines 298-306, the plot's data).
"""
import os
import sys
from math import cos, radians, sin, sqrt

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from _report import report
from dfdm_b200 import workloads
from dfdm_b200.equilibrium import EquilibriumModel, constrained_fdm, fdm
from dfdm_b200.goals import EdgeDirectionGoal, EdgeLengthGoal
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.optimization import BFGS, OptimizationRecorder


def rotate_about_axis(vector, angle, axis):
    """Rodrigues' rotation of ``vector`` by ``angle`` (radians) about ``axis`` (what compas.geometry.rotate_points does for one point)."""
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    v = np.asarray(vector, dtype=np.float64)
    return v * cos(angle) + np.cross(axis, v) * sin(angle) + axis * np.dot(axis, v) * (1.0 - cos(angle))


def build(num_sides=8, num_rings=40, offset_distance=0.01, q0_ring=-2.0, q0_cross=-0.5, pz=-0.1, length_target=0.03,
          angle_base=20.0, angle_top=30.0, angle_vector=(0.0, 0.0, 1.0)):
    """Network, loss and the goal vectors of the reference script (lines 60-218)."""
    network = workloads.dome(num_sides=num_sides, num_rings=num_rings, diameter=1.0, offset_distance=offset_distance,
                             q0_ring=q0_ring, q0_cross=q0_cross, pz=pz)
    edges_cross_rings = network.attributes["edges_cross_rings"]
    q0_scale = sqrt(network.number_of_edges()) / 2.0
    for i, cross_ring in enumerate(edges_cross_rings):            # radial force densities grow towards the supports
        for edge in cross_ring:
            network.edge_forcedensity(edge, q0_cross * q0_scale * (num_rings - i))
    goals = []
    for cross_ring in edges_cross_rings:
        for edge in cross_ring:
            goals.append(EdgeLengthGoal(edge, target=length_target, weight=1.0))
    vector_edges = []
    for i, cross_ring in enumerate(edges_cross_rings):
        angle = angle_base + (angle_top - angle_base) * (i / (num_rings - 1))
        for u, v in cross_ring:
            normal = np.cross(network.edge_vector(u, v), angle_vector)
            vector = rotate_about_axis(angle_vector, -radians(angle), normal)
            goals.append(EdgeDirectionGoal((u, v), target=vector.tolist(), weight=1.0))
            vector_edges.append((vector.tolist(), (u, v)))
    return network, Loss(SquaredError(goals=goals)), vector_edges


def main(maxiter=1000, tol=1e-3, verbose=True, **geometry):
    network, loss, _ = build(**geometry)
    networks = {"start": network, "eq": fdm(network)}
    recorder = OptimizationRecorder()
    networks["eq_g"] = constrained_fdm(network, optimizer=BFGS(disp=False), bounds=(None, None), loss=loss, maxiter=maxiter, tol=tol,
                                       callback=recorder)
    # loss and every loss term along the optimisation history: one batched evaluation per curve
    model = EquilibriumModel(network)
    history = np.asarray(recorder.history)
    curves = {}
    if len(history):
        curves[loss.name if hasattr(loss, "name") else "Loss"] = np.atleast_1d(loss(history, model))
        for term in loss.loss_terms:
            try:
                curves[term.name] = np.atleast_1d(term(model(history), model))
            except TypeError:
                curves[term.name] = np.atleast_1d(term(history, model))
    if verbose:
        for name in ("eq", "eq_g"):
            report(name, networks[name])
        for name, y in curves.items():
            print(f"{name}: {y[0]:.6f} -> {y[-1]:.6f} over {len(y)} recorded iterations")
    return networks, curves


if __name__ == "__main__":
    main()
