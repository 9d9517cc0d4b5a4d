"""
Monkey saddle, unconstrained (reference: examples/monkey_saddle.py): length goals on every edge, reaction-force goals
on the supports, the reference's PredictionError term over those same goals, an L2 regulariser; SLSQP with bounds.
The problem definition, the optimisation history and the optimised network are written as JSON (the files
examples/history.py replays).
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from _report import report
from dfdm_b200 import workloads
from dfdm_b200.equilibrium import EquilibriumModel, constrained_fdm, fdm
from dfdm_b200.goals import EdgeLengthGoal, NodeResidualForceGoal
from dfdm_b200.losses import L2Regularizer, Loss, PredictionError, SquaredError
from dfdm_b200.optimization import SLSQP, OptimizationRecorder

NAME = "monkey_saddle"


def build(n=3, q0=-1.0, weight_residual=10.0, weight_length=1.0, alpha=0.1, alpha_lp=0.01):
    network0, targets = workloads.monkey_saddle(n=n, q0=q0, load=(0.0, 0.0, -1.0), rmin=2.0, rmax=10.0, r_exp=1.0)
    goals_a = [EdgeLengthGoal(edge, network0.edge_length(*edge), weight=weight_length) for edge in network0.edges()]
    goals_b = [NodeResidualForceGoal(key, targets[key], weight=weight_residual) for key in network0.nodes_supports()]
    loss = Loss(SquaredError(goals_a, alpha=1.0, name="EdgeLengthGoal"),
                SquaredError(goals_b, alpha=1.0, name="ReactionForceGoal"),
                PredictionError(goals_b, alpha=alpha_lp, name="LoadPathGoal"),     # as in the reference (line 176): over goals_b
                L2Regularizer(alpha=alpha))
    return network0, loss, targets


def main(maxiter=500, tol=1e-3, out_dir=None, verbose=True, **kwargs):
    network0, loss, targets = build(**kwargs)
    out_dir = out_dir or os.path.join(tempfile.gettempdir(), "dfdm_b200_examples")
    os.makedirs(out_dir, exist_ok=True)
    network0.to_json(os.path.join(out_dir, f"{NAME}_base.json"))
    network0 = fdm(network0)
    recorder = OptimizationRecorder()
    network = constrained_fdm(network0, optimizer=SLSQP(disp=False), loss=loss, bounds=(-20.0, -0.01), maxiter=maxiter, tol=tol,
                              callback=recorder)
    recorder.to_json(os.path.join(out_dir, f"{NAME}_history.json"))
    network.to_json(os.path.join(out_dir, f"{NAME}_optimized.json"))
    model = EquilibriumModel(network)
    history = np.asarray(recorder.history)
    curves = {"Loss": np.atleast_1d(loss(history, model))} if len(history) else {}
    for term in loss.loss_terms if len(history) else ():
        try:
            curves[term.name] = np.atleast_1d(term(model(history), model))
        except TypeError:
            curves[term.name] = np.atleast_1d(term(history, model))
    reactions = {key: float(np.linalg.norm(network.node_residual(key))) for key in network.nodes_supports()}
    if verbose:
        report("monkey_saddle start", network0)
        report("monkey_saddle optimised", network)
        for name, y in curves.items():
            print(f"{name}: {y[0]:.6f} -> {y[-1]:.6f} over {len(y)} recorded iterations")
        print("files written to", out_dir)
    return network, curves, reactions, targets, out_dir


if __name__ == "__main__":
    main()
