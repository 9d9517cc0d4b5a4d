"""
Replay of a recorded optimisation (reference: examples/history.py:93-151 without the viewer): read the problem
definition and the history written by examples/monkey_saddle.py, and step a network through the recorded force
densities - the loop the reference's animation callback runs (model(q), network_update) - collecting what its
viewer objects would show (coordinates, reaction forces, load path).  The whole history is also evaluated in ONE
batched call and compared with the frame-by-frame loop.
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from dfdm_b200.datastructures import FDNetwork
from dfdm_b200.equilibrium import EquilibriumModel, fdm, network_update
from dfdm_b200.optimization import OptimizationRecorder

NAME = "monkey_saddle"


def main(in_dir=None, verbose=True):
    in_dir = in_dir or os.path.join(tempfile.gettempdir(), "dfdm_b200_examples")
    network0 = FDNetwork.from_json(os.path.join(in_dir, f"{NAME}_base.json"))
    model = EquilibriumModel(network0)
    network = fdm(network0)
    recorder = OptimizationRecorder.from_json(os.path.join(in_dir, f"{NAME}_history.json"))
    frames = []
    for q in recorder.history:                              # the reference's per-frame callback (lines 146-151)
        eqstate = model(np.array(q))
        network_update(network, eqstate)
        reaction = max(float(np.linalg.norm(network.node_residual(key))) for key in network.nodes_supports())
        frames.append({"loadpath": network.loadpath(), "max_reaction": reaction,
                       "z_min": min(network.node_coordinates(key)[2] for key in network.nodes())})
    batched = model(np.asarray(recorder.history))           # the same frames as one batched evaluation
    loadpaths = np.sum(np.abs(np.asarray(batched.forces) * np.asarray(batched.lengths)), axis=1)
    if verbose:
        print(f"{len(frames)} frames; load path {frames[0]['loadpath']:.3f} -> {frames[-1]['loadpath']:.3f}; "
              f"largest reaction {frames[0]['max_reaction']:.3f} -> {frames[-1]['max_reaction']:.3f}")
    return frames, loadpaths, network


if __name__ == "__main__":
    main()
