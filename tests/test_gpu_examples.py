"""The example scripts (ports of the reference's examples, viewer removed) run end to end on the GPU."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EXAMPLES = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "examples"))


def _load(name):
    import sys
    if EXAMPLES not in sys.path:
        sys.path.insert(0, EXAMPLES)
    spec = importlib.util.spec_from_file_location("example_" + name, os.path.join(EXAMPLES, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_arch():
    z_mid, reaction = _load("arch").main(verbose=False)
    assert abs(z_mid - 312500.0) < 1e-9 * 312500.0
    assert abs(reaction[2] + 0.1 * 4999 / 2) < 1e-6          # residuals sum to the loads: half of the total load at each support


def test_arches():
    heights = _load("arches").main(vertical_comps=(0.2, 0.8), verbose=False)
    assert heights[1] > heights[0] > 0                       # steeper reactions, higher arch


def test_pillow():
    networks, stats = _load("pillow").main(maxiter=40, verbose=False)
    assert stats["uncstr_opt"]["loadpath"] > 0
    cstr = networks["cstr_opt"]
    forces = np.asarray(cstr.edges_forces())
    assert np.all(forces <= -1.0 + 1e-4) and np.all(forces >= -100.0 - 1e-4)


def test_pringle():
    _, losses = _load("pringle").main(maxiter=60, verbose=False)
    assert losses[-1] < 0.05 * losses[0]


def test_dome_constraints():
    networks, angles = _load("dome_constraints").main(maxiter=40, verbose=False)
    assert min(angles["eq_g_c"]) >= 10.0 - 1e-3 and max(angles["eq_g_c"]) <= 30.0 + 1e-3


def test_monkey_saddle_constraints():
    network, reactions, targets = _load("monkey_saddle_constraints").main(maxiter=40, verbose=False)
    lengths = np.asarray(network.edges_lengths())
    assert np.all(lengths >= 0.5 - 1e-3) and np.all(lengths <= 4.5 + 1e-3)
    start = sum((np.linalg.norm([0.0, 0.0, 0.0]) - t) ** 2 for t in targets.values())
    assert sum((reactions[k] - targets[k]) ** 2 for k in targets) < start


def test_ensemble_and_cable_net():
    value, value_opt = _load("ensemble").main(batch=256, maxiter=10, verbose=False)
    assert np.all(value_opt <= value)
    loss = _load("cable_net").main(nfree=64, batch=8, verbose=False)
    assert np.all(np.isfinite(loss)) and loss.shape == (8,)


def test_sharded_ensemble_on_one_rank():
    """examples/ensemble_sharded.py with a world of one: the shard is the whole ensemble, the gathers are identities."""
    start, end, q_opt = _load("ensemble_sharded").main(batch_per_gpu=128, maxiter=8, verbose=False)
    assert start.shape == end.shape == (128,) and q_opt.shape[0] == 128
    assert np.all(end <= start) and end.mean() < 0.7 * start.mean()
    assert np.all(q_opt <= -1e-3 + 1e-12)
