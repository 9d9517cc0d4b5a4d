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


# ---- the reference's own optimisers on the same scripts ------------------------------------------------------------

GOLDEN_EXAMPLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "examples.npz")
# keyword arguments of every example's main(): the ones tests/golden/make_example_golden.py ran the reference with
EXAMPLE_RUNS = {
    "arches": dict(vertical_comps=(0.2, 0.8), verbose=False),
    "pillow": dict(maxiter=40, verbose=False),
    "pringle": dict(maxiter=60, verbose=False),
    "dome_constraints": dict(maxiter=40, verbose=False),
    "monkey_saddle_constraints": dict(maxiter=40, verbose=False),
    "dome": dict(maxiter=300, tol=1e-3, verbose=False),
    "monkey_saddle": dict(maxiter=200, tol=1e-3, verbose=False),
}
# Relative tolerances on (loss, q, xyz) per example.  The optimisers are SciPy's in both runs and the gradients agree to
# ~1e-12, so runs that CONVERGE land on the same optimum to the optimiser's own tolerance; runs cut off by maxiter
# (pillow, dome_constraints with constraints, trust-constr) follow the same path only as far as rounding lets a
# quasi-Newton iteration - their shapes are compared to the accuracy the measured differences (profiles/README.md) support.
EXAMPLE_TOL = {
    "arches": (1e-9, 1e-6, 1e-6),
    "pillow": (1e-3, 1e-3, 1e-4),
    "pringle": (0.5, 0.5, 5e-2),          # SLSQP stopped by maxiter far from the optimum: the iterates drift apart (see the test)
    "dome_constraints": (1e-3, 1e-3, 1e-4),
    "monkey_saddle_constraints": (5e-2, 5e-2, 1e-2),
    "dome": (1e-2, 5e-2, 1e-3),
    "monkey_saddle": (1e-5, 1e-2, 2e-3),   # same loss to 1e-6; q along the flat directions the regulariser leaves
}


@pytest.mark.parametrize("name", sorted(EXAMPLE_RUNS))
def test_example_reproduces_the_reference_run(name):
    """Every constrained_fdm call of the example, in order, against the same call made by the reference's code."""
    import dfdm_b200.equilibrium as EQ
    gold = np.load(GOLDEN_EXAMPLES)
    mod = _load(name)
    calls = []
    real = EQ.constrained_fdm

    def recording(network, optimizer, loss, *args, **kwargs):
        out = real(network, optimizer, loss, *args, **kwargs)
        q = np.asarray(out.edges_forcedensities(), dtype=np.float64)
        xyz = np.asarray([out.node_coordinates(k) for k in out.nodes()], dtype=np.float64)
        from dfdm_b200.equilibrium import EquilibriumModel
        model = EquilibriumModel(network)
        calls.append({"q": q, "xyz": xyz, "loss": float(loss(q, model)), "loss_fn": loss, "model": model})
        return out

    mod.constrained_fdm = recording
    mod.main(**EXAMPLE_RUNS[name])
    expected = sorted({int(k.split(".")[1]) for k in gold.files if k.startswith(name + ".")})
    assert len(calls) == len(expected) > 0
    tol_loss, tol_q, tol_x = EXAMPLE_TOL[name]
    report = []
    for i, c in enumerate(calls):
        g = {key: gold["%s.%d.%s" % (name, i, key)] for key in ("q", "xyz", "loss")}
        e_loss = abs(c["loss"] - float(g["loss"])) / max(abs(float(g["loss"])), 1e-12) if abs(float(g["loss"])) > 1e-9 else abs(c["loss"])
        e_q = float(np.max(np.abs(c["q"] - g["q"])) / np.max(np.abs(g["q"])))
        e_x = float(np.max(np.abs(c["xyz"] - g["xyz"])) / np.max(np.abs(g["xyz"])))
        report.append((i, e_loss, e_q, e_x))
        # evaluation parity, independent of the optimiser's path: OUR loss and coordinates at the REFERENCE's optimum
        at_gold = float(c["loss_fn"](g["q"], c["model"]))
        assert abs(at_gold - float(g["loss"])) <= 1e-9 * max(abs(float(g["loss"])), 1e-3), (name, i, at_gold, float(g["loss"]))
        assert np.max(np.abs(c["model"](g["q"]).xyz - g["xyz"])) <= 1e-9 * np.max(np.abs(g["xyz"]))
    print("example %s vs reference run: (call, loss, q, xyz) relative differences: %s" % (name, report))
    for i, e_loss, e_q, e_x in report:
        assert e_loss < tol_loss and e_q < tol_q and e_x < tol_x, (name, report)


def test_history_replays_the_recorded_optimisation(tmp_path):
    """examples/monkey_saddle.py writes base / history / optimised JSON; examples/history.py replays them frame by frame."""
    network, curves, reactions, targets, out_dir = _load("monkey_saddle").main(maxiter=30, tol=1e-3, out_dir=str(tmp_path), verbose=False)
    assert set(curves) == {"Loss", "EdgeLengthGoal", "ReactionForceGoal", "LoadPathGoal", "L2Regularizer"}
    total = curves["EdgeLengthGoal"] + curves["ReactionForceGoal"] + curves["LoadPathGoal"] + curves["L2Regularizer"]
    assert np.allclose(total, curves["Loss"], rtol=1e-9)          # the terms add up to the loss at every recorded iterate
    assert curves["Loss"][-1] < curves["Loss"][0]
    frames, loadpaths, last = _load("history").main(in_dir=out_dir, verbose=False)
    assert len(frames) == len(curves["Loss"]) > 1
    assert np.allclose([f["loadpath"] for f in frames], loadpaths, rtol=1e-9)     # frame-by-frame loop == one batched evaluation
    assert len(last.edges_forcedensities()) == len(network.edges_forcedensities())


def test_dome_loss_terms_along_the_history():
    networks, curves = _load("dome").main(maxiter=25, tol=1e-3, verbose=False, num_rings=12)
    assert {"Loss", "SquaredError"} <= set(curves) and len(curves["Loss"]) > 1
    assert np.allclose(curves["Loss"], curves["SquaredError"], rtol=1e-9)
    assert curves["Loss"][-1] < curves["Loss"][0]
