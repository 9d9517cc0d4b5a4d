"""
Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors of the
reference.  Tolerances are those of BASELINE.json's north_star: equilibrium coordinates within 1e-9
relative, gradients within 1e-7 relative (FP64); indexing is checked bit-exact in test_plan_cpu.py.
"""
import os

import numpy as np
import pytest

from cases import golden_cases, case_arrays, oracle_terms, program_inputs
from oracle import fdm_oracle as O
from dfdm_b200 import _lib
from dfdm_b200.engine import Plan, GoalProgram, Evaluator

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {c["name"]: c for c in golden_cases()}
# DFDM_TEST_TOL_SCALE=0.1 checks that the solver defaults leave a tenfold margin to the contract
TOL_SCALE = float(os.environ.get("DFDM_TEST_TOL_SCALE", "1"))
TOL_X = 1e-9 * TOL_SCALE
TOL_G = 1e-7 * TOL_SCALE
STATE = ("xyz", "residuals", "vectors", "lengths", "forces")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _setup(case, solver):
    edges, is_fixed, xyz, loads = case_arrays(case)
    plan = Plan(len(is_fixed), edges, is_fixed)
    ev = Evaluator(plan, loads, xyz[plan.fixed], solver=solver)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    model = O.Model(s, loads, xyz[s.fixed_nodes])
    return plan, ev, model


@pytest.mark.parametrize("solver", ["direct", "pcg"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_matches_reference_golden(name, solver):
    if name == "mixed_sign" and solver == "pcg":
        pytest.skip("CG needs a definite operator; indefinite K is the direct solver's job")
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    plan, ev, _ = _setup(CASES[name], solver)
    out = ev.forward(g["q"])
    assert int(out["status"].cpu()[0]) == 0
    for key in STATE:
        got = out[key].cpu().numpy()[0]
        assert _rel(got, g[key]) < TOL_X, (key, _rel(got, g[key]))
    # residuals on free nodes vanish: compare against the load scale, not against themselves
    scale = max(np.abs(g["loads"]).max(), np.abs(g["residuals"]).max())
    assert np.max(np.abs(out["residuals"].cpu().numpy()[0] - g["residuals"])) < 1e-9 * scale


@pytest.mark.parametrize("solver", ["direct", "pcg"])
@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c["terms"]))
def test_loss_and_grad_match_reference_golden(name, solver):
    if name == "mixed_sign" and solver == "pcg":
        pytest.skip("CG needs a definite operator")
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = CASES[name]
    plan, ev, _ = _setup(case, solver)
    prog = GoalProgram(plan, *program_inputs(case))
    out = ev.loss_and_grad(prog, g["q"])
    assert int(out["status"].cpu()[0]) == 0
    loss = float(out["loss"].cpu()[0])
    assert abs(loss - float(g["loss"])) <= 1e-9 * max(1.0, abs(float(g["loss"])))
    assert _rel(out["dq"].cpu().numpy()[0], g["grad"]) < TOL_G
    # value-only call agrees with the fused call
    out2 = ev.loss_and_grad(prog, g["q"], grad=False)
    assert out2["dq"] is None and float(out2["loss"].cpu()[0]) == loss


@pytest.mark.parametrize("solver", ["direct", "pcg"])
@pytest.mark.parametrize("name", ["pillow", "mixed", "dome"])
def test_batch_of_designs_matches_oracle(name, solver):
    case = CASES[name]
    plan, ev, model = _setup(case, solver)
    prog = GoalProgram(plan, *program_inputs(case))
    rng = np.random.default_rng(123)
    B = 37                                                     # ragged: not a multiple of the warp width
    q = case["q"][None, :] * (1.0 + 0.2 * (rng.random((B, len(case["q"]))) - 0.5))
    out = ev.loss_and_grad(prog, q, outputs=STATE)
    assert np.all(out["status"].cpu().numpy() == 0)
    terms = oracle_terms(case)
    for b in (0, 1, 17, 36):
        st = model(q[b])
        for key in STATE:
            assert _rel(out[key].cpu().numpy()[b], getattr(st, key)) < TOL_X, key
        value, grad = O.loss_and_grad(terms, q[b], model)
        assert abs(float(out["loss"].cpu()[b]) - value) <= 1e-9 * max(1.0, abs(value))
        assert _rel(out["dq"].cpu().numpy()[b], grad) < TOL_G


@pytest.mark.parametrize("solver", ["direct", "pcg"])
def test_vjp_matches_tape(solver):
    case = CASES["pringle"]
    plan, ev, model = _setup(case, solver)
    rng = np.random.default_rng(5)
    n, E = plan.n, plan.n_edges
    B = 3
    q = np.stack([case["q"] * (1.0 + 0.1 * k) for k in range(B)])
    cots = {"xyz": rng.standard_normal((B, n, 3)), "residuals": rng.standard_normal((B, n, 3)),
            "vectors": rng.standard_normal((B, E, 3)), "lengths": rng.standard_normal((B, E)),
            "forces": rng.standard_normal((B, E))}
    dq, status = ev.vjp(q, **cots)
    assert np.all(status.cpu().numpy() == 0)
    for b in range(B):
        ref = O.vjp(q[b], model, {k: v[b] for k, v in cots.items()})
        assert _rel(dq.cpu().numpy()[b], ref) < TOL_G
    # a subset of cotangents
    dq2, _ = ev.vjp(q, lengths=cots["lengths"])
    ref = O.vjp(q[1], model, {"lengths": cots["lengths"][1]})
    assert _rel(dq2.cpu().numpy()[1], ref) < TOL_G


def test_host_entry_points_match_device_entry_points():
    case = CASES["dome"]
    plan, ev, _ = _setup(case, "auto")
    prog = GoalProgram(plan, *program_inputs(case))
    B = 8
    q = np.ascontiguousarray(np.stack([case["q"] * (1.0 + 0.01 * k) for k in range(B)]))
    dev = ev.loss_and_grad(prog, q)
    loss, dq, status = np.empty(B), np.empty_like(q), np.empty(B, dtype=np.int32)
    ev.loss_and_grad_host(prog, q, loss, dq, status)
    assert np.array_equal(loss, dev["loss"].cpu().numpy())
    assert np.array_equal(dq, dev["dq"].cpu().numpy())
    assert np.all(status == 0)
    xyz = np.empty((B, plan.n, 3))
    ev.forward_host(q, xyz=xyz, status=status)
    assert np.array_equal(xyz, ev.forward(q)["xyz"].cpu().numpy())


def test_singular_design_is_flagged_not_silently_wrong():
    # two free nodes joined by q = 0 edges only: K has a zero pivot (LinAlgError in model.py:55)
    plan = Plan(4, [[0, 1], [1, 2], [2, 3]], [1, 0, 0, 1])
    ev = Evaluator(plan, np.zeros((4, 3)), np.asarray([[0.0, 0, 0], [3.0, 0, 0]]), solver="direct")
    q = np.asarray([[1.0, 1.0, 1.0], [0.0, 0.0, 0.0]])
    out = ev.forward(q)
    status = out["status"].cpu().numpy()
    assert status[0] == _lib.STATUS_OK and status[1] == _lib.STATUS_SINGULAR
    assert np.allclose(out["xyz"].cpu().numpy()[0, :, 0], [0.0, 1.0, 2.0, 3.0])
    with pytest.raises(np.linalg.LinAlgError):
        Evaluator.raise_on_status(out["status"])


def test_runs_are_bitwise_reproducible():
    case = CASES["mixed"]
    for solver in ("direct", "pcg"):
        plan, ev, _ = _setup(case, solver)
        prog = GoalProgram(plan, *program_inputs(case))
        q = np.stack([case["q"], 0.9 * case["q"]])
        a = ev.loss_and_grad(prog, q)
        b = ev.loss_and_grad(prog, q)
        assert np.array_equal(a["dq"].cpu().numpy(), b["dq"].cpu().numpy())
        assert np.array_equal(a["loss"].cpu().numpy(), b["loss"].cpu().numpy())


def test_arch_full_size_closed_form():
    """examples/arch.py at its own size (5000 segments): z_mid = p_z N^2 / (8 q) = 312 500."""
    from dfdm_b200 import workloads as W
    net = W.arch(5000)
    edges, is_fixed, xyz, loads, q = W.network_arrays(net)
    plan = Plan(len(is_fixed), edges, is_fixed)
    assert plan.auto_solver == "direct" and plan.info["bandwidth_solver"] == 1
    ev = Evaluator(plan, loads, xyz[plan.fixed])
    out = ev.forward(q)
    z = out["xyz"].cpu().numpy()[0, :, 2]
    i = np.arange(5001)
    exact = -0.1 * i * (5000 - i) / (2 * -1.0)
    assert abs(z[2500] - 312500.0) < 1e-9 * 312500.0
    assert _rel(z, exact) < TOL_X
    res = out["residuals"].cpu().numpy()[0]
    assert np.max(np.abs(res[1:-1])) < 1e-9 * np.abs(res).max()
    assert np.allclose(res.sum(axis=0), loads.sum(axis=0), atol=1e-9 * np.abs(res).max())


def test_large_grid_properties_pcg():
    """Size-independent properties at a size the dense oracle cannot reach (64 x 64 free nodes, sparse oracle beside it)."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(64)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.auto_solver == "pcg"
    ev = Evaluator(plan, loads, xyz[plan.fixed])
    q = W.sample_q(q0, 4, seed=6, spread=0.2, relative=True)
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    out = ev.loss_and_grad(prog, q, outputs=STATE)
    assert np.all(out["status"].cpu().numpy() == 0)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    for b in (0, 3):
        loss, dq, st = sm.loss_and_grad_edge_length(q[b], lengths0, 1.0, 0.1)
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < TOL_X
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-9 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < TOL_G
    res = out["residuals"].cpu().numpy()
    free = plan.free
    assert np.max(np.abs(res[:, free])) < 1e-9 * np.abs(res).max()
    assert np.allclose(res.sum(axis=1), loads.sum(axis=0)[None, :], atol=1e-9 * np.abs(res).max())
    f, a = _lib.last_pcg_iterations()
    assert 0 < f < 5000 and 0 < a < 5000


@pytest.mark.parametrize("precond", ["jacobi", "multigrid"])
@pytest.mark.parametrize("batch", [1, 5, 6])
def test_pcg_preconditioners_agree_with_sparse_oracle(precond, batch):
    """Odd batches take the one-design-per-thread V-cycle kernels, even ones the two-design kernels; B = 1 folds rows into warps."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(40)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.info["mg_levels"] >= 2
    ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", pcg_precond=precond)
    q = W.sample_q(q0, batch, seed=11, spread=0.6, relative=True)
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    out = ev.loss_and_grad(prog, q, outputs=("xyz",))
    assert np.all(out["status"].cpu().numpy() == 0)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    for b in sorted({0, batch - 1}):
        loss, dq, st = sm.loss_and_grad_edge_length(q[b], lengths0, 1.0, 0.1)
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < TOL_X
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-9 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < TOL_G
    f, a = _lib.last_pcg_iterations()
    if precond == "multigrid":
        assert f <= 60 and a <= 60          # the V-cycle cuts the iteration count by an order of magnitude
    else:
        assert f > 100


def test_pcg_on_an_indefinite_operator_is_either_right_or_flagged():
    """Mixed-sign force densities on the CG path: a design is reported as not converged, or its answer is the reference's."""
    case = CASES["mixed_sign"]
    g = np.load(os.path.join(GOLDEN, "mixed_sign.npz"))
    edges, is_fixed, xyz, loads = case_arrays(case)
    plan = Plan(len(is_fixed), edges, is_fixed)
    direct = Evaluator(plan, loads, xyz[plan.fixed], solver="direct").forward(case["q"])
    assert int(direct["status"].cpu()[0]) == _lib.STATUS_OK
    assert _rel(direct["xyz"].cpu().numpy()[0], g["xyz"]) < TOL_X
    for precond in ("jacobi", "multigrid"):
        ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", pcg_maxiter=300, pcg_precond=precond)
        out = ev.forward(np.stack([-np.abs(case["q"]), case["q"]]))
        status = out["status"].cpu().numpy()
        assert status[0] == _lib.STATUS_OK
        if status[1] == _lib.STATUS_OK:
            assert _rel(out["xyz"].cpu().numpy()[1], g["xyz"]) < 1e-7
        else:
            assert status[1] in (_lib.STATUS_NOT_CONVERGED, _lib.STATUS_NONFINITE)


def test_overcorrected_cycle_falls_back_to_jacobi_instead_of_failing():
    """A W-cycle with an absurd over-correction is not a definite preconditioner; the solve must still come back right."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(40)
    plan = Plan(len(fixed), edges, fixed)
    ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", mg_over=1.98, mg_w_levels=3)
    q = W.sample_q(q0, 4, seed=3, spread=0.2, relative=True)
    out = ev.forward(q, outputs=("xyz",))
    assert np.all(out["status"].cpu().numpy() == 0)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    assert _rel(out["xyz"].cpu().numpy()[2], sm(q[2]).xyz) < TOL_X


@pytest.mark.parametrize("name", ["dome", "monkey_saddle", "pringle"])
def test_multigrid_pcg_matches_direct_path_on_unstructured_networks(name):
    """Medium networks of the example families forced onto the CG path: same answers as the shared-memory factorisation."""
    from dfdm_b200 import workloads as W
    if name == "dome":
        net = W.dome(num_sides=24, num_rings=30, offset_distance=0.012)
    elif name == "monkey_saddle":
        net, _ = W.monkey_saddle(n=10)
    else:
        net = W.pringle(num_u=40, num_v=21)
    edges, is_fixed, xyz, loads, q0 = W.network_arrays(net)
    plan = Plan(len(is_fixed), edges, is_fixed)
    assert plan.info["mg_levels"] >= 2
    rng = np.random.default_rng(8)
    q = q0[None, :] * (1.0 + 0.6 * (rng.random((6, len(q0))) - 0.5))
    goals = [("NodePoint", int(i), 0, 1.0, list(xyz[i] + [0.0, 0.0, 0.3])) for i in plan.free[::7]]
    goals += [("EdgeLength", e, 0, 0.5, [0.9 * float(np.linalg.norm(xyz[edges[e, 1]] - xyz[edges[e, 0]]))]) for e in range(0, len(edges), 5)]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.01)
    ref = Evaluator(plan, loads, xyz[plan.fixed], solver="direct").loss_and_grad(prog, q, outputs=("xyz",))
    out = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg").loss_and_grad(prog, q, outputs=("xyz",))
    assert np.all(ref["status"].cpu().numpy() == 0) and np.all(out["status"].cpu().numpy() == 0)
    assert _rel(out["xyz"].cpu().numpy(), ref["xyz"].cpu().numpy()) < TOL_X
    assert _rel(out["loss"].cpu().numpy(), ref["loss"].cpu().numpy()) < 1e-9
    assert _rel(out["dq"].cpu().numpy(), ref["dq"].cpu().numpy()) < TOL_G
    f, a = _lib.last_pcg_iterations()
    assert f <= 120 and a <= 120


def test_disconnected_components_and_stalled_coarsening():
    """
    200 separate 3-node chains, each hanging from its own support: the multigrid hierarchy stops after one level at
    200 rows without couplings (too many for the dense coarsest factor), and several right-hand-side columns are zero.
    """
    edges, fixed, xyz, loads = [], [], [], []
    for k in range(200):
        base = 4 * k
        fixed += [1, 0, 0, 0]
        for j in range(4):
            xyz.append([float(k), 0.0, -float(j)])
            loads.append([0.0, 0.0, 0.0 if j == 0 else -1.0 - 0.01 * k])
        edges += [[base, base + 1], [base + 1, base + 2], [base + 2, base + 3]]
    edges, fixed, xyz, loads = np.asarray(edges, dtype=np.int32), np.asarray(fixed, dtype=np.uint8), np.asarray(xyz), np.asarray(loads)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.info["mg_levels"] == 1 and plan.info["mg_n_coarsest"] == 200
    rng = np.random.default_rng(4)
    q = 1.0 + rng.random((3, len(edges)))
    ref = Evaluator(plan, loads, xyz[plan.fixed], solver="direct").forward(q, outputs=("xyz", "residuals"))
    for precond in ("multigrid", "jacobi"):
        out = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", pcg_precond=precond).forward(q, outputs=("xyz", "residuals"))
        assert np.all(out["status"].cpu().numpy() == 0)
        assert _rel(out["xyz"].cpu().numpy(), ref["xyz"].cpu().numpy()) < TOL_X
        assert _rel(out["residuals"].cpu().numpy(), ref["residuals"].cpu().numpy()) < 1e-8
    # hanging chain under tension: z_j = z_{j-1} - (load below) / q
    z = ref["xyz"].cpu().numpy()[0, 1, 2]
    assert abs(z - (0.0 - 3.0 / q[0, 0])) < 1e-12


@pytest.mark.parametrize("batch", [10, 100])
def test_pcg_with_force_densities_spanning_orders_of_magnitude(batch):
    """q log-uniform in [0.05, 20]: strongly varying coefficients, batch sizes that do not fill warps or chunks."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(48)
    plan = Plan(len(fixed), edges, fixed)
    ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg")
    rng = np.random.default_rng(21)
    q = 10.0 ** rng.uniform(np.log10(0.05), np.log10(20.0), size=(batch, len(q0)))
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("MeanSquaredError", 2.0)], 0.0)
    out = ev.loss_and_grad(prog, q, outputs=("xyz",))
    assert np.all(out["status"].cpu().numpy() == 0)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    for b in (0, batch // 2, batch - 1):
        loss, dq, st = sm.loss_and_grad_edge_length(q[b], lengths0, 2.0 / len(edges), 0.0)
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < TOL_X
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-9 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < TOL_G
    f, a = _lib.last_pcg_iterations()
    assert f < 400 and a < 400           # the multigrid preconditioner holds without the Jacobi fallback


def test_multigrid_pcg_on_an_irregular_triangulation():
    """A Delaunay mesh of 6 000 random points (node degrees 3 ... 12, hull fixed): the aggregation hierarchy built by
    pairwise matching has to work without any grid structure.  Checked against the sparse CPU oracle."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(21)
    pts = rng.random((6000, 2))
    tri = Delaunay(pts)
    pairs = np.vstack([tri.simplices[:, [0, 1]], tri.simplices[:, [1, 2]], tri.simplices[:, [2, 0]]])
    edges = np.unique(np.sort(pairs, axis=1), axis=0).astype(np.int32)
    fixed = np.zeros(len(pts), dtype=np.uint8)
    fixed[np.unique(tri.convex_hull)] = 1
    xyz = np.column_stack([pts, 0.2 * np.sin(3.0 * pts[:, 0]) * np.cos(2.0 * pts[:, 1])])
    loads = np.zeros_like(xyz)
    loads[:, 2] = -1.0 / len(pts)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.auto_solver == "pcg" and plan.info["mg_levels"] >= 3
    rows = plan.array(_lib.ARR_MG_ROWS)
    assert all(rows[k + 1] < 0.5 * rows[k] for k in range(len(rows) - 1))          # coarsening does not stall
    ev = Evaluator(plan, loads, xyz[plan.fixed])
    q = 1.0 + 0.8 * (rng.random((4, len(edges))) - 0.5)
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    out = ev.loss_and_grad(prog, q, outputs=("xyz",))
    assert np.all(out["status"].cpu().numpy() == 0)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    for b in (0, 3):
        loss, dq, st = sm.loss_and_grad_edge_length(q[b], lengths0, 1.0, 0.1)
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < TOL_X
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-9 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < TOL_G
    f, a = _lib.last_pcg_iterations()
    assert f <= 80 and a <= 80, (f, a)                                             # multigrid, not the Jacobi fallback


@pytest.mark.parametrize("mesh", ["grid", "delaunay"])
def test_matrix_free_spmv_gives_the_same_answers(mesh):
    """DFDM_FLAG_SPMV_MATRIX_FREE: K p from q over the incidence lists (fixed neighbours, ragged degrees) vs the stored CSR values."""
    from dfdm_b200 import workloads as W
    if mesh == "grid":
        edges, fixed, xyz, loads, q0 = W.grid_arrays(40)
        q = W.sample_q(q0, 6, seed=3, spread=0.5, relative=True)
    else:
        from scipy.spatial import Delaunay
        rng = np.random.default_rng(5)
        pts = rng.random((3000, 2))
        tri = Delaunay(pts)
        pairs = np.vstack([tri.simplices[:, [0, 1]], tri.simplices[:, [1, 2]], tri.simplices[:, [2, 0]]])
        edges = np.unique(np.sort(pairs, axis=1), axis=0).astype(np.int32)
        fixed = np.zeros(len(pts), dtype=np.uint8)
        fixed[np.unique(tri.convex_hull)] = 1
        xyz = np.column_stack([pts, 0.1 * np.cos(4.0 * pts[:, 0])])
        loads = np.zeros_like(xyz)
        loads[:, 2] = -1.0 / len(pts)
        q = 1.0 + 0.6 * (rng.random((6, len(edges))) - 0.5)
    plan = Plan(len(fixed), edges, fixed)
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    outs = []
    for flags in (0, _lib.FLAG_SPMV_MATRIX_FREE):
        ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", flags=flags)
        out = ev.loss_and_grad(prog, q, outputs=("xyz",))
        assert np.all(out["status"].cpu().numpy() == 0)
        outs.append({k: out[k].cpu().numpy() for k in ("xyz", "loss", "dq")})
    assert _rel(outs[1]["xyz"], outs[0]["xyz"]) < TOL_X
    assert _rel(outs[1]["loss"], outs[0]["loss"]) < 1e-9
    assert _rel(outs[1]["dq"], outs[0]["dq"]) < TOL_G


def test_grid256_matches_sparse_oracle():
    """BASELINE configs[4] at its own size (258 x 258 nodes, 65 536 free): the multigrid-preconditioned solve with its
    single-precision cycle and the default stop tests, against SuperLU on the same K.  Coordinates and loss within 1e-9,
    gradients within 1e-7 (BASELINE.json); the oracle costs about a second per design."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(256)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.auto_solver == "pcg" and plan.n_free == 65536
    ev = Evaluator(plan, loads, xyz[plan.fixed])
    q = W.sample_q(q0, 4, seed=6, spread=0.2, relative=True)       # bench.py's own sampler and seed
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    out = ev.loss_and_grad(prog, q, outputs=STATE)
    assert np.all(out["status"].cpu().numpy() == 0)
    f, a = _lib.last_pcg_iterations()
    assert 0 < f <= 40 and 0 < a <= 40
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    for b in (0, 3):
        loss, dq, st = sm.loss_and_grad_edge_length(q[b], lengths0, 1.0, 0.1)
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < TOL_X
        assert _rel(out["lengths"].cpu().numpy()[b], st.lengths) < TOL_X
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-9 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < TOL_G
    res = out["residuals"].cpu().numpy()
    assert np.max(np.abs(res[:, plan.free])) < 1e-9 * np.abs(res).max()


def test_host_calls_from_several_threads_share_a_plan():
    """dfdm_loss_and_grad_host from several threads on one plan: every call in flight owns one of two cached arenas, the
    calls take turns with the kernels, a third concurrent call queues for an arena.  Different batch sizes, so the arenas
    are regrown under the threads; each must get the device path's answers.  Then the same through the helper that
    issues a stream of batches from two threads (Evaluator.loss_and_grad_host_many)."""
    import threading
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(24)
    plan = Plan(len(fixed), edges, fixed)
    ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg")
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    q = W.sample_q(q0, 200, seed=2, spread=0.3, relative=True)
    ref = ev.loss_and_grad(prog, q)
    loss_ref, dq_ref = ref["loss"].cpu().numpy(), ref["dq"].cpu().numpy()
    errors = []

    def worker(count, rounds):
        try:
            for _ in range(rounds):
                loss = np.empty(count)
                dq = np.empty((count, len(edges)))
                status = np.empty(count, dtype=np.int32)
                ev.loss_and_grad_host(prog, np.ascontiguousarray(q[:count]), loss, dq, status)
                assert np.all(status == 0)
                assert np.allclose(loss, loss_ref[:count], rtol=1e-9, atol=0)
                assert _rel(dq, dq_ref[:count]) < TOL_G
        except Exception as exc:                       # noqa: BLE001 (reported to the main thread)
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(200, 3)), threading.Thread(target=worker, args=(37, 6)),
               threading.Thread(target=worker, args=(64, 4))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    outs = [(np.empty(200), np.empty((200, len(edges))), np.empty(200, dtype=np.int32)) for _ in range(2)]
    qc = np.ascontiguousarray(q)
    ev.loss_and_grad_host_many(prog, [(qc,) + outs[k % 2] for k in range(5)], threads=2)
    for loss, dq, status in outs:
        assert np.all(status == 0)
        assert np.allclose(loss, loss_ref, rtol=1e-9, atol=0)
        assert _rel(dq, dq_ref) < TOL_G


@pytest.mark.parametrize("nfree,batch", [(40, 6), (64, 2), (96, 34)])
def test_tail_kernel_is_the_same_preconditioner_as_the_launch_per_level_path(nfree, batch):
    """k_mg_tail (one CTA per design pair walks the coarse levels) against the launch-per-level path it replaces:
    the same arithmetic in the same order, so the same iteration counts and the same answers to rounding."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(nfree)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.info["mg_levels"] >= 2
    q = W.sample_q(q0, batch, seed=9, spread=0.5, relative=True)
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    res = {}
    for name, flags in (("tail", 0), ("levels", _lib.FLAG_MG_NO_TAIL)):
        ev = Evaluator(plan, loads, xyz[plan.fixed], solver="pcg", flags=flags)
        out = ev.loss_and_grad(prog, q, outputs=("xyz",))
        assert np.all(out["status"].cpu().numpy() == 0)
        res[name] = (out["xyz"].cpu().numpy(), out["dq"].cpu().numpy(), _lib.last_pcg_iterations())
    assert res["tail"][2] == res["levels"][2]
    assert _rel(res["tail"][0], res["levels"][0]) < 1e-11
    assert _rel(res["tail"][1], res["levels"][1]) < 1e-10
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    loss, dq, st = sm.loss_and_grad_edge_length(q[batch - 1], lengths0, 1.0, 0.1)
    assert _rel(res["tail"][0][batch - 1], st.xyz) < TOL_X and _rel(res["tail"][1][batch - 1], dq) < TOL_G


def test_direct_path_flags_a_solve_that_lost_its_accuracy_to_a_tiny_pivot():
    """Unpivoted LDL^T on an indefinite K: a pivot that is tiny but not zero passes the zero-pivot test and wrecks the
    answer.  The kernel compares the residual of its own solve with the size of the terms and must flag such designs; the
    reference's pivoted LU (model.py:55) solves them, so whatever is NOT flagged must agree with it."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(6)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.auto_solver == "direct"
    ev = Evaluator(plan, loads, xyz[plan.fixed], solver="direct")
    s = O.Structure(len(fixed), edges, fixed)
    model = O.Model(s, loads, xyz[s.fixed_nodes])
    first = int(plan.free[plan.solver_perm[0]])                      # the node whose diagonal is the first pivot
    incident = [e for e, (u, v) in enumerate(edges) if u == first or v == first]
    rng = np.random.default_rng(0)
    designs = []
    for eps in (0.0, 1e-14, 1e-11, 1e-7, 1e-3):
        q = 1.0 + 0.2 * rng.random(len(edges))
        q[incident[0]] = -(q[incident[1:]].sum()) + eps               # the first diagonal entry, sum of the incident q, is eps
        designs.append(q)
    out = ev.forward(np.stack(designs))
    status = out["status"].cpu().numpy()
    assert status[0] != 0                                              # an exact zero pivot is always flagged
    for b, q in enumerate(designs):
        if status[b] == 0:
            ref = model(q)
            assert _rel(out["xyz"].cpu().numpy()[b], ref.xyz) < 1e-7, (b, status)
    assert status[-1] == 0                                             # a pivot of 1e-3 is harmless and must not be flagged


@pytest.mark.parametrize("nfree,batch", [(5, 3), (12, 70), (20, 9), (27, 5), (30, 4)])
def test_warp_and_cta_direct_kernels_agree_with_the_dense_oracle(nfree, batch):
    """Bandwidths 5 ... 27 take the warp-per-design kernel (blocked LDL^T, DMMA trailing updates, register substitutions);
    30 is beyond it and takes the CTA-per-design kernel.  Both against np.linalg.solve on the dense oracle."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(nfree)
    plan = Plan(len(fixed), edges, fixed)
    if plan.auto_solver != "direct":
        pytest.skip("band does not fit shared memory")
    ev = Evaluator(plan, loads, xyz[plan.fixed], solver="direct")
    q = -W.sample_q(q0, batch, seed=13, spread=0.5, relative=True)     # compression: negative definite K
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.05)
    out = ev.loss_and_grad(prog, q, outputs=STATE)
    assert np.all(out["status"].cpu().numpy() == 0)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    for b in sorted({0, batch // 2, batch - 1}):
        loss, dq, st = sm.loss_and_grad_edge_length(q[b], lengths0, 1.0, 0.05)
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < TOL_X
        assert _rel(out["forces"].cpu().numpy()[b], st.forces) < TOL_X
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-9 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < TOL_G


def test_direct_solver_beyond_shared_memory_and_the_rescue_of_mixed_sign_designs():
    """48 x 48 free nodes: the banded factor (about 0.9 MB) is past the shared-memory limit, so DFDM_SOLVER_AUTO means PCG.
    (1) DFDM_SOLVER_DIRECT by name takes the global-memory variant of the factorisation: against SuperLU on a definite K.
    (2) mixed-sign q makes K indefinite, CG gives up, and with rescue=True the evaluator hands exactly those designs to
    that variant: against SuperLU again (the reference's LU, model.py:55, solves such designs at any size)."""
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q0 = W.grid_arrays(48)
    plan = Plan(len(fixed), edges, fixed)
    assert plan.auto_solver == "pcg"
    lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
    goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
    prog = GoalProgram(plan, goals, [("SquaredError", 1.0)], 0.1)
    sm = O.SparseModel(len(fixed), edges, fixed, loads, xyz[plan.fixed])
    q = W.sample_q(q0, 4, seed=21, spread=0.4, relative=True)

    direct = Evaluator(plan, loads, xyz[plan.fixed], solver="direct")
    out = direct.loss_and_grad(prog, q, outputs=STATE)
    assert np.all(out["status"].cpu().numpy() == 0)
    for b in (0, 3):
        loss, dq, st = sm.loss_and_grad_edge_length(q[b], lengths0, 1.0, 0.1)
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < TOL_X
        assert _rel(out["forces"].cpu().numpy()[b], st.forces) < TOL_X
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-9 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < TOL_G

    rng = np.random.default_rng(3)
    q_mixed = q.copy()
    for b in (1, 2):                                                   # designs 0 and 3 stay definite
        flip = rng.random(len(edges)) < 0.3
        q_mixed[b, flip] *= -1.0
    auto = Evaluator(plan, loads, xyz[plan.fixed])
    plain = auto.loss_and_grad(prog, q_mixed)
    st_plain = plain["status"].cpu().numpy()
    assert st_plain[0] == 0 and st_plain[3] == 0 and st_plain[1] != 0 and st_plain[2] != 0
    out = auto.loss_and_grad(prog, q_mixed, outputs=STATE, rescue=True)
    status = out["status"].cpu().numpy()
    for b in range(4):
        loss, dq, st = sm.loss_and_grad_edge_length(q_mixed[b], lengths0, 1.0, 0.1)
        if status[b] != 0:                                             # the unpivoted factor may flag a design, never mislead
            continue
        assert _rel(out["xyz"].cpu().numpy()[b], st.xyz) < 1e-7
        assert abs(float(out["loss"].cpu()[b]) - loss) < 1e-7 * abs(loss)
        assert _rel(out["dq"].cpu().numpy()[b], dq) < 1e-5
    assert status[0] == 0 and status[3] == 0 and (status[1] == 0 or status[2] == 0)
    fwd = auto.forward(q_mixed, outputs=("xyz",), rescue=True)
    ok = fwd["status"].cpu().numpy() == 0
    assert ok[0] and ok[3] and (ok[1] or ok[2])
    for b in np.nonzero(ok)[0]:
        assert _rel(fwd["xyz"].cpu().numpy()[b], sm(q_mixed[b]).xyz) < 1e-7
