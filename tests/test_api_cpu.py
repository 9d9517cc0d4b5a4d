"""
Host-side mirror of the reference API, no GPU needed: network stand-in, goal flattening, host goal
inspection, constraint cotangents, recorder round trip, sharding helpers under gloo (world size 2).
"""
import os

import numpy as np
import pytest

from cases import golden_cases, case_arrays, oracle_terms, program_inputs, goal_objects, constraint_objects
from oracle import fdm_oracle as O
from dfdm_b200.datastructures import FDNetwork
from dfdm_b200.equilibrium import EquilibriumModel, EquilibriumState, NetworkBatch, network_updated
from dfdm_b200.goals import goals_state
from dfdm_b200.losses import Loss
from dfdm_b200.optimization import OptimizationRecorder

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {c["name"]: c for c in golden_cases()}


def _golden_state(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return g, EquilibriumState(g["xyz"], g["residuals"], g["vectors"], g["lengths"], g["forces"], g["q"])


def test_network_orders_edges_by_first_endpoint_then_insertion():
    net = FDNetwork()
    for k in range(4):
        net.add_node(x=float(k))
    net.add_edge(2, 3)
    net.add_edge(0, 1)
    net.add_edge(0, 3)
    net.add_edge(1, 2)
    assert list(net.edges()) == [(0, 1), (0, 3), (1, 2), (2, 3)]
    assert net.uv_index()[(0, 3)] == 1 and net.index_uv()[3] == (2, 3)
    net.node_support(3)
    assert list(net.nodes_free()) == [0, 1, 2] and list(net.nodes_fixed()) == [3]
    net.edges_forcedensities(-1.5, keys=[(0, 1)])
    assert net.edges_forcedensities() == [-1.5, 0.0, 0.0, 0.0]
    net.nodes_loads([0.0, 0.0, -2.0], keys=[1])
    assert net.node_load(1) == [0.0, 0.0, -2.0] and net.node_load(0) == [0.0, 0.0, 0.0]
    copy = net.copy()
    copy.node_attribute(0, "x", 9.0)
    assert net.node_attribute(0, "x") == 0.0 and list(copy.edges()) == list(net.edges())


def test_network_json_round_trip(tmp_path):
    case = CASES["pringle"]
    path = os.path.join(tmp_path, "net.json")
    case["network"].to_json(path)
    back = FDNetwork.from_json(path)
    assert list(back.nodes()) == list(case["network"].nodes())
    assert list(back.edges()) == list(case["network"].edges())
    assert back.edges_forcedensities() == case["network"].edges_forcedensities()
    assert list(back.nodes_fixed()) == list(case["network"].nodes_fixed())


def test_from_lines_merges_end_points():
    net = FDNetwork.from_lines([([0, 0, 0], [1, 0, 0]), ([1, 0, 0], [2, 0, 0]), ([2, 0, 0], [2, 1, 0])])
    assert net.number_of_nodes() == 4 and list(net.edges()) == [(0, 1), (1, 2), (2, 3)]


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c["terms"]))
def test_goal_objects_flatten_to_the_same_program(name):
    case = CASES[name]
    model = EquilibriumModel(case["network"])
    loss = Loss(*goal_objects(case))
    goals, terms, l2 = program_inputs(case)
    flat = []
    for t, term in enumerate(x for x in loss.loss_terms if hasattr(x, "goals")):
        for goal in term.goals:
            kind, index, weight, params = goal.record(model)
            flat.append((kind, index, t, weight, [float(p) for p in params]))
    assert len(flat) == len(goals)
    for got, want in zip(flat, goals):
        assert got[:4] == (want[0], want[1], want[2], float(want[3]))
        assert np.allclose(got[4], np.asarray(want[4], dtype=np.float64).ravel())
    prog = loss.program(model)
    assert prog.n_goals == len(goals) and prog.n_terms == len(terms) and prog.l2_alpha == l2


@pytest.mark.parametrize("name", ["mixed", "pringle", "arches10"])
def test_host_goal_inspection_matches_oracle(name):
    case = CASES[name]
    g, eqstate = _golden_state(name)
    model = EquilibriumModel(case["network"])
    edges, is_fixed, xyz, loads = case_arrays(case)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    ost = O.Model(s, loads, xyz[s.fixed_nodes])(g["q"])
    for term_obj, term_spec in zip(goal_objects(case), oracle_terms(case)):
        if term_spec[0] == "L2Regularizer":
            continue
        specs = [sp for sp in term_spec[1] if not (sp[0] == "NetworkLoadPath" and sp[2] is None)]
        objs = [o for o, sp in zip(term_obj.goals, term_spec[1]) if not (sp[0] == "NetworkLoadPath" and sp[2] is None)]
        gs = goals_state(objs, eqstate, model)
        p, t, w = O.goals_state(specs, ost, O._NP)
        assert np.allclose(gs.prediction, p, rtol=1e-13, atol=1e-15)
        assert np.allclose(gs.target, t, rtol=1e-13, atol=1e-15)
        assert np.array_equal(gs.weight, w)


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c.get("constraints")))
def test_constraint_values_and_cotangents_match_reference(name):
    """c(state) bit-close to the reference; cotangent rows pushed through the oracle's tape reproduce the golden Jacobians."""
    case = CASES[name]
    g, eqstate = _golden_state(name)
    model = EquilibriumModel(case["network"])
    edges, is_fixed, xyz, loads = case_arrays(case)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    omodel = O.Model(s, loads, xyz[s.fixed_nodes])
    for k, cons in enumerate(constraint_objects(case)):
        value = np.atleast_1d(cons.constraint(eqstate, model))
        assert np.allclose(value, g["constraint_%d" % k], rtol=1e-12, atol=1e-14)
        cots = cons.cotangents(eqstate, model)
        m = value.shape[0]
        rows = [0, m // 2, m - 1] if m > 3 else range(m)
        for r in rows:
            ref = O.vjp(g["q"], omodel, {key: arr[r] for key, arr in cots.items()})
            want = g["constraint_jac_%d" % k][r]
            assert np.max(np.abs(ref - want)) <= 1e-9 * max(np.max(np.abs(want)), 1e-12)


def test_network_update_writes_state_into_a_copy():
    case = CASES["pillow"]
    g, eqstate = _golden_state("pillow")
    net = case["network"]
    new = network_updated(net, eqstate)
    assert new is not net
    for idx, edge in new.index_uv().items():
        assert new.edge_attribute(edge, "length") == g["lengths"][idx]
        assert new.edge_attribute(edge, "force") == g["forces"][idx]
        assert new.edge_forcedensity(edge) == g["q"][idx]
    for idx, node in new.index_key().items():
        assert new.node_coordinates(node) == g["xyz"][idx].tolist()
        assert new.node_residual(node) == g["residuals"][idx].tolist()
    assert abs(new.loadpath() - np.sum(np.abs(g["forces"] * g["lengths"]))) < 1e-9
    assert net.node_coordinates(24) != new.node_coordinates(24)


def test_network_batch_defers_the_write_back():
    """A batch of states materialises networks only for the designs asked for (row f-3: lazy write-back)."""
    case = CASES["pillow"]
    g, eqstate = _golden_state("pillow")
    scale = np.array([1.0, 2.0, 0.5])
    batched = EquilibriumState(*[np.stack([np.asarray(a) * s for s in scale]) for a in
                                 (eqstate.xyz, eqstate.residuals, eqstate.vectors, eqstate.lengths, eqstate.forces, eqstate.force_densities)])
    batch = NetworkBatch(case["network"], batched)
    assert len(batch) == 3 and not batch._made
    second = batch[1]
    assert list(batch._made) == [1] and batch[1] is second and batch[-2] is second
    for idx, edge in second.index_uv().items():
        assert second.edge_attribute(edge, "length") == 2.0 * g["lengths"][idx]
    assert np.allclose(batch.loadpaths(), scale ** 2 * np.sum(np.abs(g["forces"] * g["lengths"])), rtol=1e-14)
    assert abs(batch[2].loadpath() - batch.loadpaths()[2]) < 1e-9
    assert len(batch[0:2]) == 2
    with pytest.raises(IndexError):
        batch[3]
    with pytest.raises(ValueError):
        NetworkBatch(case["network"], eqstate)


def test_recorder_round_trip(tmp_path):
    rec = OptimizationRecorder()
    for k in range(3):
        rec(np.arange(4) * float(k))
    path = os.path.join(tmp_path, "history.json")
    rec.to_json(path)
    back = OptimizationRecorder.from_json(path)
    assert back.history == rec.history and len(back.history) == 3


def test_model_without_gpu_raises_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    model = EquilibriumModel(CASES["pillow"]["network"])
    assert model.structure.free_nodes == np.load(os.path.join(GOLDEN, "pillow.npz"))["free"].tolist()
    with pytest.raises(RuntimeError):
        model(CASES["pillow"]["q"])


# ---- sharding under gloo, world size 2 ---------------------------------------------------------

def _gloo_worker(rank, world, port, batch, tmpdir):
    import torch
    import torch.distributed as dist
    from dfdm_b200.distributed import ShardedBatch, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E = 5
    q = torch.arange(batch * E, dtype=torch.float64).reshape(batch, E)

    def evaluate(q_local):                      # stands in for Evaluator.loss_and_grad on this rank's GPU
        return (q_local ** 2).sum(dim=1), 2.0 * q_local

    sb = ShardedBatch(evaluate)
    lo, hi = shard_bounds(batch, world, rank)
    assert sb.local_slice(q).shape[0] == hi - lo
    loss, dq = sb.loss_and_grad(q)
    mean_loss, mean_dq = sb.ensemble(q)
    # the other exchanges: root only; losses only (gradients stay with their designs); agreement on a per-rank verdict
    from dfdm_b200.distributed import all_ranks_agree
    root_loss, root_dq = ShardedBatch(evaluate, exchange="root").loss_and_grad(q)
    assert (root_loss is None) == (rank != 0) and (root_dq is None) == (rank != 0)
    own = ShardedBatch(evaluate, exchange="losses")
    all_loss, own_dq = own.loss_and_grad(q)
    assert "losses" in own.collective and own_dq.shape[0] == hi - lo and torch.equal(own_dq, 2.0 * q[lo:hi])
    assert all_ranks_agree(True) and not all_ranks_agree(rank == 0)
    torch.save({"loss": loss, "dq": dq, "mean_loss": mean_loss, "mean_dq": mean_dq, "root_loss": root_loss, "root_dq": root_dq,
                "all_loss": all_loss}, os.path.join(tmpdir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_sharded_batch_gathers_in_design_order(tmp_path, batch):
    import socket
    import torch
    import torch.multiprocessing as mp
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    mp.spawn(_gloo_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    q = torch.arange(batch * 5, dtype=torch.float64).reshape(batch, 5)
    for rank in range(2):
        out = torch.load(os.path.join(tmp_path, "rank%d.pt" % rank))
        assert torch.equal(out["loss"], (q ** 2).sum(dim=1))
        assert torch.equal(out["dq"], 2.0 * q)
        assert torch.allclose(out["mean_loss"], (q ** 2).sum(dim=1).mean())
        assert torch.allclose(out["mean_dq"], (2.0 * q).mean(dim=0))
        assert torch.equal(out["all_loss"], (q ** 2).sum(dim=1))
        if rank == 0:
            assert torch.equal(out["root_loss"], (q ** 2).sum(dim=1)) and torch.equal(out["root_dq"], 2.0 * q)


def test_peer_gather_is_not_offered_without_nccl_ranks():
    from dfdm_b200.distributed import PeerGather
    assert not PeerGather.available() and not PeerGather.available(batch=8)


def test_shard_bounds_cover_the_batch_exactly():
    from dfdm_b200.distributed import shard_bounds
    for batch in (1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_plans_are_cached_by_topology():
    from dfdm_b200.equilibrium.structure import cached_plan
    a = EquilibriumModel(CASES["dome"]["network"])
    b = EquilibriumModel(CASES["dome"]["network"].copy())
    assert a.plan is b.plan
    other = CASES["dome"]["network"].copy()
    other.node_support(40)
    assert EquilibriumModel(other).plan is not a.plan
    edges, is_fixed, _, _ = case_arrays(CASES["pillow"])
    assert cached_plan(len(is_fixed), edges, is_fixed) is cached_plan(len(is_fixed), edges.copy(), is_fixed.copy())


def test_reference_import_paths_resolve_after_compat_install():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import dfdm_b200.compat as c; c.install(); "
            "from dfdm.equilibrium import fdm, constrained_fdm, EquilibriumModel; from dfdm.datastructures import FDNetwork; "
            "from dfdm.goals import NodeLineGoal, EdgeLengthGoal, NetworkLoadPathGoal; from dfdm.losses import Loss, SquaredError, L2Regularizer; "
            "from dfdm.constraints import NetworkEdgesLengthConstraint, NodeCurvatureConstraint; "
            "from dfdm.optimization import SLSQP, BFGS, TrustRegionConstrained, OptimizationRecorder; print('ok')") % os.path.join(os.path.dirname(__file__), "..")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_batched_lbfgs_matches_scipy_on_a_batch_of_bounded_quadratics():
    """The lock-step optimiser against SciPy's L-BFGS-B, design by design, on problems with a known gradient."""
    from scipy.optimize import minimize
    from dfdm_b200.optimization import BatchedLBFGS
    rng = np.random.default_rng(0)
    B, E = 6, 12
    A = rng.standard_normal((B, E, E))
    H = np.einsum("bij,bkj->bik", A, A) + 0.5 * np.eye(E)
    c = rng.standard_normal((B, E))

    def fun(q):
        Hq = np.einsum("bij,bj->bi", H, q)
        return 0.5 * np.sum(q * Hq, axis=1) + np.sum(c * q, axis=1) + 0.1 * np.sum(q ** 4, axis=1), Hq + c + 0.4 * q ** 3

    q0 = rng.standard_normal((B, E))
    q, f = BatchedLBFGS(memory=10, maxiter=200, gtol=1e-9).minimize_fun(fun, q0, bounds=(-0.3, 0.5))
    for b in range(B):
        def one(x, b=b):
            full = np.zeros((B, E))
            full[b] = x
            v, g = fun(full)
            return v[b], g[b]
        ref = minimize(one, q0[b].clip(-0.3, 0.5), jac=True, method="L-BFGS-B", bounds=[(-0.3, 0.5)] * E, options={"ftol": 1e-15, "gtol": 1e-10})
        assert abs(f[b] - ref.fun) <= 1e-8 * max(1.0, abs(ref.fun))
        assert np.max(np.abs(q[b] - ref.x)) < 1e-4
    assert np.all(q >= -0.3) and np.all(q <= 0.5)


def test_bind_to_gpu_cpus_changes_nothing_without_a_gpu():
    from dfdm_b200.distributed import bind_to_gpu_cpus
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_cpus(0) is None
    assert os.sched_getaffinity(0) == before


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: one JSON line on stdout with the keys the driver reads, timed on the host cores."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "pillow",
                          "--cpu-budget", "1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "equilibrium+VJP evals/sec" and line["unit"] == "evals/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "pillow"


def test_coarse_cycle_bytes_counts_every_visit():
    """bench.py's algorithmic bytes of the coarse levels: V-cycle = one down + one up per level; W levels double the visits below."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from dfdm_b200 import workloads as W
    from dfdm_b200.engine import Plan
    edges, fixed, xyz, loads, q = W.grid_arrays(64)
    plan = Plan(len(fixed), edges, fixed)
    from dfdm_b200 import _lib
    rows = [int(v) for v in plan.array(_lib.ARR_MG_ROWS)]
    nnzs = [int(v) for v in plan.array(_lib.ARR_MG_NNZ)]
    v = bench.coarse_cycle_bytes(plan, -1)
    want = sum(8.0 * nnzs[l] + 64.0 * rows[l] + 24.0 * rows[l + 1] for l in range(1, len(rows) - 1)) + 4.0 * rows[-1] ** 2 + 24.0 * rows[-1]
    assert abs(v - want) < 1e-6 * want
    assert bench.coarse_cycle_bytes(plan, 1) > 1.9 * v - 1e-9 or len(rows) < 3          # level 1 and everything below it twice
    assert bench.coarse_cycle_bytes(plan, 0) == bench.coarse_cycle_bytes(plan, 3)       # 0 means the default depth


def test_compas_data_json_fixtures_are_read_and_written_back():
    """Files in the layout compas.json_dump gives Data objects ({"dtype": "module/Class", "value": obj.data, "guid": ...};
    Network.data with repr()-ed keys, 'dna' / 'dea' defaults, 'adjacency', 'max_node'), written by hand from COMPAS 1.x's
    published serialisation - COMPAS itself is not installable here - as read by examples/history.py:95,104."""
    import json
    net = FDNetwork.from_json(os.path.join(GOLDEN, "compas_network_fixture.json"))
    assert list(net.nodes()) == [0, 1, 2, 3] and list(net.edges()) == [(0, 1), (1, 2), (2, 3)]
    assert list(net.nodes_fixed()) == [0, 3] and net.edges_forcedensities() == [-1.0, -2.0, -1.0]
    assert net.node_load(1) == [0.0, 0.0, -0.5] and net.node_attribute(2, "rx") == 0.0        # defaults come from 'dna'
    rec = OptimizationRecorder.from_json(os.path.join(GOLDEN, "compas_history_fixture.json"))
    assert len(rec.history) == 3 and rec.history[1] == [-1.1, -1.9, -1.05]
    # written back in the same envelope, keys as repr() strings
    payload = json.loads(net.to_jsonstring())
    assert payload["dtype"] == "dfdm.datastructures/FDNetwork" and set(payload["value"]) >= {"attributes", "dna", "dea", "node", "edge", "adjacency", "max_node"}
    assert list(payload["value"]["node"]) == ["0", "1", "2", "3"] and payload["value"]["edge"]["1"] == {"2": {"q": -2.0}}
    assert payload["value"]["adjacency"]["1"] == {"0": None, "2": None}
