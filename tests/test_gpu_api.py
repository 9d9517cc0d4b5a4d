"""
The reference-facing Python API on the GPU: fdm(), EquilibriumModel, Loss, constraints and the
optimiser loop, against golden vectors of the reference and against the same SciPy call fed by the oracle.
"""
import os

import numpy as np
import pytest
from scipy.optimize import Bounds, minimize

from cases import golden_cases, case_arrays, oracle_terms, goal_objects, constraint_objects
from oracle import fdm_oracle as O
from dfdm_b200.equilibrium import EquilibriumModel, fdm, constrained_fdm
from dfdm_b200.losses import Loss, SquaredError
from dfdm_b200.goals import NodeResidualDirectionGoal
from dfdm_b200.optimization import SLSQP, BFGS, TrustRegionConstrained, OptimizationRecorder, BatchedGradientDescent, BatchedLBFGS

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {c["name"]: c for c in golden_cases()}


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _oracle_model(case):
    edges, is_fixed, xyz, loads = case_arrays(case)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    return O.Model(s, loads, xyz[s.fixed_nodes])


@pytest.mark.parametrize("name", ["pillow", "dome", "arch200"])
def test_fdm_updates_a_copy_of_the_network(name):
    case = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    net = case["network"].copy()
    for edge, value in zip(list(net.edges()), g["q"]):
        net.edge_forcedensity(edge, float(value))
    eq = fdm(net)
    assert eq is not net
    xyz = np.asarray([eq.node_coordinates(k) for k in eq.nodes()])
    assert _rel(xyz, g["xyz"]) < 1e-9
    assert _rel(eq.edges_lengths(), g["lengths"]) < 1e-9 and _rel(eq.edges_forces(), g["forces"]) < 1e-9
    assert abs(eq.loadpath() - np.sum(np.abs(g["forces"] * g["lengths"]))) < 1e-9 * max(1.0, eq.loadpath())


def test_fdm_batch_equals_fdm_design_by_design():
    from dfdm_b200.equilibrium import fdm_batch
    case = CASES["dome"]
    g = np.load(os.path.join(GOLDEN, "dome.npz"))
    net = case["network"].copy()
    q = np.stack([g["q"], 1.1 * g["q"], 0.9 * g["q"]])
    batch = fdm_batch(net, q)
    assert len(batch) == 3
    for b in (0, 2):
        one = net.copy()
        for edge, value in zip(list(one.edges()), q[b]):
            one.edge_forcedensity(edge, float(value))
        eq = fdm(one)
        assert _rel([batch[b].node_coordinates(k) for k in net.nodes()], [eq.node_coordinates(k) for k in net.nodes()]) < 1e-12
        assert _rel(batch[b].edges_forces(), eq.edges_forces()) < 1e-12
    assert _rel([batch[0].node_coordinates(k) for k in net.nodes()], g["xyz"]) < 1e-9
    assert _rel(batch.loadpaths()[0], np.sum(np.abs(g["forces"] * g["lengths"]))) < 1e-9


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c["terms"]))
def test_loss_objects_match_reference(name):
    case = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = EquilibriumModel(case["network"])
    loss = Loss(*goal_objects(case))
    value = loss(g["q"], model)
    assert isinstance(value, float) and abs(value - float(g["loss"])) <= 1e-9 * max(1.0, abs(float(g["loss"])))
    v2, grad = loss.value_and_grad(g["q"], model)
    assert abs(v2 - value) <= 1e-12 * max(1.0, abs(value)) and _rel(grad, g["grad"]) < 1e-7
    # the terms evaluated one by one (examples/dome.py:298-306) add up to the loss
    eqstate = model(g["q"])
    parts = sum(term(eqstate, model) for term in loss.loss_terms)
    assert abs(parts - value) <= 1e-9 * max(1.0, abs(value))


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c.get("constraints")))
def test_constraints_and_jacobians_match_reference(name):
    case = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = EquilibriumModel(case["network"])
    for k, cons in enumerate(constraint_objects(case)):
        assert _rel(np.atleast_1d(cons(g["q"], model)), g["constraint_%d" % k]) < 1e-9
        assert _rel(cons.jacobian(g["q"], model), g["constraint_jac_%d" % k]) < 1e-7


def test_singular_network_raises_linalg_error():
    from dfdm_b200.datastructures import FDNetwork
    net = FDNetwork()
    for k in range(4):
        net.add_node(x=float(k))
    for k in range(3):
        net.add_edge(k, k + 1)
    net.node_support(0)
    net.node_support(3)
    net.edges_forcedensities(0.0)        # getter: q stays 0 everywhere -> K = 0
    with pytest.raises(np.linalg.LinAlgError):
        fdm(net)


def test_slsqp_follows_the_same_path_as_scipy_on_the_oracle():
    """examples/arches.py: SLSQP with bounds (-inf, 0) on residual-direction goals."""
    case = CASES["arches10"]
    net = case["network"]
    loss = Loss(*goal_objects(case))
    recorder = OptimizationRecorder()
    out = constrained_fdm(net, optimizer=SLSQP(disp=False), loss=loss, bounds=(-np.inf, 0.0), maxiter=200, tol=1e-9, callback=recorder)
    q_opt = np.asarray(out.edges_forcedensities())
    omodel = _oracle_model(case)
    terms = oracle_terms(case)
    q0 = np.asarray(net.edges_forcedensities(), dtype=np.float64)
    ref = minimize(lambda x: O.loss_and_grad(terms, x, omodel), q0, jac=True, method="SLSQP", tol=1e-9,
                   bounds=Bounds(-np.inf, 0.0), options={"maxiter": 200})
    assert _rel(q_opt, ref.x) < 1e-5
    model = EquilibriumModel(net)
    assert loss(q_opt, model) < 1e-6 and loss(q_opt, model) <= loss(q0, model)
    assert len(recorder.history) > 1
    # the supports now pull along the target directions
    res = np.asarray(out.node_residual(0))
    assert np.allclose(res / np.linalg.norm(res), np.asarray([-1.0, 0.0, -0.4]) / np.linalg.norm([-1.0, 0.0, -0.4]), atol=1e-3)


def test_bfgs_reduces_the_loss_and_constrained_slsqp_matches_scipy_on_the_oracle():
    from scipy.optimize import NonlinearConstraint
    case = CASES["pillow"]
    net = case["network"].copy()
    for edge, value in zip(list(net.edges()), case["q"]):
        net.edge_forcedensity(edge, float(value))
    loss = Loss(*goal_objects(case))
    model = EquilibriumModel(net)
    start = loss(case["q"], model)
    q_bfgs = BFGS(disp=False).minimize(net, loss, maxiter=30, tol=1e-6, verbose=False)
    assert loss(q_bfgs, model) < start
    # examples/pillow.py: length and force constraints on every edge
    cons = constraint_objects(case)[:2]
    specs = case["constraints"][:2]
    lengths = model(case["q"]).lengths
    cons[0].bound_low, cons[0].bound_up = 0.9 * lengths.min(), 1.02 * lengths.max()
    cons[1].bound_low, cons[1].bound_up = -100.0, -3.0
    q_c = SLSQP(disp=False).minimize(net, loss, constraints=cons, maxiter=40, tol=1e-8, verbose=False)
    omodel = _oracle_model(case)
    terms = oracle_terms(case)
    ocons = [NonlinearConstraint(fun=lambda x, sp=sp: O.constraint_jacobian(sp, x, omodel)[0],
                                 jac=lambda x, sp=sp: O.constraint_jacobian(sp, x, omodel)[1], lb=c.bound_low, ub=c.bound_up)
             for sp, c in zip(specs, cons)]
    ref = minimize(lambda x: O.loss_and_grad(terms, x, omodel), case["q"], jac=True, method="SLSQP", tol=1e-8,
                   bounds=Bounds(-np.inf, np.inf), constraints=ocons, options={"maxiter": 40})
    assert _rel(q_c, ref.x) < 1e-4
    state = model(q_c)
    assert np.all(state.lengths >= cons[0].bound_low - 1e-6) and np.all(state.lengths <= cons[0].bound_up + 1e-6)
    assert np.all(state.forces <= -3.0 + 1e-6)
    # trust-constr runs through the same callbacks
    q_t = TrustRegionConstrained(disp=False).minimize(net, loss, constraints=cons, maxiter=10, tol=1e-3, verbose=False)
    assert np.all(np.isfinite(q_t)) and np.isfinite(loss(q_t, model))


def test_batched_lockstep_descent_improves_every_design():
    case = CASES["dome"]
    model = EquilibriumModel(case["network"])
    loss = Loss(*goal_objects(case))
    rng = np.random.default_rng(3)
    q0 = case["q"][None, :] * (1.0 + 0.1 * (rng.random((16, len(case["q"]))) - 0.5))
    start = loss(q0, model)
    q, value = BatchedGradientDescent(step=1e-1, maxiter=25).minimize(model, loss, q0, bounds=(None, -1e-3))
    assert value.shape == (16,) and np.all(value <= start) and np.mean(value) < 0.7 * np.mean(start)
    assert np.all(q <= -1e-3)


def test_batched_lbfgs_reaches_scipy_optima_design_by_design():
    """An ensemble optimised in lock step ends where SciPy's L-BFGS-B ends for each design on its own."""
    case = CASES["pringle"]
    model = EquilibriumModel(case["network"])
    loss = Loss(*goal_objects(case))
    rng = np.random.default_rng(9)
    q0 = np.clip(case["q"][None, :] * (1.0 + 0.3 * (rng.random((32, len(case["q"]))) - 0.5)), -5.0, -0.1)
    opt = BatchedLBFGS(maxiter=150, gtol=1e-7)
    q, value = opt.minimize(model, loss, q0, bounds=(-5.0, -0.1))
    assert np.all(value <= loss(q0, model)) and np.all(q <= -0.1) and np.all(q >= -5.0)
    for b in (0, 17):
        ref = minimize(lambda x: loss.value_and_grad(x, model), q0[b], jac=True, method="L-BFGS-B", bounds=[(-5.0, -0.1)] * q0.shape[1],
                       options={"maxiter": 150, "ftol": 1e-14, "gtol": 1e-8})
        assert value[b] <= 1.15 * ref.fun + 1e-9          # same iteration budget, ill-conditioned problem: same ballpark
    assert opt.evaluations < 400


def test_stacked_constraints_equal_the_individual_ones():
    from dfdm_b200.optimization import StackedConstraints
    case = CASES["pillow"]
    g = np.load(os.path.join(GOLDEN, "pillow.npz"))
    model = EquilibriumModel(case["network"])
    cons = constraint_objects(case)
    stacked = StackedConstraints(cons, model)
    values = stacked.values(g["q"])
    jac = stacked.jacobian(g["q"])
    want_v = np.concatenate([g["constraint_%d" % k] for k in range(len(cons))])
    want_j = np.concatenate([g["constraint_jac_%d" % k] for k in range(len(cons))])
    assert values.shape == want_v.shape and jac.shape == want_j.shape
    assert _rel(values, want_v) < 1e-9 and _rel(jac, want_j) < 1e-7
    lb, ub = stacked.bounds(g["q"])
    assert lb.shape == want_v.shape and np.all(lb == 0.0) and np.all(ub == 1.0)


def test_peer_push_kernel_on_one_device():
    """dfdm_peer_alloc / _push / _free inside one process: the kernel that stores a rank's rows into every mapped buffer
    (here two buffers of this process), odd double counts included (16-byte body + 8-byte tail)."""
    import ctypes as C
    import torch
    from dfdm_b200 import _lib
    from dfdm_b200.distributed import _DeviceView
    dev = torch.device("cuda", 0)
    n_total, offset = 4099, 1024                                       # doubles; the offset is 16-byte aligned
    ptrs, views = [], []
    for _ in range(2):
        p, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(_lib.lib.dfdm_peer_alloc(n_total * 8, C.byref(p), handle))
        assert any(handle.raw)                                         # an IPC handle was produced
        ptrs.append(p.value)
        views.append(torch.as_tensor(_DeviceView(p.value, (n_total,), "<f8"), device=dev))
        views[-1].zero_()
    for count in (3075, 2048, 1):
        src = torch.arange(1, count + 1, dtype=torch.float64, device=dev)
        bases = (C.c_void_p * 2)(*ptrs)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib.dfdm_peer_push(C.c_void_p(src.data_ptr()), count * 8, offset * 8, bases, 2, stream))
        torch.cuda.synchronize()
        for v in views:
            assert torch.equal(v[offset:offset + count], src)
            assert float(v[:offset].abs().sum()) == 0.0 and float(v[offset + 3075:].abs().sum()) == 0.0
    with pytest.raises(_lib.DfdmError):
        _lib.check(_lib.lib.dfdm_peer_push(C.c_void_p(src.data_ptr()), 8, 8, bases, 2, stream))   # misaligned destination
    del views
    for p in ptrs:
        _lib.check(_lib.lib.dfdm_peer_free(C.c_void_p(p)))


def test_user_defined_goal_and_loss_term_run_through_the_cuda_adjoint():
    """The reference differentiates any Goal / LossTerm subclass through autograd (goals/goal.py:40-68, losses.py:22-31).
    Here user code runs on a torch tape over EquilibriumFunction (forward dfdm_forward, backward dfdm_vjp); value and
    gradient must equal the same functions on the oracle's CPU tape."""
    import torch
    from dfdm_b200.goals import NodeGoal, EdgeGoal, ScalarGoal, EdgeLengthGoal
    from dfdm_b200.losses import LossTerm, SquaredError, L2Regularizer
    from dfdm_b200.autograd import equilibrium

    class NodeHeightGoal(ScalarGoal, NodeGoal):                 # user-defined: no KIND, torch prediction
        def __init__(self, key, target, weight=1.0):
            super().__init__(key=key, target=target, weight=weight)

        def prediction(self, eq_state, index):
            return eq_state.xyz[index, 2:3]

    class EdgeStrainEnergyGoal(ScalarGoal, EdgeGoal):
        def __init__(self, key, target, weight=1.0):
            super().__init__(key=key, target=target, weight=weight)

        def prediction(self, eq_state, index):
            return torch.atleast_1d(eq_state.forces[index] ** 2 * eq_state.lengths[index])

    class LogCoshError(LossTerm):                               # user-defined term: overrides loss(gstate)
        def loss(self, gstate):
            return self.alpha * torch.sum(gstate.weight * torch.log(torch.cosh(gstate.prediction - gstate.target)))

    case = CASES["pringle"]
    network = case["network"]
    model = EquilibriumModel(network)
    q = np.asarray(case["q"], dtype=np.float64)
    free = list(network.nodes_free())
    edges = list(network.edges())
    heights = [NodeHeightGoal(k, target=0.3 + 0.01 * i, weight=1.0 + 0.1 * i) for i, k in enumerate(free[:7])]
    energies = [EdgeStrainEnergyGoal(e, target=0.05, weight=2.0) for e in edges[3:9]]
    lengths = [EdgeLengthGoal(e, target=0.4) for e in edges[:10]]
    loss = Loss(SquaredError(lengths), SquaredError(heights, alpha=0.5), LogCoshError(energies, alpha=3.0), L2Regularizer(0.01))
    value, grad = loss.value_and_grad(q, model)

    # the same on the oracle's torch-f64 CPU tape
    edges_a, is_fixed, xyz, loads = case_arrays(case)
    s = O.Structure(len(is_fixed), edges_a, is_fixed)
    omodel = O.Model(s, loads, xyz[s.fixed_nodes])
    xp = O._torch_backend()
    qt = torch.tensor(q, requires_grad=True)
    st = omodel(qt, xp)
    ni, ei = model.structure.node_index, model.structure.edge_index
    ref = sum((st.lengths[ei[e]] - 0.4) ** 2 for e in edges[:10])
    ref = ref + 0.5 * sum((1.0 + 0.1 * i) * (st.xyz[ni[k], 2] - (0.3 + 0.01 * i)) ** 2 for i, k in enumerate(free[:7]))
    ref = ref + 3.0 * sum(2.0 * torch.log(torch.cosh(st.forces[ei[e]] ** 2 * st.lengths[ei[e]] - 0.05)) for e in edges[3:9])
    ref = ref + 0.01 * torch.sum(qt ** 2)
    gref, = torch.autograd.grad(ref, qt)
    assert abs(value - float(ref)) < 1e-9 * abs(float(ref))
    assert np.max(np.abs(grad - gref.numpy())) < 1e-7 * np.max(np.abs(gref.numpy()))

    # the differentiable calculator on its own, batched: d/dq of a scalar of every state array
    def scalar_of(state, dim):
        return (state.xyz ** 2).sum() + 0.3 * state.lengths.sum() + (state.residuals[..., 2] * state.forces.mean(dim=dim, keepdim=True)).sum() \
            + (state.vectors[..., 0] * state.forces).sum()

    qb = torch.tensor(np.stack([q, 0.9 * q]), dtype=torch.float64, device="cuda", requires_grad=True)
    scalar_of(equilibrium(qb, model), 1).backward()
    for b, scale in enumerate((1.0, 0.9)):
        qo = torch.tensor(scale * q, requires_grad=True)
        go, = torch.autograd.grad(scalar_of(omodel(qo, xp), 0), qo)
        assert np.max(np.abs(qb.grad[b].cpu().numpy() - go.numpy())) < 1e-7 * np.max(np.abs(go.numpy()))


def test_ensemble_optimisation_stays_on_the_device_and_sits_behind_minimize():
    """Optimizer.minimize(q0=q[B, E]) dispatches to the device-resident lock-step L-BFGS: tensors in, tensors out, and the
    same optima as the NumPy-facing call; the evaluation it uses returns device tensors."""
    import torch
    case = CASES["dome"]
    net = case["network"]
    model = EquilibriumModel(net)
    loss = Loss(*goal_objects(case))
    rng = np.random.default_rng(4)
    q0 = np.clip(case["q"][None, :] * (1.0 + 0.2 * (rng.random((24, len(case["q"]))) - 0.5)), -50.0, -1e-3)
    f_dev, g_dev = loss.value_and_grad_device(torch.as_tensor(q0, device="cuda"), model)
    f_host, g_host = loss.value_and_grad(q0, model)
    assert f_dev.is_cuda and g_dev.is_cuda and np.allclose(f_dev.cpu().numpy(), f_host, rtol=1e-12) and np.allclose(g_dev.cpu().numpy(), g_host, rtol=1e-9, atol=1e-12)
    opt = SLSQP(disp=False)
    q_np = opt.minimize(net, loss, bounds=(-50.0, -1e-3), maxiter=40, tol=1e-7, verbose=False, q0=q0)
    assert isinstance(q_np, np.ndarray) and q_np.shape == q0.shape and opt.result["nit"] > 0
    q_t = BFGS(disp=False).minimize(net, loss, bounds=(-50.0, -1e-3), maxiter=40, tol=1e-7, verbose=False, q0=torch.as_tensor(q0, device="cuda"))
    assert isinstance(q_t, torch.Tensor) and q_t.is_cuda
    assert np.allclose(q_t.cpu().numpy(), q_np, rtol=1e-9, atol=1e-12)          # the same iteration, whatever the container
    end = loss(q_np, model)
    assert np.all(end <= f_host + 1e-12) and end.mean() < 0.5 * f_host.mean()
    assert np.all(q_np <= -1e-3) and np.all(q_np >= -50.0)
