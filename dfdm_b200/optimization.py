"""
Gradient-based optimisers (reference: src/dfdm/optimization.py:26-215): same classes, same
``minimize`` signature, SciPy drives the iteration on the host exactly as in the reference.

What changes is where the numbers come from: ``fun`` and ``jac`` are one fused CUDA evaluation
(forward + loss + closed-form adjoint) instead of two tape replays, and constraint Jacobians are
batched VJPs instead of ``autograd.jacobian``.  ``BatchedGradientDescent`` is the lock-step optimiser
for ensembles of designs, which SciPy's sequential ``minimize`` cannot serve.
"""
import json
from time import time

import numpy as np
from scipy.optimize import Bounds
from scipy.optimize import NonlinearConstraint
from scipy.optimize import minimize

from dfdm_b200.equilibrium.model import EquilibriumModel

__all__ = ["Optimizer", "ConstrainedOptimizer", "BFGS", "SLSQP", "TrustRegionConstrained", "OptimizationRecorder",
           "BatchedGradientDescent", "BatchedLBFGS", "StackedConstraints"]


class Optimizer:
    def __init__(self, name, disp=True, **kwargs):
        self.name = name
        self.disp = disp
        self.model = None

    def constraints(self, constraints, model):
        if constraints:
            print(f"Warning! {self.__class__.__name__} does not support constraints. I am ignoring them.")
        return ()

    def minimize(self, network, loss, bounds=(None, None), constraints=None, maxiter=100, tol=1e-6, verbose=True, callback=None, q0=None):
        """
        Minimise ``loss`` over the edge force densities of ``network``; returns the optimal q[E].

        ``q0`` (not in the reference): a batch of starting points ``q0[B, E]`` - NumPy array or torch tensor - turns the call
        into an ENSEMBLE optimisation: all designs are improved in lock step by the device-resident projected L-BFGS
        (``BatchedLBFGS``; SciPy's sequential drivers cannot serve a batch), one batched CUDA evaluation per trial point,
        and ``q_opt[B, E]`` comes back in the type it was given.  Constraint objects are not supported there.
        """
        name = self.name
        if q0 is not None and np.ndim(q0) == 2:
            if constraints:
                raise NotImplementedError("ensemble optimisation (q0[B, E]) handles bounds, not constraint objects")
            model = self.model = EquilibriumModel(network)
            batched = BatchedLBFGS(maxiter=maxiter, gtol=tol, ftol=min(tol * 1e-3, 1e-9))
            q_opt, value = batched.minimize(model, loss, q0, bounds=bounds, callback=callback)
            self.result = {"fun": value, "nit": batched.iterations, "nfev": batched.evaluations}
            if verbose:
                print(f"Lock-step L-BFGS over {len(q_opt)} designs: {batched.iterations} iterations, {batched.evaluations} batched evaluations")
            return q_opt
        q = np.asarray(network.edges_forcedensities(), dtype=np.float64) if q0 is None else np.asarray(q0, dtype=np.float64)
        model = self.model = EquilibriumModel(network)

        def fun(x):
            value, grad = loss.value_and_grad(x, model)
            return value, grad

        start_time = time()
        loss(q, model)
        if verbose:
            print(f"Loss warmup time: {round(time() - start_time, 4)} seconds")
        start_time = time()
        fun(q)
        if verbose:
            print(f"Gradient warmup time: {round(time() - start_time, 4)} seconds")

        lb, ub = bounds
        if lb is None:
            lb = -np.inf
        if ub is None:
            ub = +np.inf
        bounds = Bounds(lb=lb, ub=ub)
        constraints = self.constraints(constraints, model)

        if verbose:
            print(f"Optimization with {self.name} started...")
        start_time = time()
        res_q = minimize(fun=fun, jac=True, method=name, x0=q, tol=tol, bounds=bounds, constraints=constraints,
                         callback=callback, options={"maxiter": maxiter, "disp": self.disp})
        if verbose:
            print(res_q.message)
            print(f"Final loss in {res_q.nit} iterations: {res_q.fun}")
            print(f"Elapsed time: {time() - start_time} seconds")
        self.result = res_q
        return res_q.x


class ConstrainedOptimizer(Optimizer):
    def constraints(self, constraints, model):
        """
        The reference hands SciPy one NonlinearConstraint per constraint object, each with its own
        ``autograd.jacobian`` (optimization.py:115-130): every object re-solves the equilibrium for its value
        and once more per Jacobian row.  Here all constraints are stacked into ONE vector constraint
        (equivalent for SLSQP and trust-constr): its value reads the cached state of the current ``q`` and
        its Jacobian is a single batched VJP with one design per constraint row.
        """
        if not constraints:
            return ()
        return [StackedConstraints(constraints, model).as_scipy()]


class StackedConstraints:
    """Several constraint objects as one vector-valued constraint with a batched Jacobian."""
    STATE = ("xyz", "residuals", "vectors", "lengths", "forces")

    def __init__(self, constraints, model):
        self.constraints = list(constraints)
        self.model = model

    def values(self, q):
        return np.concatenate([np.atleast_1d(np.asarray(c(q, self.model), dtype=np.float64)) for c in self.constraints])

    def bounds(self, q):
        lows, ups = [], []
        for c in self.constraints:
            m = np.atleast_1d(np.asarray(c(q, self.model))).shape[0]
            lows.append(np.broadcast_to(np.asarray(c.bound_low, dtype=np.float64), (m,)))
            ups.append(np.broadcast_to(np.asarray(c.bound_up, dtype=np.float64), (m,)))
        return np.concatenate(lows), np.concatenate(ups)

    def jacobian(self, q):
        from dfdm_b200.constraints.constraint import _cached_state
        from dfdm_b200.engine import Evaluator
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
        model = self.model
        eqstate = _cached_state(q, model)
        parts = [c.cotangents(eqstate, model) for c in self.constraints]
        sizes = [next(iter(p.values())).shape[0] for p in parts]
        total = sum(sizes)
        n, E = model.plan.n, model.plan.n_edges
        shapes = {"xyz": (total, n, 3), "residuals": (total, n, 3), "vectors": (total, E, 3), "lengths": (total, E), "forces": (total, E)}
        stacked = {}
        row = 0
        for p, m in zip(parts, sizes):
            for name, arr in p.items():
                if name not in stacked:
                    stacked[name] = np.zeros(shapes[name])
                stacked[name][row:row + m] = arr
            row += m
        dq, status = model.evaluator.vjp(np.repeat(q[None, :], total, axis=0), rescue=True, **stacked)
        Evaluator.raise_on_status(status)
        return dq.cpu().numpy()

    def as_scipy(self):
        q0 = np.asarray(self.model.structure.network.edges_forcedensities(), dtype=np.float64)
        lb, ub = self.bounds(q0)
        return NonlinearConstraint(fun=self.values, jac=self.jacobian, lb=lb, ub=ub)


class BFGS(Optimizer):
    def __init__(self, **kwargs):
        super().__init__(name="BFGS", **kwargs)


class TrustRegionConstrained(ConstrainedOptimizer):
    def __init__(self, **kwargs):
        super().__init__(name="trust-constr", **kwargs)


class SLSQP(ConstrainedOptimizer):
    def __init__(self, **kwargs):
        super().__init__(name="SLSQP", **kwargs)


class OptimizationRecorder:
    """SciPy callback that records the force densities of every iteration; JSON round trip as in the reference."""
    def __init__(self):
        self.history = []

    def record(self, value):
        self.history.append(value)

    def __call__(self, q, *args, **kwargs):
        self.record(np.array(q, dtype=np.float64).tolist())

    @property
    def data(self):
        return {"history": self.history}

    @data.setter
    def data(self, data):
        self.history = data["history"]

    def to_json(self, filepath):
        with open(filepath, "w") as fp:
            json.dump({"dtype": "dfdm.optimization/OptimizationRecorder", "value": self.data}, fp)

    @classmethod
    def from_json(cls, filepath):
        with open(filepath, "r") as fp:
            payload = json.load(fp)
        recorder = cls()
        recorder.data = payload.get("value", payload)
        return recorder


class BatchedGradientDescent:
    """
    Lock-step projected gradient descent with Barzilai-Borwein steps over a batch of designs q[B, E]:
    every iteration is one batched CUDA evaluation of all designs (loss + gradient); the step and the
    projection onto the bounds are elementwise torch operations on the evaluator's GPU - the state never
    leaves the device (``Loss.value_and_grad_device``).  For ensembles / parameter sweeps on one or many GPUs.
    """
    def __init__(self, step=1e-2, maxiter=100, tol=1e-6):
        self.step, self.maxiter, self.tol = step, maxiter, tol

    def minimize(self, model, loss, q0, bounds=(None, None), callback=None):
        import torch
        dev = model.evaluator.device
        host = not isinstance(q0, torch.Tensor)
        lb = -float("inf") if bounds[0] is None else float(bounds[0])
        ub = float("inf") if bounds[1] is None else float(bounds[1])
        q = (torch.as_tensor(np.asarray(q0, dtype=np.float64)) if host else q0.detach()).to(dev, torch.float64).clamp(lb, ub).clone()
        value, grad = loss.value_and_grad_device(q, model)
        value, grad = value.clone(), grad.clone()
        step = torch.full((q.shape[0],), float(self.step), dtype=torch.float64, device=dev)
        history = [value.clone()]
        for it in range(self.maxiter):
            q_new = (q - step[:, None] * grad).clamp(lb, ub)
            value_new, grad_new = loss.value_and_grad_device(q_new, model)
            worse = value_new > value
            # Barzilai-Borwein step per design; halve where the loss went up and keep the old iterate there
            s, y = q_new - q, grad_new - grad
            sy = (s * y).sum(dim=1)
            bb = torch.where(sy > 0, (s * s).sum(dim=1) / torch.where(sy > 0, sy, torch.ones_like(sy)), step * 2.0)
            step = torch.where(worse, step * 0.5, bb.clamp(1e-12, 1e6))
            keep = ~worse
            q = torch.where(keep[:, None], q_new, q)
            grad = torch.where(keep[:, None], grad_new, grad)
            value = torch.where(keep, value_new, value)
            history.append(value.clone())
            if callback is not None:
                callback(q, value)
            done = ((history[-2] - history[-1]).abs() <= self.tol * history[-1].abs().clamp_min(1.0)).all() & ~worse.any()
            if bool(done.item()):                      # the one host read of an iteration
                break
        self.history = [h.cpu().numpy() for h in history]
        if host:
            return q.cpu().numpy(), value.cpu().numpy()
        return q, value


class BatchedLBFGS:
    """
    Lock-step projected L-BFGS over a batch of designs ``q[B, E]`` with box bounds, DEVICE RESIDENT.

    Every design keeps its own curvature pairs, step length and convergence flag; what is shared is the
    evaluation: one batched CUDA call per trial point returns the loss and gradient of all designs.  The
    search direction is the two-loop recursion on the free variables (those not pinned at a bound with
    the gradient pointing outward), the step is found by backtracking on the Armijo condition along the
    projected path, and designs that have converged stop moving.

    The whole state - ``q``, gradients, the curvature pairs, step lengths, masks - lives in torch tensors on the
    evaluator's GPU and the evaluation hands back device tensors (``Loss.value_and_grad_device``): nothing of size
    ``B x E`` crosses PCIe during the optimisation; the host reads one flag per iteration ("any design still active?").
    In a sharded run every rank optimises its own slice of the ensemble with its own state, so there is no traffic
    between GPUs either (``group``: the flag is all-reduced so that the ranks leave the loop together).

    ``fun(q) -> (loss[B], grad[B, E])`` is any callable with that contract on torch tensors (NumPy results are
    accepted and converted: CPU tests); ``minimize(model, loss, q0, ...)`` builds it from a ``Loss``.
    """
    def __init__(self, memory=8, maxiter=100, gtol=1e-6, ftol=1e-10, armijo=1e-4, max_backtracks=12, group=None):
        self.memory, self.maxiter, self.gtol, self.ftol = memory, maxiter, gtol, ftol
        self.armijo, self.max_backtracks = armijo, max_backtracks
        self.group = group

    def minimize(self, model, loss, q0, bounds=(None, None), callback=None, as_numpy=None):
        import torch
        device = model.evaluator.device
        host = not isinstance(q0, torch.Tensor)
        q0_t = torch.as_tensor(np.asarray(q0, dtype=np.float64)) if host else q0
        q, f = self.minimize_fun(lambda x: loss.value_and_grad_device(x, model), q0_t.to(device=device, dtype=torch.float64), bounds, callback)
        if as_numpy if as_numpy is not None else host:
            return q.cpu().numpy(), f.cpu().numpy()
        return q, f

    def minimize_fun(self, fun, q0, bounds=(None, None), callback=None):
        import torch
        host = not isinstance(q0, torch.Tensor)
        q = (torch.as_tensor(np.array(q0, dtype=np.float64)) if host else q0.detach().clone()).to(torch.float64)
        dev = q.device
        lb = -float("inf") if bounds[0] is None else float(bounds[0])
        ub = float("inf") if bounds[1] is None else float(bounds[1])
        q = q.clamp(lb, ub)
        B, E = q.shape

        def call(x):
            ft, gt = fun(x.cpu().numpy() if host else x)
            ft = ft if isinstance(ft, torch.Tensor) else torch.as_tensor(np.asarray(ft, dtype=np.float64))
            gt = gt if isinstance(gt, torch.Tensor) else torch.as_tensor(np.asarray(gt, dtype=np.float64))
            return ft.to(dev, torch.float64).clone(), gt.to(dev, torch.float64).clone()

        def any_true(mask):
            flag = mask.any().to(torch.int32).reshape(1)
            if self.group is not None or _dist_ready():
                import torch.distributed as dist
                if dev.type != "cuda" and dist.get_backend(self.group) == "nccl":
                    return bool(flag.item())
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            return bool(flag.item())                      # the one host read of an iteration

        f, g = call(q)
        S, Y = [], []                                    # curvature pairs, each [B, E] on the device
        active = torch.ones(B, dtype=torch.bool, device=dev)
        zero = torch.zeros((), dtype=torch.float64, device=dev)
        one = torch.ones((), dtype=torch.float64, device=dev)
        self.evaluations = 1
        history = [f.clone()]

        def projected_gradient(q, g):
            return torch.where(((q <= lb) & (g > 0)) | ((q >= ub) & (g < 0)), zero, g)

        for it in range(self.maxiter):
            pg = projected_gradient(q, g)
            active &= pg.abs().amax(dim=1) > self.gtol
            if not any_true(active):
                break
            free = pg != 0.0
            # two-loop recursion, vectorised over designs
            d = pg.clone()
            alphas = []
            for s_k, y_k in zip(reversed(S), reversed(Y)):
                sy = (s_k * y_k).sum(dim=1)
                ok = sy > 1e-300
                a = torch.where(ok, (s_k * d).sum(dim=1) / torch.where(ok, sy, one), zero)
                d -= a[:, None] * y_k
                alphas.append((a, ok, sy))
            if S:
                sy = (S[-1] * Y[-1]).sum(dim=1)
                yy = (Y[-1] * Y[-1]).sum(dim=1)
                gamma = torch.where((sy > 1e-300) & (yy > 0), sy / torch.where(yy > 0, yy, one), one)
                d *= gamma[:, None]
            for (a, ok, sy), s_k, y_k in zip(reversed(alphas), S, Y):
                b = torch.where(ok, (y_k * d).sum(dim=1) / torch.where(ok, sy, one), zero)
                d += (a - b)[:, None] * s_k
            d = -torch.where(free, d, zero)
            slope = (pg * d).sum(dim=1)
            bad = slope >= 0                              # not a descent direction: fall back to steepest descent
            d = torch.where(bad[:, None], -pg, d)
            # backtracking along the projected path; per-design step lengths, one batched evaluation per trial
            if S:
                t = torch.ones(B, dtype=torch.float64, device=dev)
            else:
                t = torch.minimum(one, 1.0 / pg.abs().amax(dim=1).clamp_min(1e-300))
            todo = active.clone()
            q_new, f_new, g_new = q.clone(), f.clone(), g.clone()
            for _ in range(self.max_backtracks):
                trial = torch.where(todo[:, None], (q + t[:, None] * d).clamp(lb, ub), q_new)
                ft, gt = call(trial)
                self.evaluations += 1
                decrease = (pg * (trial - q)).sum(dim=1)
                accept = todo & torch.isfinite(ft) & (ft <= f + self.armijo * decrease)
                q_new = torch.where(accept[:, None], trial, q_new)
                g_new = torch.where(accept[:, None], gt, g_new)
                f_new = torch.where(accept, ft, f_new)
                todo &= ~accept
                if not any_true(todo):
                    break
                t = torch.where(todo, 0.5 * t, t)
            active &= ~todo                               # no acceptable step: that design is done
            s_k, y_k = q_new - q, g_new - g
            moved = (s_k != 0.0).any(dim=1)
            small = (f - f_new).abs() <= self.ftol * f_new.abs().clamp_min(1.0)
            active &= ~(moved & small)
            S.append(torch.where(moved[:, None], s_k, zero))
            Y.append(torch.where(moved[:, None], y_k, zero))
            if len(S) > self.memory:
                S.pop(0)
                Y.pop(0)
            q, f, g = q_new, f_new, g_new
            history.append(f.clone())
            if callback is not None:
                callback(q, f)
        self.history = [h.cpu().numpy() for h in history]
        self.iterations = len(history) - 1
        if host:
            return q.cpu().numpy(), f.cpu().numpy()
        return q, f


def _dist_ready():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False
