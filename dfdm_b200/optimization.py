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

    def minimize(self, network, loss, bounds=(None, None), constraints=None, maxiter=100, tol=1e-6, verbose=True, callback=None):
        """Minimise ``loss`` over the edge force densities of ``network``; returns the optimal q[E]."""
        name = self.name
        q = np.asarray(network.edges_forcedensities(), dtype=np.float64)
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
        dq, status = model.evaluator.vjp(np.repeat(q[None, :], total, axis=0), **stacked)
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
    projection onto the bounds are elementwise.  For ensembles / parameter sweeps on one or many GPUs.
    """
    def __init__(self, step=1e-2, maxiter=100, tol=1e-6):
        self.step, self.maxiter, self.tol = step, maxiter, tol

    def minimize(self, model, loss, q0, bounds=(None, None), callback=None):
        lb = -np.inf if bounds[0] is None else bounds[0]
        ub = np.inf if bounds[1] is None else bounds[1]
        q = np.clip(np.array(q0, dtype=np.float64, copy=True), lb, ub)
        value, grad = loss.value_and_grad(q, model)
        step = np.full(q.shape[0], self.step)
        history = [np.array(value)]
        for it in range(self.maxiter):
            q_new = np.clip(q - step[:, None] * grad, lb, ub)
            value_new, grad_new = loss.value_and_grad(q_new, model)
            worse = value_new > value
            # Barzilai-Borwein step per design; halve where the loss went up and keep the old iterate there
            s, y = q_new - q, grad_new - grad
            sy = np.sum(s * y, axis=1)
            bb = np.where(sy > 0, np.sum(s * s, axis=1) / np.where(sy > 0, sy, 1.0), step * 2.0)
            step = np.where(worse, step * 0.5, np.clip(bb, 1e-12, 1e6))
            keep = ~worse
            q[keep], grad[keep] = q_new[keep], grad_new[keep]
            value = np.where(keep, value_new, value)
            history.append(np.array(value))
            if callback is not None:
                callback(q, value)
            if np.all(np.abs(history[-2] - history[-1]) <= self.tol * np.maximum(1.0, np.abs(history[-1]))) and not np.any(worse):
                break
        self.history = history
        return q, value


class BatchedLBFGS:
    """
    Lock-step projected L-BFGS over a batch of designs ``q[B, E]`` with box bounds.

    Every design keeps its own curvature pairs, step length and convergence flag; what is shared is the
    evaluation: one batched CUDA call per trial point returns the loss and gradient of all designs.  The
    search direction is the two-loop recursion on the free variables (those not pinned at a bound with
    the gradient pointing outward), the step is found by backtracking on the Armijo condition along the
    projected path, and designs that have converged stop moving.  ``fun(q) -> (loss[B], grad[B, E])`` is
    any callable with that contract; ``minimize(model, loss, q0, ...)`` builds it from a ``Loss``.
    """
    def __init__(self, memory=8, maxiter=100, gtol=1e-6, ftol=1e-10, armijo=1e-4, max_backtracks=12):
        self.memory, self.maxiter, self.gtol, self.ftol = memory, maxiter, gtol, ftol
        self.armijo, self.max_backtracks = armijo, max_backtracks

    def minimize(self, model, loss, q0, bounds=(None, None), callback=None):
        return self.minimize_fun(lambda q: loss.value_and_grad(q, model), q0, bounds, callback)

    def minimize_fun(self, fun, q0, bounds=(None, None), callback=None):
        lb = -np.inf if bounds[0] is None else bounds[0]
        ub = np.inf if bounds[1] is None else bounds[1]
        q = np.clip(np.array(q0, dtype=np.float64, copy=True), lb, ub)
        B, E = q.shape
        f, g = fun(q)
        f, g = np.array(f, dtype=np.float64), np.array(g, dtype=np.float64)
        S, Y = [], []                                    # curvature pairs, each [B, E]; rho per design
        active = np.ones(B, dtype=bool)
        self.evaluations = 1
        history = [f.copy()]

        def projected_gradient(q, g):
            pg = g.copy()
            pg[(q <= lb) & (g > 0)] = 0.0
            pg[(q >= ub) & (g < 0)] = 0.0
            return pg

        for it in range(self.maxiter):
            pg = projected_gradient(q, g)
            active &= np.max(np.abs(pg), axis=1) > self.gtol
            if not active.any():
                break
            free = pg != 0.0
            # two-loop recursion, vectorised over designs
            d = pg.copy()
            alphas = []
            for s_k, y_k in zip(reversed(S), reversed(Y)):
                sy = np.sum(s_k * y_k, axis=1)
                ok = sy > 1e-300
                a = np.where(ok, np.sum(s_k * d, axis=1) / np.where(ok, sy, 1.0), 0.0)
                d -= a[:, None] * y_k
                alphas.append((a, ok, sy))
            if S:
                sy = np.sum(S[-1] * Y[-1], axis=1)
                yy = np.sum(Y[-1] * Y[-1], axis=1)
                gamma = np.where((sy > 1e-300) & (yy > 0), sy / np.where(yy > 0, yy, 1.0), 1.0)
                d *= gamma[:, None]
            for (a, ok, sy), s_k, y_k in zip(reversed(alphas), S, Y):
                b = np.where(ok, np.sum(y_k * d, axis=1) / np.where(ok, sy, 1.0), 0.0)
                d += (a - b)[:, None] * s_k
            d = -np.where(free, d, 0.0)
            slope = np.sum(pg * d, axis=1)
            bad = slope >= 0                              # not a descent direction: fall back to steepest descent
            d[bad] = -pg[bad]
            slope = np.sum(pg * d, axis=1)
            # backtracking along the projected path; per-design step lengths, one batched evaluation per trial
            t = np.ones(B) if S else np.minimum(1.0, 1.0 / np.maximum(np.max(np.abs(pg), axis=1), 1e-300))
            todo = active.copy()
            q_new, f_new, g_new = q.copy(), f.copy(), g.copy()
            for _ in range(self.max_backtracks):
                trial = np.where(todo[:, None], np.clip(q + t[:, None] * d, lb, ub), q_new)
                ft, gt = fun(trial)
                self.evaluations += 1
                ft, gt = np.asarray(ft, dtype=np.float64), np.asarray(gt, dtype=np.float64)
                decrease = np.sum(pg * (trial - q), axis=1)
                accept = todo & np.isfinite(ft) & (ft <= f + self.armijo * decrease)
                q_new[accept], f_new[accept], g_new[accept] = trial[accept], ft[accept], gt[accept]
                todo &= ~accept
                if not todo.any():
                    break
                t = np.where(todo, 0.5 * t, t)
            active &= ~todo                               # no acceptable step: that design is done
            s_k, y_k = q_new - q, g_new - g
            moved = np.any(s_k != 0.0, axis=1)
            small = np.abs(f - f_new) <= self.ftol * np.maximum(1.0, np.abs(f_new))
            active &= ~(moved & small)
            S.append(np.where(moved[:, None], s_k, 0.0))
            Y.append(np.where(moved[:, None], y_k, 0.0))
            if len(S) > self.memory:
                S.pop(0)
                Y.pop(0)
            q, f, g = q_new, f_new, g_new
            history.append(f.copy())
            if callback is not None:
                callback(q, f)
        self.history = history
        self.iterations = len(history) - 1
        return q, f
