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
           "BatchedGradientDescent"]


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
        if not constraints:
            return ()
        clist = []
        for constraint in constraints:
            fun = (lambda x, c=constraint: np.atleast_1d(c(x, model)))
            jac = (lambda x, c=constraint: c.jacobian(x, model))
            clist.append(NonlinearConstraint(fun=fun, jac=jac, lb=constraint.bound_low, ub=constraint.bound_up))
        return clist


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
