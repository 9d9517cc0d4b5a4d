"""
Loss functions (reference: src/dfdm/losses/losses.py:10-87, losses/regularizers.py:8-20).

``Loss(*terms)(q, model)`` keeps the reference's signature and meaning — equilibrium for ``q`` then
the sum of the terms — but the whole evaluation is one call into the CUDA path: the terms and their
goals are flattened once per model into a device goal program, and value and gradient come from the
same fused pass (``value_and_grad``), replacing ``autograd.grad(loss)`` of optimization.py:61.
"""
import weakref

import numpy as np

from dfdm_b200.engine import Evaluator, GoalProgram

__all__ = ["LossTerm", "SquaredError", "MeanSquaredError", "PredictionError", "Loss", "Regularizer", "L2Regularizer"]


class LossTerm:
    TERM_KIND = None

    def __init__(self, goals, alpha=1.0, name=None, *args, **kwargs):
        self.goals = goals
        self.alpha = alpha
        self._name = name

    @property
    def name(self):
        if not self._name:
            self._name = self.__class__.__name__
        return self._name

    def __call__(self, eqstate, model):
        """Value of this term for the design of an already computed state (one GPU evaluation of the term alone)."""
        return Loss(self)(eqstate.force_densities, model)

    def goals_state(self, eqstate, model):
        from dfdm_b200.goals import goals_state
        return goals_state(self.goals, eqstate, model)


class SquaredError(LossTerm):
    """alpha * sum(weight * (prediction - target)^2)"""
    TERM_KIND = "SquaredError"


class MeanSquaredError(LossTerm):
    """alpha * mean(weight * (prediction - target)^2) over all entries of the term"""
    TERM_KIND = "MeanSquaredError"


class PredictionError(LossTerm):
    """alpha * sum(prediction)"""
    TERM_KIND = "PredictionError"


class Regularizer:
    pass


class L2Regularizer(Regularizer):
    """alpha * sum(q^2)"""
    def __init__(self, alpha, name=None):
        self.alpha = alpha
        self._name = name

    @property
    def name(self):
        if not self._name:
            self._name = self.__class__.__name__
        return self._name

    def __call__(self, eqstate, model):
        return Loss(self)(eqstate.force_densities, model)


class Loss:
    def __init__(self, *args, name=None):
        self.loss_terms = args
        self._name = name
        self._programs = weakref.WeakKeyDictionary()
        self._last = None

    @property
    def name(self):
        if not self._name:
            self._name = self.__class__.__name__
        return self._name

    # -- compilation --------------------------------------------------------------------------------

    def program(self, model):
        """The device goal program of this loss for ``model`` (built once, cached)."""
        prog = self._programs.get(model)
        if prog is None:
            goals, terms, l2 = [], [], 0.0
            for term in self.loss_terms:
                if isinstance(term, Regularizer):
                    l2 += float(term.alpha)
                    continue
                if getattr(term, "TERM_KIND", None) is None:
                    raise NotImplementedError("loss term %r has no device kernel" % type(term).__name__)
                t = len(terms)
                terms.append((term.TERM_KIND, float(term.alpha)))
                for goal in term.goals:
                    kind, index, weight, params = goal.record(model)
                    goals.append((kind, index, t, weight, params))
            prog = GoalProgram(model.plan, goals, terms, l2)
            self._programs[model] = prog
        return prog

    # -- evaluation ---------------------------------------------------------------------------------

    def _evaluate(self, q, model, grad):
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
        key = (id(model), q.shape, q.tobytes())
        if self._last is not None and self._last[0] == key and (self._last[2] is not None or not grad):
            return self._last[1], self._last[2]
        out = model.evaluator.loss_and_grad(self.program(model), q, grad=grad)
        Evaluator.raise_on_status(out["status"])
        loss = out["loss"].cpu().numpy()
        dq = out["dq"].cpu().numpy() if grad else None
        if q.ndim == 1:
            loss, dq = float(loss[0]), (dq[0] if grad else None)
        self._last = (key, loss, dq)
        return loss, dq

    def __call__(self, q, model):
        """The loss of design ``q[E]`` (a float) or of a batch ``q[B, E]`` (an array)."""
        return self._evaluate(q, model, grad=False)[0]

    def value_and_grad(self, q, model):
        """Loss and dL/dq from one fused forward + adjoint pass."""
        return self._evaluate(q, model, grad=True)

    def grad(self, q, model):
        return self._evaluate(q, model, grad=True)[1]
