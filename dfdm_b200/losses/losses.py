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


_CACHE_BYTES = 1 << 20      # largest q whose last evaluation is remembered (a megabyte: any single design)


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

    # -- user-defined goals / terms: the torch route (dfdm_b200.autograd) -------------------------------------------

    def on_device(self):
        """True when the CUDA goal program can evaluate this term: a library term kind whose goals are all library goals."""
        return type(self).TERM_KIND is not None and type(self).loss is LossTerm.loss and \
            all(getattr(goal, "KIND", None) is not None for goal in self.goals)

    def loss(self, gstate):
        """The term as a function of the collated goal state (losses.py:22-31 in the reference).  Library terms have a
        closed form on the GPU; this torch version serves terms whose goals are user-defined, and is what a user-defined
        term overrides."""
        import torch
        err = gstate.weight * torch.square(gstate.prediction - gstate.target)
        if self.TERM_KIND == "SquaredError":
            return self.alpha * torch.sum(err)
        if self.TERM_KIND == "MeanSquaredError":
            return self.alpha * torch.mean(err)
        if self.TERM_KIND == "PredictionError":
            return self.alpha * torch.sum(gstate.prediction)
        raise NotImplementedError("%s must implement loss(gstate) with torch operations" % type(self).__name__)

    def torch_value(self, eqstate, model):
        """Value of the term for a differentiable state (torch tensors): goals collated in list order, then ``loss``."""
        import torch
        from dfdm_b200.goals.state import GoalState
        as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.atleast_1d(np.asarray(a, dtype=np.float64)),  # noqa: E731
                                                                                device=eqstate.xyz.device)
        preds, targets, weights = [], [], []
        for goal in self.goals:
            if getattr(goal, "KIND", None) is not None:
                raise NotImplementedError("loss term %r mixes library goals with user-defined ones: library goals are evaluated by "
                                          "the CUDA goal program, user-defined ones on the torch tape - put them in separate terms" % self.name)
            gstate = goal(eqstate, model)
            preds.append(torch.atleast_1d(as_t(gstate.prediction)))
            targets.append(torch.atleast_1d(as_t(gstate.target)))
            weights.append(torch.atleast_1d(as_t(gstate.weight)))
        return self.loss(GoalState(prediction=torch.cat(preds), target=torch.cat(targets), weight=torch.cat(weights)))


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
                if not term.on_device():
                    continue                            # user-defined: evaluated on the torch tape (_evaluate)
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
        # SciPy asks for fun(q) and jac(q) separately at the same iterate: remember the last single-design evaluation.
        # Batches are never cached (comparing them would mean a pass over hundreds of megabytes per call).
        key = (id(model), q.shape, q.tobytes()) if q.nbytes <= _CACHE_BYTES else None
        if key is not None and self._last is not None and self._last[0] == key and (self._last[2] is not None or not grad):
            return self._last[1], self._last[2]
        out = model.evaluator.loss_and_grad(self.program(model), q, grad=grad, rescue=True)
        Evaluator.raise_on_status(out["status"])
        loss = out["loss"].cpu().numpy()
        dq = out["dq"].cpu().numpy() if grad else None
        custom = [term for term in self.loss_terms if not isinstance(term, Regularizer) and not term.on_device()]
        if custom:
            # user-defined goals / terms: torch differentiates their code down to the state, dfdm_vjp takes it to q
            import torch
            from dfdm_b200.autograd import equilibrium
            for b, qb in enumerate(np.atleast_2d(q)):
                qt = torch.tensor(qb, dtype=torch.float64, device=model.evaluator.device, requires_grad=grad)
                state = equilibrium(qt, model)
                value = sum(term.torch_value(state, model) for term in custom)
                loss[b] += float(value.detach())
                if grad:
                    value.backward()
                    dq[b] += qt.grad.cpu().numpy()
        if q.ndim == 1:
            loss, dq = float(loss[0]), (dq[0] if grad else None)
        self._last = (key, loss, dq) if key is not None else None
        return loss, dq

    def value_and_grad_device(self, q, model):
        """``q[B, E]`` torch tensor on the model's GPU -> (loss[B], dq[B, E]) torch tensors on the same GPU: nothing crosses
        PCIe but the B status words.  The device-resident optimisers call this (optimization.BatchedLBFGS)."""
        custom = [term for term in self.loss_terms if not isinstance(term, Regularizer) and not term.on_device()]
        if custom:
            raise NotImplementedError("user-defined goals / terms are evaluated design by design on the torch tape: use value_and_grad")
        out = model.evaluator.loss_and_grad(self.program(model), q, grad=True)
        Evaluator.raise_on_status(out["status"])
        return out["loss"], out["dq"]

    def __call__(self, q, model):
        """The loss of design ``q[E]`` (a float) or of a batch ``q[B, E]`` (an array)."""
        return self._evaluate(q, model, grad=False)[0]

    def value_and_grad(self, q, model):
        """Loss and dL/dq from one fused forward + adjoint pass."""
        return self._evaluate(q, model, grad=True)

    def grad(self, q, model):
        return self._evaluate(q, model, grad=True)[1]
