"""
Goals: what a designer wants a node, an edge or the whole network to do in equilibrium.

Class names and constructor signatures are those of the reference (src/dfdm/goals/*.py).  On the hot
path a goal is not evaluated in Python at all: ``record(model)`` flattens it into one row of the
device goal table (kind, element row, weight, parameters) and the CUDA kernels compute prediction,
target, loss contribution and cotangents (include/dfdm_b200.h, DFDM_GOAL_*).

``goal(eqstate, model)`` (goal.py:23-33 in the reference) is kept for inspecting an already computed
state on the host; for the library goals it plays no part in ``Loss.__call__`` or in the gradient.

USER-DEFINED goals (the reference differentiates any ``Goal`` subclass through autograd, goal.py:40-68): subclass
``NodeGoal`` / ``EdgeGoal`` / ``NetworkGoal`` with the ``ScalarGoal`` / ``VectorGoal`` mix-ins, leave ``KIND`` unset
and write ``prediction(eq_state, index)`` with torch operations - ``eq_state`` then holds torch tensors on the GPU
that are attached to the CUDA path through ``dfdm_b200.autograd.EquilibriumFunction`` (forward: dfdm_forward,
backward: dfdm_vjp).  ``Loss`` sends such goals down that route by itself; see losses.py.
"""
import numpy as np

from dfdm_b200.goals.state import GoalState

__all__ = ["Goal", "ScalarGoal", "VectorGoal", "NodeGoal", "EdgeGoal", "NetworkGoal"]


class Goal:
    KIND = None          # name of the device goal kind (dfdm_b200._lib.GOAL_KINDS)

    def __init__(self, key, target, weight):
        self._key = key
        self._target = target
        self._weight = weight

    def __call__(self, eqstate, model):
        prediction = self.prediction(eqstate, self.index(model))
        target = self.target(prediction)
        weight = self.weight()
        return GoalState(target=target, prediction=prediction, weight=weight)

    def key(self):
        return self._key

    def weight(self):
        raise NotImplementedError

    def prediction(self, eq_state, index):
        raise NotImplementedError

    def target(self, prediction):
        raise NotImplementedError

    def index(self, model):
        raise NotImplementedError

    def params(self):
        """Parameters of the device record (at most 6 doubles)."""
        raise NotImplementedError

    def record(self, model):
        """``(kind, element row or None, weight, params)`` for dfdm_b200.engine.GoalProgram."""
        if self.KIND is None:
            raise NotImplementedError("%s has no record in the device goal table; user-defined goals are evaluated on the torch "
                                      "tape over the CUDA forward / adjoint (Loss does that by itself)" % type(self).__name__)
        return self.KIND, self.index(model), float(self._weight), self.params()


class ScalarGoal:
    """A goal on one scalar quantity (goal.py:72-90)."""
    def weight(self):
        return np.atleast_1d(self._weight)

    def target(self, prediction):
        return np.atleast_1d(self._target)

    def params(self):
        return [float(self._target)]


class VectorGoal:
    """A goal on a 3-vector; the weight applies to each component (goal.py:93-111)."""
    def weight(self):
        return np.asarray([self._weight] * 3, dtype=np.float64)

    def target(self, prediction):
        return np.asarray(self._target, dtype=np.float64)

    def params(self):
        return [float(c) for c in self._target]


class NodeGoal(Goal):
    def __init__(self, key, target, weight):
        super().__init__(key=key, target=target, weight=weight)

    def index(self, model):
        return model.structure.node_index[self.key()]


class EdgeGoal(Goal):
    def __init__(self, key, target, weight):
        super().__init__(key=key, target=target, weight=weight)

    def index(self, model):
        return model.structure.edge_index[self.key()]


class NetworkGoal(Goal):
    def __init__(self, key, target, weight):
        super().__init__(key=key, target=target, weight=weight)

    def index(self, model):
        return None

    def key(self):
        return None
