"""Edge goals (reference: src/dfdm/goals/edgegoal/*.py)."""
import numpy as np

from dfdm_b200.goals.goal import EdgeGoal, ScalarGoal, VectorGoal

__all__ = ["EdgeLengthGoal", "EdgeForceGoal", "EdgeLoadPathGoal", "EdgeVectorAngleGoal", "EdgeDirectionGoal"]


class EdgeLengthGoal(ScalarGoal, EdgeGoal):
    """Make an edge reach a length."""
    KIND = "EdgeLength"

    def __init__(self, key, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        return np.atleast_1d(eq_state.lengths[index])


class EdgeForceGoal(ScalarGoal, EdgeGoal):
    """Make an edge reach a force."""
    KIND = "EdgeForce"

    def __init__(self, key, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        return np.atleast_1d(eq_state.forces[index])


class EdgeLoadPathGoal(ScalarGoal, EdgeGoal):
    """Make |force x length| of an edge reach a magnitude."""
    KIND = "EdgeLoadPath"

    def __init__(self, key, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        return np.atleast_1d(np.abs(eq_state.lengths[index] * eq_state.forces[index]))


class EdgeVectorAngleGoal(ScalarGoal, EdgeGoal):
    """Make the angle (degrees) between an edge and a vector reach a target (anglegoal.py:7-32)."""
    KIND = "EdgeVectorAngle"

    def __init__(self, key, vector, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)
        self.vector_other = np.asarray(vector, dtype=np.float64)

    def prediction(self, eq_state, index):
        return np.atleast_1d(self._angle_vectors_numpy(eq_state.vectors[index, :], self.vector_other, deg=True))

    @staticmethod
    def _angle_vectors_numpy(u, v, deg=False):
        a = np.dot(u, v) / (np.linalg.norm(u) * np.linalg.norm(v))
        a = max(min(a, 1), -1)
        return np.degrees(np.arccos(a)) if deg else np.arccos(a)

    def params(self):
        return [float(self._target)] + [float(c) for c in self.vector_other]


class EdgeDirectionGoal(VectorGoal, EdgeGoal):
    """Make an edge parallel to a vector (both unitised, directiongoal.py:7-25)."""
    KIND = "EdgeDirection"

    def __init__(self, key, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        vector = eq_state.vectors[index, :]
        return vector / np.linalg.norm(vector)

    def target(self, prediction):
        return np.asarray(self._target, dtype=np.float64) / np.linalg.norm(self._target)
