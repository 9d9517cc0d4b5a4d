"""Node goals (reference: src/dfdm/goals/nodegoal/pointgoal.py, residualgoal.py)."""
import numpy as np

from dfdm_b200.goals.goal import NodeGoal, ScalarGoal, VectorGoal

__all__ = ["NodePointGoal", "NodeLineGoal", "NodePlaneGoal", "NodeResidualForceGoal", "NodeResidualVectorGoal",
           "NodeResidualDirectionGoal"]


def _as_point(p):
    return [float(p[0]), float(p[1]), float(p[2])]


class NodePointGoal(VectorGoal, NodeGoal):
    """Pull a node to target xyz coordinates."""
    KIND = "NodePoint"

    def __init__(self, key, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        return eq_state.xyz[index, :]


class NodeLineGoal(NodePointGoal):
    """Pull a node to its closest point on a line given by two points (pointgoal.py:31-44)."""
    KIND = "NodeLine"

    def target(self, prediction):
        a, b = (np.asarray(_as_point(p)) for p in self._line())
        ab = b - a
        return a + ab * (np.dot(np.asarray(prediction) - a, ab) / np.dot(ab, ab))

    def _line(self):
        line = self._target
        if hasattr(line, "start") and hasattr(line, "end"):
            return line.start, line.end
        return line[0], line[1]

    def params(self):
        a, b = self._line()
        return _as_point(a) + _as_point(b)


class NodePlaneGoal(NodePointGoal):
    """Pull a node to its closest point on a plane given as (origin, normal) (pointgoal.py:47-61)."""
    KIND = "NodePlane"

    def target(self, prediction):
        origin, normal = (np.asarray(_as_point(p)) for p in self._plane())
        n = normal / np.linalg.norm(normal)
        p = np.asarray(prediction)
        return p - n * ((np.dot(n, p) - np.dot(n, origin)) / np.dot(n, n))

    def _plane(self):
        plane = self._target
        if hasattr(plane, "point") and hasattr(plane, "normal"):
            return plane.point, plane.normal
        return plane[0], plane[1]

    def params(self):
        origin, normal = self._plane()
        return _as_point(origin) + _as_point(normal)


class NodeResidualForceGoal(ScalarGoal, NodeGoal):
    """Make the magnitude of the residual (reaction) force at a node match a non-negative value."""
    KIND = "NodeResidualForce"

    def __init__(self, key, target, weight=1.0):
        assert target >= 0.0, "Only non-negative target forces are supported!"
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        return np.atleast_1d(np.linalg.norm(eq_state.residuals[index, :]))


class NodeResidualVectorGoal(VectorGoal, NodeGoal):
    """Make the residual force at a node match a vector."""
    KIND = "NodeResidualVector"

    def __init__(self, key, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        return eq_state.residuals[index, :]


class NodeResidualDirectionGoal(VectorGoal, NodeGoal):
    """Make the residual force at a node parallel to a vector (both are unitised, residualgoal.py:42-70)."""
    KIND = "NodeResidualDirection"

    def __init__(self, key, target, weight=1.0):
        super().__init__(key=key, target=target, weight=weight)

    def prediction(self, eq_state, index):
        residual = eq_state.residuals[index, :]
        return residual / np.linalg.norm(residual)

    def target(self, prediction):
        return np.asarray(self._target, dtype=np.float64) / np.linalg.norm(self._target)
