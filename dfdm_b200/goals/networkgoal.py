"""Network goals (reference: src/dfdm/goals/networkgoal/loadpathgoal.py)."""
import numpy as np

from dfdm_b200.goals.goal import NetworkGoal, ScalarGoal

__all__ = ["NetworkLoadPathGoal"]


class NetworkLoadPathGoal(ScalarGoal, NetworkGoal):
    """Make the total load path, sum |force x length| over all edges, reach a magnitude."""
    KIND = "NetworkLoadPath"

    def __init__(self, target=None, weight=1.0):
        super().__init__(key=None, target=target, weight=weight)

    def prediction(self, eq_state, *args, **kwargs):
        return np.atleast_1d(np.sum(np.abs(eq_state.lengths * eq_state.forces)))

    def params(self):
        # target=None is only meaningful under PredictionError, which ignores the target (losses.py:64)
        return [0.0 if self._target is None else float(self._target)]
