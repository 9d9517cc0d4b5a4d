import numpy as np

from dfdm_b200.goals.state import GoalState

__all__ = ["goals_state"]


def goals_state(goals, eqstate, model):
    """Collate goal states into flat vectors in goal-list order (goals/helpers.py:19-37); host-side inspection only."""
    predictions, targets, weights = [], [], []
    for goal in goals:
        gstate = goal(eqstate, model)
        predictions.append(np.atleast_1d(gstate.prediction))
        targets.append(np.atleast_1d(gstate.target))
        weights.append(np.atleast_1d(gstate.weight))
    return GoalState(prediction=np.concatenate(predictions, axis=0), target=np.concatenate(targets, axis=0),
                     weight=np.concatenate(weights, axis=0))
