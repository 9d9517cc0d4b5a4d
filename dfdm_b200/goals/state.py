from dataclasses import dataclass

import numpy as np

__all__ = ["GoalState"]


@dataclass
class GoalState:
    prediction: np.ndarray
    target: np.ndarray
    weight: np.ndarray
