"""Result record of an equilibrium evaluation (reference: src/dfdm/equilibrium/state.py:11-18)."""
from dataclasses import dataclass

import numpy as np

__all__ = ["EquilibriumState"]


@dataclass
class EquilibriumState:
    """
    Arrays in node / edge row order.  For a single design: xyz n x 3, residuals n x 3, vectors E x 3,
    lengths E, forces E, force_densities E.  For a batch every array gains a leading dimension B.
    """
    xyz: np.ndarray
    residuals: np.ndarray
    vectors: np.ndarray
    lengths: np.ndarray
    forces: np.ndarray
    force_densities: np.ndarray

    @property
    def is_batched(self):
        return np.ndim(self.force_densities) == 2

    def design(self, b):
        """The state of design ``b`` of a batched state."""
        if not self.is_batched:
            raise ValueError("not a batched state")
        return EquilibriumState(self.xyz[b], self.residuals[b], self.vectors[b], self.lengths[b], self.forces[b],
                                self.force_densities[b])
