"""
The calculator (reference: src/dfdm/equilibrium/model.py:12-79), evaluated on the GPU.

``model(q)`` with ``q[E]`` returns one ``EquilibriumState`` of NumPy arrays; ``q[B, E]`` returns a
batched state.  A singular stiffness matrix raises ``numpy.linalg.LinAlgError`` as in model.py:55.
"""
import numpy as np

from dfdm_b200.engine import Evaluator
from dfdm_b200.equilibrium.state import EquilibriumState
from dfdm_b200.equilibrium.structure import EquilibriumStructure

__all__ = ["EquilibriumModel"]


class EquilibriumModel:
    def __init__(self, network, device=None, solver="auto", **solver_options):
        self.structure = EquilibriumStructure(network)
        self.loads = np.asarray(list(network.nodes_loads()), dtype=np.float64).reshape(-1, 3)
        self.xyz_fixed = np.asarray([network.node_coordinates(node) for node in network.nodes_fixed()], dtype=np.float64).reshape(-1, 3)
        self._device = device
        self._solver = solver
        self._solver_options = solver_options
        self._evaluator = None

    @property
    def plan(self):
        return self.structure.plan

    @property
    def evaluator(self):
        """The CUDA evaluator, created on first use (raises without a GPU: there is no CPU path)."""
        if self._evaluator is None:
            self._evaluator = Evaluator(self.plan, self.loads, self.xyz_fixed, device=self._device, solver=self._solver,
                                        **self._solver_options)
        return self._evaluator

    def __call__(self, q):
        q = np.asarray(q, dtype=np.float64)
        single = q.ndim == 1
        out = self.evaluator.forward(q, rescue=True)
        Evaluator.raise_on_status(out["status"])
        arrays = {k: out[k].cpu().numpy() for k in Evaluator.STATE}
        if single:
            arrays = {k: v[0] for k, v in arrays.items()}
        return EquilibriumState(xyz=arrays["xyz"], residuals=arrays["residuals"], vectors=arrays["vectors"],
                                lengths=arrays["lengths"], forces=arrays["forces"], force_densities=q)
