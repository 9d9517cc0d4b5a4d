"""
Topology and indexing of a network (reference: src/dfdm/equilibrium/structure.py:11-106).

Same lazy properties as the reference; in addition ``plan`` is the compiled topology the CUDA kernels
consume.  ``connectivity`` is materialised as a dense array only when somebody asks for it — the
kernels never do (a 66 564-node grid would need 35 GB).
"""
from collections import OrderedDict

import numpy as np

from dfdm_b200.engine import Plan

__all__ = ["EquilibriumStructure", "cached_plan"]

_PLAN_CACHE = OrderedDict()
_PLAN_CACHE_SIZE = 16


def cached_plan(n, edges, is_fixed):
    """
    Plans are pure functions of the topology, and the reference rebuilds its model in every ``fdm()``,
    ``minimize()`` and example loop (fdm.py:15, optimization.py:50): keep the last few compiled plans,
    keyed by the exact bytes of ``(n, edges, is_fixed)``.
    """
    edges = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
    is_fixed = np.ascontiguousarray(np.asarray(is_fixed).astype(np.uint8))
    key = (int(n), edges.tobytes(), is_fixed.tobytes())
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        plan = Plan(n, edges, is_fixed)
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _PLAN_CACHE.popitem(last=False)
    else:
        _PLAN_CACHE.move_to_end(key)
    return plan


class EquilibriumStructure:
    def __init__(self, network):
        self._network = network
        self._node_index = None
        self._edge_index = None
        self._edges = None
        self._plan = None
        self._connectivity = None

    @property
    def network(self):
        return self._network

    @property
    def node_index(self):
        """node key -> row (structure.py:36-43)."""
        if not self._node_index:
            self._node_index = self.network.key_index()
        return self._node_index

    @property
    def edge_index(self):
        """edge key (u, v) -> row (structure.py:45-52)."""
        if not self._edge_index:
            self._edge_index = self.network.uv_index()
        return self._edge_index

    @property
    def edges(self):
        """int32[E, 2] rows (u, v) in node rows: the branch-node matrix in coordinate form."""
        if self._edges is None:
            node_idx = self.node_index
            self._edges = np.asarray([(node_idx[u], node_idx[v]) for u, v in self.network.edges()], dtype=np.int32).reshape(-1, 2)
        return self._edges

    @property
    def plan(self):
        if self._plan is None:
            network = self.network
            is_fixed = np.asarray([1 if network.node_attribute(key, "is_support") else 0 for key in network.nodes()], dtype=np.uint8)
            self._plan = cached_plan(len(is_fixed), self.edges, is_fixed)
        return self._plan

    @property
    def connectivity(self):
        """Dense E x n int32 branch-node matrix, C[e,u] = -1, C[e,v] = +1 (structure.py:63-72)."""
        if self._connectivity is None:
            edges = self.edges
            C = np.zeros((edges.shape[0], self.plan.n), dtype=np.int32)
            rows = np.arange(edges.shape[0])
            np.add.at(C, (rows, edges[:, 0]), -1)
            np.add.at(C, (rows, edges[:, 1]), 1)
            self._connectivity = C
        return self._connectivity

    @property
    def free_nodes(self):
        return self.plan.free.tolist()

    @property
    def fixed_nodes(self):
        return self.plan.fixed.tolist()

    @property
    def freefixed_nodes(self):
        return self.plan.freefixed.tolist()
