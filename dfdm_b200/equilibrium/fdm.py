"""Entry points (reference: src/dfdm/equilibrium/fdm.py:10-84): same names, same signatures."""
import numpy as np

from dfdm_b200.equilibrium.model import EquilibriumModel

__all__ = ["fdm", "constrained_fdm", "network_update", "network_updated", "fdm_batch", "NetworkBatch"]


def _fdm(network, q, model=None):
    model = model or EquilibriumModel(network)
    eq_state = model(q)
    return network_updated(network, eq_state)


def fdm(network, model=None):
    """Equilibrium of a network for the force densities stored on its edges."""
    q = np.asarray(network.edges_forcedensities(), dtype=np.float64)
    return _fdm(network, q, model)


def constrained_fdm(network, optimizer, loss, bounds=(None, None), constraints=None, maxiter=100, tol=1e-6, callback=None):
    """Minimise ``loss`` over the force densities, then return the network in the optimal equilibrium state."""
    q_opt = optimizer.minimize(network, loss, bounds, constraints, maxiter, tol, callback=callback)
    return _fdm(network, q_opt, getattr(optimizer, "model", None))


def network_updated(network, eq_state):
    network = network.copy()
    network_update(network, eq_state)
    return network


def network_update(network, eq_state):
    """Write an equilibrium state into the attributes of a network, in place."""
    xyz = np.asarray(eq_state.xyz).tolist()
    lengths = np.asarray(eq_state.lengths).tolist()
    residuals = np.asarray(eq_state.residuals).tolist()
    forces = np.asarray(eq_state.forces).tolist()
    forcedensities = np.asarray(eq_state.force_densities).tolist()

    for idx, edge in network.index_uv().items():
        network.edge_attribute(edge, name="length", value=lengths[idx])
        network.edge_attribute(edge, name="force", value=forces[idx])
        network.edge_attribute(edge, name="q", value=forcedensities[idx])

    for idx, node in network.index_key().items():
        network.node_attributes(node, "xyz", xyz[idx])
        network.node_attributes(node, ["rx", "ry", "rz"], residuals[idx])


class NetworkBatch:
    """
    The equilibrium states of a batch of designs of ONE network, with the write-back deferred.

    The reference copies the network and walks every node and edge in Python for each result
    (fdm.py:48-84: O(n + E) attribute writes per design).  For an ensemble of thousands of designs
    that loop would dwarf the evaluation, so the batch keeps the arrays as they came off the GPU and
    materialises ``network_updated(network, state_b)`` only for the designs that are asked for.
    """

    def __init__(self, network, eq_state):
        if not eq_state.is_batched:
            raise ValueError("NetworkBatch needs a batched equilibrium state")
        self.network = network
        self.state = eq_state
        self._made = {}

    def __len__(self):
        return len(self.state.force_densities)

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[k] for k in range(*b.indices(len(self)))]
        b = int(b)
        if b < 0:
            b += len(self)
        if not 0 <= b < len(self):
            raise IndexError("design %d of a batch of %d" % (b, len(self)))
        if b not in self._made:
            self._made[b] = network_updated(self.network, self.state.design(b))
        return self._made[b]

    def __iter__(self):
        return (self[b] for b in range(len(self)))

    def loadpaths(self):
        """Load path sum(|force * length|) of every design (datastructures.py:119-139), without building networks."""
        return np.sum(np.abs(np.asarray(self.state.forces) * np.asarray(self.state.lengths)), axis=1)


def fdm_batch(network, q, model=None):
    """Equilibrium of ``network`` for every row of ``q[B, E]`` in one batched evaluation; networks are built on demand."""
    q = np.asarray(q, dtype=np.float64)
    if q.ndim != 2:
        raise ValueError("fdm_batch expects q[B, E]; use fdm() for the force densities stored on the network")
    model = model or EquilibriumModel(network)
    return NetworkBatch(network, model(q))
