"""Entry points (reference: src/dfdm/equilibrium/fdm.py:10-84): same names, same signatures."""
import numpy as np

from dfdm_b200.equilibrium.model import EquilibriumModel

__all__ = ["fdm", "constrained_fdm", "network_update", "network_updated"]


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
