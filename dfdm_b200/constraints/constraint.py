"""
Non-linear constraints lb <= c(q) <= ub for SLSQP / trust-constr (reference: src/dfdm/constraints/*.py).

Values come from one GPU forward evaluation (shared with the other constraints of the same ``q``
through a one-entry cache on the model); Jacobians are reverse-mode: the small map state -> c is
differentiated on the host (a handful of nodes or edges) and its cotangents go through the CUDA VJP,
one adjoint solve per constraint row, batched in a single call (optimization.py:115-130 used
autograd.jacobian, which re-solved the system once per row on the CPU).
"""
from math import pi

import numpy as np

from dfdm_b200.engine import Evaluator

__all__ = ["Constraint", "NodeConstraint", "EdgeConstraint", "NetworkConstraint", "EdgeLengthConstraint",
           "EdgeVectorAngleConstraint", "NodeNormalAngleConstraint", "NodeCurvatureConstraint",
           "NetworkEdgesLengthConstraint", "NetworkEdgesForceConstraint"]


def _cached_state(q, model):
    q = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    if q.nbytes > (1 << 20):                 # batches: no comparison pass over hundreds of megabytes, just evaluate
        return model(q)
    cache = getattr(model, "_state_cache", None)
    key = q.tobytes()
    if cache is None or cache[0] != key:
        cache = (key, model(q))
        model._state_cache = cache
    return cache[1]


class Constraint:
    def __init__(self, bound_low, bound_up, **kwargs):
        self.bound_low = bound_low
        self.bound_up = bound_up

    def __call__(self, q, model):
        return self.constraint(_cached_state(q, model), model)

    def constraint(self, eqstate, model, **kwargs):
        raise NotImplementedError

    def cotangents(self, eqstate, model):
        """Rows of d c / d state as a dict name -> array [m, ...] (zero arrays omitted)."""
        raise NotImplementedError

    def jacobian(self, q, model):
        """d c / d q, shape [m, E]: all rows in one batched VJP on the GPU."""
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
        eqstate = _cached_state(q, model)
        cots = self.cotangents(eqstate, model)
        m = next(iter(cots.values())).shape[0]
        dq, status = model.evaluator.vjp(np.repeat(q[None, :], m, axis=0), rescue=True, **cots)
        Evaluator.raise_on_status(status)
        return dq.cpu().numpy()


class NodeConstraint(Constraint):
    def __init__(self, key, bound_low, bound_up, **kwargs):
        super().__init__(bound_low=bound_low, bound_up=bound_up)
        self._key = key

    def index(self, model):
        return model.structure.node_index[self.key()]

    def key(self):
        return self._key


class EdgeConstraint(Constraint):
    def __init__(self, key, bound_low, bound_up, **kwargs):
        super().__init__(bound_low=bound_low, bound_up=bound_up)
        self._key = key

    def index(self, model):
        return model.structure.edge_index[self.key()]

    def key(self):
        return self._key


class NetworkConstraint(Constraint):
    pass


def _angle(u, v, deg):
    """Smallest angle between two vectors and its derivative with respect to u (zero when the cosine is clipped)."""
    nu, nv = np.linalg.norm(u), np.linalg.norm(v)
    a = np.dot(u, v) / (nu * nv)
    scale = 180.0 / pi if deg else 1.0
    if a > 1 or a < -1:
        return scale * np.arccos(max(min(a, 1), -1)), np.zeros(3)
    dadu = v / (nu * nv) - a * u / (nu * nu)
    return scale * np.arccos(a), -scale / np.sqrt(1.0 - a * a) * dadu


class EdgeLengthConstraint(EdgeConstraint):
    """Length of one edge."""
    def constraint(self, eqstate, model):
        return eqstate.lengths[self.index(model)]

    def cotangents(self, eqstate, model):
        lb = np.zeros((1, model.plan.n_edges))
        lb[0, self.index(model)] = 1.0
        return {"lengths": lb}


class EdgeVectorAngleConstraint(EdgeConstraint):
    """Angle in degrees between one edge and a vector (edgeconstraint.py:38-62)."""
    def __init__(self, key, vector, bound_low, bound_up):
        super().__init__(key=key, bound_low=bound_low, bound_up=bound_up)
        self.vector_other = np.asarray(vector, dtype=np.float64)

    def constraint(self, eqstate, model):
        return _angle(eqstate.vectors[self.index(model)], self.vector_other, deg=True)[0]

    def cotangents(self, eqstate, model):
        e = self.index(model)
        ub = np.zeros((1, model.plan.n_edges, 3))
        ub[0, e] = _angle(eqstate.vectors[e], self.vector_other, deg=True)[1]
        return {"vectors": ub}


class NodeNormalAngleConstraint(NodeConstraint):
    """Angle in radians between the normal of the polygon around a node and a vector (nodeconstraint.py:29-73)."""
    def __init__(self, key, polygon, vector, bound_low, bound_up):
        super().__init__(key, bound_low, bound_up)
        self.polygon = polygon
        self.vector_other = np.asarray(vector, dtype=np.float64)

    def _rows(self, model):
        return [model.structure.node_index[nbr] for nbr in self.polygon]

    @staticmethod
    def _normal_polygon(p):
        o = np.mean(p, axis=0)
        op = p - o
        op_off = np.roll(op, 1, axis=0)
        n = np.sum(0.5 * np.cross(op_off, op), axis=0)
        return n, n / np.linalg.norm(n)

    def constraint(self, eqstate, model):
        _, normal = self._normal_polygon(eqstate.xyz[self._rows(model), :])
        return _angle(normal, self.vector_other, deg=False)[0]

    def cotangents(self, eqstate, model):
        rows = self._rows(model)
        p = eqstate.xyz[rows, :]
        k = len(rows)
        n, normal = self._normal_polygon(p)
        nn = np.linalg.norm(n)
        g_normal = _angle(normal, self.vector_other, deg=False)[1]
        g_n = (g_normal - np.dot(g_normal, normal) * normal) / nn          # through n / |n|
        o = np.mean(p, axis=0)
        op = p - o
        # n = 1/2 sum_i op_{i-1} x op_i  =>  d n . g / d op_i = 1/2 (g x op_{i-1}) ... assembled by cyclic neighbours
        g_op = np.zeros_like(op)
        for i in range(k):
            prev_, next_ = op[(i - 1) % k], op[(i + 1) % k]
            g_op[i] = 0.5 * (np.cross(g_n, prev_) - np.cross(g_n, next_))
        g_p = g_op - np.mean(g_op, axis=0)                                  # through op = p - mean(p)
        xb = np.zeros((1, model.plan.n, 3))
        for i, row in enumerate(rows):
            xb[0, row] += g_p[i]
        return {"xyz": xb}


class NodeCurvatureConstraint(NodeConstraint):
    """Discrete curvature 2 pi - sum of angles between successive edges around a node (nodeconstraint.py:76-115)."""
    def __init__(self, key, polygon, bound_low, bound_up):
        super().__init__(key, bound_low, bound_up)
        self.polygon = polygon

    def _rows(self, model):
        return [model.structure.node_index[nbr] for nbr in self.polygon]

    def constraint(self, eqstate, model):
        o = eqstate.xyz[self.index(model), :]
        op = eqstate.xyz[self._rows(model), :] - o
        op = op / np.linalg.norm(op, axis=1).reshape(-1, 1)
        dot = np.clip(np.sum(np.roll(op, 1, axis=0) * op, axis=1), -1, 1)
        return 2 * pi - np.sum(np.arccos(dot))

    def cotangents(self, eqstate, model):
        rows = self._rows(model)
        center = self.index(model)
        o = eqstate.xyz[center, :]
        d = eqstate.xyz[rows, :] - o
        norms = np.linalg.norm(d, axis=1)
        h = d / norms.reshape(-1, 1)
        k = len(rows)
        g_h = np.zeros_like(h)
        for i in range(k):
            j = (i - 1) % k
            a = np.dot(h[j], h[i])
            if a >= 1 or a <= -1:          # np.minimum / np.maximum clip: zero gradient outside (-1, 1)
                continue
            s = 1.0 / np.sqrt(1.0 - a * a)   # d(-arccos a) / da
            g_h[i] += s * h[j]
            g_h[j] += s * h[i]
        g_d = (g_h - np.sum(g_h * h, axis=1).reshape(-1, 1) * h) / norms.reshape(-1, 1)
        xb = np.zeros((1, model.plan.n, 3))
        for i, row in enumerate(rows):
            xb[0, row] += g_d[i]
        xb[0, center] -= np.sum(g_d, axis=0)
        return {"xyz": xb}


class NetworkEdgesLengthConstraint(NetworkConstraint):
    """Lengths of all edges."""
    def constraint(self, eqstate, *args, **kwargs):
        return eqstate.lengths

    def cotangents(self, eqstate, model):
        return {"lengths": np.eye(model.plan.n_edges)}


class NetworkEdgesForceConstraint(NetworkConstraint):
    """Forces of all edges."""
    def constraint(self, eqstate, *args, **kwargs):
        return eqstate.forces

    def cotangents(self, eqstate, model):
        return {"forces": np.eye(model.plan.n_edges)}
