"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under ``dfdm_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

Op-for-op restatement, on explicit arrays, of the reference's force-density hot path:

    indexing   /root/reference/src/dfdm/equilibrium/structure.py:36-106
    forward    /root/reference/src/dfdm/equilibrium/model.py:16-79
    goals      /root/reference/src/dfdm/goals/goal.py:23-111, goals/helpers.py:19-37,
               goals/nodegoal/pointgoal.py:11-61, goals/nodegoal/residualgoal.py:12-70,
               goals/edgegoal/{lengthgoal,forcegoal,loadpathgoal,anglegoal,directiongoal}.py,
               goals/networkgoal/loadpathgoal.py:7-23
    losses     /root/reference/src/dfdm/losses/losses.py:10-87, losses/regularizers.py:8-20
    constraints /root/reference/src/dfdm/constraints/*.py (library classes)

The dense path makes the same NumPy calls as the reference (``np.diag``, ``@``, ``np.linalg.solve``,
``np.linalg.norm``) so BLAS/LAPACK see identical operands.  Gradients come from the same op sequence
replayed on a ``torch.float64`` CPU autograd tape (the reference uses HIPS autograd, which is not
installable here) and are cross-checked against central finite differences in the tests.

Third-party behaviour restated from its published algorithm because the package is absent from
/root/reference (requirements.txt:1 pins only ``COMPAS>=0.17.1``):
    compas.numerical.connectivity_matrix      -> C[e, u] = -1, C[e, v] = +1        (structure.py:71)
    compas.geometry.closest_point_on_line     -> a + ((p-a).ab / |ab|^2) ab        (pointgoal.py:42)
    compas.geometry.closest_point_on_plane    -> p - ((p-o).n^) n^, n^ = n/|n|     (pointgoal.py:59)

PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path
(tests/PLACEHOLDER only).  This oracle is pinned instead against (i) outputs of the reference's OWN
source files executed in the build container with shims for the absent third-party packages
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) and (ii) analytic known answers that follow
from model.py.  COMPAS ordering rules themselves rest on recall: "parity unpinned" for
those; parity is therefore defined on explicit ``(edges, is_fixed, loads, xyz_fixed, q)`` arrays.
"""
from math import pi

import numpy as np


# ======================================================================================
# indexing (structure.py)
# ======================================================================================

def connectivity_matrix(edges, n=None):
    """structure.py:63-72 via compas.numerical.connectivity_matrix: dense E x n int32."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    m = edges.shape[0]
    if n is None:
        n = int(edges.max()) + 1 if m else 0
    C = np.zeros((m, n), dtype=np.int32)
    for e in range(m):
        C[e, edges[e, 0]] += -1
        C[e, edges[e, 1]] += 1
    return C


def free_fixed(is_fixed):
    """structure.py:74-106: ascending free / fixed node indices and the inverse permutation."""
    is_fixed = np.asarray(is_fixed).astype(bool)
    free = [i for i in range(len(is_fixed)) if not is_fixed[i]]
    fixed = [i for i in range(len(is_fixed)) if is_fixed[i]]
    freefixed_nodes = free + fixed
    indices = {node: index for index, node in enumerate(freefixed_nodes)}
    freefixed = [index for _, index in sorted(indices.items(), key=lambda item: item[0])]
    return free, fixed, freefixed


class Structure:
    def __init__(self, n, edges, is_fixed):
        self.n = int(n)
        self.edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        self.connectivity = connectivity_matrix(self.edges, self.n)
        self.free_nodes, self.fixed_nodes, self.freefixed_nodes = free_fixed(is_fixed)


# ======================================================================================
# array backends: the same op sequence runs on numpy (values) and torch (tape)
# ======================================================================================

class _NP:
    name = "numpy"
    asarray = staticmethod(lambda a: np.asarray(a, dtype=np.float64))
    intmat = staticmethod(lambda a: a)          # the reference multiplies the int32 matrix as is
    diag = staticmethod(np.diag)
    solve = staticmethod(np.linalg.solve)
    concatenate = staticmethod(lambda seq: np.concatenate(seq, axis=0))
    sum = staticmethod(np.sum)
    mean = staticmethod(np.mean)
    square = staticmethod(np.square)
    abs = staticmethod(np.abs)
    arccos = staticmethod(np.arccos)
    degrees = staticmethod(np.degrees)
    dot = staticmethod(np.dot)
    cross = staticmethod(np.cross)
    atleast_1d = staticmethod(np.atleast_1d)
    transpose = staticmethod(np.transpose)

    @staticmethod
    def norm(a, axis=None):
        return np.linalg.norm(a, axis=axis)

    @staticmethod
    def roll(a, k):
        return np.roll(a, k, axis=0)

    @staticmethod
    def clamp_vec(a):
        return np.maximum(np.minimum(a, np.full(a.shape, 1)), np.full(a.shape, -1))

    @staticmethod
    def stack(seq):
        return np.array(seq, dtype=np.float64)


def _torch_backend():
    import torch

    class _TH:
        name = "torch"
        asarray = staticmethod(lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=np.float64)))
        intmat = staticmethod(lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)))
        diag = staticmethod(torch.diag)
        solve = staticmethod(torch.linalg.solve)
        concatenate = staticmethod(lambda seq: torch.cat(list(seq), dim=0))
        sum = staticmethod(torch.sum)
        mean = staticmethod(torch.mean)
        square = staticmethod(torch.square)
        abs = staticmethod(torch.abs)
        arccos = staticmethod(torch.arccos)
        degrees = staticmethod(lambda a: a * (180.0 / pi))
        dot = staticmethod(torch.dot)
        atleast_1d = staticmethod(torch.atleast_1d)
        transpose = staticmethod(lambda a: a.T)

        @staticmethod
        def cross(a, b):
            return torch.linalg.cross(a, b)

        @staticmethod
        def norm(a, axis=None):
            return torch.linalg.norm(a) if axis is None else torch.linalg.norm(a, dim=axis)

        @staticmethod
        def roll(a, k):
            return torch.roll(a, k, dims=0)

        @staticmethod
        def clamp_vec(a):
            return torch.maximum(torch.minimum(a, torch.ones_like(a)), -torch.ones_like(a))

        @staticmethod
        def stack(seq):
            return torch.stack([s if isinstance(s, torch.Tensor) else torch.as_tensor(float(s), dtype=torch.float64) for s in seq])

    return _TH


# ======================================================================================
# forward (model.py)
# ======================================================================================

class State:
    """state.py:11-18"""
    def __init__(self, xyz, residuals, vectors, lengths, forces, force_densities):
        self.xyz = xyz
        self.residuals = residuals
        self.vectors = vectors
        self.lengths = lengths
        self.forces = forces
        self.force_densities = force_densities


class Model:
    """model.py:12-79 on explicit arrays: loads (n x 3, node order), xyz_fixed (n_x x 3, fixed order)."""

    def __init__(self, structure, loads, xyz_fixed):
        self.structure = structure
        self.loads = np.asarray(loads, dtype=np.float64).reshape(-1, 3)
        self.xyz_fixed = np.asarray(xyz_fixed, dtype=np.float64).reshape(-1, 3)

    def __call__(self, q, xp=_NP):
        s = self.structure
        free, fixed = s.free_nodes, s.fixed_nodes
        loads = xp.asarray(self.loads)
        xyz_fixed = xp.asarray(self.xyz_fixed)
        # model.py:44-47 (the int32 matrix is upcast by the matmul; values -1/0/+1 are exact)
        c_matrix = xp.intmat(s.connectivity)
        c_fixed = c_matrix[:, fixed]
        c_free = c_matrix[:, free]
        c_free_t = xp.transpose(c_free)
        # model.py:50-55
        q_matrix = xp.diag(q)
        A = c_free_t @ q_matrix @ c_free
        b = loads[free, :] - c_free_t @ q_matrix @ c_fixed @ xyz_fixed
        xyz_free = xp.solve(A, b)
        # model.py:57-61
        xyz = xp.concatenate((xyz_free, xyz_fixed))[s.freefixed_nodes, :]
        # model.py:21-34
        vectors = c_matrix @ xyz
        residuals = loads - xp.transpose(c_matrix) @ xp.diag(q) @ vectors
        lengths = xp.norm(vectors, axis=1)
        forces = q * lengths
        return State(xyz, residuals, vectors, lengths, forces, q)

    def stiffness(self, q):
        """A and b of model.py:53-54 (for assembly parity)."""
        s = self.structure
        c = s.connectivity
        c_free_t = np.transpose(c[:, s.free_nodes])
        Q = np.diag(np.asarray(q, dtype=np.float64))
        A = c_free_t @ Q @ c[:, s.free_nodes]
        b = self.loads[s.free_nodes, :] - c_free_t @ Q @ c[:, s.fixed_nodes] @ self.xyz_fixed
        return A, b


# ======================================================================================
# goals: plain specs ``(kind, index, target, weight, extra)``; index is a node / edge ROW index
# ======================================================================================

SCALAR_GOALS = ("NodeResidualForce", "EdgeLength", "EdgeForce", "EdgeLoadPath", "EdgeVectorAngle", "NetworkLoadPath")
VECTOR_GOALS = ("NodePoint", "NodeLine", "NodePlane", "NodeResidualVector", "NodeResidualDirection", "EdgeDirection")


def closest_point_on_line(point, line, xp):
    a, b = (xp.asarray(line[0]), xp.asarray(line[1]))
    ab = b - a
    ap = point - a
    c = ab * (xp.dot(ap, ab) / xp.dot(ab, ab))
    return a + c


def closest_point_on_plane(point, plane, xp):
    base, normal = np.asarray(plane[0], dtype=np.float64), np.asarray(plane[1], dtype=np.float64)
    n = normal / np.linalg.norm(normal)
    d = float(np.dot(n, base))
    n = xp.asarray(n)
    k = (xp.dot(n, point) - d) / xp.dot(n, n)
    return point - n * k


def angle_vectors(u, v, xp, deg):
    """anglegoal.py:24-32 / edgeconstraint.py:55-63 (Python max/min clamp => zero gradient when clipped)."""
    a = xp.dot(u, v) / (xp.norm(u) * xp.norm(v))
    a = max(min(a, 1), -1)
    if not hasattr(a, "shape"):
        a = xp.asarray(float(a))
    if deg:
        return xp.degrees(xp.arccos(a))
    return xp.arccos(a)


def goal_state(goal, eqstate, xp):
    """One goal -> (prediction, target, weight) as 1-vectors or 3-vectors (goal.py:23-33)."""
    kind, index, target, weight = goal[:4]
    extra = goal[4] if len(goal) > 4 else None

    if kind == "NodePoint":
        prediction = eqstate.xyz[index, :]
        tgt = xp.asarray(target)
    elif kind == "NodeLine":
        prediction = eqstate.xyz[index, :]
        tgt = closest_point_on_line(prediction, target, xp)
    elif kind == "NodePlane":
        prediction = eqstate.xyz[index, :]
        tgt = closest_point_on_plane(prediction, target, xp)
    elif kind == "NodeResidualForce":
        assert target >= 0.0
        prediction = xp.atleast_1d(xp.norm(eqstate.residuals[index, :]))
        tgt = xp.asarray(np.atleast_1d(target))
    elif kind == "NodeResidualVector":
        prediction = eqstate.residuals[index, :]
        tgt = xp.asarray(target)
    elif kind == "NodeResidualDirection":
        residual = eqstate.residuals[index, :]
        prediction = residual / xp.norm(residual)
        t = np.asarray(target, dtype=np.float64)
        tgt = xp.asarray(t / np.linalg.norm(t))
    elif kind == "EdgeLength":
        prediction = xp.atleast_1d(eqstate.lengths[index])
        tgt = xp.asarray(np.atleast_1d(target))
    elif kind == "EdgeForce":
        prediction = xp.atleast_1d(eqstate.forces[index])
        tgt = xp.asarray(np.atleast_1d(target))
    elif kind == "EdgeLoadPath":
        prediction = xp.atleast_1d(xp.abs(eqstate.lengths[index] * eqstate.forces[index]))
        tgt = xp.asarray(np.atleast_1d(target))
    elif kind == "EdgeVectorAngle":
        vector = eqstate.vectors[index, :]
        prediction = xp.atleast_1d(angle_vectors(vector, xp.asarray(extra), xp, deg=True))
        tgt = xp.asarray(np.atleast_1d(target))
    elif kind == "EdgeDirection":
        vector = eqstate.vectors[index, :]
        prediction = vector / xp.norm(vector)
        t = np.asarray(target, dtype=np.float64)
        tgt = xp.asarray(t / np.linalg.norm(t))
    elif kind == "NetworkLoadPath":
        prediction = xp.atleast_1d(xp.sum(xp.abs(eqstate.lengths * eqstate.forces)))
        tgt = xp.asarray(np.atleast_1d(0.0 if target is None else target))
    else:
        raise ValueError(kind)

    if kind in SCALAR_GOALS:
        wgt = xp.asarray(np.atleast_1d(weight))
    else:
        wgt = xp.asarray(np.asarray([weight] * 3, dtype=np.float64))
    return prediction, tgt, wgt


def goals_state(goals, eqstate, xp):
    """helpers.py:19-37"""
    predictions, targets, weights = [], [], []
    for goal in goals:
        p, t, w = goal_state(goal, eqstate, xp)
        predictions.append(p)
        targets.append(t)
        weights.append(w)
    return xp.concatenate(predictions), xp.concatenate(targets), xp.concatenate(weights)


# ======================================================================================
# losses: terms are ``(kind, goals, alpha)``; kind in SquaredError / MeanSquaredError /
# PredictionError / L2Regularizer (goals ignored)
# ======================================================================================

def loss_term(term, eqstate, xp):
    kind, goals, alpha = term
    if kind == "L2Regularizer":
        return alpha * xp.sum(xp.square(eqstate.force_densities))
    prediction, target, weight = goals_state(goals, eqstate, xp)
    if kind == "SquaredError":
        return alpha * xp.sum(weight * xp.square(prediction - target))
    if kind == "MeanSquaredError":
        return alpha * xp.mean(weight * xp.square(prediction - target))
    if kind == "PredictionError":
        return alpha * xp.sum(prediction)
    raise ValueError(kind)


def loss_value(terms, q, model, xp=_NP):
    """losses.py:76-81"""
    eqstate = model(q, xp)
    error = 0.0
    for term in terms:
        error = error + loss_term(term, eqstate, xp)
    return error


def loss_and_grad(terms, q, model):
    """value + dL/dq on a torch.float64 CPU tape (stands in for autograd.grad, optimization.py:61)."""
    import torch
    xp = _torch_backend()
    qt = torch.tensor(np.asarray(q, dtype=np.float64), requires_grad=True)
    value = loss_value(terms, qt, model, xp)
    grad, = torch.autograd.grad(value, qt)
    return float(value.detach()), grad.numpy().copy()


def vjp(q, model, cot):
    """q-bar for given cotangents of (xyz, residuals, vectors, lengths, forces), via the torch tape."""
    import torch
    xp = _torch_backend()
    qt = torch.tensor(np.asarray(q, dtype=np.float64), requires_grad=True)
    st = model(qt, xp)
    total = 0.0
    for name in ("xyz", "residuals", "vectors", "lengths", "forces"):
        if cot.get(name) is not None:
            total = total + torch.sum(getattr(st, name) * torch.as_tensor(np.asarray(cot[name], dtype=np.float64)))
    grad, = torch.autograd.grad(total, qt)
    return grad.numpy().copy()


def fd_grad(fun, q, h=1e-6):
    """Central finite differences with one Richardson step (third opinion on the gradient)."""
    q = np.asarray(q, dtype=np.float64)
    g = np.zeros_like(q)
    for k in range(q.size):
        def d(step):
            qp, qm = q.copy(), q.copy()
            qp[k] += step
            qm[k] -= step
            return (fun(qp) - fun(qm)) / (2 * step)
        g[k] = (4.0 * d(h) - d(2 * h)) / 3.0
    return g


# ======================================================================================
# constraints (library classes of constraints/*.py); spec ``(kind, index, extra)``
# ======================================================================================

def constraint_value(spec, eqstate, xp=_NP):
    kind = spec[0]
    if kind == "EdgeLength":
        return eqstate.lengths[spec[1]]
    if kind == "EdgeVectorAngle":
        return angle_vectors(eqstate.vectors[spec[1]], xp.asarray(spec[2]), xp, deg=True)
    if kind == "NetworkEdgesLength":
        return eqstate.lengths
    if kind == "NetworkEdgesForce":
        return eqstate.forces
    if kind == "NodeNormalAngle":
        # nodeconstraint.py:29-73 (radians)
        _, index, polygon, vector = spec
        p = eqstate.xyz[list(polygon), :]
        o = xp.mean(p, 0) if xp.name == "torch" else np.mean(p, axis=0)
        op = p - o
        op_off = xp.roll(op, 1)
        ns = xp.cross(op_off, op) * 0.5
        n = xp.sum(ns, 0)
        normal = n / xp.norm(n)
        return angle_vectors(normal, xp.asarray(vector), xp, deg=False)
    if kind == "NodeCurvature":
        # nodeconstraint.py:76-115
        _, index, polygon = spec
        o = eqstate.xyz[index, :]
        p = eqstate.xyz[list(polygon), :]
        op = p - o
        norm_op = xp.norm(op, axis=1).reshape(-1, 1)
        op = op / norm_op
        op_off = xp.roll(op, 1)
        dot = xp.sum(op_off * op, 1)
        dot = xp.clamp_vec(dot)
        angles = xp.arccos(dot)
        return 2 * pi - xp.sum(angles)
    raise ValueError(kind)


def constraint_jacobian(spec, q, model):
    """autograd.jacobian stand-in (optimization.py:125): one reverse pass per output row."""
    import torch
    xp = _torch_backend()
    qt = torch.tensor(np.asarray(q, dtype=np.float64), requires_grad=True)
    value = torch.atleast_1d(constraint_value(spec, model(qt, xp), xp))
    rows = []
    for k in range(value.numel()):
        g, = torch.autograd.grad(value[k], qt, retain_graph=True)
        rows.append(g.numpy().copy())
    return value.detach().numpy().copy(), np.stack(rows)


# ======================================================================================
# sparse restatement for networks whose dense C does not fit (NOT the reference's code path)
# ======================================================================================

class SparseModel:
    """Same maths as model.py with scipy.sparse CSR + SuperLU; used for the 256x256 grid config."""

    def __init__(self, n, edges, is_fixed, loads, xyz_fixed):
        import scipy.sparse as sp
        self.n = int(n)
        self.edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        m = self.edges.shape[0]
        rows = np.concatenate([np.arange(m), np.arange(m)])
        cols = np.concatenate([self.edges[:, 0], self.edges[:, 1]])
        data = np.concatenate([-np.ones(m), np.ones(m)])
        self.C = sp.coo_matrix((data, (rows, cols)), shape=(m, self.n)).tocsr()
        self.free, self.fixed, self.freefixed = free_fixed(is_fixed)
        self.Cf = self.C[:, self.free].tocsr()
        self.Cx = self.C[:, self.fixed].tocsr()
        self.CfT = self.Cf.T.tocsr()
        self.loads = np.asarray(loads, dtype=np.float64).reshape(-1, 3)
        self.xyz_fixed = np.asarray(xyz_fixed, dtype=np.float64).reshape(-1, 3)
        self._sp = sp

    def factor(self, q):
        from scipy.sparse.linalg import splu
        Q = self._sp.diags(q)
        A = (self.CfT @ Q @ self.Cf).tocsc()
        return A, splu(A)

    def __call__(self, q, lu=None):
        q = np.asarray(q, dtype=np.float64)
        if lu is None:
            _, lu = self.factor(q)
        Q = self._sp.diags(q)
        b = self.loads[self.free, :] - self.CfT @ (Q @ (self.Cx @ self.xyz_fixed))
        xyz_free = lu.solve(b)
        xyz = np.concatenate((xyz_free, self.xyz_fixed))[self.freefixed, :]
        vectors = self.C @ xyz
        residuals = self.loads - self.C.T @ (q[:, None] * vectors)
        lengths = np.linalg.norm(vectors, axis=1)
        forces = q * lengths
        return State(xyz, residuals, vectors, lengths, forces, q)

    def loss_and_grad_edge_length(self, q, targets, weight=1.0, alpha_reg=0.0):
        """SquaredError(EdgeLengthGoal for every edge) + L2Regularizer, hand adjoint."""
        q = np.asarray(q, dtype=np.float64)
        _, lu = self.factor(q)
        st = self(q, lu)
        diff = st.lengths - targets
        loss = float(np.sum(weight * diff * diff) + alpha_reg * np.sum(q * q))
        lbar = 2.0 * weight * diff
        ubar = (lbar / st.lengths)[:, None] * st.vectors
        xbar = self.C.T @ ubar
        lam = lu.solve(xbar[self.free, :])          # K symmetric
        dq = -np.sum((self.Cf @ lam) * st.vectors, axis=1) + 2.0 * alpha_reg * q
        return loss, dq, st
