"""
Generate golden vectors by executing the REFERENCE'S OWN SOURCE FILES (/root/reference/src/dfdm)
in the build container.  Run once here (``python tests/golden/make_golden.py``); the ``.npz`` files it
writes are committed, because /root/reference does not exist on the GPU box.

The reference cannot be imported as is: its third-party dependencies ``autograd`` and ``compas`` are
neither in /root/reference nor installable offline.  This script installs small shims for exactly the
symbols the hot path touches, and leaves every line of reference code (model.py, structure.py,
goals/*, losses/*, constraints/*) untouched:

    autograd.numpy  -> NumPy itself for values ("numpy" mode); the same functions mapped onto
                       torch.float64 tensors ("torch" mode) so that ``autograd.grad`` can be served by
                       torch's tape.  Gradients therefore come from the reference's op sequence under a
                       different AD engine than the one it ships with.
    compas.datastructures.Network            -> dfdm_b200.datastructures.Network (COMPAS ordering rules
                                                restated from recall; see that module)
    compas.numerical.connectivity_matrix     -> C[e,u] = -1, C[e,v] = +1 (published COMPAS algorithm)
    compas.geometry.closest_point_on_line / closest_point_on_plane -> published COMPAS algorithms
    compas.data.Data                         -> JSON (de)serialisable base

What the goldens therefore pin: the oracle's (and the CUDA path's) agreement with the reference's
forward, goal, loss and constraint code.  What they cannot pin: COMPAS' own behaviour.
"""
import json
import os
import sys
import types
from math import sqrt

import numpy as _np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REFERENCE_SRC = "/root/reference/src"
sys.path.insert(0, ROOT)

MODE = {"value": "numpy"}


# --------------------------------------------------------------------------------------
# autograd shim
# --------------------------------------------------------------------------------------

def _t(a):
    if isinstance(a, torch.Tensor):
        return a.to(torch.float64)
    if isinstance(a, (list, tuple)) and any(isinstance(x, torch.Tensor) for x in a):
        return torch.stack([_t(x) for x in a])
    return torch.as_tensor(_np.asarray(a, dtype=_np.float64))


def _torch_mode(*args):
    if MODE["value"] == "torch":
        return True
    for a in args:
        if isinstance(a, torch.Tensor):
            return True
        if isinstance(a, (list, tuple)) and any(isinstance(x, torch.Tensor) for x in a):
            return True
    return False


def _make_anp():
    anp = types.ModuleType("autograd.numpy")
    for name in ("float64", "int32", "inf", "pi", "ndarray"):
        setattr(anp, name, getattr(_np, name))

    def asarray(a, dtype=None):
        return _t(a) if _torch_mode(a) else _np.asarray(a, dtype=dtype)

    def array(a, dtype=None):
        return _t(a) if _torch_mode(a) else _np.array(a, dtype=dtype)

    def diag(a):
        return torch.diag(_t(a)) if _torch_mode(a) else _np.diag(a)

    def transpose(a):
        return _t(a).T if _torch_mode(a) else _np.transpose(a)

    def concatenate(seq, axis=0):
        if _torch_mode(*seq):
            return torch.cat([_t(s) for s in seq], dim=axis)
        return _np.concatenate(seq, axis=axis)

    def _reduce(tfun, nfun):
        def fun(a, axis=None):
            if _torch_mode(a):
                return tfun(_t(a)) if axis is None else tfun(_t(a), dim=axis)
            return nfun(a, axis=axis)
        return fun

    def _unary(tfun, nfun):
        def fun(a):
            return tfun(_t(a)) if _torch_mode(a) else nfun(a)
        return fun

    def atleast_1d(a):
        return torch.atleast_1d(_t(a)) if _torch_mode(a) else _np.atleast_1d(a)

    def dot(a, b):
        return torch.dot(_t(a), _t(b)) if _torch_mode(a, b) else _np.dot(a, b)

    def cross(a, b):
        return torch.linalg.cross(_t(a), _t(b)) if _torch_mode(a, b) else _np.cross(a, b)

    def multiply(a, b):
        return _t(a) * b if _torch_mode(a, b) else _np.multiply(a, b)

    def roll(a, shift, axis=None):
        return torch.roll(_t(a), shift, dims=axis) if _torch_mode(a) else _np.roll(a, shift, axis=axis)

    def reshape(a, shape):
        return _t(a).reshape(shape) if _torch_mode(a) else _np.reshape(a, shape)

    def full(shape, value):
        return torch.full(tuple(shape), float(value), dtype=torch.float64) if _torch_mode() else _np.full(shape, value)

    def maximum(a, b):
        return torch.maximum(_t(a), _t(b)) if _torch_mode(a, b) else _np.maximum(a, b)

    def minimum(a, b):
        return torch.minimum(_t(a), _t(b)) if _torch_mode(a, b) else _np.minimum(a, b)

    anp.asarray, anp.array, anp.diag, anp.transpose, anp.concatenate = asarray, array, diag, transpose, concatenate
    anp.sum = _reduce(torch.sum, _np.sum)
    anp.mean = _reduce(torch.mean, _np.mean)
    anp.square = _unary(torch.square, _np.square)
    anp.abs = _unary(torch.abs, _np.abs)
    anp.arccos = _unary(torch.arccos, _np.arccos)
    anp.degrees = _unary(torch.rad2deg, _np.degrees)
    anp.atleast_1d, anp.dot, anp.cross, anp.multiply, anp.roll = atleast_1d, dot, cross, multiply, roll
    anp.reshape, anp.full, anp.maximum, anp.minimum = reshape, full, maximum, minimum

    linalg = types.ModuleType("autograd.numpy.linalg")

    def solve(a, b):
        return torch.linalg.solve(a, b) if _torch_mode(a, b) else _np.linalg.solve(a, b)

    def norm(a, axis=None):
        if isinstance(a, torch.Tensor):
            return torch.linalg.norm(a) if axis is None else torch.linalg.norm(a, dim=axis)
        return _np.linalg.norm(a, axis=axis)          # plain lists stay NumPy (targets: list / float)

    linalg.solve, linalg.norm = solve, norm
    anp.linalg = linalg
    return anp


def _install_shims():
    anp = _make_anp()
    autograd = types.ModuleType("autograd")
    autograd.numpy = anp

    def grad(fun):
        def gradfun(q, *args, **kwargs):
            qt = torch.tensor(_np.asarray(q, dtype=_np.float64), requires_grad=True)
            value = fun(qt, *args, **kwargs)
            g, = torch.autograd.grad(value, qt)
            return g.numpy().copy()
        return gradfun

    def jacobian(fun):
        def jacfun(q, *args, **kwargs):
            qt = torch.tensor(_np.asarray(q, dtype=_np.float64), requires_grad=True)
            value = torch.atleast_1d(fun(qt, *args, **kwargs))
            rows = [torch.autograd.grad(value[k], qt, retain_graph=True)[0].numpy().copy() for k in range(value.numel())]
            return _np.stack(rows)
        return jacfun

    autograd.grad, autograd.jacobian = grad, jacobian
    sys.modules["autograd"] = autograd
    sys.modules["autograd.numpy"] = anp
    sys.modules["autograd.numpy.linalg"] = anp.linalg

    from dfdm_b200.datastructures import Network

    compas = types.ModuleType("compas")
    ds = types.ModuleType("compas.datastructures")
    ds.Network = Network

    def network_adjacency_matrix(network, rtype="array"):
        key_index = network.key_index()
        n = network.number_of_nodes()
        A = [[0] * n for _ in range(n)]
        for key in network.nodes():
            for nbr in network.neighbors(key):
                A[key_index[key]][key_index[nbr]] = 1
        return A

    ds.network_adjacency_matrix = network_adjacency_matrix

    numerical = types.ModuleType("compas.numerical")

    def connectivity_matrix(edges, rtype="array"):
        m = len(edges)
        n = max(max(u, v) for u, v in edges) + 1
        C = [[0] * n for _ in range(m)]
        for e, (u, v) in enumerate(edges):
            C[e][u] += -1
            C[e][v] += 1
        return C

    numerical.connectivity_matrix = connectivity_matrix

    geometry = types.ModuleType("compas.geometry")

    def closest_point_on_line(point, line):
        a, b = line
        ab = [b[i] - a[i] for i in range(3)]
        ap = [point[i] - a[i] for i in range(3)]
        l2 = sum(c * c for c in ab)
        x = sum(ap[i] * ab[i] for i in range(3)) / l2
        return [a[i] + ab[i] * x for i in range(3)]

    def closest_point_on_plane(point, plane):
        base, normal = plane
        x, y, z = base
        ln = sqrt(sum(c * c for c in normal))
        a, b, c = (comp / ln for comp in normal)
        x1, y1, z1 = point[0], point[1], point[2]
        d = a * x + b * y + c * z
        k = (a * x1 + b * y1 + c * z1 - d) / (a ** 2 + b ** 2 + c ** 2)
        return [x1 - k * a, y1 - k * b, z1 - k * c]

    geometry.closest_point_on_line = closest_point_on_line
    geometry.closest_point_on_plane = closest_point_on_plane

    data = types.ModuleType("compas.data")

    class Data:
        def to_json(self, filepath):
            with open(filepath, "w") as fp:
                json.dump({"dtype": type(self).__name__, "value": self.data}, fp)

    data.Data = Data
    compas.datastructures, compas.numerical, compas.geometry, compas.data = ds, numerical, geometry, data
    for name, mod in (("compas", compas), ("compas.datastructures", ds), ("compas.numerical", numerical),
                      ("compas.geometry", geometry), ("compas.data", data)):
        sys.modules[name] = mod


# --------------------------------------------------------------------------------------
# cases
# --------------------------------------------------------------------------------------

def goal_specs_to_reference(specs, index_key, index_uv, G):
    """Oracle-style specs ``(kind, row index, target, weight[, extra])`` -> reference goal objects."""
    out = []
    for spec in specs:
        kind, index, target, weight = spec[:4]
        if kind.startswith("Node"):
            key = index_key[index]
        elif kind.startswith("Edge"):
            key = index_uv[index]
        if kind == "NodePoint":
            out.append(G.NodePointGoal(key, target, weight))
        elif kind == "NodeLine":
            out.append(G.NodeLineGoal(key, target, weight))
        elif kind == "NodePlane":
            out.append(G.NodePlaneGoal(key, target, weight))
        elif kind == "NodeResidualForce":
            out.append(G.NodeResidualForceGoal(key, target, weight))
        elif kind == "NodeResidualVector":
            out.append(G.NodeResidualVectorGoal(key, target, weight))
        elif kind == "NodeResidualDirection":
            out.append(G.NodeResidualDirectionGoal(key, target, weight))
        elif kind == "EdgeLength":
            out.append(G.EdgeLengthGoal(key, target, weight))
        elif kind == "EdgeForce":
            out.append(G.EdgeForceGoal(key, target, weight))
        elif kind == "EdgeLoadPath":
            out.append(G.EdgeLoadPathGoal(key, target, weight))
        elif kind == "EdgeVectorAngle":
            out.append(G.EdgeVectorAngleGoal(key, spec[4], target, weight))
        elif kind == "EdgeDirection":
            out.append(G.EdgeDirectionGoal(key, target, weight))
        elif kind == "NetworkLoadPath":
            out.append(G.NetworkLoadPathGoal(target, weight))
        else:
            raise ValueError(kind)
    return out


def main():
    assert os.path.isdir(REFERENCE_SRC), "the reference tree is only present in the build container"
    _install_shims()
    sys.path.insert(0, REFERENCE_SRC)
    import dfdm.equilibrium as REQ          # the reference itself
    import dfdm.goals as RG
    import dfdm.losses as RL
    import dfdm.constraints as RC
    from dfdm.datastructures import FDNetwork as RefFDNetwork

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import golden_cases        # shared case definitions (networks, q, goal specs, terms)

    for case in golden_cases():
        name = case["name"]
        net = RefFDNetwork.from_data(case["network"].data)
        q = _np.asarray(case["q"], dtype=_np.float64)

        # forward through the reference's EquilibriumModel with real NumPy
        MODE["value"] = "numpy"
        model = REQ.EquilibriumModel(net)
        st = model(q)
        s = model.structure
        out = {"q": q,
               "xyz": st.xyz, "residuals": st.residuals, "vectors": st.vectors,
               "lengths": st.lengths, "forces": st.forces,
               "free": _np.asarray(s.free_nodes, dtype=_np.int64),
               "fixed": _np.asarray(s.fixed_nodes, dtype=_np.int64),
               "freefixed": _np.asarray(s.freefixed_nodes, dtype=_np.int64),
               "connectivity": _np.asarray(s.connectivity, dtype=_np.int8),
               "loads": _np.asarray(model.loads), "xyz_fixed": _np.asarray(model.xyz_fixed)}

        index_key, index_uv = net.index_key(), net.index_uv()

        def build_loss():
            terms = []
            for kind, specs, alpha in case["terms"]:
                if kind == "L2Regularizer":
                    terms.append(RL.L2Regularizer(alpha))
                else:
                    goals = goal_specs_to_reference(specs, index_key, index_uv, RG)
                    terms.append(getattr(RL, kind)(goals, alpha=alpha))
            return RL.Loss(*terms)

        if case["terms"]:
            loss = build_loss()
            out["loss"] = _np.float64(loss(q, model))
            # gradient: the same reference code on torch tensors
            MODE["value"] = "torch"
            tmodel = REQ.EquilibriumModel(net)
            tloss = build_loss()
            qt = torch.tensor(q, requires_grad=True)
            value = tloss(qt, tmodel)
            g, = torch.autograd.grad(value, qt)
            out["loss_torch"] = _np.float64(value.item())
            out["grad"] = g.numpy().copy()
            MODE["value"] = "numpy"

        for k, spec in enumerate(case.get("constraints", [])):
            kind = spec[0]
            if kind == "EdgeLength":
                c = RC.EdgeLengthConstraint(index_uv[spec[1]], 0.0, 1.0)
            elif kind == "EdgeVectorAngle":
                c = RC.EdgeVectorAngleConstraint(index_uv[spec[1]], spec[2], 0.0, 1.0)
            elif kind == "NetworkEdgesLength":
                c = RC.NetworkEdgesLengthConstraint(0.0, 1.0)
            elif kind == "NetworkEdgesForce":
                c = RC.NetworkEdgesForceConstraint(0.0, 1.0)
            elif kind == "NodeNormalAngle":
                c = RC.NodeNormalAngleConstraint(index_key[spec[1]], [index_key[p] for p in spec[2]], spec[3], 0.0, 1.0)
            elif kind == "NodeCurvature":
                c = RC.NodeCurvatureConstraint(index_key[spec[1]], [index_key[p] for p in spec[2]], 0.0, 1.0)
            MODE["value"] = "numpy"
            out["constraint_%d" % k] = _np.atleast_1d(_np.asarray(c(q, model), dtype=_np.float64))
            MODE["value"] = "torch"
            tmodel = REQ.EquilibriumModel(net)
            if kind == "EdgeVectorAngle":
                c = RC.EdgeVectorAngleConstraint(index_uv[spec[1]], spec[2], 0.0, 1.0)
            elif kind == "NodeNormalAngle":
                c = RC.NodeNormalAngleConstraint(index_key[spec[1]], [index_key[p] for p in spec[2]], spec[3], 0.0, 1.0)
            qt = torch.tensor(q, requires_grad=True)
            value = torch.atleast_1d(c(qt, tmodel))
            rows = [torch.autograd.grad(value[r], qt, retain_graph=True)[0].numpy().copy() for r in range(value.numel())]
            out["constraint_jac_%d" % k] = _np.stack(rows)
            MODE["value"] = "numpy"

        path = os.path.join(HERE, name + ".npz")
        _np.savez_compressed(path, **out)
        print("wrote", path, {k: getattr(v, "shape", ()) for k, v in out.items() if k in ("xyz", "grad", "loss")})


if __name__ == "__main__":
    main()
