"""
Shared parity cases: networks of the BASELINE.json shapes, seeded force densities, goal / loss /
constraint definitions as plain specs with ROW indices.

    goal spec        (kind, row index | None, target, weight[, extra])
    term spec        (loss kind, [goal specs], alpha)      kinds: SquaredError, MeanSquaredError,
                                                          PredictionError, L2Regularizer
    constraint spec  (kind, row index, ...)

``golden_cases()`` feeds tests/golden/make_golden.py (reference source executed under shims), the
oracle tests and the GPU parity tests, so all three see identical inputs.
"""
import numpy as np

from dfdm_b200 import workloads as W


def _rng(seed):
    return np.random.default_rng(seed)


def _line_goals(network, weight=1.0):
    key_index = network.key_index()
    goals = []
    for node in network.nodes_free():
        xyz = network.node_coordinates(node)
        line = (xyz, [xyz[0], xyz[1], xyz[2] + 1.0])
        goals.append(("NodeLine", key_index[node], line, weight))
    return goals


def _interior_polygon(nx, i, j):
    """Ordered 4-neighbourhood of vertex (i, j) in a row-major (nx+1)-wide grid."""
    vid = lambda a, b: b * (nx + 1) + a
    return [vid(i + 1, j), vid(i, j + 1), vid(i - 1, j), vid(i, j - 1)]


def case_arches():
    net = W.arch(10)
    E = net.number_of_edges()
    q = -1.0 + 0.1 * (_rng(1).random(E) - 0.5)
    goals = [("NodeResidualDirection", 0, [-1.0, 0.0, -0.4], 1.0),
             ("NodeResidualDirection", 10, [1.0, 0.0, -0.4], 1.0)]
    return {"name": "arches10", "network": net, "q": q, "terms": [("SquaredError", goals, 1.0)]}


def case_pillow():
    net = W.pillow()
    E = net.number_of_edges()
    q = -2.0 + 0.1 * (_rng(2).random(E) - 0.5)
    constraints = [("NetworkEdgesLength",), ("NetworkEdgesForce",),
                   ("NodeCurvature", 3 * 7 + 3, _interior_polygon(6, 3, 3)),
                   ("NodeCurvature", 2 * 7 + 4, _interior_polygon(6, 4, 2)),
                   ("NodeNormalAngle", 3 * 7 + 2, _interior_polygon(6, 2, 3), [0.0, 0.0, 1.0])]
    return {"name": "pillow", "network": net, "q": q,
            "terms": [("SquaredError", _line_goals(net), 1.0)], "constraints": constraints}


def case_pringle():
    net = W.pringle()
    key_index, uv_index = net.key_index(), net.uv_index()
    E = net.number_of_edges()
    q = np.clip(-0.25 + 0.1 * (_rng(3).random(E) - 0.5), -5.0, -0.1)
    arches = net.attributes["arches"]
    num_v = len(arches)
    rz_min, rz_max = 0.45, 2.0
    num_steps = (num_v - 1) / 2.0
    step = (rz_max - rz_min) / num_steps
    rzs = [rz_min + i * step for i in range(int(num_steps) + 1)]
    rzs = rzs + rzs[0:-1][::-1]
    goals = []
    for rz, row in zip(rzs, arches):
        goals.append(("NodeResidualForce", key_index[row[0]], rz, 100.0))
        goals.append(("NodeResidualForce", key_index[row[-1]], rz, 100.0))
    for node in net.nodes_free():
        goals.append(("NodePlane", key_index[node], (net.node_coordinates(node), [1.0, 0.0, 0.0]), 10.0))
    for edge in net.attributes["cross_edges"]:
        goals.append(("EdgeLength", uv_index[tuple(edge)], net.edge_length(*edge), 1.0))
    return {"name": "pringle", "network": net, "q": q, "terms": [("SquaredError", goals, 1.0)]}


def case_monkey_saddle():
    net, targets = W.monkey_saddle()
    key_index = net.key_index()
    E = net.number_of_edges()
    q = -1.0 * (0.5 + _rng(4).random(E))
    goals = [("NodeResidualForce", key_index[k], targets[k], 1.0) for k in net.nodes_supports()]
    return {"name": "monkey_saddle", "network": net, "q": q, "terms": [("SquaredError", goals, 1.0)],
            "constraints": [("NetworkEdgesLength",)]}


def case_dome():
    net = W.dome()
    uv_index = net.uv_index()
    edges = list(net.edges())
    rings = set(tuple(e) for e in net.attributes["edges_rings"])
    u = _rng(5).random(len(edges))
    q = np.asarray([(-2.0 if e in rings else -0.5) * (1.0 + 0.2 * (u[k] - 0.5)) for k, e in enumerate(edges)])
    cross = net.attributes["edges_cross"]
    constraints = [("EdgeVectorAngle", uv_index[tuple(cross[0])], [0.0, 0.0, 1.0]),
                   ("EdgeVectorAngle", uv_index[tuple(cross[17])], [0.0, 0.0, 1.0]),
                   ("EdgeLength", uv_index[tuple(cross[5])])]
    return {"name": "dome", "network": net, "q": q,
            "terms": [("SquaredError", _line_goals(net), 1.0)], "constraints": constraints}


def case_mixed():
    """Every goal kind, every loss kind, several goals on one element, mixed weights."""
    net = W.grid(4, q0=1.0)
    edges, fixed, xyz, loads, q0 = W.network_arrays(net)
    rng = _rng(7)
    E, n = len(edges), len(fixed)
    q = -1.5 + 0.6 * (rng.random(E) - 0.5)
    free = [i for i in range(n) if not fixed[i]]
    fixd = [i for i in range(n) if fixed[i]]
    t1 = [("NodePoint", free[0], [0.3, 0.2, 0.5], 2.0),
          ("NodePoint", free[5], [0.6, 0.5, -0.1], 0.5),
          ("NodeLine", free[5], ([0.1, 0.2, 0.0], [0.4, 0.1, 1.0]), 1.5),
          ("NodePlane", free[7], ([0.0, 0.0, 0.3], [0.2, -0.1, 1.0]), 3.0),
          ("NodeResidualVector", fixd[2], [0.01, -0.02, 0.03], 4.0),
          ("NodeResidualForce", fixd[3], 0.05, 2.5),
          ("NodeResidualDirection", fixd[6], [0.5, 0.2, -1.0], 1.25),
          ("NodeResidualForce", fixd[6], 0.02, 0.75),
          ("EdgeForce", 3, -0.4, 1.0),
          ("EdgeLoadPath", 9, 0.1, 2.0),
          ("EdgeVectorAngle", 12, 40.0, 0.01, [0.0, 0.3, 1.0]),
          ("EdgeDirection", 12, [1.0, 0.1, 0.2], 1.0),
          ("EdgeDirection", 20, [0.0, 1.0, -0.3], 0.5),
          ("EdgeLength", 20, 0.3, 3.0)]
    t2 = [("NodePoint", free[9], [0.5, 0.5, 0.2], 1.0), ("EdgeLength", 15, 0.25, 2.0), ("EdgeForce", 16, -0.3, 1.0)]
    t3 = [("NetworkLoadPath", None, None, 1.0), ("EdgeLength", 30, 0.0, 1.0), ("NodePoint", free[2], [0.0, 0.0, 0.0], 1.0)]
    t4 = [("NetworkLoadPath", None, 1.5, 0.2)]
    terms = [("SquaredError", t1, 1.0), ("MeanSquaredError", t2, 0.7), ("PredictionError", t3, 0.05),
             ("SquaredError", t4, 2.0), ("L2Regularizer", None, 0.01)]
    return {"name": "mixed", "network": net, "q": q, "terms": terms}


def case_cable_grid(nfree=6, name="cable_grid"):
    """The 256x256 configuration's loss on a small grid: EdgeLengthGoal on every edge + L2 regulariser, q > 0."""
    net = W.grid(nfree, q0=1.0)
    E = net.number_of_edges()
    q = 1.0 + 0.2 * (_rng(6).random(E) - 0.5)
    goals = [("EdgeLength", k, net.edge_length(*edge), 1.0) for k, edge in enumerate(net.edges())]
    return {"name": name, "network": net, "q": q,
            "terms": [("SquaredError", goals, 1.0), ("L2Regularizer", None, 0.1)]}


def case_arch_forward():
    net = W.arch(200)
    E = net.number_of_edges()
    q = -1.0 + 0.1 * (_rng(1).random(E) - 0.5)
    return {"name": "arch200", "network": net, "q": q, "terms": []}


def case_mixed_sign():
    """Mixed-sign force densities: K indefinite (bounds of None allow it, examples/pillow.py:70-71)."""
    net = W.pringle(6, 5)
    E = net.number_of_edges()
    rng = _rng(11)
    q = np.where(rng.random(E) < 0.3, 0.6, -1.0) * (1.0 + 0.3 * rng.random(E))
    key_index = net.key_index()
    goals = [("NodePoint", key_index[k], [0.0, 0.0, 1.0], 1.0) for k in list(net.nodes_free())[:4]]
    return {"name": "mixed_sign", "network": net, "q": q, "terms": [("SquaredError", goals, 1.0)]}


def golden_cases():
    return [case_arches(), case_pillow(), case_pringle(), case_monkey_saddle(), case_dome(), case_mixed(),
            case_cable_grid(), case_arch_forward(), case_mixed_sign()]


def case_arrays(case):
    """Explicit arrays of a case: edges, is_fixed, xyz, loads (node / edge row order)."""
    edges, is_fixed, xyz, loads, _ = W.network_arrays(case["network"])
    return edges, is_fixed, xyz, loads


def oracle_terms(case):
    """Term specs in the oracle's form ``(kind, goals, alpha)``."""
    return [(kind, goals, alpha) for kind, goals, alpha in case["terms"]]


def program_inputs(case):
    """Case term specs -> (goal rows, term rows, l2_alpha) for dfdm_b200.engine.GoalProgram."""
    goals, terms, l2 = [], [], 0.0
    for kind, specs, alpha in case["terms"]:
        if kind == "L2Regularizer":
            l2 += alpha
            continue
        t = len(terms)
        terms.append((kind, alpha))
        for spec in specs:
            gkind, index, target, weight = spec[:4]
            if gkind in ("NodeLine", "NodePlane"):
                params = list(target[0]) + list(target[1])
            elif gkind == "EdgeVectorAngle":
                params = [target] + list(spec[4])
            elif gkind == "NetworkLoadPath":
                params = [0.0 if target is None else target]
            elif np.ndim(target) == 0:
                params = [target]
            else:
                params = list(target)
            goals.append((gkind, index, t, weight, params))
    return goals, terms, l2


def goal_objects(case):
    """Case term specs -> dfdm_b200 loss-term objects keyed by network keys (as a user of the reference API writes them)."""
    from dfdm_b200 import goals as G
    from dfdm_b200 import losses as L
    net = case["network"]
    index_key, index_uv = net.index_key(), net.index_uv()
    terms = []
    for kind, specs, alpha in case["terms"]:
        if kind == "L2Regularizer":
            terms.append(L.L2Regularizer(alpha))
            continue
        goals = []
        for spec in specs:
            gkind, index, target, weight = spec[:4]
            key = None if index is None else (index_key[index] if gkind.startswith("Node") else index_uv[index])
            if gkind == "EdgeVectorAngle":
                goals.append(G.EdgeVectorAngleGoal(key, spec[4], target, weight))
            elif gkind == "NetworkLoadPath":
                goals.append(G.NetworkLoadPathGoal(target, weight))
            else:
                goals.append(getattr(G, gkind + "Goal")(key, target, weight))
        terms.append(getattr(L, kind)(goals, alpha=alpha))
    return terms


def constraint_objects(case):
    from dfdm_b200 import constraints as C
    net = case["network"]
    index_key, index_uv = net.index_key(), net.index_uv()
    out = []
    for spec in case.get("constraints", []):
        kind = spec[0]
        if kind == "EdgeLength":
            out.append(C.EdgeLengthConstraint(index_uv[spec[1]], 0.0, 1.0))
        elif kind == "EdgeVectorAngle":
            out.append(C.EdgeVectorAngleConstraint(index_uv[spec[1]], spec[2], 0.0, 1.0))
        elif kind == "NetworkEdgesLength":
            out.append(C.NetworkEdgesLengthConstraint(0.0, 1.0))
        elif kind == "NetworkEdgesForce":
            out.append(C.NetworkEdgesForceConstraint(0.0, 1.0))
        elif kind == "NodeNormalAngle":
            out.append(C.NodeNormalAngleConstraint(index_key[spec[1]], [index_key[p] for p in spec[2]], spec[3], 0.0, 1.0))
        elif kind == "NodeCurvature":
            out.append(C.NodeCurvatureConstraint(index_key[spec[1]], [index_key[p] for p in spec[2]], 0.0, 1.0))
    return out
