"""
Synthetic networks of the named benchmark shapes (BASELINE.json ``configs``).

Each builder restates the geometry set-up of one reference example as an ``FDNetwork`` with supports,
loads and initial force densities; viewers, plotting and export of the example scripts are not part
of the hot path.  Mesh helpers of ``compas_singular`` (absent here) are replaced by explicit,
deterministic numbering.

    arch          /root/reference/examples/arch.py:21-54
    pillow        /root/reference/examples/pillow.py:56-159
    pringle       /root/reference/examples/pringle.py:44-141   (script's own name: "monkey_vault")
    monkey_saddle /root/reference/examples/monkey_saddle_constraints.py:39-133
    dome          /root/reference/examples/dome_constraints.py:64-162, examples/dome.py:69-70
    grid          synthetic (N+2)x(N+2)-node quad cable net, boundary fixed (BASELINE.json configs[4])
"""
import json
from math import cos
from math import pi
from math import sin

import numpy as np

from dfdm_b200.datastructures import FDNetwork


__all__ = ["arch", "pillow", "pringle", "monkey_saddle", "dome", "grid", "grid_arrays",
           "network_arrays", "sample_q", "MONKEY_SADDLE_COARSE"]


# examples/arch.py ---------------------------------------------------------------------

def arch(num_segments=5000, length=5.0, q_init=-1.0, pz=-0.1):
    points = [[length * i / num_segments, 0.0, 0.0] for i in range(num_segments + 1)]
    lines = [(points[i], points[i + 1]) for i in range(num_segments)]
    precision = "3f" if num_segments <= 5000 else "9f"
    network = FDNetwork.from_lines(lines, precision=precision)
    network.node_support(key=0)
    network.node_support(key=len(points) - 1)
    network.edges_forcedensities(q_init, keys=list(network.edges()))
    network.nodes_loads([0.0, 0.0, pz], keys=list(network.nodes_free()))
    return network


# examples/pillow.py -------------------------------------------------------------------

def _quad_grid(nx, ny, lx, ly):
    """(nx+1) x (ny+1) vertices, row-major; quads as vertex 4-tuples."""
    vertices = [[lx * i / nx, ly * j / ny, 0.0] for j in range(ny + 1) for i in range(nx + 1)]
    vid = lambda i, j: j * (nx + 1) + i
    faces = [[vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)] for j in range(ny) for i in range(nx)]
    return vertices, faces


def _mesh_edges(faces):
    edges, seen = [], set()
    for face in faces:
        for a, b in zip(face, face[1:] + face[:1]):
            if (a, b) not in seen and (b, a) not in seen:
                seen.add((a, b))
                edges.append((a, b))
    return edges


def _boundary_vertices(faces):
    count = {}
    for face in faces:
        for a, b in zip(face, face[1:] + face[:1]):
            key = (min(a, b), max(a, b))
            count[key] = count.get(key, 0) + 1
    boundary = set()
    for (a, b), c in count.items():
        if c == 1:
            boundary.update((a, b))
    return boundary, {k for k, c in count.items() if c == 1}


def _vertex_areas(vertices, faces):
    """Tributary area: a quarter of every incident quad (planar quads)."""
    area = [0.0] * len(vertices)
    total = 0.0
    for face in faces:
        p = np.asarray([vertices[k] for k in face])
        a = 0.5 * np.linalg.norm(np.cross(p[2] - p[0], p[3] - p[1]))
        total += a
        for k in face:
            area[k] += a / len(face)
    return area, total


def pillow(divisions=6, l1=10.0, l2=10.0, q0=-2.0, pz=-100.0):
    vertices, faces = _quad_grid(divisions, divisions, l1, l2)
    network = FDNetwork.from_nodes_and_edges(vertices, _mesh_edges(faces))
    boundary, _ = _boundary_vertices(faces)
    for key in network.nodes():
        if key in boundary:
            network.node_support(key)
    for edge in network.edges():
        network.edge_forcedensity(edge, q0)
    areas, total = _vertex_areas(vertices, faces)
    for key in network.nodes():
        network.node_load(key, load=[0.0, 0.0, pz * areas[key] / total])
    return network


# examples/pringle.py ------------------------------------------------------------------

def pringle(num_u=10, num_v=9, length_vault=6.0, width_vault=3.0, q_init=-0.25, pz=-0.1):
    network = FDNetwork()
    length_u = length_vault / (num_u - 1)
    length_v = width_vault / (num_v - 1)
    arches, long_edges, cross_edges = [], [], []
    for i in range(num_v):
        row = []
        for j in range(num_u):
            row.append(network.add_node(x=j * length_u - length_vault / 2.0, y=i * length_v - width_vault / 2.0, z=0.0))
        arches.append(row)
        a = i * num_u
        b = a + num_u - 1
        for u, v in zip(range(a, b), range(a + 1, b + 1)):
            long_edges.append(network.add_edge(u, v))
    for i in range(1, num_u - 1):
        seq = [row[i] for row in arches]
        for u, v in zip(seq[:-1], seq[1:]):
            cross_edges.append(network.add_edge(u, v))
    for row in arches:
        network.node_support(row[0])
        network.node_support(row[-1])
    for node in list(network.nodes_free()):
        network.node_load(node, load=[0.0, 0.0, pz])
    for edge in network.edges():
        network.edge_forcedensity(edge, q_init)
    network.attributes["arches"] = arches
    network.attributes["cross_edges"] = cross_edges
    return network


# examples/monkey_saddle_constraints.py ------------------------------------------------

MONKEY_SADDLE_COARSE = {
    "vertices": [[-9.6602544784545898, -3.267949104309082, 0.0], [-7.7942285537719727, -4.5, 0.0],
                 [-7.6602540016174316, -6.732050895690918, 0.0], [-3.4641015529632568, 2.0, 0.0],
                 [-2.0, 10.0, 0.0], [-4.4408920985006262e-16, -4.0, 0.0], [0.0, 0.0, 0.0], [0.0, 9.0, 0.0],
                 [2.0, 10.0, 0.0], [3.4641015529632568, 2.0, 0.0], [7.6602540016174316, -6.732050895690918, 0.0],
                 [7.7942285537719727, -4.5, 0.0], [9.6602544784545898, -3.267949104309082, 0.0]],
    "faces": [[6, 9, 8, 7], [6, 11, 12, 9], [3, 6, 7, 4], [0, 1, 6, 3], [1, 2, 5, 6], [5, 10, 11, 6]],
}


def load_coarse_mesh(filepath):
    """Read a COMPAS mesh JSON of the shape of data/json/monkey_saddle.json (vertex / face dicts)."""
    with open(filepath, "r") as fp:
        data = json.load(fp)
    data = data.get("value", data)
    keys = sorted(data["vertex"], key=int)
    index = {k: i for i, k in enumerate(keys)}
    vertices = [[data["vertex"][k]["x"], data["vertex"][k]["y"], data["vertex"][k]["z"]] for k in keys]
    faces = [[index[str(v)] for v in data["face"][f]] for f in sorted(data["face"], key=int)]
    return {"vertices": vertices, "faces": faces}


def _densify(coarse, n):
    """Split every coarse quad into n x n by bilinear interpolation and weld coincident vertices."""
    vertices, faces, index = [], [], {}

    def vid(p):
        key = "{:.9f},{:.9f},{:.9f}".format(*(0.0 if c == 0 else c for c in p))
        if key not in index:
            index[key] = len(vertices)
            vertices.append([float(c) for c in p])
        return index[key]

    for face in coarse["faces"]:
        a, b, c, d = (np.asarray(coarse["vertices"][k], dtype=np.float64) for k in face)
        grid = [[vid((1 - s) * (1 - t) * a + s * (1 - t) * b + s * t * c + (1 - s) * t * d)
                 for s in (i / n for i in range(n + 1))] for t in (j / n for j in range(n + 1))]
        for j in range(n):
            for i in range(n):
                faces.append([grid[j][i], grid[j][i + 1], grid[j + 1][i + 1], grid[j + 1][i]])
    return vertices, faces


def monkey_saddle(n=3, q0=-1.0, load=(0.0, 0.0, -1.0), coarse=None, rmin=2.0, rmax=10.0, r_exp=1.0):
    """Returns (network, reaction targets {support key: magnitude})."""
    coarse = coarse or MONKEY_SADDLE_COARSE
    vertices, faces = _densify(coarse, n)
    edges = _mesh_edges(faces)
    boundary, boundary_edges = _boundary_vertices(faces)
    degree = {}
    adjacency = {k: set() for k in range(len(vertices))}
    for a, b in edges:
        adjacency[a].add(b)
        adjacency[b].add(a)
    for k, nbrs in adjacency.items():
        degree[k] = len(nbrs)
    corners = sorted(k for k in boundary if degree[k] == 2)
    # boundary polyedges run corner to corner
    bnbrs = {k: [] for k in boundary}
    for a, b in sorted(boundary_edges):
        bnbrs[a].append(b)
        bnbrs[b].append(a)
    polyedges, done = [], set()
    for corner in corners:
        for start in bnbrs[corner]:
            if (corner, start) in done:
                continue
            chain = [corner, start]
            while chain[-1] not in corners:
                nxt = [k for k in bnbrs[chain[-1]] if k != chain[-2]][0]
                chain.append(nxt)
            done.add((chain[0], chain[1]))
            done.add((chain[-1], chain[-2]))
            polyedges.append(chain)
    vx = np.asarray(vertices)
    lengths = [sum(np.linalg.norm(vx[a] - vx[b]) for a, b in zip(p[:-1], p[1:])) for p in polyedges]
    mean_length = sum(lengths) / len(lengths)
    supports = set()
    for chain, length in zip(polyedges, lengths):
        if length < mean_length:
            supports.update(chain)
    # assembly "steps": hop distance to the nearest corner, reversed (examples/...constraints.py:103-121)
    def hops(src):
        dist, frontier = {src: 0}, [src]
        while frontier:
            nxt = []
            for k in frontier:
                for m in adjacency[k]:
                    if m not in dist:
                        dist[m] = dist[k] + 1
                        nxt.append(m)
            frontier = nxt
        return dist
    corner_dist = {c: hops(c) for c in corners}
    steps = {k: min(corner_dist[c][k] for c in corners) for k in supports}
    max_step = max(steps.values())
    steps = {k: max_step - s for k, s in steps.items()}

    net_edges = [(u, v) for u, v in edges if u not in supports or v not in supports]
    network = FDNetwork.from_nodes_and_edges(vertices, net_edges)
    network.nodes_supports(sorted(supports))
    network.nodes_loads(list(load), keys=list(network.nodes()))
    network.edges_forcedensities(q=q0)
    targets = {}
    for key in network.nodes_supports():
        targets[key] = (1 - steps[key] / max_step) ** r_exp * (rmax - rmin) + rmin
    return network, targets


# examples/dome_constraints.py, examples/dome.py ---------------------------------------

def dome(num_sides=16, num_rings=6, diameter=1.0, offset_distance=0.03, q0_ring=-2.0, q0_cross=-0.5, pz=-0.1):
    network = FDNetwork()
    apothem0 = diameter / 2.0 * cos(pi / num_sides)
    rings = []
    for i in range(num_rings + 1):
        radius = (apothem0 - (i + 1) * offset_distance) / cos(pi / num_sides)
        ring = [network.add_node(x=radius * cos(-2.0 * pi * k / num_sides), y=radius * sin(-2.0 * pi * k / num_sides), z=0.0)
                for k in range(num_sides)]
        rings.append(ring)
    edges_rings, edges_cross = [], []
    for ring in rings[1:]:
        closed = ring + ring[:1]
        for u, v in zip(closed[:-1], closed[1:]):
            edges_rings.append(network.add_edge(u, v))
    for k in range(num_sides):
        radial = [ring[k] for ring in rings]
        for u, v in zip(radial[:-1], radial[1:]):
            edges_cross.append(network.add_edge(u, v))
    for node in rings[0]:
        network.node_support(node)
    for node in list(network.nodes_free()):
        network.node_load(node, load=[0.0, 0.0, pz])
    for edge in edges_rings:
        network.edge_forcedensity(edge, q0_ring)
    for edge in edges_cross:
        network.edge_forcedensity(edge, q0_cross)
    network.attributes["edges_rings"] = edges_rings
    network.attributes["edges_cross"] = edges_cross
    # the radial edges between ring i and ring i + 1 (examples/dome.py:137-142), lower node first
    network.attributes["edges_cross_rings"] = [[(a, b) for a, b in zip(lower, upper)] for lower, upper in zip(rings[:-1], rings[1:])]
    return network


# synthetic quad cable net -------------------------------------------------------------

def grid_arrays(nfree=256, q0=1.0):
    """
    (nfree+2)^2-node quad grid on the unit square, boundary fixed on z = 0.1 (x^2 - y^2), load
    p_z = -1/nfree^2 on free nodes.  Returned as plain arrays (no Python objects per element) so the
    65k-node configuration builds in milliseconds:
    ``edges int32[E,2], is_fixed uint8[n], xyz float64[n,3], loads float64[n,3], q float64[E]``.
    Edge order follows the COMPAS rule (grouped by first endpoint in node order): for node (i, j) the
    +x edge precedes the +y edge.
    """
    m = nfree + 2
    idx = np.arange(m * m, dtype=np.int64).reshape(m, m)          # [j, i] -> node id, row-major
    h = 1.0 / (m - 1)
    jj, ii = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
    x = ii * h
    y = jj * h
    on_boundary = (ii == 0) | (ii == m - 1) | (jj == 0) | (jj == m - 1)
    z = np.where(on_boundary, 0.1 * (x * x - y * y), 0.0)
    xyz = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    right = np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], axis=1)
    up = np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], axis=1)
    edges = np.concatenate([right, up])
    fixed = on_boundary.ravel()
    order = np.lexsort((edges[:, 1], edges[:, 0]))
    edges = edges[order].astype(np.int32)
    loads = np.zeros((m * m, 3))
    loads[~fixed, 2] = -1.0 / (nfree * nfree)
    q = np.full(edges.shape[0], q0, dtype=np.float64)
    return edges, fixed.astype(np.uint8), xyz, loads, q


def grid(nfree=8, q0=1.0):
    edges, fixed, xyz, loads, q = grid_arrays(nfree, q0)
    network = FDNetwork.from_nodes_and_edges([list(p) for p in xyz], [tuple(int(k) for k in e) for e in edges])
    for key in network.nodes():
        if fixed[key]:
            network.node_support(key)
        network.node_load(key, load=list(loads[key]))
    for edge, value in zip(list(network.edges()), q):
        network.edge_forcedensity(edge, float(value))
    return network


# arrays + batches -----------------------------------------------------------------------

def network_arrays(network):
    """``edges int32[E,2], is_fixed uint8[n], xyz[n,3], loads[n,3], q[E]`` in node / edge row order."""
    key_index = network.key_index()
    edges = np.asarray([(key_index[u], key_index[v]) for u, v in network.edges()], dtype=np.int32).reshape(-1, 2)
    is_fixed = np.asarray([1 if network.node_attribute(k, "is_support") else 0 for k in network.nodes()], dtype=np.uint8)
    xyz = np.asarray([network.node_coordinates(k) for k in network.nodes()], dtype=np.float64).reshape(-1, 3)
    loads = np.asarray(network.nodes_loads(), dtype=np.float64).reshape(-1, 3)
    q = np.asarray(network.edges_forcedensities(), dtype=np.float64)
    return edges, is_fixed, xyz, loads, q


def sample_q(q0, batch, seed, spread=0.1, relative=False, lo=None, hi=None):
    """Batch of force-density vectors around ``q0[E]``: q0 + spread (u - 1/2), or q0 (1 + spread (u - 1/2))."""
    rng = np.random.default_rng(seed)
    u = rng.random((batch, len(q0)))
    q = q0[None, :] * (1.0 + spread * (u - 0.5)) if relative else q0[None, :] + spread * (u - 0.5)
    if lo is not None or hi is not None:
        q = np.clip(q, lo, hi)
    return np.ascontiguousarray(q, dtype=np.float64)
