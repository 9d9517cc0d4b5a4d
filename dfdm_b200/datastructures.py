"""
Force density networks.

``FDNetwork`` keeps the method surface of the reference's ``dfdm.datastructures.FDNetwork``
(/root/reference/src/dfdm/datastructures.py:9-139).  The reference derives from
``compas.datastructures.Network``; COMPAS is a third-party dependency that is not part of the
reference tree and is not installable here, so ``Network`` below is a minimal stand-in that
supplies exactly the COMPAS methods the reference and its examples call:
ordering rules are those of COMPAS 1.x — nodes iterate in insertion order, edges iterate grouped
by their first endpoint in node order and then by insertion order of the second endpoint.

This is host-side bookkeeping only; no arithmetic of the hot path lives here.
"""
import copy as _copy
import json
from ast import literal_eval
from math import fabs
from math import sqrt


__all__ = ["Network", "FDNetwork"]


class Network:
    """
    A graph of nodes and directed edge keys with default attributes (COMPAS ``Network`` subset).
    """

    DTYPE = "compas.datastructures/Network"

    def __init__(self, name=None, **kwargs):
        self._max_node = -1
        self.node = {}
        self.edge = {}
        self.adjacency = {}
        self.attributes = {"name": name or "Network"}
        self.default_node_attributes = {"x": 0.0, "y": 0.0, "z": 0.0}
        self.default_edge_attributes = {}

    def __str__(self):
        return "<{} with {} nodes, {} edges>".format(type(self).__name__, self.number_of_nodes(), self.number_of_edges())

    # ----------------------------------------------------------------------
    # defaults
    # ----------------------------------------------------------------------

    def update_default_node_attributes(self, attr_dict=None, **kwattr):
        attr = dict(attr_dict or {})
        attr.update(kwattr)
        self.default_node_attributes.update(attr)

    def update_default_edge_attributes(self, attr_dict=None, **kwattr):
        attr = dict(attr_dict or {})
        attr.update(kwattr)
        self.default_edge_attributes.update(attr)

    # ----------------------------------------------------------------------
    # builders
    # ----------------------------------------------------------------------

    def add_node(self, key=None, attr_dict=None, **kwattr):
        if key is None:
            key = self._max_node = self._max_node + 1
        try:
            if int(key) == key and key > self._max_node:
                self._max_node = int(key)
        except (ValueError, TypeError):
            pass
        if key not in self.node:
            self.node[key] = {}
            self.edge[key] = {}
            self.adjacency[key] = {}
        attr = dict(attr_dict or {})
        attr.update(kwattr)
        self.node[key].update(attr)
        return key

    def add_edge(self, u, v, attr_dict=None, **kwattr):
        attr = dict(attr_dict or {})
        attr.update(kwattr)
        if u not in self.node:
            self.add_node(u)
        if v not in self.node:
            self.add_node(v)
        data = self.edge[u].get(v, {})
        data.update(attr)
        self.edge[u][v] = data
        if v not in self.adjacency[u]:
            self.adjacency[u][v] = None
        if u not in self.adjacency[v]:
            self.adjacency[v][u] = None
        return u, v

    @classmethod
    def from_nodes_and_edges(cls, nodes, edges):
        network = cls()
        if isinstance(nodes, dict):
            for key, (x, y, z) in nodes.items():
                network.add_node(key, x=x, y=y, z=z)
        else:
            for x, y, z in nodes:
                network.add_node(x=x, y=y, z=z)
        for u, v in edges:
            network.add_edge(u, v)
        return network

    @classmethod
    def from_lines(cls, lines, precision=None):
        """
        Nodes are the distinct end points (merged on a fixed-precision geometric key), numbered in
        order of first appearance; one edge per line.
        """
        fmt = "{{0:.{0}}},{{1:.{0}}},{{2:.{0}}}".format(precision or "3f")

        def gkey(xyz):
            x, y, z = (0.0 if c == 0 else c for c in xyz)  # fold -0.0 into 0.0
            return fmt.format(x, y, z)

        node = {}
        edges = []
        for sp, ep in lines:
            a, b = gkey(sp), gkey(ep)
            node[a] = sp
            node[b] = ep
            edges.append((a, b))
        key_index = {k: i for i, k in enumerate(node)}
        network = cls()
        for key, xyz in node.items():
            network.add_node(key_index[key], x=xyz[0], y=xyz[1], z=xyz[2])
        for a, b in edges:
            network.add_edge(key_index[a], key_index[b])
        return network

    # ----------------------------------------------------------------------
    # accessors
    # ----------------------------------------------------------------------

    def number_of_nodes(self):
        return len(self.node)

    def number_of_edges(self):
        return sum(len(nbrs) for nbrs in self.edge.values())

    def _node_view(self, key):
        view = dict(self.default_node_attributes)
        view.update(self.node[key])
        return view

    def _edge_view(self, u, v):
        view = dict(self.default_edge_attributes)
        view.update(self.edge[u][v])
        return view

    def nodes(self, data=False):
        for key in self.node:
            if data:
                yield key, self._node_view(key)
            else:
                yield key

    def edges(self, data=False):
        for u, nbrs in self.edge.items():
            for v in nbrs:
                if data:
                    yield (u, v), self._edge_view(u, v)
                else:
                    yield u, v

    def nodes_where(self, conditions, data=False):
        for key in self.node:
            attr = self._node_view(key)
            is_match = True
            for name, value in conditions.items():
                if name not in attr:
                    is_match = False
                    break
                current = attr[name]
                if isinstance(current, list):
                    if value not in current:
                        is_match = False
                        break
                elif isinstance(value, (tuple, list)):
                    minval, maxval = value
                    if current < minval or current > maxval:
                        is_match = False
                        break
                elif value != current:
                    is_match = False
                    break
            if is_match:
                if data:
                    yield key, attr
                else:
                    yield key

    def key_index(self):
        return {key: index for index, key in enumerate(self.nodes())}

    def index_key(self):
        return dict(enumerate(self.nodes()))

    def uv_index(self):
        return {(u, v): index for index, (u, v) in enumerate(self.edges())}

    def index_uv(self):
        return dict(enumerate(self.edges()))

    def has_edge(self, u, v, directed=True):
        if directed:
            return u in self.edge and v in self.edge[u]
        return (u in self.edge and v in self.edge[u]) or (v in self.edge and u in self.edge[v])

    def neighbors(self, key):
        return list(self.adjacency[key])

    # ----------------------------------------------------------------------
    # attributes
    # ----------------------------------------------------------------------

    def node_attribute(self, key, name, value=None):
        if key not in self.node:
            raise KeyError(key)
        if value is not None:
            self.node[key][name] = value
            return None
        if name in self.node[key]:
            return self.node[key][name]
        return self.default_node_attributes.get(name)

    def node_attributes(self, key, names=None, values=None):
        if key not in self.node:
            raise KeyError(key)
        if values is not None:
            for name, value in zip(names, values):
                self.node[key][name] = value
            return None
        if not names:
            return self._node_view(key)
        return [self.node_attribute(key, name) for name in names]

    def nodes_attribute(self, name, value=None, keys=None):
        if not keys:
            keys = self.nodes()
        if value is not None:
            for key in keys:
                self.node_attribute(key, name, value)
            return None
        return [self.node_attribute(key, name) for key in keys]

    def nodes_attributes(self, names=None, values=None, keys=None):
        if not keys:
            keys = self.nodes()
        if values is not None:
            for key in keys:
                self.node_attributes(key, names, values)
            return None
        return [self.node_attributes(key, names) for key in keys]

    def edge_attribute(self, key, name, value=None):
        u, v = key
        if u not in self.edge or v not in self.edge[u]:
            raise KeyError(key)
        if value is not None:
            self.edge[u][v][name] = value
            return None
        if name in self.edge[u][v]:
            return self.edge[u][v][name]
        return self.default_edge_attributes.get(name)

    def edge_attributes(self, key, names=None, values=None):
        u, v = key
        if u not in self.edge or v not in self.edge[u]:
            raise KeyError(key)
        if values is not None:
            for name, value in zip(names, values):
                self.edge[u][v][name] = value
            return None
        if not names:
            return self._edge_view(u, v)
        return [self.edge_attribute(key, name) for name in names]

    def edges_attribute(self, name, value=None, keys=None):
        if not keys:
            keys = self.edges()
        if value is not None:
            for key in keys:
                self.edge_attribute(key, name, value)
            return None
        return [self.edge_attribute(key, name) for key in keys]

    def edges_attributes(self, names=None, values=None, keys=None):
        if not keys:
            keys = self.edges()
        if values is not None:
            for key in keys:
                self.edge_attributes(key, names, values)
            return None
        return [self.edge_attributes(key, names) for key in keys]

    # ----------------------------------------------------------------------
    # geometry
    # ----------------------------------------------------------------------

    def node_coordinates(self, key, axes="xyz"):
        return [self.node_attribute(key, axis) for axis in axes]

    def edge_coordinates(self, u, v, axes="xyz"):
        return self.node_coordinates(u, axes), self.node_coordinates(v, axes)

    def edge_vector(self, u, v):
        a, b = self.edge_coordinates(u, v)
        return [b[0] - a[0], b[1] - a[1], b[2] - a[2]]

    def edge_length(self, u, v):
        x, y, z = self.edge_vector(u, v)
        return sqrt(x * x + y * y + z * z)

    def transformed(self, transformation):
        """
        Copy with node coordinates mapped by a 4x4 (nested list) affine matrix or a callable xyz -> xyz.
        """
        network = self.copy()
        network.transform(transformation)
        return network

    def transform(self, transformation):
        matrix = getattr(transformation, "matrix", transformation)
        for key in self.node:
            x, y, z = self.node_coordinates(key)
            if callable(matrix):
                x, y, z = matrix([x, y, z])
            else:
                x, y, z = (matrix[i][0] * x + matrix[i][1] * y + matrix[i][2] * z + matrix[i][3] for i in range(3))
            self.node_attributes(key, "xyz", [x, y, z])

    # ----------------------------------------------------------------------
    # data
    # ----------------------------------------------------------------------

    @property
    def data(self):
        data = {"attributes": self.attributes,
                "dna": self.default_node_attributes,
                "dea": self.default_edge_attributes,
                "node": {},
                "edge": {},
                "adjacency": {},
                "max_node": self._max_node}
        for key in self.node:
            data["node"][repr(key)] = self.node[key]
        for u in self.edge:
            data["edge"][repr(u)] = {repr(v): attr for v, attr in self.edge[u].items()}
        for u in self.adjacency:
            data["adjacency"][repr(u)] = {repr(v): None for v in self.adjacency[u]}
        return data

    @data.setter
    def data(self, data):
        self.attributes.update(data.get("attributes") or {})
        self.default_node_attributes.update(data.get("dna") or {})
        self.default_edge_attributes.update(data.get("dea") or {})
        self.node, self.edge, self.adjacency = {}, {}, {}
        for key, attr in (data.get("node") or {}).items():
            self.add_node(literal_eval(key), attr_dict=attr)
        for u, nbrs in (data.get("edge") or {}).items():
            for v, attr in nbrs.items():
                self.add_edge(literal_eval(u), literal_eval(v), attr_dict=attr)
        self._max_node = data.get("max_node", self._max_node)

    def copy(self):
        network = type(self)()
        network.data = _copy.deepcopy(self.data)
        return network

    def to_data(self):
        return self.data

    @classmethod
    def from_data(cls, data):
        network = cls()
        network.data = data
        return network

    def to_jsonstring(self, pretty=False):
        payload = {"dtype": self.DTYPE, "value": self.data}
        return json.dumps(payload, indent=4 if pretty else None)

    @classmethod
    def from_jsonstring(cls, string):
        payload = json.loads(string)
        return cls.from_data(payload["value"] if "value" in payload else payload)

    def to_json(self, filepath, pretty=False):
        with open(filepath, "w") as fp:
            fp.write(self.to_jsonstring(pretty))

    @classmethod
    def from_json(cls, filepath):
        with open(filepath, "r") as fp:
            return cls.from_jsonstring(fp.read())


class FDNetwork(Network):
    """
    A force density network (reference: datastructures.py:9-139; same defaults, same accessors).
    """
    # compas.data.Data.dtype = first two levels of the defining module + "/" + class name: the reference class lives in
    # dfdm.datastructures.datastructures, so its files say "dfdm.datastructures/FDNetwork"
    DTYPE = "dfdm.datastructures/FDNetwork"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

        self.update_default_node_attributes({"x": 0.0, "y": 0.0, "z": 0.0,
                                             "px": 0.0, "py": 0.0, "pz": 0.0,
                                             "rx": 0.0, "ry": 0.0, "rz": 0.0,
                                             "is_support": False})

        self.update_default_edge_attributes({"q": 0.0, "length": 0.0, "force": 0.0})

    # nodes ----------------------------------------------------------------

    def nodes_coordinates(self, keys=None, axes="xyz"):
        keys = keys or self.nodes()
        return [self.node_coordinates(node, axes) for node in keys]

    def node_support(self, key):
        return self.node_attribute(key=key, name="is_support", value=True)

    def nodes_supports(self, keys=None):
        if keys is None:
            return self.nodes_where({"is_support": True})
        return self.nodes_attribute(name="is_support", value=True, keys=keys)

    def nodes_free(self):
        return self.nodes_where({"is_support": False})

    def nodes_fixed(self):
        return self.nodes_where({"is_support": True})

    def node_load(self, key, load=None):
        return self.node_attributes(key=key, names=("px", "py", "pz"), values=load)

    def nodes_loads(self, load=None, keys=None):
        return self.nodes_attributes(names=("px", "py", "pz"), values=load, keys=keys)

    def nodes_residual(self, keys=None):
        return self.nodes_attributes(names=("rx", "ry", "rz"), keys=keys)

    def node_residual(self, key):
        return self.node_attributes(key=key, names=("rx", "ry", "rz"))

    # edges ----------------------------------------------------------------

    def edge_forcedensity(self, key, q=None):
        return self.edge_attribute(name="q", value=q, key=key)

    def edges_forcedensities(self, q=None, keys=None):
        return self.edges_attribute(name="q", value=q, keys=keys)

    def edge_force(self, key):
        return self.edge_attribute(key=key, name="force")

    def edges_forces(self, keys=None):
        return self.edges_attribute(keys=keys, name="force")

    def edges_lengths(self, keys=None):
        return self.edges_attribute(keys=keys, name="length")

    def edge_loadpath(self, key):
        force = self.edge_attribute(key=key, name="force")
        length = self.edge_attribute(key=key, name="length")
        return fabs(force * length)

    def edges_loadpaths(self, keys=None):
        keys = keys or self.edges()
        for key in keys:
            yield self.edge_loadpath(key)

    def loadpath(self):
        return sum(list(self.edges_loadpaths()))
