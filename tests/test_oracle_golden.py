"""
The oracle against (i) golden vectors produced by the reference's own source under shims
(tests/golden/make_golden.py) and (ii) analytic known answers that follow from model.py.
"""
import os

import numpy as np
import pytest

from cases import golden_cases, case_arrays, oracle_terms
from oracle import fdm_oracle as O
from dfdm_b200 import workloads as W

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {c["name"]: c for c in golden_cases()}


def _model(case):
    edges, is_fixed, xyz, loads = case_arrays(case)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    return O.Model(s, loads, xyz[s.fixed_nodes])


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("name", sorted(CASES))
def test_indexing_bit_exact(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = _model(CASES[name])
    s = m.structure
    assert np.array_equal(np.asarray(s.free_nodes), g["free"])
    assert np.array_equal(np.asarray(s.fixed_nodes), g["fixed"])
    assert np.array_equal(np.asarray(s.freefixed_nodes), g["freefixed"])
    assert np.array_equal(s.connectivity.astype(np.int8), g["connectivity"])
    assert np.array_equal(m.loads, g["loads"]) and np.array_equal(m.xyz_fixed, g["xyz_fixed"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    st = _model(CASES[name])(g["q"])
    for key in ("xyz", "residuals", "vectors", "lengths", "forces"):
        assert np.array_equal(getattr(st, key), g[key]), key     # same NumPy calls => same bits


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c["terms"]))
def test_loss_and_grad_match_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = CASES[name]
    m = _model(case)
    terms = oracle_terms(case)
    value = float(O.loss_value(terms, g["q"], m))
    assert abs(value - float(g["loss"])) <= 1e-13 * max(1.0, abs(float(g["loss"])))
    tvalue, grad = O.loss_and_grad(terms, g["q"], m)
    assert abs(tvalue - float(g["loss"])) <= 1e-11 * max(1.0, abs(float(g["loss"])))
    assert _rel(grad, g["grad"]) < 1e-11


@pytest.mark.parametrize("name", ["arches10", "mixed", "cable_grid"])
def test_grad_vs_finite_differences(name):
    case = CASES[name]
    m = _model(case)
    terms = oracle_terms(case)
    _, grad = O.loss_and_grad(terms, case["q"], m)
    fd = O.fd_grad(lambda q: float(O.loss_value(terms, q, m)), case["q"], h=1e-5)
    assert _rel(grad, fd) < 1e-6


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c.get("constraints")))
def test_constraints_match_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = CASES[name]
    m = _model(case)
    for k, spec in enumerate(case["constraints"]):
        value, jac = O.constraint_jacobian(spec, g["q"], m)
        assert _rel(value, g["constraint_%d" % k]) < 1e-12
        assert _rel(jac, g["constraint_jac_%d" % k]) < 1e-10


# analytic known answers -----------------------------------------------------------------

def test_chain_closed_form():
    """Uniform chain, constant q, uniform load: z_i = p_z i (N - i) h^0 / (2 q) (tridiagonal q[-1,2,-1])."""
    N, qv, pz = 40, -1.0, -0.1
    net = W.arch(N, q_init=qv, pz=pz)
    edges, is_fixed, xyz, loads, q = W.network_arrays(net)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    st = O.Model(s, loads, xyz[s.fixed_nodes])(q)
    i = np.arange(N + 1)
    assert np.allclose(st.xyz[:, 2], pz * i * (N - i) / (2 * qv), rtol=1e-11, atol=1e-12)
    assert np.allclose(st.xyz[:, 0], xyz[:, 0], atol=1e-12)


def test_equilibrium_invariants():
    case = CASES["pringle"]
    m = _model(case)
    q = case["q"]
    st = m(q)
    s = m.structure
    assert np.max(np.abs(st.residuals[s.free_nodes])) < 1e-12            # R_free = 0
    assert np.allclose(st.residuals.sum(axis=0), m.loads.sum(axis=0), atol=1e-12)   # reactions balance loads
    assert np.array_equal(st.forces, q * st.lengths)
    # scaling q -> cq, P -> cP leaves X unchanged
    m2 = O.Model(s, 3.0 * m.loads, m.xyz_fixed)
    assert np.allclose(m2(3.0 * q).xyz, st.xyz, rtol=1e-12, atol=1e-13)


def test_unloaded_planar_supports_stay_planar():
    net = W.pillow(pz=0.0)
    edges, is_fixed, xyz, loads, q = W.network_arrays(net)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    st = O.Model(s, loads, xyz[s.fixed_nodes])(q)
    assert np.max(np.abs(st.xyz[:, 2])) < 1e-13


def test_sparse_restatement_matches_dense():
    case = CASES["cable_grid"]
    edges, is_fixed, xyz, loads = case_arrays(case)
    m = _model(case)
    sm = O.SparseModel(len(is_fixed), edges, is_fixed, loads, xyz[m.structure.fixed_nodes])
    st, sst = m(case["q"]), sm(case["q"])
    for key in ("xyz", "residuals", "vectors", "lengths", "forces"):
        assert _rel(getattr(sst, key), getattr(st, key)) < 1e-12
    targets = np.asarray([g[2] for g in case["terms"][0][1]])
    loss, dq, _ = sm.loss_and_grad_edge_length(case["q"], targets, 1.0, 0.1)
    tvalue, grad = O.loss_and_grad(oracle_terms(case), case["q"], m)
    assert abs(loss - tvalue) < 1e-12 * max(1, abs(tvalue))
    assert _rel(dq, grad) < 1e-10
