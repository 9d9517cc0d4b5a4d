"""
Host logic without a GPU: the C-ABI library loads and exports every declared symbol; the plan's index
arrays are bit-exact against the reference's indexing (golden vectors) and the oracle's dense matrices.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from cases import golden_cases, case_arrays, program_inputs
from oracle import fdm_oracle as O
from dfdm_b200 import _lib
from dfdm_b200.engine import Plan, GoalProgram

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {c["name"]: c for c in golden_cases()}
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dfdm_b200.h")).read()
    declared = set(re.findall(r"DFDM_API[^;(]*?\b(dfdm_\w+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert _lib.lib.dfdm_abi_version() == 3


@pytest.mark.parametrize("name", sorted(CASES))
def test_free_fixed_bit_exact(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    edges, is_fixed, xyz, loads = case_arrays(CASES[name])
    plan = Plan(len(is_fixed), edges, is_fixed)
    assert np.array_equal(plan.free, g["free"])
    assert np.array_equal(plan.fixed, g["fixed"])
    assert np.array_equal(plan.freefixed, g["freefixed"])
    # node incidence signs are the entries of the reference's branch-node matrix
    C = g["connectivity"].astype(np.int32)
    ptr, inc_e, inc_s = plan.array(_lib.ARR_NODE_INC_PTR), plan.array(_lib.ARR_NODE_INC_EDGE), plan.array(_lib.ARR_NODE_INC_SIGN)
    rebuilt = np.zeros_like(C)
    for i in range(plan.n):
        for k in range(ptr[i], ptr[i + 1]):
            rebuilt[inc_e[k], i] += inc_s[k]
    assert np.array_equal(rebuilt, C)


@pytest.mark.parametrize("name", sorted(CASES))
def test_csr_pattern_and_scatter_map(name):
    case = CASES[name]
    edges, is_fixed, xyz, loads = case_arrays(case)
    s = O.Structure(len(is_fixed), edges, is_fixed)
    m = O.Model(s, loads, xyz[s.fixed_nodes])
    rng = np.random.default_rng(0)
    q = 1.0 + rng.random(len(edges))                 # strictly positive: no accidental cancellation
    A, _ = m.stiffness(q)
    plan = Plan(len(is_fixed), edges, is_fixed)
    rowptr, colidx = plan.rowptr, plan.colidx
    assert plan.nnz == int(np.count_nonzero(A))
    rows = np.repeat(np.arange(plan.n_free), np.diff(rowptr))
    pattern = np.zeros_like(A, dtype=bool)
    pattern[rows, colidx] = True
    assert np.array_equal(pattern, A != 0)           # sparsity indexing bit-exact
    for f in range(plan.n_free):
        assert np.all(np.diff(colidx[rowptr[f]:rowptr[f + 1]]) > 0)
    # edge-parallel scatter-add through the slot map reproduces A
    vals = np.zeros(plan.nnz)
    slots = plan.edge_slots
    for e in range(len(edges)):
        uu, vv, uv, vu = slots[e]
        if uu >= 0: vals[uu] += q[e]
        if vv >= 0: vals[vv] += q[e]
        if uv >= 0: vals[uv] -= q[e]
        if vu >= 0: vals[vu] -= q[e]
    B = np.zeros_like(A)
    B[rows, colidx] = vals
    assert np.allclose(B, A, rtol=1e-14, atol=0)


def test_solver_ordering_is_a_permutation_and_reduces_bandwidth():
    case = CASES["dome"]
    edges, is_fixed, _, _ = case_arrays(case)
    plan = Plan(len(is_fixed), edges, is_fixed)
    perm = plan.solver_perm
    assert sorted(perm.tolist()) == list(range(plan.n_free))
    assert plan.info["bandwidth_solver"] <= plan.info["bandwidth_natural"]
    assert plan.info["direct_fits"] == 1 and plan.auto_solver == "direct"


def test_large_grid_plan_counts():
    from dfdm_b200 import workloads as W
    edges, fixed, xyz, loads, q = W.grid_arrays(256)
    plan = Plan(len(fixed), edges, fixed)
    assert (plan.n, plan.n_edges, plan.n_free, plan.n_fixed, plan.nnz) == (66564, 132612, 65536, 1028, 326656)
    assert plan.auto_solver == "pcg"
    # the multigrid hierarchy of the plan: every level about a quarter of the one above, dense solve at the bottom
    rows, nnzs = plan.array(_lib.ARR_MG_ROWS), plan.array(_lib.ARR_MG_NNZ)
    assert len(rows) == plan.info["mg_levels"] + 1 == len(nnzs)
    assert rows[0] == plan.n_free and nnzs[0] == plan.nnz
    assert rows[1] == plan.info["mg_n_coarse"] and rows[-1] == plan.info["mg_n_coarsest"] <= 96
    assert all(3.0 < rows[k] / rows[k + 1] <= 4.0 for k in range(len(rows) - 1))
    assert all(nnzs[k] >= rows[k] for k in range(len(rows)))


def test_bad_arguments_are_rejected():
    with pytest.raises(_lib.DfdmError):
        Plan(3, [[0, 5]], [0, 0, 1])
    with pytest.raises(_lib.DfdmError):
        Plan(2, [[0, 1]], [1, 1])
    plan = Plan(3, [[0, 1], [1, 2]], [1, 0, 1])
    with pytest.raises(_lib.DfdmError):
        GoalProgram(plan, [("EdgeLength", 7, 0, 1.0, [1.0])], [("SquaredError", 1.0)])
    with pytest.raises(_lib.DfdmError):
        GoalProgram(plan, [("EdgeLength", 0, 3, 1.0, [1.0])], [("SquaredError", 1.0)])


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if c["terms"]))
def test_goal_programs_compile(name):
    case = CASES[name]
    edges, is_fixed, _, _ = case_arrays(case)
    plan = Plan(len(is_fixed), edges, is_fixed)
    goals, terms, l2 = program_inputs(case)
    prog = GoalProgram(plan, goals, terms, l2)
    assert prog.n_goals == len(goals)


def test_evaluator_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dfdm_b200.engine import Evaluator
    plan = Plan(3, [[0, 1], [1, 2]], [1, 0, 1])
    with pytest.raises(RuntimeError):
        Evaluator(plan, np.zeros((3, 3)), np.zeros((2, 3)))
