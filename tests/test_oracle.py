"""CPU-only: the oracle against every known answer available for the hot path.

The reference ships one golden vector (input/sample.txt:15-16).  Everything else is
pinned by independent solvers (HiGHS via scipy, Hungarian) and closed forms
(Klee-Minty), as SURVEY.md 8(c4) lays out; GLPK is not installed in this image.
"""
import json
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment, linprog

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _sample(dtype):
    from simplex_method_gpu_b200 import read_lp
    return read_lp(os.path.join(GOLDEN, "sample.txt"), dtype=dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sample_known_answer(oracle, dtype):
    A, b, c = _sample(dtype)
    s = oracle.solve(A, b, c, eps=1e-4, max_iter=5)          # the reference's own constants (v4:18-19)
    assert s.status == oracle.OPTIMUM
    assert s.iterations == 3 and s.pivots == 2
    assert s.z == 9.0                                         # input/sample.txt:15
    assert s.trace_p.tolist() == [0, 1] and s.trace_q.tolist() == [1, 0]
    assert s.b_ixs.tolist() == [1, 0] and s.x_b.tolist() == [3.0, 1.0]   # x0 = 1, x1 = 3 (sample.txt:16)
    assert s.x(4).tolist() == [1.0, 3.0, 0.0, 0.0]


def test_max_iter_status(oracle):
    A, b, c = _sample(np.float64)
    s = oracle.solve(A, b, c, max_iter=1)
    assert s.status == oracle.MAX_ITER and s.iterations == 1 and s.pivots == 1
    s = oracle.solve(A, b, c, max_iter=2)
    assert s.status == oracle.MAX_ITER and s.iterations == 2 and s.pivots == 2
    s = oracle.solve(A, b, c, max_iter=3)
    assert s.status == oracle.OPTIMUM and s.iterations == 3


def test_unbounded(oracle):
    # max x0 s.t. -x0 + x1 <= 1: column 0 has no positive entry
    A = np.array([[-1.0, 1.0, 1.0]], order="F")
    s = oracle.solve(A, np.array([1.0]), np.array([1.0, 0.0, 0.0]), max_iter=10)
    assert s.status == oracle.UNBOUNDED and s.iterations == 1 and s.pivots == 0


@pytest.mark.parametrize("d", [3, 6, 10, 14])
def test_klee_minty_exact(oracle, d):
    A, b, c = oracle.gen_klee_minty(d)
    s = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 22, trace_cap=4)
    assert s.status == oracle.OPTIMUM
    assert s.pivots == 2 ** d - 1                  # Dantzig's rule visits every vertex
    assert s.z == 5.0 ** d                         # exact: all data are integers < 2^53


@pytest.mark.parametrize("k", [4, 8, 16, 32])
def test_assignment_vs_hungarian(oracle, k):
    A, b, c, w = oracle.gen_assignment(k, seed=1)
    s = oracle.solve(A, b, c, eps=1e-4, max_iter=100000)
    r, cc = linear_sum_assignment(-w)
    assert s.status == oracle.OPTIMUM
    assert s.z == w[r, cc].sum()                   # integer data, +-1 pivots: exact
    assert (s.gap_q == 0).any()                    # the ratio test really does tie (degenerate)


@pytest.mark.parametrize("m,seed", [(32, 1), (64, 2), (128, 3), (256, 1)])
def test_dense_vs_highs(oracle, m, seed):
    A, b, c = oracle.gen_dense(m, 2 * m, seed)
    s = oracle.solve(A, b, c, eps=1e-9, max_iter=100000)
    ref = linprog(-c[:m], A_ub=A[:, :m], b_ub=b, method="highs-ds")
    assert s.status == oracle.OPTIMUM and ref.status == 0
    assert abs(s.z + ref.fun) <= 1e-9 * abs(ref.fun)
    x = s.x(2 * m)
    assert np.all(A @ x <= b * (1 + 1e-9) + 1e-9) and np.all(x >= -1e-9)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_engine_summation_order_same_pivots(oracle, dtype):
    """order=1 (the B200 engine's association of the sums) walks the same vertices."""
    m = 96
    A, b, c = oracle.gen_dense(m, 2 * m, 5, dtype=dtype)
    eps = 1e-4 if dtype == np.float32 else 1e-9
    s0 = oracle.solve(A, b, c, eps=eps, max_iter=10000, order=0)
    s1 = oracle.solve(A, b, c, eps=eps, max_iter=10000, order=1)
    assert s0.status == s1.status == oracle.OPTIMUM
    assert np.array_equal(s0.trace_p, s1.trace_p) and np.array_equal(s0.trace_q, s1.trace_q)
    assert abs(s0.z - s1.z) <= (1e-4 if dtype == np.float32 else 1e-11) * abs(s0.z)


def test_generator_is_counter_based(oracle):
    """Same (seed, index) -> same number regardless of problem shape or thread count."""
    A, b, c = oracle.gen_dense(8, 16, 7)
    for i in range(8):
        for j in range(8):
            assert A[i, j] == oracle.u01(7, 0, i * 8 + j)
    assert np.array_equal(A[:, 8:], np.eye(8))
    assert np.all((b >= 4.0) & (b < 8.0)) and np.all((c[:8] >= 0.5) & (c[:8] < 1.5)) and np.all(c[8:] == 0)


def test_golden_traces(oracle):
    """Committed pivot traces (tests/golden/make_golden.py): the oracle must not drift."""
    with open(os.path.join(GOLDEN, "traces.json")) as f:
        gold = json.load(f)
    for name, g in gold.items():
        if g["kind"] == "dense":
            A, b, c = oracle.gen_dense(g["m"], g["n"], g["seed"])
        elif g["kind"] == "klee_minty":
            A, b, c = oracle.gen_klee_minty(g["d"])
        else:
            A, b, c, _ = oracle.gen_assignment(g["k"], g["seed"])
        s = oracle.solve(A, b, c, eps=g["eps"], max_iter=g["max_iter"])
        assert s.status == g["status"], name
        assert s.pivots == g["pivots"] and s.iterations == g["iterations"], name
        assert s.trace_p[:len(g["p_head"])].tolist() == g["p_head"], name
        assert s.trace_q[:len(g["q_head"])].tolist() == g["q_head"], name
        assert abs(s.z - g["z"]) <= 1e-12 * max(1.0, abs(g["z"])), name


@pytest.mark.parametrize("kw", [dict(pricing_rule=1), dict(ratio_mode=1), dict(ratio_mode=2, harris_delta=1e-9),
                                dict(pivot_tol=1e-9), dict(pricing_rule=1, ratio_mode=2, harris_delta=1e-9)])
def test_optional_modes_reach_the_reference_optimum(oracle, kw):
    """Modes outside the parity contract (simplex_oracle.h): a different pivot sequence, the same optimum as the
    reference's rule and as HiGHS; steepest edge needs far fewer pivots; both summation orders agree on the optimum."""
    from scipy.optimize import linprog
    m = 256
    A, b, c = oracle.gen_dense(m, 2 * m, 1)
    base = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    hi = linprog(-c[:m], A_ub=A[:, :m], b_ub=b, method="highs-ds")
    for order in (0, 1):
        s = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20, order=order, **kw)
        assert s.status == oracle.OPTIMUM
        assert abs(s.z - base.z) <= 1e-9 * abs(base.z) and abs(s.z + hi.fun) <= 1e-9 * abs(hi.fun)
        if kw.get("pricing_rule") == 1:
            assert s.pivots < base.pivots // 2
    km = oracle.gen_klee_minty(10)
    s = oracle.solve(*km, eps=1e-4, max_iter=1 << 20, **kw)
    assert s.status == oracle.OPTIMUM and s.z == 5.0 ** 10
