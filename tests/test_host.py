"""CPU-only: host logic and the C-ABI surface (no compute without a GPU)."""
import ctypes as C
import io
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(engine_lib):
    from simplex_method_gpu_b200 import capi
    with open(os.path.join(ROOT, "include", "b200lp.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(b200lp_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200lp_engine"}
    assert declared, "no prototypes found in include/b200lp.h"
    for name in sorted(declared):
        assert hasattr(engine_lib, name), f"{name} declared in b200lp.h but not exported"
    assert declared == set(capi.EXPORTS)
    assert b"sm_100a" in engine_lib.b200lp_version()


def test_default_options_are_the_reference_constants(engine_lib):
    from simplex_method_gpu_b200 import capi
    o = capi.default_options()
    assert o.eps == 1e-4 and o.max_iter == 5 and o.device == 0 and o.check_slack == 1   # v4:18-19


def test_no_gpu_fails_loudly(engine_lib):
    """Without a CUDA device the product path refuses to run: there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import simplex_method_gpu_b200 as s
    from simplex_method_gpu_b200 import capi
    A = np.asfortranarray(np.hstack([np.ones((2, 2)), np.eye(2)]))
    with pytest.raises(capi.B200LPError) as ei:
        s.solve(A, np.ones(2), np.ones(4))
    assert ei.value.code == capi.ERR_NO_GPU
    with pytest.raises(capi.B200LPError):
        s.Engine(2, 4)


def test_argument_errors(engine_lib):
    from simplex_method_gpu_b200 import capi
    h = C.c_void_p()
    o = capi.default_options()
    assert engine_lib.b200lp_create(capi.F64, 4, 2, C.byref(o), C.byref(h)) == capi.ERR_ARG      # m > n (v4:402)
    assert engine_lib.b200lp_create(capi.F64, 0, 2, C.byref(o), C.byref(h)) == capi.ERR_ARG
    assert engine_lib.b200lp_create(7, 2, 4, C.byref(o), C.byref(h)) == capi.ERR_ARG
    assert engine_lib.b200lp_run(None, 1, None) == capi.ERR_ARG
    assert b"NULL" in engine_lib.b200lp_last_error()
    import simplex_method_gpu_b200 as s
    with pytest.raises(ValueError):
        s.solve(np.ones((3, 2)), np.ones(3), np.ones(2))
    with pytest.raises(TypeError):
        s.solve(np.ones((2, 4), dtype=np.int32), np.ones(2), np.ones(4), dtype=np.int32)


def test_read_lp_sample_and_roundtrip(tmp_path):
    import simplex_method_gpu_b200 as s
    A, b, c = s.read_lp(os.path.join(GOLDEN, "sample.txt"))
    assert A.dtype == np.float32 and A.flags.f_contiguous          # real = float (v4:12), column-major (v4:98)
    assert A.tolist() == [[1, 1, 1, 0], [2, 1, 0, 1]] and b.tolist() == [4, 5] and c.tolist() == [3, 2, 0, 0]
    rng = np.random.default_rng(0)
    A2 = np.asfortranarray(rng.random((5, 9)))
    b2, c2 = rng.random(5), rng.random(9)
    p = tmp_path / "lp.txt"
    s.write_lp(p, A2, b2, c2)
    A3, b3, c3 = s.read_lp(p, dtype=np.float64)
    assert np.array_equal(A2, A3) and np.array_equal(b2, b3) and np.array_equal(c2, c3)


def test_read_lp_errors():
    import simplex_method_gpu_b200 as s
    with pytest.raises(ValueError, match="m > n"):
        s.read_lp(io.StringIO("3 2\n1 2 3 4 5 6\n"))
    with pytest.raises(ValueError, match="m > n"):
        s.read_lp(io.StringIO("x y"))
    with pytest.raises(ValueError, match=r"\(1,0\) for A"):
        s.read_lp(io.StringIO("2 2\n1 2\n"))
    with pytest.raises(ValueError, match="for b"):
        s.read_lp(io.StringIO("1 2\n1 2\n"))
    with pytest.raises(ValueError, match=r"\(0,1\) for c"):
        s.read_lp(io.StringIO("1 2\n1 2\n3\n4\n"))


def test_format_result_matches_reference_stdout():
    """Exact text v4 prints for input/sample.txt before its timing block (v4:136-142, 426-445)."""
    import simplex_method_gpu_b200 as s
    sol = s.Solution(9.0, s.SolveStatus.OptimumFound, np.array([3, 1], np.float32), np.array([1, 0], np.int32), 3, 2,
                     np.array([[0, 1], [1, 0]], np.int32))
    assert s.format_result(sol) == "# Iteration 1\n# Iteration 2\n# Iteration 3\nOptimum found: 9\n\tx_1 = 3\n\tx_0 = 1\n\n"
    sol.status, sol.iterations = s.SolveStatus.MaxIter, 5
    assert s.format_result(sol).endswith("# Iteration 5\nMAX_ITER exceeded.\n\n")
    sol.status = s.SolveStatus.Unbounded
    assert "Problem unbounded.\n" in s.format_result(sol)
    sol = s.Solution(1234567.0, s.SolveStatus.OptimumFound, np.array([0.1 + 0.2]), np.array([7], np.int32), 1, 0,
                     np.zeros((0, 2), np.int32))
    assert s.format_result(sol) == "# Iteration 1\nOptimum found: 1.23457e+06\n\tx_7 = 0.3\n\n"   # ostream default: 6 sig. digits
