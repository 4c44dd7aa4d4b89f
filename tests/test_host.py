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
    header = ""
    for h in ("b200lp.h", "b200lp_io.h"):
        with open(os.path.join(ROOT, "include", h)) as f:
            header += f.read()
    declared = set(re.findall(r"\b(b200lp_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200lp_engine"}
    assert declared, "no prototypes found in include/*.h"
    for name in sorted(declared):
        assert hasattr(engine_lib, name), f"{name} declared in b200lp.h but not exported"
    assert declared == set(capi.EXPORTS)
    # the ctypes mirrors of the two structs have the library's layout size (field order is checked by the option tests)
    assert engine_lib.b200lp_sizeof_options() == C.sizeof(capi.Options)
    assert engine_lib.b200lp_sizeof_result() == C.sizeof(capi.Result)
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


# ---------------------------------------------------------------- include/b200lp_io.h (host code, runs without a GPU)

def test_native_reader_matches_python_reader_on_the_reference_fixture(engine_lib):
    import simplex_method_gpu_b200 as s
    from simplex_method_gpu_b200.solver import read_lp_native
    path = os.path.join(GOLDEN, "sample.txt")
    for dt in (np.float32, np.float64):
        A0, b0, c0 = s.read_lp(path, dtype=dt)
        A1, b1, c1 = read_lp_native(path, dtype=dt)
        assert A1.flags.f_contiguous and A1.dtype == dt
        assert np.array_equal(A0, A1) and np.array_equal(b0, b1) and np.array_equal(c0, c1)


def test_native_text_and_binary_round_trip(engine_lib, oracle, tmp_path):
    from simplex_method_gpu_b200.solver import read_lp_native, write_lp_native
    A, b, c = oracle.gen_dense(37, 90, 5)
    A[3, 4], b[2], c[1] = -1.25e-300, 1e300, -0.0          # awkward values must survive the text form
    txt, binf = str(tmp_path / "lp.txt"), str(tmp_path / "lp.b200lp")
    write_lp_native(txt, A, b, c)
    write_lp_native(binf, A, b, c, binary=True)
    for path in (txt, binf):
        A1, b1, c1 = read_lp_native(path, dtype=np.float64)
        assert np.array_equal(A, A1) and np.array_equal(b, b1) and np.array_equal(c, c1), path
    A32, b32, c32 = read_lp_native(binf, dtype=np.float32)                 # binary of the other dtype is converted
    assert np.array_equal(A32, A.astype(np.float32)) and np.array_equal(c32, c.astype(np.float32))
    with open(txt) as f:
        assert f.readline().split() == ["37", "90"]                           # v4:401 header


def test_native_reader_is_parallel_safe_on_a_larger_file(engine_lib, oracle, tmp_path):
    """Big enough that the reader cuts the text into several per-thread pieces."""
    from simplex_method_gpu_b200.solver import read_lp_native, write_lp_native
    A, b, c = oracle.gen_dense(300, 700, 8)
    path = str(tmp_path / "big.txt")
    write_lp_native(path, A, b, c)
    assert os.path.getsize(path) > 1 << 20
    A1, b1, c1 = read_lp_native(path, dtype=np.float64)
    assert np.array_equal(A, A1) and np.array_equal(b, b1) and np.array_equal(c, c1)
    with open(path, "a") as f:
        f.write("\nOptimum: 12.5 at x0 = 1\n")                            # trailing text is ignored (input/sample.txt:15-16)
    A2, _, c2 = read_lp_native(path, dtype=np.float64)
    assert np.array_equal(A, A2) and np.array_equal(c, c2)


def test_native_reader_errors_use_the_reference_messages(engine_lib, tmp_path):
    from simplex_method_gpu_b200.solver import read_lp_native
    with pytest.raises(ValueError, match=r"Could not open .*nope\.txt\."):        # v4:397
        read_lp_native(str(tmp_path / "nope.txt"))
    p = tmp_path / "bad.txt"
    p.write_text("3 2\n1 2 3 4 5 6\n1 1 1\n1 1\n")
    with pytest.raises(ValueError, match="Either failed to read m and n, or m > n."):   # v4:403
        read_lp_native(str(p))
    p.write_text("2 4\n1 1 1 0\n1 x 0 1\n3 4\n1 1 0 0\n")
    with pytest.raises(ValueError, match=r"Failed to read \(1,1\) for A"):           # v4:99
        read_lp_native(str(p))
    p.write_text("2 4\n1 1 1 0\n1 2 0 1\n3\n")
    with pytest.raises(ValueError, match=r"Failed to read \(1,0\) for b"):
        read_lp_native(str(p))
    p.write_text("2 4\n1 1 1 0\n1 2 0 1\n3 4\n1 1 0\n")
    with pytest.raises(ValueError, match=r"Failed to read \(0,3\) for c"):
        read_lp_native(str(p))


def test_native_io_round_trip_fuzz(engine_lib, tmp_path):
    """Random shapes and awkward values (denormals, huge exponents, negative zero, integers) through text and binary."""
    from simplex_method_gpu_b200.solver import read_lp_native, write_lp_native
    import simplex_method_gpu_b200 as s
    rng = np.random.default_rng(7)
    specials = np.array([0.0, -0.0, 1.0, -1.0, 5e-324, -2.2250738585072014e-308, 1.7976931348623157e308, 123456789.0,
                         1e-7, 0.1, 1 / 3])
    for case in range(25):
        m = int(rng.integers(1, 30))
        n = m + int(rng.integers(0, 40))
        for dt in (np.float64, np.float32):
            A = rng.standard_normal((m, n)) * 10.0 ** rng.integers(-12, 12, (m, n))
            pick = rng.random((m, n)) < 0.2
            A[pick] = rng.choice(specials, int(pick.sum()))
            with np.errstate(over="ignore", under="ignore"):
                A = np.asfortranarray(A.astype(dt))
                b = (rng.standard_normal(m) * 1e3).astype(dt)
                c = rng.choice(specials, n).astype(dt)
            A[~np.isfinite(A)] = 1.0
            c[~np.isfinite(c)] = 1.0
            for binary in (False, True):
                path = str(tmp_path / f"f{case}_{dt.__name__}_{int(binary)}")
                write_lp_native(path, A, b, c, binary=binary)
                A1, b1, c1 = read_lp_native(path, dtype=dt)
                assert A1.tobytes() == A.tobytes() and b1.tobytes() == b.tobytes() and c1.tobytes() == c.tobytes(), \
                    (case, dt.__name__, binary)                                  # bit-exact, sign of zero included
                if not binary:
                    A2, b2, c2 = s.read_lp(path, dtype=dt)                        # the pure-Python reader agrees
                    assert np.array_equal(A2, A) and np.array_equal(b2, b) and np.array_equal(c2, c)
