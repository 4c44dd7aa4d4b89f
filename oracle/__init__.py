"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes front-end of the CPU oracle (``liboracle.so``, built by ``make -C oracle``):
a restatement of the reference's per-pivot loop (src/v4_cub_reduction.cu:219-380)
and the synthetic LP generators.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / reference legs may import this package; the product
path (``simplex_method_gpu_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX_ITER, OPTIMUM, UNBOUNDED, THETA_OVERFLOW = 0, 1, 2, 3


class _Opts(C.Structure):
    _fields_ = [("pivot_tol", C.c_double), ("harris_delta", C.c_double), ("ratio_mode", C.c_int), ("pricing_rule", C.c_int),
                ("nranks", C.c_int), ("reserved", C.c_int)]


class _Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_long), ("pivots", C.c_long), ("z", C.c_double)]


def build(force: bool = False) -> str:
    """Compile liboracle.so in-tree (gcc, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("simplex_oracle.c", "simplex_oracle_impl.h", "simplex_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        for name, real in (("oracle_solve_ex_f64", C.c_double), ("oracle_solve_ex_f32", C.c_float)):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, real, C.c_long, C.c_int,
                           C.POINTER(_Opts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(_Result)]
        L.lpgen_u01.restype = C.c_double
        L.lpgen_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        for name in ("lpgen_dense_f64", "lpgen_dense_f32"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_uint64]
        L.lpgen_klee_minty_f64.restype = None
        L.lpgen_klee_minty_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        L.lpgen_assignment_f64.restype = None
        L.lpgen_assignment_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_uint64, C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


@dataclass
class OracleSolution:
    status: int
    iterations: int
    pivots: int
    z: float
    x_b: np.ndarray
    b_ixs: np.ndarray
    y: np.ndarray
    trace_p: np.ndarray
    trace_q: np.ndarray
    gap_p: np.ndarray
    gap_q: np.ndarray
    Binv: np.ndarray | None = field(default=None, repr=False)

    def x(self, n: int) -> np.ndarray:
        """Full primal vector (non-basic variables are 0)."""
        out = np.zeros(n, dtype=self.x_b.dtype)
        out[self.b_ixs] = self.x_b
        return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def solve(A, b, c, eps=1e-4, max_iter=5, order=0, want_Binv=False, trace_cap=None,
          pivot_tol=0.0, ratio_mode=0, harris_delta=0.0, pricing_rule=0, nranks=1) -> OracleSolution:
    """A: (m, n) array in Fortran (column-major) order, dtype float32/float64; slack block last.
    pivot_tol / ratio_mode / harris_delta / pricing_rule: the modes outside the reference's contract
    (simplex_oracle.h); all zero = the reference's loop."""
    dt = np.dtype(A.dtype)
    assert dt in (np.float32, np.float64)
    A = np.asfortranarray(A, dtype=dt)
    b = np.ascontiguousarray(b, dtype=dt)
    c = np.ascontiguousarray(c, dtype=dt)
    m, n = A.shape
    cap = int(trace_cap if trace_cap is not None else min(max_iter, 1 << 22))
    x_b = np.zeros(m, dt)
    y = np.zeros(m, dt)
    b_ixs = np.zeros(m, np.int32)
    Binv = np.zeros((m, m), dt, order="F") if want_Binv else None
    tp = np.full(cap, -1, np.int32)
    tq = np.full(cap, -1, np.int32)
    gp = np.zeros(cap, np.float64)
    gq = np.zeros(cap, np.float64)
    res = _Result()
    fn = lib().oracle_solve_ex_f64 if dt == np.float64 else lib().oracle_solve_ex_f32
    opts = _Opts(float(pivot_tol), float(harris_delta), int(ratio_mode), int(pricing_rule), int(nranks), 0)
    rc = fn(_ptr(A), _ptr(b), _ptr(c), m, n, eps, int(max_iter), int(order), C.byref(opts),
            _ptr(x_b), _ptr(b_ixs), _ptr(y), _ptr(Binv), _ptr(tp), _ptr(tq), _ptr(gp), _ptr(gq), cap, C.byref(res))
    if rc != 0:
        raise ValueError(f"oracle_solve failed with code {rc} (m={m}, n={n})")
    k = min(res.pivots, cap)
    return OracleSolution(res.status, res.iterations, res.pivots, res.z, x_b, b_ixs, y,
                          tp[:k], tq[:k], gp[:k], gq[:k], Binv)


def gen_dense(m: int, n: int, seed: int = 1, dtype=np.float64):
    dt = np.dtype(dtype)
    A = np.empty((m, n), dt, order="F")
    b = np.empty(m, dt)
    c = np.empty(n, dt)
    fn = lib().lpgen_dense_f64 if dt == np.float64 else lib().lpgen_dense_f32
    fn(_ptr(A), _ptr(b), _ptr(c), m, n, seed)
    return A, b, c


def gen_dense_into(A_ptr: int, b_ptr: int, c_ptr: int, m: int, n: int, seed: int = 1, dtype=np.float64):
    """Same LP written into caller-owned host memory (e.g. pinned buffers): A column-major m x n, b (m), c (n)."""
    fn = lib().lpgen_dense_f64 if np.dtype(dtype) == np.float64 else lib().lpgen_dense_f32
    fn(C.c_void_p(A_ptr), C.c_void_p(b_ptr), C.c_void_p(c_ptr), m, n, seed)


def gen_klee_minty(d: int):
    A = np.empty((d, 2 * d), np.float64, order="F")
    b = np.empty(d, np.float64)
    c = np.empty(2 * d, np.float64)
    lib().lpgen_klee_minty_f64(_ptr(A), _ptr(b), _ptr(c), d)
    return A, b, c


def gen_assignment(k: int, seed: int = 1):
    m, n = 2 * k, k * k + 2 * k
    A = np.empty((m, n), np.float64, order="F")
    b = np.empty(m, np.float64)
    c = np.empty(n, np.float64)
    w = np.empty(k * k, np.float64)
    lib().lpgen_assignment_f64(_ptr(A), _ptr(b), _ptr(c), k, seed, _ptr(w))
    return A, b, c, w.reshape(k, k)


def u01(seed: int, stream: int, idx: int) -> float:
    return lib().lpgen_u01(seed, stream, idx)


def num_threads() -> int:
    return lib().oracle_num_threads()


# ---------------------------------------------------------------- the reference's own v4 build (GPU only)

_REF_DIR = os.path.join(_HERE, "_ref")
_ref_libs: dict = {}


def build_ref() -> bool:
    """Run make_ref.sh when /root/reference is mounted (build container); no-op on the GPU box."""
    src = "/root/reference/src/v4_cub_reduction.cu"
    if not os.path.exists(src):
        return ref_available()
    outs = [os.path.join(_REF_DIR, f) for f in ("libv4ref_f64.so", "libv4ref_f32.so", "v4_stock.out", "v4_shim.out",
                                                "v4_shim_env_f64.out", "v4_shim_env_f32.out", "v4_cli_f64.out", "v4_cli_f32.out")]
    deps = [src, os.path.join(_HERE, "make_ref.sh"), os.path.join(_HERE, "ref_harness.cu"), os.path.join(_HERE, "v4_shim.cu"),
            os.path.join(_HERE, "ref_cli.cu"), os.path.join(os.path.dirname(_HERE), "integration", "v4_b200.inc"),
            os.path.join(os.path.dirname(_HERE), "include", "b200lp.h")]
    if all(os.path.exists(o) for o in outs) and min(map(os.path.getmtime, outs)) > max(map(os.path.getmtime, deps)):
        return True
    subprocess.run([os.path.join(_HERE, "make_ref.sh")], check=True, capture_output=True)
    return ref_available()


def ref_available(dtype=np.float64) -> bool:
    name = "libv4ref_f64.so" if np.dtype(dtype) == np.float64 else "libv4ref_f32.so"
    return os.path.exists(os.path.join(_REF_DIR, name))


def ref_stock_binary() -> str | None:
    p = os.path.join(_REF_DIR, "v4_stock.out")
    return p if os.path.exists(p) else None


def ref_binary(name: str) -> str | None:
    """v4_shim.out (stock main() + the INTEGRATION.md binding), v4_shim_env_f{32,64}.out, v4_cli_f{32,64}.out."""
    p = os.path.join(_REF_DIR, name)
    return p if os.path.exists(p) else None


def _ref_lib(dtype):
    dt = np.dtype(dtype)
    if dt not in _ref_libs:
        name = "libv4ref_f64.so" if dt == np.float64 else "libv4ref_f32.so"
        L = C.CDLL(os.path.join(_REF_DIR, name))
        L.ref_solve.restype = C.c_int
        L.ref_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_sizeof_real.restype = C.c_int
        assert L.ref_sizeof_real() == dt.itemsize
        L.ref_set_always_readback.restype = None
        L.ref_set_always_readback.argtypes = [C.c_int]
        _ref_libs[dt] = L
    return _ref_libs[dt]


@dataclass
class RefSolution:
    status: int
    iterations: int
    pivots: int
    z: float
    x_b: np.ndarray
    b_ixs: np.ndarray
    trace_p: np.ndarray
    trace_q: np.ndarray
    secs_total: float
    secs_loop: float


def ref_solve(A, b, c, eps=1e-4, max_iter=5, trace_cap=None, always_readback=False) -> RefSolution:
    """The reference's own solve() (v4:219, patched per make_ref.sh) on the current CUDA device.
    `always_readback`: also return z / x_b / b_ixs when the run stops at max_iter (patch P8; v4:363 reads them
    back on OptimumFound only).  Arrays are passed as they are (no copy when already column-major of one dtype),
    so pinned host buffers stay pinned like the reference's own cudaMallocHost arrays (v4:408-414)."""
    dt = np.dtype(A.dtype)
    A = np.asfortranarray(A, dtype=dt)
    b = np.ascontiguousarray(b, dtype=dt)
    c = np.ascontiguousarray(c, dtype=dt)
    m, n = A.shape
    _ref_lib(dt).ref_set_always_readback(1 if always_readback else 0)
    cap = int(trace_cap if trace_cap is not None else min(max_iter, 1 << 22))
    x_b = np.zeros(m, dt)
    b_ixs = np.zeros(m, np.int32)
    tr = np.full((max(cap, 1), 2), -1, np.int32)
    z, it, st, sl = C.c_double(0), C.c_long(0), C.c_double(0), C.c_double(0)
    status = _ref_lib(dt).ref_solve(_ptr(A), _ptr(b), _ptr(c), m, n, eps, int(max_iter), _ptr(x_b), _ptr(b_ixs),
                                    _ptr(tr), cap, C.byref(z), C.byref(it), C.byref(st), C.byref(sl))
    if status < 0:
        raise RuntimeError("reference solve failed (CUDA error)")
    piv = it.value - 1 if status in (OPTIMUM, UNBOUNDED) else it.value
    k = min(piv, cap)
    return RefSolution(status, it.value, piv, z.value, x_b, b_ixs, tr[:k, 0].copy(), tr[:k, 1].copy(), st.value, sl.value)
