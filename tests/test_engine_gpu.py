"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bars (BASELINE.json north_star): entering/leaving
index sequences identical except at ties within 1e-12, objective and x within 1e-9
relative (fp64); integer-exact problems (Klee-Minty, assignment) bit-exact."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
RTOL_F64 = 1e-9          # north_star: objective and x agree within 1e-9 relative
TIE = 1e-12              # north_star: index sequences may differ only at ties within 1e-12


@pytest.fixture(scope="module")
def lp(engine_lib):
    import simplex_method_gpu_b200 as s
    return s


def assert_same_path(sol, ref, n, rtol):
    """Pivot sequence identical (or first divergence is a tie), objective / x within rtol."""
    k = min(len(sol.trace), len(ref.trace_p))
    same = (sol.trace[:k, 0] == ref.trace_p[:k]) & (sol.trace[:k, 1] == ref.trace_q[:k])
    if not same.all() or len(sol.trace) != len(ref.trace_p):
        first = int(np.argmin(same)) if not same.all() else k
        gap = min(ref.gap_p[first], ref.gap_q[first]) if first < len(ref.gap_p) else 0.0
        assert gap <= TIE, f"pivot {first}: engine {sol.trace[first].tolist()} vs oracle " \
                           f"({ref.trace_p[first]}, {ref.trace_q[first]}), runner-up gap {gap:g} is not a tie"
        # after a genuine tie the paths may differ; the optimum must not
        assert abs(sol.z - ref.z) <= rtol * max(1.0, abs(ref.z))
        return
    assert int(sol.status) == ref.status and sol.iterations == ref.iterations and sol.pivots == ref.pivots
    assert np.array_equal(sol.b_ixs, ref.b_ixs)
    assert abs(sol.z - ref.z) <= rtol * max(1.0, abs(ref.z))
    scale = max(1.0, float(np.abs(ref.x_b).max()))
    assert np.abs(sol.x_b.astype(np.float64) - ref.x_b.astype(np.float64)).max() <= rtol * scale


# ---------------------------------------------------------------- config 1: sample.txt

@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sample(lp, oracle, dtype):
    A, b, c = lp.read_lp(os.path.join(GOLDEN, "sample.txt"), dtype=dtype)
    sol = lp.solve(A, b, c)                                   # reference constants: eps 1e-4, MAX_ITER 5
    assert sol.status == lp.SolveStatus.OptimumFound and sol.iterations == 3 and sol.pivots == 2
    assert sol.z == 9.0 and sol.b_ixs.tolist() == [1, 0] and sol.x_b.tolist() == [3.0, 1.0]
    assert sol.trace.tolist() == [[0, 1], [1, 0]]
    assert lp.format_result(sol) == "# Iteration 1\n# Iteration 2\n# Iteration 3\nOptimum found: 9\n\tx_1 = 3\n\tx_0 = 1\n\n"
    assert sol.kernel_launches >= 1


def test_cli_stdout_matches_reference_contract(lp):
    exe = os.path.join(ROOT, "bin", "solver.out")
    out = subprocess.run([exe, os.path.join(GOLDEN, "sample.txt")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    head = "# Iteration 1\n# Iteration 2\n# Iteration 3\nOptimum found: 9\n\tx_1 = 3\n\tx_0 = 1\n\n"
    assert out.stdout.startswith(head)
    labels = [ln.split(":")[0].strip() for ln in out.stdout[len(head):].splitlines() if ln.strip()]
    assert labels == ["Total", "y", "p", "B_inv", "x_b", "Alloc", "Init", "Dealloc", "Host alloc", "Read file",
                      "Solve call", "Print result", "Host free"]                      # v4:456-471
    assert subprocess.run([exe], capture_output=True, text=True).returncode == 1      # v4:387-390
    r = subprocess.run([exe, "/nonexistent"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open /nonexistent." in r.stderr           # v4:396-399


def test_cli_reads_the_binary_twin_and_reports_parse_errors(lp, tmp_path):
    """bin/solver.out goes through b200lp_read_lp: binary LP files give the same stdout, short files the reference's message."""
    from simplex_method_gpu_b200.solver import write_lp_native
    exe = os.path.join(ROOT, "bin", "solver.out")
    A, b, c = lp.read_lp(os.path.join(GOLDEN, "sample.txt"), dtype=np.float32)
    binf = str(tmp_path / "sample.b200lp")
    write_lp_native(binf, A, b, c, binary=True)
    head = "# Iteration 1\n# Iteration 2\n# Iteration 3\nOptimum found: 9\n\tx_1 = 3\n\tx_0 = 1\n\n"
    for extra in ([], ["--f64"]):
        out = subprocess.run([exe, binf] + extra, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and out.stdout.startswith(head), out.stderr
    short = tmp_path / "short.txt"
    short.write_text("2 4\n1 1 1 0\n2 1 0\n")
    r = subprocess.run([exe, str(short)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "Failed to read (1,3) for A" in r.stderr                 # v4:99-100
    bad = tmp_path / "bad.txt"
    bad.write_text("5 4\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "Either failed to read m and n, or m > n." in r.stderr   # v4:403


def test_memory_cache_reuses_the_engine_without_changing_results(lp, oracle):
    A, b, c = oracle.gen_dense(300, 700, 2)
    base = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    assert lp.set_memory_cache(True) is False
    try:
        for _ in range(3):                                   # same shape: the cached engine is reused
            sol = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
            assert np.array_equal(sol.trace, base.trace) and np.array_equal(sol.x_b, base.x_b) and sol.z == base.z
        A2, b2, c2 = oracle.gen_dense(64, 160, 3)            # another shape: a fresh engine replaces it
        ref = oracle.solve(A2, b2, c2, eps=1e-9, max_iter=1 << 20, order=1)
        sol = lp.solve(A2, b2, c2, eps=1e-9, max_iter=1 << 20)
        assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.z == ref.z
        sol = lp.solve(A, b, c, eps=1e-9, max_iter=7)         # other options: not reused either
        assert sol.pivots == 7 and np.array_equal(sol.trace, base.trace[:7])
    finally:
        assert lp.set_memory_cache(False) is True


def test_max_iter_and_unbounded(lp, oracle):
    A, b, c = lp.read_lp(os.path.join(GOLDEN, "sample.txt"), dtype=np.float64)
    for k in (1, 2, 3):
        sol, ref = lp.solve(A, b, c, max_iter=k), oracle.solve(A, b, c, max_iter=k)
        assert int(sol.status) == ref.status and sol.iterations == ref.iterations and sol.pivots == ref.pivots
        assert np.array_equal(sol.x_b, ref.x_b) and np.array_equal(sol.b_ixs, ref.b_ixs) and sol.z == ref.z
    A = np.array([[-1.0, 1.0, 1.0]], order="F")
    sol = lp.solve(A, np.array([1.0]), np.array([1.0, 0.0, 0.0]), max_iter=10)
    assert sol.status == lp.SolveStatus.Unbounded and sol.iterations == 1 and sol.pivots == 0


# ---------------------------------------------------------------- configs 2-3 (scaled to oracle-seconds)

@pytest.mark.parametrize("m,n,seed", [(17, 40, 3), (64, 128, 1), (100, 228, 2), (256, 512, 1), (512, 1024, 1),
                                      (1024, 2048, 1)])
def test_dense_f64_matches_oracle(lp, oracle, m, n, seed):
    A, b, c = oracle.gen_dense(m, n, seed)
    ref = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    sol = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    assert ref.status == oracle.OPTIMUM
    assert_same_path(sol, ref, n, RTOL_F64)
    x = sol.x(n).astype(np.float64)
    assert np.all(A @ x <= b * (1 + 1e-9)) and np.all(x >= -1e-9)          # primal feasible


@pytest.mark.parametrize("m,seed", [(64, 1), (200, 4)])
def test_dense_f32_matches_oracle(lp, oracle, m, seed):
    A, b, c = oracle.gen_dense(m, 2 * m, seed, dtype=np.float32)
    ref = oracle.solve(A, b, c, eps=1e-4, max_iter=100000)
    sol = lp.solve(A, b, c, eps=1e-4, max_iter=100000)
    # fp32 has no 1e-12 tie rule: sums differ in the last bits, so only the optimum is compared
    assert sol.status == lp.SolveStatus.OptimumFound
    assert abs(sol.z - ref.z) <= 2e-4 * abs(ref.z)


def test_bit_exact_against_engine_order_oracle(lp, oracle):
    """order=1 replays the engine's summation order on the CPU: everything must be identical."""
    for m, n, seed in [(64, 128, 1), (300, 700, 2), (512, 1024, 3)]:
        A, b, c = oracle.gen_dense(m, n, seed)
        ref = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20, order=1)
        sol = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
        assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.trace[:, 1].tolist() == ref.trace_q.tolist()
        assert np.array_equal(sol.x_b, ref.x_b) and np.array_equal(sol.b_ixs, ref.b_ixs) and sol.z == ref.z


# ---------------------------------------------------------------- config 5: exact arithmetic, ties

@pytest.mark.parametrize("d", [4, 10, 14])
def test_klee_minty_exact(lp, oracle, d):
    A, b, c = oracle.gen_klee_minty(d)
    ref = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 22)
    sol = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 22)
    assert sol.status == lp.SolveStatus.OptimumFound and sol.pivots == 2 ** d - 1 and sol.z == 5.0 ** d
    assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.trace[:, 1].tolist() == ref.trace_q.tolist()
    assert np.array_equal(sol.x_b, ref.x_b) and np.array_equal(sol.b_ixs, ref.b_ixs)


def test_tiny_kernel_windows_and_state_round_trip(lp, oracle):
    """The shared-memory kernel reads and writes the same global state as the general one: uneven windows,
    B^-1 download, reset and the general kernel (explicit grid) all agree bit for bit."""
    A, b, c = oracle.gen_klee_minty(9)
    one = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 20, grid_ctas=1)          # general kernel
    assert one.pivots == 2 ** 9 - 1
    m, n = A.shape
    with lp.Engine(m, n, np.float64, eps=1e-4, max_iter=1 << 20) as e:         # auto -> shared-memory kernel
        e.upload(A, b, c)
        while True:
            r = e.run(37)
            if r["status"] != lp.SolveStatus.MaxIter:
                break
        assert r["pivots"] == one.pivots and r["z"] == one.z and r["iterations"] == one.iterations
        x_b, b_ixs, y = e.download()
        assert np.array_equal(x_b, one.x_b) and np.array_equal(b_ixs, one.b_ixs) and np.array_equal(e.trace(), one.trace)
        Binv = e.download_binv()
        assert np.abs(A[:, b_ixs] @ Binv - np.eye(m)).max() < 1e-9
        err, scale = e.check_basis()
        assert err <= 1e-12 * scale
        e.reset()
        assert e.run(1 << 20)["z"] == one.z


def test_klee_minty_20_full_config(lp, oracle):
    """BASELINE config 5a at full size: 2^20 - 1 pivots, optimum 5^20, every pivot identical to the oracle's."""
    A, b, c = oracle.gen_klee_minty(20)
    sol = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 22)
    assert sol.status == lp.SolveStatus.OptimumFound and sol.pivots == 2 ** 20 - 1 and sol.z == 5.0 ** 20
    assert sol.kernel_launches <= 4                                   # one persistent launch, not one per pivot
    ref = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 22)
    assert ref.pivots == sol.pivots and ref.z == sol.z
    assert np.array_equal(sol.trace[:, 0], ref.trace_p) and np.array_equal(sol.trace[:, 1], ref.trace_q)
    assert np.array_equal(sol.x_b, ref.x_b) and np.array_equal(sol.b_ixs, ref.b_ixs)


@pytest.mark.parametrize("k", [8, 16, 32, 64])
def test_assignment_exact_with_ties(lp, oracle, k):
    A, b, c, w = oracle.gen_assignment(k, 1)
    ref = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 20)
    sol = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 20)
    assert (ref.gap_q == 0).any()                   # lowest-index tie-break is exercised
    assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.trace[:, 1].tolist() == ref.trace_q.tolist()
    assert sol.z == ref.z and np.array_equal(sol.x_b, ref.x_b) and np.array_equal(sol.b_ixs, ref.b_ixs)


def test_golden_traces_on_gpu(lp, oracle):
    import json
    with open(os.path.join(GOLDEN, "traces.json")) as f:
        gold = json.load(f)
    for name, g in gold.items():
        if g["kind"] == "dense":
            A, b, c = oracle.gen_dense(g["m"], g["n"], g["seed"])
        elif g["kind"] == "klee_minty":
            A, b, c = oracle.gen_klee_minty(g["d"])
        else:
            A, b, c, _ = oracle.gen_assignment(g["k"], g["seed"])
        sol = lp.solve(A, b, c, eps=g["eps"], max_iter=g["max_iter"])
        assert int(sol.status) == g["status"] and sol.pivots == g["pivots"] and sol.iterations == g["iterations"], name
        h = len(g["p_head"])
        assert sol.trace[:h, 0].tolist() == g["p_head"] and sol.trace[:h, 1].tolist() == g["q_head"], name
        assert abs(sol.z - g["z"]) <= 1e-9 * max(1.0, abs(g["z"])), name


def test_small_shape_fuzz_bit_exact(lp, oracle):
    """Ragged tiny shapes, including n == m (no structural column), optimal-at-start, unbounded and degenerate
    LPs: status, iteration count, pivot sequence and every bit of x_b / z equal the engine-order oracle."""
    rng = np.random.default_rng(20261018)
    seen = set()
    for case in range(60):
        m = int(rng.integers(1, 41))
        n = m + (0 if case in (3, 17, 42) else int(rng.integers(0, 61)))    # n == m: nothing but the slack block
        ns = n - m
        A = np.zeros((m, n), order="F")
        A[:, :ns] = rng.uniform(-0.3 if case % 3 == 0 else 0.0, 1.0, (m, ns))
        A[:, ns:] = np.eye(m)
        b = rng.uniform(1.0, 2.0, m) * max(ns, 1)
        if case % 5 == 0:
            b[rng.integers(0, m)] = 0.0                                   # degenerate vertex, zero-length steps
        c = np.concatenate([rng.uniform(-0.5 if case % 4 == 0 else 0.1, 1.0, ns), np.zeros(m)])
        if case % 7 == 0:
            c[:ns] = -np.abs(c[:ns])                                      # optimal at the slack basis
        for dt, eps in ((np.float64, 1e-9), (np.float32, 1e-4)):
            Ad, bd, cd = A.astype(dt), b.astype(dt), c.astype(dt)
            ref = oracle.solve(Ad, bd, cd, eps=eps, max_iter=500, order=1)
            # auto = the shared-memory kernel for these sizes; an explicit grid = the general kernel (1 CTA, 3 CTAs)
            for kw in (dict(), dict(grid_ctas=1), dict(grid_ctas=3)):
                sol = lp.solve(Ad, bd, cd, eps=eps, max_iter=500, **kw)
                tag = (case, m, n, dt.__name__, kw)
                assert int(sol.status) == ref.status and sol.iterations == ref.iterations and sol.pivots == ref.pivots, tag
                assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.trace[:, 1].tolist() == ref.trace_q.tolist(), tag
                assert np.array_equal(sol.x_b, ref.x_b) and np.array_equal(sol.b_ixs, ref.b_ixs) and sol.z == ref.z, tag
            seen.add(ref.status)
    assert {oracle.OPTIMUM, oracle.UNBOUNDED} <= seen                     # both outcomes were exercised


# ---------------------------------------------------------------- engine properties

def test_geometry_independence(lp, oracle):
    """Grid size, tile shape and launch mode never change a bit of the result."""
    m, n = 384, 900
    A, b, c = oracle.gen_dense(m, n, 9)
    base = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    for kw in (dict(grid_ctas=1), dict(grid_ctas=7), dict(grid_ctas=148, tile_shape=1), dict(tile_shape=2),
               dict(tile_shape=4), dict(tile_shape=8), dict(mode=1), dict(check_slack=0), dict(price_cols=2),
               dict(price_mode=1), dict(price_mode=2), dict(price_mode=1, grid_ctas=5), dict(price_mode=2, grid_ctas=5),
               dict(price_mode=1, mode=1), dict(price_mode=2, mode=1)):
        sol = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20, **kw)
        assert np.array_equal(sol.trace, base.trace), kw
        assert np.array_equal(sol.x_b, base.x_b) and sol.z == base.z and np.array_equal(sol.b_ixs, base.b_ixs), kw


def test_windows_equal_one_shot_and_binv_is_inverse(lp, oracle):
    m, n = 256, 512
    A, b, c = oracle.gen_dense(m, n, 1)
    one = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20) as e:
        e.upload(A, b, c)
        assert e.dense_columns == n - m                      # slack block recognised, never stored
        total = 0
        while True:
            r = e.run(37)
            total += 1
            if r["status"] != lp.SolveStatus.MaxIter:
                break
        assert r["pivots"] == one.pivots and r["iterations"] == one.iterations and r["z"] == one.z
        x_b, b_ixs, y = e.download()
        assert np.array_equal(x_b, one.x_b) and np.array_equal(b_ixs, one.b_ixs)
        assert np.array_equal(e.trace(), one.trace)
        Binv = e.download_binv()
        ref = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20, want_Binv=True)
        Bmat = A[:, b_ixs]
        assert np.abs(Bmat @ Binv - np.eye(m)).max() < 1e-8              # B B^-1 = I
        assert np.abs(Binv - ref.Binv).max() <= 1e-9 * max(1.0, np.abs(ref.Binv).max())
        assert np.abs(y - ref.y).max() <= 1e-9 * max(1.0, np.abs(ref.y).max())
        # reset returns to the slack basis and the run repeats bit for bit
        e.reset()
        r2 = e.run(1 << 20)
        assert r2["pivots"] == one.pivots and r2["z"] == one.z


def test_check_basis_reports_drift_and_leaves_the_run_untouched(lp, oracle):
    m, n = 512, 1024
    A, b, c = oracle.gen_dense(m, n, 1)
    one = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20) as e:
        e.upload(A, b, c)
        err0, scale0 = e.check_basis()
        assert err0 == 0.0 and scale0 == np.abs(b).max()                    # slack basis: B^-1 = I, x_b = b
        e.run(150)
        err, scale = e.check_basis()                                         # mid-run, with a pending update
        assert 0 <= err <= 1e-9 * scale
        r = e.run(1 << 20)                                                   # the check must not disturb the path
        assert r["pivots"] == one.pivots and r["z"] == one.z and np.array_equal(e.trace(), one.trace)
        err, scale = e.check_basis()
        x_b, b_ixs, _ = e.download()
        want = np.abs(np.linalg.solve(A[:, b_ixs], b) - x_b).max()
        assert err <= 1e-9 * scale and want <= 1e-9 * scale


def test_phase_entry_points_against_numpy(lp, oracle):
    """One pivot through the per-phase C entry points, each checked on its own."""
    m, n = 200, 520
    A, b, c = oracle.gen_dense(m, n, 11)
    with lp.Engine(m, n, np.float64, eps=1e-9) as e:
        e.upload(A, b, c)
        Binv = np.eye(m)
        y = c[n - m:].copy()
        x_b = b.copy()
        c_b = c[n - m:].copy()
        for _ in range(6):
            p, mn = e.phase_price()
            red = y @ A - c                                       # v4:289-290
            assert p == int(np.argmin(red)) and abs(mn - red.min()) <= 1e-12 * max(1, abs(red.min()))
            e.phase_update_ftran(p)
            q, elig = e.phase_ratio()
            alpha = Binv @ A[:, p]                                # v4:307-308
            assert np.abs(e.vector("alpha") - alpha).max() <= 1e-12 * max(1.0, np.abs(alpha).max())
            theta = np.where(alpha > 0, x_b / np.where(alpha > 0, alpha, 1), np.inf)
            assert elig == int((alpha > 0).sum()) and q == int(np.argmin(theta))
            e.phase_pivot_update(p, q)
            row_q = Binv[q].copy()
            E = -alpha / alpha[q]
            E[q] = 1.0 / alpha[q] - 1.0
            Binv += np.outer(E, row_q)                            # v4:333
            c_b_q, c_b[q] = c_b[q], c[p]
            x_b += (row_q @ b) * E                                # v4:347-348
            y += ((c_b @ E) + (c[p] - c_b_q)) * row_q             # v4:354-356
            for name, want in (("E_q", E), ("row_q", row_q), ("x_b", x_b), ("y", y), ("c_b", c_b)):
                got = e.vector(name)
                assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max()), name
        got = e.download_binv()
        assert np.abs(got - Binv).max() <= 1e-11 * max(1.0, np.abs(Binv).max())


def test_non_identity_slack_block_is_priced_as_dense(lp, oracle):
    """The reference never checks its identity assumption (v4:272); the engine does and then
    treats all n columns as data, which is what the reference's GEMM/GEMV read."""
    m, n = 48, 120
    A, b, c = oracle.gen_dense(m, n, 2)
    A[:, n - m:] *= 1.0          # still identity
    A2 = A.copy(order="F")
    A2[3, n - 1] = 0.25          # break one slack entry
    ref = oracle.solve(A2, b, c, eps=1e-9, max_iter=50)
    with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=50) as e:
        e.upload(A2, b, c)
        assert e.dense_columns == n
        r = e.run(50)
        assert r["pivots"] == ref.pivots and int(r["status"]) == ref.status
        tr = e.trace()
        assert tr[:, 0].tolist() == ref.trace_p.tolist() and tr[:, 1].tolist() == ref.trace_q.tolist()
    # the one-call solve() starts pivoting before its host-side check of the slack block has finished and must
    # start over with all n columns as data when the block turns out not to be the identity
    sol = lp.solve(A2, b, c, eps=1e-9, max_iter=50)
    assert sol.pivots == ref.pivots and int(sol.status) == ref.status
    assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.trace[:, 1].tolist() == ref.trace_q.tolist()
    assert np.abs(sol.x_b - ref.x_b).max() <= 1e-9 * max(1.0, np.abs(ref.x_b).max())


def test_device_generator_matches_oracle_generator(lp, oracle):
    m, n = 128, 320
    A, b, c = oracle.gen_dense(m, n, 5)
    host = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20) as e:
        e.generate_dense(5)
        r = e.run(1 << 20)
        x_b, b_ixs, _ = e.download()
        assert r["pivots"] == host.pivots and r["z"] == host.z
        assert np.array_equal(x_b, host.x_b) and np.array_equal(b_ixs, host.b_ixs)


# ---------------------------------------------------------------- BASELINE's full sizes: size-independent properties

def _column(m, n, j, seed=1):
    """Column j of the synthetic dense LP [A_s, I] without materialising the matrix."""
    from simplex_method_gpu_b200.solver import lpgen_dense_into
    col = np.empty(m, np.float64)
    lpgen_dense_into(col.ctypes.data, 0, 0, m, n, j, 1, seed)
    return col


@pytest.mark.parametrize("m,n,head,tail", [(8192, 16384, 48, 400), (32768, 65536, 0, 40)])
def test_full_size_invariants(lp, oracle, m, n, head, tail):
    """C3 / C4 on one GPU: first pivots bit-exact against the engine-order oracle (C3 only: the oracle needs the
    matrix on the host), then properties that hold at any size: windows replay bit for bit, the objective never
    decreases, x_b stays feasible, B^-1 times a basic column is a unit vector, y = c_b B^-1."""
    rng = np.random.default_rng(5)
    with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 30) as e:
        e.generate_dense(1)
        if head:
            A, b, c = oracle.gen_dense(m, n, 1)
            ref = oracle.solve(A, b, c, eps=1e-9, max_iter=head, order=1)
            r = e.run(head)
            tr = e.trace()
            assert tr[:, 0].tolist() == ref.trace_p.tolist() and tr[:, 1].tolist() == ref.trace_q.tolist()
            x_b, b_ixs, y = e.download()
            assert np.array_equal(x_b, ref.x_b) and np.array_equal(b_ixs, ref.b_ixs) and r["z"] == ref.z
            del A
        z_prev = e.run(0)["z"]
        for _ in range(4):
            r = e.run(tail // 4)
            assert r["z"] >= z_prev and r["status"] == lp.SolveStatus.MaxIter
            z_prev = r["z"]
        x_b, b_ixs, y = e.download()
        trace1 = e.trace()
        assert x_b.min() >= -1e-9 * max(1.0, np.abs(x_b).max())
        assert len(set(b_ixs.tolist())) == m                                  # a basis: m distinct columns
        # B^-1 . A[:, b_ixs[k]] = e_k on a sample of basis positions (rows of B^-1 come back one window at a time)
        Binv = e.download_binv()
        for k in rng.choice(m, 6, replace=False):
            col = _column(m, n, int(b_ixs[k]))
            u = Binv @ col
            u[k] -= 1.0
            assert np.abs(u).max() < 1e-8, (k, np.abs(u).max())
        # y = c_b . B^-1 on a sample of columns (v4:353-356 keeps it by a linear update)
        c = np.array([0.0] * n)
        from simplex_method_gpu_b200.solver import lpgen_dense_into
        bb = np.empty(m)
        lpgen_dense_into(0, bb.ctypes.data, c.ctypes.data, m, n, 0, 0, 1)
        cb = c[b_ixs]
        for j in rng.choice(m, 6, replace=False):
            assert abs(cb @ Binv[:, j] - y[j]) <= 1e-8 * max(1.0, abs(y[j]))
        assert abs(cb @ x_b - z_prev) <= 1e-9 * max(1.0, abs(z_prev))         # z = c_b . x_b (v4:365)
        del Binv
        # the same windows again from the slack basis: identical pivots, identical bits
        e.reset()
        if head:
            e.run(head)
        for _ in range(4):
            r2 = e.run(tail // 4)
        assert r2["z"] == z_prev and np.array_equal(e.trace(), trace1)
        x_b2, b_ixs2, _ = e.download()
        assert np.array_equal(x_b2, x_b) and np.array_equal(b_ixs2, b_ixs)


@pytest.mark.parametrize("m,n,seed,dtype,eps", [
    (128, 256, 1, np.float64, 1e-9), (300, 700, 2, np.float64, 1e-9), (160, 1500, 3, np.float64, 1e-9),
    (777, 1600, 4, np.float64, 1e-9), (1024, 2048, 1, np.float64, 1e-9), (1100, 2150, 5, np.float64, 1e-9),
    (520, 1200, 6, np.float32, 1e-4), (1024, 2048, 2, np.float32, 1e-4)])
def test_resident_kernel_bit_identical_to_general(oracle, engine_lib, m, n, seed, dtype, eps):
    """simplex_resident (A and B^-1 in the shared memory of the grid, m ~ 128..1500) against the general persistent
    kernel (options.resident = -1) and the engine-order oracle: same pivots, x_b, b_ixs, z, y and B^-1 bit for bit;
    uneven windows carry the deferred rank-1 update across launches and across the two kernels."""
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(m, n, seed, dtype=dtype)
    ref = oracle.solve(A, b, c, eps=eps, max_iter=1 << 20, order=1, want_Binv=True)
    out = {}
    for tag, kw in (("resident", {}), ("general", {"resident": -1})):
        with lp.Engine(m, n, dtype, eps=eps, max_iter=1 << 20, **kw) as e:
            e.upload(A, b, c)
            r = e.run(3)
            while r["status"] == lp.SolveStatus.MaxIter:
                r = e.run(101)
            out[tag] = (r, e.download(), e.trace(), e.download_binv())
    (r1, (x1, i1, y1), t1, B1), (r2, (x2, i2, y2), t2, B2) = out["resident"], out["general"]
    assert r1["pivots"] == r2["pivots"] == ref.pivots and r1["z"] == r2["z"] == ref.z and r1["iterations"] == ref.iterations
    assert np.array_equal(t1, t2) and t1[:, 0].tolist() == ref.trace_p.tolist() and t1[:, 1].tolist() == ref.trace_q.tolist()
    assert np.array_equal(x1, x2) and np.array_equal(i1, i2) and np.array_equal(y1, y2) and np.array_equal(B1, B2)
    assert np.array_equal(x1, ref.x_b) and np.array_equal(i1, ref.b_ixs) and np.array_equal(B1, ref.Binv)
    # hand-over between the kernels in the middle of a solve: resident windows, then the general kernel finishes
    with lp.Engine(m, n, dtype, eps=eps, max_iter=1 << 20) as e:
        e.upload(A, b, c)
        e.run(min(17, ref.pivots))
        mid_x, mid_i, _ = e.download()
        k = e.trace().shape[0]
        assert e.trace()[:, 0].tolist() == ref.trace_p[:k].tolist()
        p, _ = e.phase_price()                       # the phase entry points see the resident kernel's state
        if k < ref.pivots:
            assert p == ref.trace_p[k]


def test_resident_kernel_degenerate_and_unbounded(oracle, engine_lib):
    import simplex_method_gpu_b200 as lp
    A, b, c, w = oracle.gen_assignment(64, 1)              # m = 128: ties everywhere
    ref = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 20, order=1)
    sol = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 20)
    gen = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 20, resident=-1)
    assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.trace[:, 1].tolist() == ref.trace_q.tolist()
    assert np.array_equal(sol.trace, gen.trace) and sol.z == gen.z == ref.z and np.array_equal(sol.x_b, gen.x_b)
    A, b, c = oracle.gen_dense(200, 440, 8)
    A[:, 5] = -np.abs(A[:, 5])                              # a column with no positive entry and an attractive cost: unbounded
    c[5] = 1e3
    ref = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20, order=1)
    sol = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    assert ref.status == oracle.UNBOUNDED and int(sol.status) == ref.status and sol.pivots == ref.pivots and sol.iterations == ref.iterations
