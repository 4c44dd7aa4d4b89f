"""Modes outside the reference's parity contract: they change the pivot sequence, so the
checker is the oracle running the SAME rule with the engine's summation order (bit-exact traces), plus the
mode-independent facts: same optimum as the reference rule, optimality certificate."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(sol, ref, tag):
    assert int(sol.status) == ref.status and sol.pivots == ref.pivots and sol.iterations == ref.iterations, \
        (tag, sol.status, sol.pivots, ref.status, ref.pivots)
    assert sol.trace[:, 0].tolist() == ref.trace_p.tolist() and sol.trace[:, 1].tolist() == ref.trace_q.tolist(), tag
    assert np.array_equal(sol.x_b, ref.x_b) and np.array_equal(sol.b_ixs, ref.b_ixs) and sol.z == ref.z, tag


@pytest.mark.parametrize("m,n,seed,dtype,eps", [
    (64, 128, 1, np.float64, 1e-9), (192, 448, 1, np.float64, 1e-9), (300, 700, 2, np.float64, 1e-9),
    (77, 300, 5, np.float64, 1e-9), (1024, 2048, 1, np.float64, 1e-9), (2048, 4096, 1, np.float64, 1e-9),
    (200, 520, 3, np.float32, 1e-4)])
def test_steepest_edge_bit_exact_against_oracle(oracle, engine_lib, m, n, seed, dtype, eps):
    """pricing_rule = 1 (README.md:16-17 'steepest edge with a recurrence'): same pivots, x_b, b_ixs and z as the
    oracle's steepest edge in the engine's summation order; same optimum as Dantzig; far fewer pivots."""
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(m, n, seed, dtype=dtype)
    ref = oracle.solve(A, b, c, eps=eps, max_iter=1 << 20, order=1, pricing_rule=1)
    sol = lp.solve(A, b, c, eps=eps, max_iter=1 << 20, pricing_rule=1)
    _same(sol, ref, f"steepest edge {m}x{n}")
    dz = lp.solve(A, b, c, eps=eps, max_iter=1 << 20)
    assert sol.status == lp.SolveStatus.OptimumFound
    assert abs(sol.z - dz.z) <= (1e-9 if dtype == np.float64 else 5e-4) * abs(dz.z)
    if m >= 192:
        assert sol.pivots < dz.pivots


def test_steepest_edge_windows_geometry_and_exact_problems(oracle, engine_lib):
    """The pending weight recurrence crosses launch boundaries (windows of uneven length), the result does not
    depend on the grid / fused-prologue choice, and Klee-Minty / assignment stay exact."""
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(640, 1400, 5)
    m, n = A.shape
    ref = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20, order=1, pricing_rule=1)
    for kw in ({}, {"grid_ctas": 3}, {"fuse_book2": -1}, {"grid_ctas": 148, "tile_shape": 2}):
        with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20, pricing_rule=1, **kw) as e:
            e.upload(A, b, c)
            r = e.run(5)
            while r["status"] == lp.SolveStatus.MaxIter:
                r = e.run(37)
            x_b, b_ixs, _ = e.download()
            tr = e.trace()
            assert r["pivots"] == ref.pivots and r["z"] == ref.z, kw
            assert tr[:, 0].tolist() == ref.trace_p.tolist() and tr[:, 1].tolist() == ref.trace_q.tolist(), kw
            assert np.array_equal(x_b, ref.x_b) and np.array_equal(b_ixs, ref.b_ixs), kw
            e.reset()                                    # weights start over with the slack basis
            r2 = e.run(1 << 20)
            assert r2["pivots"] == ref.pivots and r2["z"] == ref.z, kw
    A, b, c = oracle.gen_klee_minty(12)
    ref = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 20, order=1, pricing_rule=1)
    sol = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 20, pricing_rule=1, grid_ctas=2)
    _same(sol, ref, "klee-minty 12")
    assert sol.z == 5.0 ** 12
    A, b, c, w = oracle.gen_assignment(16, 1)
    ref = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 20, order=1, pricing_rule=1)
    sol = lp.solve(A, b, c, eps=1e-4, max_iter=1 << 20, pricing_rule=1)
    _same(sol, ref, "assignment 16")


def test_steepest_edge_rejected_where_unsupported(engine_lib):
    import simplex_method_gpu_b200 as lp
    from simplex_method_gpu_b200 import capi
    with pytest.raises(capi.B200LPError):
        lp.Engine(64, 128, pricing_rule=1, mode=1)
    with pytest.raises(capi.B200LPError):
        lp.Engine(64, 128, pricing_rule=7)


@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_steepest_edge_sharded_bit_exact_against_oracle(oracle, engine_lib, ranks):
    """Steepest edge on several ranks (emulated on one device here, real peers where there are any): gamma is
    column-sharded like A, v = B^-T alpha is the rank-ordered sum of the ranks' partial column sums (exchange X4).  The
    oracle replays exactly that association (nranks), so the traces must be equal bit for bit; the optimum equals the
    single-GPU steepest-edge and Dantzig optima to 1e-9."""
    import torch
    import simplex_method_gpu_b200 as lp
    layouts = [[0] * ranks]
    if torch.cuda.device_count() >= ranks:
        layouts.append(list(range(ranks)))
    for name, (A, b, c), eps in (("dense 300x700", oracle.gen_dense(300, 700, 2), 1e-9),
                                 ("dense 1024x2048", oracle.gen_dense(1024, 2048, 1), 1e-9),
                                 ("dense 96x1000", oracle.gen_dense(96, 1000, 4), 1e-9),
                                 ("assignment 16", oracle.gen_assignment(16, 1)[:3], 1e-4)):
        ref = oracle.solve(A, b, c, eps=eps, max_iter=1 << 20, order=1, pricing_rule=1, nranks=ranks)
        one = lp.solve(A, b, c, eps=eps, max_iter=1 << 20, pricing_rule=1)
        for devs in layouts:
            sol = lp.solve(A, b, c, eps=eps, max_iter=1 << 20, pricing_rule=1, devices=devs)
            _same(sol, ref, f"{name} devices={devs}")
            assert abs(sol.z - one.z) <= 1e-9 * max(1.0, abs(one.z))
    # windows: the pending recurrence and the exchange epochs cross launch boundaries
    A, b, c = oracle.gen_dense(640, 1400, 5)
    ref = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20, order=1, pricing_rule=1, nranks=ranks)
    with lp.Engine(640, 1400, np.float64, eps=1e-9, max_iter=1 << 20, pricing_rule=1, devices=[0] * ranks) as e:
        e.upload(A, b, c)
        r = e.run(3)
        while r["status"] == lp.SolveStatus.MaxIter:
            r = e.run(41)
        assert r["pivots"] == ref.pivots and r["z"] == ref.z
        tr = e.trace()
        assert tr[:, 0].tolist() == ref.trace_p.tolist() and tr[:, 1].tolist() == ref.trace_q.tolist()


@pytest.mark.parametrize("tol", [1e-9, 1e-3])
def test_pivot_tol_matches_oracle(oracle, engine_lib, tol):
    """pivot_tol: eligibility alpha > pivot_tol (v4:203 is the strict alpha > 0 = default)."""
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(256, 600, 4)
    ref = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20, order=1, pivot_tol=tol)
    sol = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20, pivot_tol=tol)
    _same(sol, ref, f"pivot_tol {tol}")
    multi = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20, pivot_tol=tol, devices=[0, 0])
    _same(multi, ref, f"pivot_tol {tol} multi")


@pytest.mark.parametrize("kw", [dict(ratio_mode=1), dict(ratio_mode=2, harris_delta=1e-9), dict(ratio_mode=2, harris_delta=1e-6),
                                dict(ratio_mode=2, harris_delta=1e-9, pricing_rule=1), dict(ratio_mode=1, pivot_tol=1e-7)])
def test_ratio_modes_bit_exact_against_oracle(oracle, engine_lib, kw):
    """The reference's open items (README.md:29-30): bounded ratio test (x_b clamped at 0) and the Harris two-pass
    test with the largest pivot element; same rule in the oracle with the engine's summation order -> same pivots."""
    import simplex_method_gpu_b200 as lp
    for name, (A, b, c), eps in (("dense 300x700", oracle.gen_dense(300, 700, 2), 1e-9),
                                 ("dense 1024x2048", oracle.gen_dense(1024, 2048, 1), 1e-9),
                                 ("assignment 16", oracle.gen_assignment(16, 1)[:3], 1e-4),
                                 ("klee-minty 10", oracle.gen_klee_minty(10), 1e-4)):
        ref = oracle.solve(A, b, c, eps=eps, max_iter=1 << 20, order=1, **kw)
        sol = lp.solve(A, b, c, eps=eps, max_iter=1 << 20, **kw)
        _same(sol, ref, f"{name} {kw}")
        base = lp.solve(A, b, c, eps=eps, max_iter=1 << 20)
        if sol.status == lp.SolveStatus.OptimumFound:
            assert abs(sol.z - base.z) <= 1e-9 * max(1.0, abs(base.z)), (name, kw)


def test_harris_rejected_where_unsupported(engine_lib):
    import simplex_method_gpu_b200 as lp
    from simplex_method_gpu_b200 import capi
    for kw in (dict(ratio_mode=2, devices=[0, 0]), dict(ratio_mode=2, mode=1), dict(ratio_mode=3)):
        with pytest.raises(capi.B200LPError):
            lp.Engine(64, 128, **kw)


def test_refactorisation_rebuilds_the_inverse(oracle, engine_lib):
    """b200lp_refactor (the reference lists the numerical guards as open, README.md:29-30): B^-1 rebuilt from the
    basis alone must invert the basis matrix, x_b / y must be B^-1 b / c_b^T B^-1 (checked with numpy, an independent
    implementation), and the solve continues to the same optimum."""
    import simplex_method_gpu_b200 as lp
    for (m, n, seed, k) in ((300, 700, 2, 60), (1024, 2048, 1, 500)):
        A, b, c = oracle.gen_dense(m, n, seed)
        full = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
        for kw in ({}, {"pricing_rule": 1}):
            with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20, **kw) as e:
                e.upload(A, b, c)
                r = e.run(k)
                x0, ix0, y0 = e.download()
                replayed = e.refactor()
                x1, ix1, y1 = e.download()
                Binv = e.download_binv()
                assert np.array_equal(ix0, ix1)
                assert replayed == int(np.sum(ix1 != np.arange(n - m, n)))
                B = A[:, ix1]
                assert np.abs(B @ Binv - np.eye(m)).max() <= 1e-9
                assert np.abs(x1 - np.linalg.solve(B, b)).max() <= 1e-9 * np.abs(x1).max()
                assert np.abs(y1 - np.linalg.solve(B.T, c[ix1])).max() <= 1e-9 * max(1.0, np.abs(y1).max())
                assert np.abs(x1 - x0).max() <= 1e-9 * np.abs(x0).max() and np.abs(y1 - y0).max() <= 1e-9 * max(1.0, np.abs(y0).max())
                assert e.check_basis()[0] <= 1e-10 * np.abs(x1).max()
                r = e.run(1 << 20)
                assert r["status"] == lp.SolveStatus.OptimumFound and abs(r["z"] - full.z) <= 1e-9 * abs(full.z), kw
                # the rebuilt inverse differs from the product-form one in the last bits only, and a pending steepest-edge
                # weight recurrence survives the refactorisation: the next pivots are the ones of the undisturbed run
                plain = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20, **kw)
                tr = e.trace()
                assert np.array_equal(tr[:k + 12], plain.trace[:k + 12]), kw


def test_guarded_run_refactorises_on_drift(oracle, engine_lib):
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(512, 1200, 3)
    full = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    with lp.Engine(512, 1200, np.float64, eps=1e-9, max_iter=1 << 20) as e:
        e.upload(A, b, c)
        r, nref = e.run_guarded(1 << 20, 64, 0.0)        # drift_tol 0: every window triggers a refactorisation
        assert r["status"] == lp.SolveStatus.OptimumFound and nref >= 1
        assert abs(r["z"] - full.z) <= 1e-9 * abs(full.z)
    with lp.Engine(512, 1200, np.float64, eps=1e-9, max_iter=1 << 20) as e:
        e.upload(A, b, c)
        r, nref = e.run_guarded(1 << 20, 64, 1e-6)       # a healthy inverse never triggers one
        assert nref == 0 and r["pivots"] == full.pivots and r["z"] == full.z
