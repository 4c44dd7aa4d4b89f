"""Several GPUs behind the C ABI (b200lp_solve_*_multi / b200lp_create_multi, SURVEY.md 8(b2) `devices, ndev`).

On a single-GPU box the ranks are emulated: a repeated device ordinal makes the library run all ranks inside ONE
cooperative launch (slices of one grid) with the same sharded loop, mailboxes and flag protocol as on real peers —
so the driver's 1-GPU test box exercises the exchange code too.  With >= 2 GPUs the same tests also run on real peers.
The bar is the sharded engine's contract: bit-identical to the single-GPU engine in everything."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases(oracle):
    yield "dense 300x700", oracle.gen_dense(300, 700, 2), 1e-9
    yield "dense 1024x2048", oracle.gen_dense(1024, 2048, 1), 1e-9
    yield "dense 96x1000 (wide)", oracle.gen_dense(96, 1000, 4), 1e-9
    yield "klee-minty 10", oracle.gen_klee_minty(10), 1e-4
    yield "assignment 16", oracle.gen_assignment(16, 1)[:3], 1e-4
    yield "dense f32 200x520", oracle.gen_dense(200, 520, 3, dtype=np.float32), 1e-4


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0], [0] * 8]          # emulated ranks on one device
    if n >= 2:
        lists.append([0, 1])
    if n >= 4:
        lists.append([0, 1, 2, 3])
    if n >= 8:
        lists.append(list(range(8)))
    return lists


def test_one_call_multi_matches_single_gpu(oracle, engine_lib):
    import simplex_method_gpu_b200 as lp
    for name, (A, b, c), eps in _cases(oracle):
        one = lp.solve(A, b, c, eps=eps, max_iter=1 << 20)
        for devs in _device_lists():
            sol = lp.solve(A, b, c, eps=eps, max_iter=1 << 20, devices=devs)
            tag = f"{name} devices={devs}"
            assert sol.status == one.status and sol.pivots == one.pivots and sol.iterations == one.iterations, tag
            assert np.array_equal(sol.trace, one.trace), tag
            assert np.array_equal(sol.x_b, one.x_b) and np.array_equal(sol.b_ixs, one.b_ixs) and sol.z == one.z, tag


def test_multi_handle_windows_binv_and_drift(oracle, engine_lib):
    """Handle API on a multi-GPU engine: uneven windows carry the state over, B^-1 assembled from the row blocks equals
    the single-GPU one bit for bit, check_basis works on sharded rows, reset + rerun reproduces."""
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(640, 1400, 5)
    m, n = A.shape
    with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20) as e1:
        e1.upload(A, b, c)
        r1 = e1.run(1 << 20)
        B1 = e1.download_binv()
        x1, ix1, y1 = e1.download()
        tr1 = e1.trace()
        drift1 = e1.check_basis()
    for devs in _device_lists():
        with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20, devices=devs) as e:
            e.upload(A, b, c)
            r = e.run(7)
            while r["status"] == lp.SolveStatus.MaxIter:
                r = e.run(333)
            assert r["pivots"] == r1["pivots"] and r["z"] == r1["z"] and r["iterations"] == r1["iterations"], devs
            x, ix, y = e.download()
            assert np.array_equal(x, x1) and np.array_equal(ix, ix1) and np.array_equal(y, y1), devs
            assert np.array_equal(e.trace(), tr1), devs
            assert np.array_equal(e.download_binv(), B1), devs
            drift = e.check_basis()
            assert drift[1] == drift1[1] and drift[0] <= 1e-6 * drift[1], (devs, drift, drift1)
            e.reset()
            r2 = e.run(1 << 20)
            assert r2["pivots"] == r1["pivots"] and r2["z"] == r1["z"], devs


def test_multi_non_identity_slack_block(oracle, engine_lib):
    """The last m columns are data, not the identity the reference assumes (v4:272): all n columns are stored and
    priced on every layout, like the single-GPU engine does."""
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(128, 320, 7)
    A[:, -128:] += 0.25 * np.eye(128)          # still a valid starting basis for the loop's arithmetic, not the identity
    one = lp.solve(A, b, c, eps=1e-9, max_iter=400)
    for devs in _device_lists()[:2]:
        sol = lp.solve(A, b, c, eps=1e-9, max_iter=400, devices=devs)
        assert np.array_equal(sol.trace, one.trace) and sol.z == one.z and np.array_equal(sol.x_b, one.x_b), devs


def test_multi_argument_errors(engine_lib):
    import simplex_method_gpu_b200 as lp
    from simplex_method_gpu_b200 import capi
    A = np.zeros((4, 8), order="F")
    with pytest.raises(capi.B200LPError):
        lp.solve(A, np.ones(4), np.ones(8), devices=[0, 0, 99])
    with pytest.raises(capi.B200LPError):
        lp.solve(A, np.ones(4), np.ones(8), devices=[])
    with pytest.raises(capi.B200LPError):
        lp.Engine(4, 8, devices=[0] * 9)


@pytest.mark.parametrize("devices", [None, [0, 0]])
def test_abort_stops_at_an_iteration_boundary_and_can_continue(oracle, engine_lib, devices):
    """b200lp_abort: the loop polls one word per iteration (sharded: the request rides in the pricing record, so all
    ranks stop in the same iteration); the state stays consistent and the run can be continued to the same optimum."""
    import simplex_method_gpu_b200 as lp
    A, b, c = oracle.gen_dense(2048, 4096, 3)
    m, n = A.shape
    full = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    kw = {} if devices is None else {"devices": devices}
    with lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 20, **kw) as e:
        e.upload(A, b, c)
        e.run_async(1 << 20)
        time.sleep(0.01)
        e.abort()
        r = e.wait()
        if r["status"] == lp.SolveStatus.MaxIter:          # (a very fast box may already be done)
            assert r["aborted"] and 0 < r["pivots"] < full.pivots
        r = e.run(1 << 20)
        assert not r["aborted"] and r["status"] == lp.SolveStatus.OptimumFound
        assert r["pivots"] == full.pivots and r["z"] == full.z
        x_b, b_ixs, _ = e.download()
        assert np.array_equal(x_b, full.x_b) and np.array_equal(b_ixs, full.b_ixs)
        assert np.array_equal(e.trace(), full.trace)


def test_cli_multi_gpu_flag(oracle, engine_lib, tmp_path):
    """bin/solver.out --devices 0,0 / --gpus N: same stdout as the single-GPU run up to the timing values."""
    import os
    import subprocess
    import torch
    import simplex_method_gpu_b200 as lp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "bin", "solver.out")
    A, b, c = oracle.gen_dense(96, 224, 9)
    path = str(tmp_path / "lp.txt")
    lp.write_lp(path, A, b, c)
    base = [cli, path, "--f64", "--eps", "1e-9", "--max-iter", "100000"]
    one = subprocess.run(base, capture_output=True, text=True, timeout=300)
    two = subprocess.run(base + ["--devices", "0,0"], capture_output=True, text=True, timeout=300)
    assert one.returncode == 0 and two.returncode == 0, (one.stderr, two.stderr)
    assert "Optimum found" in one.stdout
    assert one.stdout.split("\n\n")[0] == two.stdout.split("\n\n")[0]
    if torch.cuda.device_count() >= 2:
        real = subprocess.run(base + ["--gpus", "2"], capture_output=True, text=True, timeout=300)
        assert real.returncode == 0 and one.stdout.split("\n\n")[0] == real.stdout.split("\n\n")[0]
