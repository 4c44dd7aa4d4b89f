"""CPU-only: tools/lp_convert.py — the working replacement of the reference's GLPK side tools
(glpk_interface.cpp:16-104 MPS -> text, solver_glpk.cpp:15-39 MPS solve)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def conv():
    spec = importlib.util.spec_from_file_location("lp_convert", os.path.join(ROOT, "tools", "lp_convert.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_sample_round_trip_and_optimum(conv, tmp_path, capsys):
    mps, txt = str(tmp_path / "s.mps"), str(tmp_path / "s.txt")
    assert conv.main(["to-mps", os.path.join(GOLDEN, "sample.txt"), mps]) == 0
    assert conv.main(["from-mps", mps, txt]) == 0
    A0, b0, c0 = conv.read_text(os.path.join(GOLDEN, "sample.txt"))
    A1, b1, c1 = conv.read_text(txt)
    assert np.array_equal(A0, A1) and np.array_equal(b0, b1) and np.array_equal(c0, c1)
    capsys.readouterr()
    assert conv.main(["solve", mps]) == 0
    out = capsys.readouterr().out
    assert out == "x[1] = 1\nx[2] = 3\nOptimal objective: 9\n"          # input/sample.txt:15-16, solver_glpk.cpp:30-36 format


def test_dense_lp_survives_mps_and_matches_oracle(conv, oracle, tmp_path):
    A, b, c = oracle.gen_dense(24, 60, 3)
    mps = str(tmp_path / "d.mps")
    conv.write_mps(mps, A, b, c)
    A1, b1, c1 = conv.mps_to_standard(*conv.read_mps(mps)[:5])
    assert A1.shape == A.shape and np.allclose(A1, A, rtol=1e-8, atol=0) and np.allclose(b1, b, rtol=1e-8) \
        and np.allclose(c1, c, rtol=1e-8)                                    # 12-character MPS number fields
    status, z, _ = conv.solve_highs(A, b, c)
    ref = oracle.solve(A, b, c, eps=1e-9, max_iter=10000)
    assert status == 0 and ref.status == oracle.OPTIMUM and abs(z - ref.z) <= 1e-9 * abs(ref.z)


def test_row_senses_are_converted(conv, tmp_path):
    """G rows are negated, E rows split, the slack identity is appended (what glpk_interface.cpp never did)."""
    p = tmp_path / "r.mps"
    p.write_text("NAME T\nROWS\n N COST\n L R0\n G R1\n E R2\nCOLUMNS\n X0 COST 1 R0 1\n X0 R1 1 R2 1\n X1 COST 2 R0 1\n"
                 " X1 R2 -1\nRHS\n RHS R0 4 R1 1\n RHS R2 0\nENDATA\n")
    A, sense, b, c, neg, names = conv.read_mps(str(p))
    assert sense == ["L", "G", "E"] and names == ["X0", "X1"] and not neg
    F, rhs, cc = conv.mps_to_standard(A, sense, b, c, neg)
    assert F.shape == (4, 6) and np.array_equal(F[:, 2:], np.eye(4))
    assert F[:, :2].tolist() == [[1, 1], [-1, 0], [1, -1], [-1, 1]] and rhs.tolist() == [4, -1, 0, 0]
    assert cc.tolist() == [1, 2, 0, 0, 0, 0]
    with pytest.raises(ValueError):
        q = tmp_path / "b.mps"
        q.write_text("NAME T\nROWS\n N COST\n L R0\nCOLUMNS\n X0 COST 1 R0 1\nRHS\n RHS R0 4\nBOUNDS\n UP BND X0 3\nENDATA\n")
        conv.read_mps(str(q))


def test_text_binary_conversion(conv, engine_lib, tmp_path):
    binf, txt = str(tmp_path / "s.b200lp"), str(tmp_path / "s.txt")
    assert conv.main(["to-bin", os.path.join(GOLDEN, "sample.txt"), binf]) == 0
    assert conv.main(["from-bin", binf, txt]) == 0
    A0, b0, c0 = conv.read_text(os.path.join(GOLDEN, "sample.txt"))
    A1, b1, c1 = conv.read_text(txt)
    assert np.array_equal(A0, A1) and np.array_equal(b0, b1) and np.array_equal(c0, c1)
    with open(binf, "rb") as f:
        assert f.read(8) == b"B200LP1\0"


def test_solve_with_engine_needs_a_gpu(conv, engine_lib, capsys):
    import torch
    rc = conv.main(["solve", os.path.join(GOLDEN, "sample.txt"), "--engine"])
    out = capsys.readouterr()
    if torch.cuda.is_available():
        assert rc == 0 and out.out == "x[1] = 1\nx[2] = 3\nOptimal objective: 9\n"
    else:
        assert rc == 2 and "no CUDA device" in out.err          # no CPU fallback behind --engine


def test_glpk_harness_builds_and_degrades_loudly(conv, tmp_path):
    """tools/solver_glpk_harness.cpp = solver_glpk.cpp's contract (/root/reference/solver_glpk.cpp:15-39), compiled
    against GLPK only where <glpk.h> exists.  Here GLPK is absent: the binary must say so and exit 3 (never pretend);
    where GLPK exists it must print the optimum of the sample LP like solver_glpk.cpp does."""
    import subprocess
    from simplex_method_gpu_b200 import _build
    _build.build()
    assert os.path.exists(_build.GLPK_PATH)
    mps = str(tmp_path / "sample.mps")
    assert conv.main(["to-mps", os.path.join(GOLDEN, "sample.txt"), mps]) == 0
    r = subprocess.run([_build.GLPK_PATH, mps], capture_output=True, text=True, timeout=60)
    have = any(os.path.exists(os.path.join(d, "glpk.h")) for d in ("/usr/include", "/usr/local/include"))
    if have:
        assert r.returncode == 0 and "Optimal objective: " in r.stdout
        assert abs(abs(float(r.stdout.split("Optimal objective: ")[1].split()[0])) - 9.0) < 1e-9
    else:
        assert r.returncode == 3 and "built without GLPK" in r.stderr
    assert subprocess.run([_build.GLPK_PATH], capture_output=True, text=True).returncode == 1
