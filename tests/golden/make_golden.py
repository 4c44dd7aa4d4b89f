"""Generates tests/golden/traces.json from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  Each entry is cross-checked against an
independent answer before it is written: HiGHS (scipy) for the dense LPs, 5^d / 2^d-1
for Klee-Minty, the Hungarian optimum for the assignment LPs.  The reference itself
can only run on the GPU box; its outputs are compared live in tests/test_reference_gpu.py.
"""
import json
import os
import sys

import numpy as np
from scipy.optimize import linear_sum_assignment, linprog

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402

HEAD = 64
out = {}


def entry(kind, s, **kw):
    return dict(kind=kind, status=int(s.status), pivots=int(s.pivots), iterations=int(s.iterations), z=float(s.z),
                p_head=s.trace_p[:HEAD].tolist(), q_head=s.trace_q[:HEAD].tolist(), **kw)


for m, seed in [(64, 1), (128, 2), (256, 1), (512, 1)]:
    A, b, c = oracle.gen_dense(m, 2 * m, seed)
    s = oracle.solve(A, b, c, eps=1e-9, max_iter=1 << 20)
    ref = linprog(-c[:m], A_ub=A[:, :m], b_ub=b, method="highs-ds")
    assert abs(s.z + ref.fun) <= 1e-9 * abs(ref.fun)
    out[f"dense_m{m}_s{seed}"] = entry("dense", s, m=m, n=2 * m, seed=seed, eps=1e-9, max_iter=1 << 20)

for d in (8, 12):
    A, b, c = oracle.gen_klee_minty(d)
    s = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 22)
    assert s.pivots == 2 ** d - 1 and s.z == 5.0 ** d
    out[f"klee_minty_d{d}"] = entry("klee_minty", s, d=d, eps=1e-4, max_iter=1 << 22)

for k in (8, 16):
    A, b, c, w = oracle.gen_assignment(k, 1)
    s = oracle.solve(A, b, c, eps=1e-4, max_iter=1 << 20)
    r, cc = linear_sum_assignment(-w)
    assert s.z == w[r, cc].sum()
    out[f"assignment_k{k}"] = entry("assignment", s, k=k, seed=1, eps=1e-4, max_iter=1 << 20)

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traces.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out), "entries")
