"""Profiling target (GPU box): L consecutive persistent launches of P pivots each on one dense LP, no reset in
between, so every launch after the first is steady state (pending rank-1 update, warm TLB).  With --phases the
same pivots run in one-launch-per-phase mode (k_price / k_update_ftran / k_ratio / k_book1 / k_book2).

    python tools/prof_target.py --lp 32768x65536 --pivots 8 --launches 3
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_method_gpu_b200 as lp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lp", default="8192x16384")
ap.add_argument("--pivots", type=int, default=16)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--phases", action="store_true")
ap.add_argument("--rule", type=int, default=0, help="pricing_rule: 1 = steepest edge")
ap.add_argument("--emu", type=int, default=0, help="R > 1: R sharded ranks emulated on device 0 in one cooperative launch")
ap.add_argument("--general", action="store_true", help="resident = -1: force the general kernel on a mid-size LP")
ap.add_argument("--km", type=int, default=0, help="Klee-Minty cube of this dimension instead of a dense LP (tiny kernel)")
a = ap.parse_args()
m, n = (int(x) for x in a.lp.lower().split("x"))
kw = dict(eps=1e-9, max_iter=1 << 30, mode=1 if a.phases else 0, pricing_rule=a.rule)
if a.general:
    kw["resident"] = -1
if a.emu > 1:
    kw["devices"] = [0] * a.emu
if a.km:
    dk = a.km                                  # Klee-Minty cube (BASELINE config 5a), all data exact integers
    A = np.zeros((dk, 2 * dk), order="F")
    for i in range(dk):
        for j in range(i):
            A[i, j] = 2.0 ** (i - j + 1)
        A[i, i] = A[i, dk + i] = 1.0
    b = 5.0 ** np.arange(1, dk + 1)
    c = np.concatenate([2.0 ** np.arange(dk - 1, -1, -1), np.zeros(dk)])
    m, n = A.shape
    kw["eps"] = 1e-4
    e = lp.Engine(m, n, np.float64, **kw)
    e.upload(A, b, c)
else:
    e = lp.Engine(m, n, np.float64, **kw)
    e.generate_dense(1)
for k in range(a.launches):
    r = e.run(a.pivots)
    print(f"launch {k}: {r['pivots']} pivots total, {r['ms_solve']:.3f} ms, status {int(r['status'])}", flush=True)
e.close()
