"""Dev tool (GPU box): time to optimal of a dense synthetic LP with either pricing rule.

    python tools/tto.py --lp 8192x16384 --rule 1
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_method_gpu_b200 as lp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lp", default="8192x16384")
ap.add_argument("--rule", type=int, default=1)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-pivots", type=int, default=1 << 40)
a = ap.parse_args()
m, n = (int(x) for x in a.lp.lower().split("x"))
e = lp.Engine(m, n, np.float64, eps=1e-9, max_iter=1 << 40, pricing_rule=a.rule)
e.generate_dense(a.seed)
e.run(4)
e.reset()
t0 = time.perf_counter()
r = e.run(a.max_pivots)
wall = time.perf_counter() - t0
drift = e.check_basis()
print(f"{a.lp} rule={a.rule}: status={int(r['status'])} pivots={r['pivots']} z={r['z']!r} "
      f"device {r['ms_solve']:.1f} ms ({r['ms_solve'] * 1e3 / max(r['pivots'], 1):.1f} us/pivot) wall {wall:.3f} s "
      f"drift {drift[0]:.3g} of {drift[1]:.3g}", flush=True)
e.close()
