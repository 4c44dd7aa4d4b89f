"""Dev tool (GPU box): sweep persistent-grid size and update+FTRAN tile shape on one workload,
and optionally run a few pivots in one-launch-per-phase mode (for an ncu launch list)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_method_gpu_b200 as lp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=8192)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--pivots", type=int, default=200)
ap.add_argument("--grids", default="148,296")
ap.add_argument("--shapes", default="1,2,4,8")
ap.add_argument("--phases", type=int, default=0, help="run this many pivots in mode 1 only")
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
dt = np.float64 if a.dtype == "f64" else np.float32

if a.phases:
    e = lp.Engine(a.m, a.n, dt, eps=1e-9, max_iter=1 << 30, mode=1)
    e.generate_dense(1)
    r = e.run(a.phases)
    print("mode1", r["pivots"], "pivots", r["ms_solve"], "ms", r["kernel_launches"], "launches")
    sys.exit(0)

bpp = np.dtype(dt).itemsize * (2 * a.m * a.m + a.m * (a.n - a.m))
for g in [int(x) for x in a.grids.split(",")]:
    for wc in [int(x) for x in a.shapes.split(",")]:
        e = lp.Engine(a.m, a.n, dt, eps=1e-9, max_iter=1 << 30, grid_ctas=g, tile_shape=wc)
        e.generate_dense(1)
        e.run(20)
        r0 = e.run(0)
        r = e.run(a.pivots)
        piv = r["pivots"] - r0["pivots"]
        ms = r["ms_solve"]
        print(f"grid={e.grid_ctas:4d} wc={wc} {piv} pivots {ms:9.2f} ms  {piv / ms * 1e3:9.1f} pivots/s  "
              f"{bpp * piv / ms / 1e6:8.1f} GB/s  {ms / piv * 1e3:8.1f} us/pivot", flush=True)
        e.close()
