"""Time to optimal on G GPUs (BASELINE.json's metric: "pivots/sec & time-to-optimal, dense m=32k LP, 1/2/4/8 B200").
Run under torchrun, one rank per GPU; with one process it uses the single-GPU engine.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
        tools/solve_sharded.py --lp 32768x65536 --out gpurun_out/tto_c4_8gpu.json

The LP is generated on the devices, solved from the slack basis in windows of --window pivots (device time summed,
max over ranks), and rank 0 proves optimality on the host (bench.optimality_certificate: primal and dual
feasibility, zero duality gap) — no CPU LP solver could do a dense 32768 x 32768 LP in reasonable time.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplex_method_gpu_b200 as lp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lp", default="8192x16384")
ap.add_argument("--window", type=int, default=20000)
ap.add_argument("--max-pivots", type=int, default=1 << 40)
ap.add_argument("--no-certificate", action="store_true")
ap.add_argument("--rule", type=int, default=0, help="pricing_rule: 0 Dantzig (the reference's), 1 steepest edge")
ap.add_argument("--out", default="")
a = ap.parse_args()
m, n = (int(x) for x in a.lp.lower().split("x"))
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
EPS, SEED = 1e-9, 1

if world > 1:
    import torch
    import torch.distributed as dist
    from simplex_method_gpu_b200.sharded import ShardedEngine
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    e = ShardedEngine(m, n, np.float64, rank=rank, world=world, device=local, eps=EPS, max_iter=1 << 40, pricing_rule=a.rule)
    e.generate_dense(SEED)
    e.connect()
    dist.barrier()
else:
    e = lp.Engine(m, n, np.float64, eps=EPS, max_iter=1 << 40, pricing_rule=a.rule)
    e.generate_dense(SEED)

ms, t0, windows = 0.0, time.perf_counter(), []
while True:
    r = e.run(a.window)
    ms += r["ms_solve"]
    windows.append((int(r["pivots"]), round(ms * 1e-3, 3), r["z"]))
    if rank == 0:
        print(f"  {r['pivots']:>9d} pivots  {ms * 1e-3:9.2f} s device  z = {r['z']:.12g}", flush=True)
    if r["status"] != lp.SolveStatus.MaxIter or r["pivots"] >= a.max_pivots:
        break
wall = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
x_b, b_ixs, y = e.download()
res = {"workload": f"dense LP m={m} n={n}, seed {SEED}, eps {EPS}, slack basis to optimum", "n_gpus": world,
       "pricing_rule": "steepest edge (Goldfarb-Reid recurrence)" if a.rule else "Dantzig (the reference's rule)",
       "status": int(r["status"]), "pivots": int(r["pivots"]), "iterations": int(r["iterations"]), "z": r["z"],
       "seconds_device": ms * 1e-3, "seconds_wall": wall, "pivots_per_s": r["pivots"] / (ms * 1e-3),
       "window": a.window, "progress": windows[:: max(1, len(windows) // 12)]}
e.close()
if rank == 0:
    if not a.no_certificate:
        import bench
        res["certificate"] = bench.optimality_certificate(m, n, x_b, b_ixs, y, r["z"])
    print(json.dumps(res), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
