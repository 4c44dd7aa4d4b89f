"""Measure what the per-pivot exchanges would cost as NCCL collectives (north_star: "via NCCL or direct peer
stores, whichever measures lower latency").  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        tools/nccl_latency.py --rows 32768

Per pivot the sharded engine exchanges: X1 a 32-byte candidate per rank (all-gather), X2 the alpha slice
(rows/R doubles per rank, all-gather) + a 32-byte candidate, X3 row q (rows doubles, broadcast from its owner).
Timed with CUDA events over a CUDA-graph replay of 100 back-to-back collectives (no host launch gaps) and eagerly.
"""
import argparse
import json
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=32768)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--out", default="")
a = ap.parse_args()

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)

cand_in = torch.zeros(4, dtype=torch.float64, device=dev)                  # 32 B
cand_out = torch.zeros(4 * world, dtype=torch.float64, device=dev)
slice_in = torch.zeros(a.rows // world, dtype=torch.float64, device=dev)   # alpha slice
slice_out = torch.zeros(a.rows // world * world, dtype=torch.float64, device=dev)
row = torch.zeros(a.rows, dtype=torch.float64, device=dev)                 # row q


def ops():
    return {
        "all_gather 32 B/rank (X1 / X2 candidate)": lambda: dist.all_gather_into_tensor(cand_out, cand_in),
        f"all_gather {slice_in.numel() * 8} B/rank (X2 alpha slice)": lambda: dist.all_gather_into_tensor(slice_out, slice_in),
        f"broadcast {row.numel() * 8} B (X3 row q)": lambda: dist.broadcast(row, src=0),
    }


def timed(fn, iters, graph):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        s.record()
        g.replay()
        e.record()
    else:
        s.record()
        for _ in range(iters):
            fn()
        e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) * 1e3 / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {"world": world, "rows": a.rows, "unit": "us per collective (max over ranks)", "ops": {}}
for name, fn in ops().items():
    entry = {"eager": timed(fn, a.iters, False)}
    try:
        entry["cuda_graph"] = timed(fn, a.iters, True)
    except Exception as exc:                                   # graph capture of NCCL not available
        entry["cuda_graph"] = None
        entry["graph_error"] = str(exc)[:120]
    res["ops"][name] = entry
g = [min(x for x in (v["cuda_graph"], v["eager"]) if x is not None) for v in res["ops"].values()]
res["per_pivot_if_nccl_us"] = 2 * g[0] + g[1] + g[2]
res["note"] = "per pivot = 2 candidate all-gathers + 1 alpha all-gather + 1 row broadcast, before any local barrier"
if rank == 0:
    print(json.dumps(res, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
dist.barrier()
dist.destroy_process_group()
