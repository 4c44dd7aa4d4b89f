"""Run under torchrun (one rank per GPU): the sharded engine must reproduce the single-GPU
engine bit for bit (trace, x_b, b_ixs, z, B^-1 rows) and therefore the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/sharded_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (test infrastructure)
import simplex_method_gpu_b200 as lp  # noqa: E402
from simplex_method_gpu_b200.sharded import ShardedEngine  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("dense", 300, 700, 2, 1e-9), ("dense", 1024, 2048, 1, 1e-9), ("dense", 96, 1000, 4, 1e-9),
             ("km", 10, 20, 0, 1e-4), ("assign", 16, 0, 1, 1e-4), ("dense32", 200, 520, 3, 1e-4)]
    if os.environ.get("SHARDED_STRESS"):       # exchange-dominated sizes, several seeds: a race would show up here first
        cases = [("dense", 2048, 4096, s, 1e-9) for s in (1, 2, 3)] + [("dense", 640, 1400, s, 1e-9) for s in (4, 5, 6, 7)] \
            + [("dense32", 1024, 2304, 8, 1e-4)]
    for kind, m, n, seed, eps in cases:
        dt = np.float64
        if kind == "dense":
            A, b, c = oracle.gen_dense(m, n, seed)
        elif kind == "dense32":
            dt = np.float32
            A, b, c = oracle.gen_dense(m, n, seed, dtype=np.float32)
        elif kind == "km":
            A, b, c = oracle.gen_klee_minty(m)
        else:
            A, b, c, _ = oracle.gen_assignment(m, seed)
        m, n = A.shape
        single = lp.solve(A, b, c, eps=eps, max_iter=1 << 20, device=local)
        eng = ShardedEngine(m, n, dt, rank=rank, world=world, device=local, eps=eps, max_iter=1 << 20)
        eng.upload(A, b, c)
        eng.connect()
        dist.barrier()
        # windows of uneven length: the state (pending update, barrier epochs) must carry over
        r = eng.run(7)
        while r["status"] == lp.SolveStatus.MaxIter:
            r = eng.run(1000)
        x_b, b_ixs, y = eng.download()
        tr = eng.trace()
        assert int(r["status"]) == int(single.status) and r["pivots"] == single.pivots, (kind, r, single.pivots)
        assert r["iterations"] == single.iterations
        assert np.array_equal(tr, single.trace), kind
        assert np.array_equal(x_b, single.x_b) and np.array_equal(b_ixs, single.b_ixs) and r["z"] == single.z, kind
        # B^-1 row block against the single-GPU engine's B^-1
        with lp.Engine(m, n, dt, eps=eps, max_iter=1 << 20, device=local) as e1:
            e1.upload(A, b, c)
            e1.run(1 << 20)
            full = e1.download_binv()
        r0, rows = eng.shard_rows()
        mine = eng.download_binv_rows()
        assert np.array_equal(mine, full[r0:r0 + rows]), kind
        # reset + rerun reproduces
        eng.reset()
        dist.barrier()
        r2 = eng.run(1 << 20)
        assert r2["pivots"] == single.pivots and r2["z"] == single.z
        eng.close()
        dist.barrier()
        if rank == 0:
            print(f"sharded ok: {kind} m={m} n={n} world={world}: {single.pivots} pivots, z={single.z!r}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
