"""CPU-only coverage of the N>1 host logic: partition arithmetic and the handle exchange,
with a world_size-2 gloo group (no GPU, no compute calls)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("m,n,world,itemsize", [(32768, 65536, 8, 8), (8192, 16384, 4, 8), (1000, 2600, 3, 8),
                                                 (20, 40, 8, 8), (300, 700, 2, 4), (64, 64, 2, 8)])
def test_shard_plan_partitions(m, n, world, itemsize):
    from simplex_method_gpu_b200.sharded import ShardPlan
    p = ShardPlan(m, n, world, itemsize)
    q = 32 * (16 // itemsize)
    assert p.ld % q == 0 and p.ld >= m and p.ld - m < q
    rows = [p.rows(r) for r in range(world)]
    cols = [p.cols(r) for r in range(world)]
    assert rows[0][0] == 0 and rows[-1][1] == p.ld and cols[0][0] == 0 and cols[-1][1] == n - m
    for r in range(world):
        assert rows[r][0] % q == 0 and rows[r][0] <= rows[r][1]            # warp-vector aligned, possibly empty
        if r:
            assert rows[r][0] == rows[r - 1][1] and cols[r][0] == cols[r - 1][1]   # contiguous, disjoint
    assert sum(p.bytes_per_pivot(r) for r in range(world)) == itemsize * (2 * m * m + m * (n - m))
    for j in (0, (n - m) // 2, n - m - 1):
        if n > m:
            c0, c1 = p.cols(p.owner_of_col(j))
            assert c0 <= j < c1
    for i in (0, m // 2, m - 1):
        r0, r1 = p.rows(p.owner_of_row(i))
        assert r0 <= i < r1
    sl = [p.slack_cols(r) for r in range(world)]
    assert sl[0][0] == 0 and sl[-1][1] == m and all(sl[r][0] == sl[r - 1][1] for r in range(1, world))


def test_plan_matches_native_partition(engine_lib):
    """ShardPlan is the Python mirror of the C++ partition; the native side reports its own through
    b200lp_shard_rows / b200lp_shard_columns on the GPU box (tests/sharded_check.py)."""
    from simplex_method_gpu_b200 import capi
    import ctypes as C
    h = C.c_void_p()
    o = capi.default_options()
    assert engine_lib.b200lp_create_sharded(capi.F64, 64, 128, 2, 2, C.byref(o), C.byref(h)) == capi.ERR_ARG  # rank >= nranks
    assert engine_lib.b200lp_create_sharded(capi.F64, 64, 128, 0, 9, C.byref(o), C.byref(h)) == capi.ERR_ARG  # > 8 ranks
    assert engine_lib.b200lp_ipc_handle_bytes() == 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from simplex_method_gpu_b200.sharded import ShardPlan, exchange_blobs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = bytes([rank * 16 + i for i in range(16)]) * 8             # 128-byte stand-in for the IPC handles
        allb = exchange_blobs(blob)
        ok = len(allb) == 128 * world and all(allb[128 * r:128 * (r + 1)] == bytes([r * 16 + i for i in range(16)]) * 8
                                              for r in range(world))
        # every rank derives the same plan and agrees on owners
        plan = ShardPlan(1000, 2600, world, 8)
        import torch
        mine = torch.tensor([plan.owner_of_col(777), plan.owner_of_row(555), plan.rows(rank)[0], plan.cols(rank)[1]])
        got = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        ok = ok and all(int(g[0]) == int(mine[0]) and int(g[1]) == int(mine[1]) for g in got)
        ok = ok and [int(g[2]) for g in got] == [plan.rows(r)[0] for r in range(world)]
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_handle_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_reference_arm_runs_on_rank0_only(tmp_path):
    """bench.py --impl reference under torchrun: non-zero ranks exit 0 without work or output."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "C2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
