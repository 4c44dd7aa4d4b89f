#!/usr/bin/env python
"""Benchmark of the hot path: pivots/s of the dense revised simplex loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload C2|C3|C4] [--pivots P]

A "step" is one window of P pivots of the loop at src/v4_cub_reduction.cu:286-359 on
a synthetic dense LP (oracle/lpgen_dense: A_s ~ U(0,1), b = (n_s/2)U(1,2),
c_s ~ U(0.5,1.5), slack block last).  Default workload = the configuration
BASELINE.json quotes its metric on: m=32768, n=65536, fp64 (fits one B200).

  value ..... pivots/s with the LP resident in HBM (CUDA events on the engine's stream)
  e2e ....... pivots/s through b200lp_solve_f64() with HOST buffers: per step the H2D of
              the LP, the pivots and the D2H of the result are all inside the timed region
  roofline .. algorithmic bytes per pivot 8*(2 m^2 + m (n-m)) / measured time, against
              the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of the reference loop) on the box's host cores,
              bounded sample, rank 0, N=1 only

--impl reference times the reference's own v4 CUDA solver (oracle/_ref, built from
/root/reference by oracle/make_ref.sh with the documented minimal patches) on the same
LP and window; when that library is unavailable it falls back to the CPU oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "C2": dict(m=1024, n=2048, pivots=1000),
    "C3": dict(m=8192, n=16384, pivots=1000),
    "C4": dict(m=32768, n=65536, pivots=192),
}
SEED = 1
EPS = 1e-9


def profiled_traffic(workload, pivots_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the persistent kernel from the committed ncu capture
    (profiles/r01_traffic.json, written by tools/ncu_extract.py), scaled to this launch's pivot count."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)[workload]
        return t["dram_bytes_per_pivot"] * pivots_per_launch, t
    except Exception:
        return None, None


def time_to_optimal(name, wl, dev=0):
    """Whole solve from the slack basis to the optimum on one GPU (the other half of BASELINE.json's metric)."""
    import simplex_method_gpu_b200 as lp
    m, n = wl["m"], wl["n"]
    eng = lp.Engine(m, n, np.float64, eps=EPS, max_iter=1 << 40, device=dev)
    eng.generate_dense(SEED)
    eng.run(8)                       # warm-up launch
    eng.reset()
    t0 = time.perf_counter()
    r = eng.run(1 << 40)
    wall = time.perf_counter() - t0
    x_b, b_ixs, y = eng.download()
    drift, scale = eng.check_basis()          # |B^-1 b - x_b| after all those rank-1 updates without a refactorisation
    eng.close()
    out = {"workload": f"{name}: dense LP m={m} n={n}, seed {SEED}, slack basis to optimum", "status": int(r["status"]),
           "basis_drift": {"max_abs_Binv_b_minus_x_b": drift, "max_abs_x_b": scale},
           "pivots": int(r["pivots"]), "iterations": int(r["iterations"]), "z": r["z"],
           "seconds": r["ms_solve"] * 1e-3, "wall_seconds": wall, "pivots_per_s": r["pivots"] / (r["ms_solve"] * 1e-3)}
    out["certificate"] = optimality_certificate(m, n, x_b, b_ixs, y, r["z"])
    return out


def optimality_certificate(m, n, x_b, b_ixs, y, z):
    """Solver-independent proof of optimality (GLPK is absent, HiGHS is too slow on a dense 8192 x 8192 LP):
    primal feasibility A_s x <= b, x >= 0, dual feasibility y >= 0, y'A_s >= c_s, and zero duality gap c'x = b'y,
    all evaluated on the host in fp64 from the engine's x_b / b_ixs / y and the regenerated LP."""
    import simplex_method_gpu_b200 as lp
    ns = n - m
    A = np.empty((m, ns), np.float64, order="F")
    b = np.empty(m, np.float64)
    c = np.empty(n, np.float64)
    lp.solver.lpgen_dense_into(A.ctypes.data, b.ctypes.data, c.ctypes.data, m, n, 0, ns, SEED)
    x = np.zeros(n)
    x[b_ixs] = x_b
    xs = x[:ns]
    primal_res = float(np.max(A @ xs - b))                 # <= 0 up to rounding
    dual_res = float(np.min(y @ A - c[:ns]))               # >= -eps: reduced costs of the structural columns
    cx, by = float(c[:ns] @ xs), float(b @ y)
    scale = max(1.0, abs(cx))
    return {"max_Ax_minus_b": primal_res, "min_x": float(x.min()), "min_y": float(y.min()),
            "min_reduced_cost": dual_res, "c_x": cx, "b_y": by, "rel_duality_gap": abs(cx - by) / scale,
            "rel_z_minus_c_x": abs(z - cx) / scale,
            "optimal_within": {"feasibility": 1e-7 * float(np.abs(b).max()), "eps": EPS, "gap": 1e-9},
            "holds": bool(primal_res <= 1e-7 * np.abs(b).max() and x.min() >= -1e-7 and y.min() >= -1e-7
                          and dual_res >= -1e-7 and abs(cx - by) <= 1e-9 * scale)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bytes_per_pivot(m, n, itemsize=8):
    return itemsize * (2 * m * m + m * (n - m))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_start(self):
        """Start of the timed region (the sampler itself is started earlier: nvidia-smi needs ~1 s to come up)."""
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        if self.t0 is not None:
            t1 = self.t1 if self.t1 is not None else time.time()
            inside = [r for t, r in rows if self.t0 <= t <= t1 + 0.06]
            # a window shorter than the sampling period: take the samples that bracket it
            rows = inside if inside else [r for t, r in rows if self.t0 - 0.25 <= t <= t1 + 0.25]
        else:
            rows = [r for _, r in rows]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------- our arm, one GPU

def run_b200_single(args, wl):
    import torch

    import simplex_method_gpu_b200 as lp
    m, n, P = wl["m"], wl["n"], args.pivots or wl["pivots"]
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)

    # ---- device-resident throughput ("value"): LP generated in HBM, windows of P pivots
    eng = lp.Engine(m, n, np.float64, eps=EPS, max_iter=1 << 40, device=dev)
    eng.generate_dense(SEED)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    # every step: back to the slack basis (untimed), then a window of P pivots (timed)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pivots_timed, launches, internal_ms = 0, 0, 0.0
    with ClockSampler(dev) as clk:
        for _ in range(args.warmup):
            eng.reset()
            r = eng.run(P)
        torch.cuda.synchronize()
        clk.mark_start()
        for a, b_ in ev:
            eng.reset()
            l0 = eng.run(0)["kernel_launches"]
            torch.cuda.synchronize()
            a.record(stream)
            eng.run_async(P)
            b_.record(stream)
            r = eng.wait()
            pivots_timed += r["pivots"]
            launches += r["kernel_launches"] - l0
            internal_ms += r["ms_solve"]
        torch.cuda.synchronize()
        clk.mark_end()
    step_ms = [a.elapsed_time(b_) for a, b_ in ev]
    total_ms = float(sum(step_ms))
    value = pivots_timed / (total_ms * 1e-3)
    status_after = int(r["status"])
    grid = eng.grid_ctas
    eng.close()

    # ---- end to end through the C ABI with host buffers ("e2e")
    e2e = None
    if not args.no_e2e:
        A_pin = torch.empty((n, m), dtype=torch.float64, pin_memory=True)    # column-major m x n
        b_pin = torch.empty(m, dtype=torch.float64, pin_memory=True)
        c_pin = torch.empty(n, dtype=torch.float64, pin_memory=True)
        A_np = A_pin.numpy().T                                                # Fortran-ordered view
        lp.solver.lpgen_dense_into(A_pin.data_ptr(), b_pin.data_ptr(), c_pin.data_ptr(), m, n, 0, n, SEED)
        times, piv = [], 0
        # a caller that solves one LP after another keeps the device buffers between calls
        # (b200lp_set_memory_cache): otherwise every call pays 40-500 ms of cudaMalloc / cudaFree for 16 GB
        lp.set_memory_cache(True)
        for s in range(args.warmup + args.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sol = lp.solve(A_np, b_pin.numpy(), c_pin.numpy(), eps=EPS, max_iter=P, device=dev, trace_cap=1)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                piv += sol.pivots
        lp.set_memory_cache(False)
        h2d = 8 * (m * (n - m) + m + n)         # dense columns + b + c (the slack block is verified on the host, not copied)
        d2h = 8 * m + 4 * m + 64                # x_b, b_ixs, result block
        e2e = {"value": piv / sum(times), "unit": "pivots/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * sum(times) / len(times), "pivots_per_step": P,
               "last_ms": {"upload": sol.ms_upload, "solve": sol.ms_solve, "download": sol.ms_download},
               "note": "b200lp_solve_f64 per step with pinned host buffers; device buffers kept between calls "
                       "(b200lp_set_memory_cache(1)); every step still uploads the whole LP and downloads the result"}
        del A_np, A_pin

    peak, peak_src = measured_peak_gbs()
    bpp = bytes_per_pivot(m, n)
    achieved = bpp * pivots_timed / (total_ms * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(args.workload, P)
    out = {
        "metric": "pivots/s, dense revised simplex (fp64)", "value": value, "unit": "pivots/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: dense LP m={m} n={n} (n counts the slack block), seed {SEED}, "
                               f"window of {P} pivots per step from the slack basis",
                   "m": m, "n": n, "pivots_per_step": P, "eps": EPS, "grid_ctas": grid,
                   "l2": "working set (A_N + B^-1) larger than L2, no flush needed" if bpp > 2 * 126e6 else
                         "working set fits the 126 MB L2 (L2-resident; roofline fraction may exceed 1)",
                   "parallelism": "1 GPU, persistent cooperative kernel"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_pivot": bpp,
                     "bytes_per_launch": bpp * P, "kernel": "simplex_persistent<double>"},
        "gpu_launches": int(launches), "clocks": clk.summary(), "e2e": e2e,
        "status_after": status_after, "pivots_timed": int(pivots_timed), "engine_event_ms": internal_ms,
    }
    return out


def cpu_baseline(wl, budget_s=20.0):
    """CPU oracle (port of the reference loop) on the host cores, bounded sample of the same LP."""
    import oracle
    m, n = wl["m"], wl["n"]
    A, b, c = oracle.gen_dense(m, n, SEED)
    k, done, t_used = 2, 0, 0.0
    while True:
        t0 = time.perf_counter()
        s = oracle.solve(A, b, c, eps=EPS, max_iter=k, trace_cap=1)
        dt = time.perf_counter() - t0
        done, t_used = s.pivots, dt
        if dt > budget_s / 3 or s.status != oracle.MAX_ITER or k >= 1 << 16:
            break
        k *= 4
    return {"value": done / t_used, "unit": "pivots/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": f"first {done} pivots of the same LP (m={m}, n={n}), {t_used:.1f} s, OpenMP over all host threads"}


# ---------------------------------------------------------------- reference arm

def run_reference(args, wl):
    """The reference's own v4 CUDA solver on the same LP; falls back to the CPU port."""
    import oracle
    m, n, P = wl["m"], wl["n"], args.pivots or wl["pivots"]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    A, b, c = oracle.gen_dense(m, n, SEED)
    use_v4 = oracle.ref_available(np.float64) and not args.ref_cpu
    if use_v4:
        try:
            import torch
            use_v4 = torch.cuda.is_available()
        except Exception:
            use_v4 = False
    times, piv, loop_s = [], 0, 0.0
    kind = "reference"
    if use_v4:
        P_ref = min(P, args.ref_pivots or P)
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            r = oracle.ref_solve(A, b, c, eps=EPS, max_iter=P_ref, trace_cap=1)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                piv += r.pivots
                loop_s += r.secs_loop
        sample = (f"reference v4 CUDA solve() (oracle/_ref/libv4ref_f64.so: fp64 retarget + init-grid/length/pointer-mode "
                  f"fixes) on the B200, {P_ref} iterations per step from the slack basis, host buffers in/out")
        cores = 1
        extra = {"device_loop_pivots_per_s": piv / loop_s if loop_s > 0 else None, "runs_on": "B200 (cuBLAS + CUB)"}
    else:
        kind = "port"
        P_ref = min(P, args.ref_pivots or 8)
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            r = oracle.solve(A, b, c, eps=EPS, max_iter=P_ref, trace_cap=1)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
                piv += r.pivots
        cores = oracle.num_threads()
        sample = f"CPU oracle port, {P_ref} iterations per step from the slack basis, {cores} OpenMP threads"
        extra = {"runs_on": "host CPU"}
    value = piv / sum(times)
    return {
        "impl": "reference", "metric": "pivots/s, dense revised simplex (fp64)", "value": value, "unit": "pivots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: dense LP m={m} n={n}, seed {SEED}, {P_ref} pivots per step", "m": m, "n": n,
                   "pivots_per_step": P_ref, "eps": EPS},
        "cpu_baseline": {"value": value, "unit": "pivots/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pivots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        **extra,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--pivots", type=int, default=0, help="pivots per step (0 = workload default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--extras", default="C3", help="comma list of extra single-GPU workloads reported under 'extra'")
    ap.add_argument("--tto", default="C2,C3", help="comma list of workloads solved to optimality (time-to-optimal), '' = none")
    ap.add_argument("--ref-cpu", action="store_true", help="reference arm: force the CPU port")
    ap.add_argument("--ref-pivots", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        out = run_reference(args, wl)
        if out is not None:
            print(json.dumps(out), flush=True)
        return

    if args.gpus > 1 or world > 1:
        from simplex_method_gpu_b200 import sharded_bench
        out = sharded_bench.run(args, wl, SEED, EPS)
    else:
        out = run_b200_single(args, wl)
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(wl)
        extra = {}
        for name in [x for x in args.extras.split(",") if x and x != args.workload]:
            sub = argparse.Namespace(**vars(args))
            sub.workload, sub.pivots, sub.no_e2e = name, 0, True
            r = run_b200_single(sub, WORKLOADS[name])
            extra[name] = {k: r[k] for k in ("value", "unit", "ms_per_step", "roofline", "config", "clocks", "gpu_launches")}
        tto = {}
        for name in [x for x in args.tto.split(",") if x]:
            tto[name] = time_to_optimal(name, WORKLOADS[name], int(os.environ.get("LOCAL_RANK", "0")))
        if tto:
            extra["time_to_optimal"] = tto
        if extra:
            out["extra"] = extra
    if rank == 0 and out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
