"""Dev tool (GPU box): N back-to-back b200lp_solve_f64 calls on the m=32768 LP with pinned host buffers; prints every
call's wall time (the spread is cudaMalloc / cudaFree of 16 GB on a shared host: 909 ms ... 2.5 s, median 915-945 ms).

    python tools/e2e_loop.py 14 tag
"""
import sys, time, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplex_method_gpu_b200 as lp
m, n, P = 32768, 65536, 192
A_pin = torch.empty((n, m), dtype=torch.float64, pin_memory=True)
b_pin = torch.empty(m, dtype=torch.float64, pin_memory=True); c_pin = torch.empty(n, dtype=torch.float64, pin_memory=True)
lp.solver.lpgen_dense_into(A_pin.data_ptr(), b_pin.data_ptr(), c_pin.data_ptr(), m, n, 0, n, 1)
A = A_pin.numpy().T
ts = []
for k in range(int(sys.argv[1])):
    t0 = time.perf_counter(); lp.solve(A, b_pin.numpy(), c_pin.numpy(), eps=1e-9, max_iter=P, trace_cap=1); ts.append(time.perf_counter() - t0)
print(sys.argv[2], " ".join(f"{1e3*t:.0f}" for t in ts), "mean", round(1e3*np.mean(ts[2:])), "median", round(1e3*np.median(ts[2:])))
