"""Generates tests/golden/trace_digests.json with the CPU oracle (plain left-to-right sums, order=0):
sha256 of the (p, q) int32 pivot trace of the benchmark LPs over several window lengths, the objective
after each window and the smallest runner-up gap seen (how far the sequence is from a tie).

    python tests/golden/make_trace_digests.py [C2 C3 C4]      # C4 needs ~30 GB of RAM and minutes of CPU

bench.py prints the same digest for the engine (every GPU count) and for the reference's own v4 build;
tests/test_reference_gpu.py compares all three.  The digest of a window of P pivots is
sha256(int32[P][2] little-endian, row = (p, q)).
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

SEED, EPS = 1, 1e-9
CASES = {"C2": (1024, 2048, [64, 192, 1000, 1089]), "C3": (8192, 16384, [64, 192, 1000]),
         "C4": (32768, 65536, [64, 192])}


def digest(tp, tq, k):
    a = np.stack([np.asarray(tp[:k], np.int32), np.asarray(tq[:k], np.int32)], axis=1)
    return hashlib.sha256(np.ascontiguousarray(a).astype("<i4").tobytes()).hexdigest()


def main():
    path = os.path.join(HERE, "trace_digests.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name in (sys.argv[1:] or ["C2", "C3"]):
        m, n, windows = CASES[name]
        A, b, c = oracle.gen_dense(m, n, SEED)
        ent = {"m": m, "n": n, "seed": SEED, "eps": EPS, "windows": {}}
        for P in windows:
            t0 = time.time()
            s = oracle.solve(A, b, c, eps=EPS, max_iter=P, trace_cap=P)
            k = int(min(s.pivots, P))
            ent["windows"][str(P)] = {"pivots": k, "status": int(s.status), "z": float(s.z),
                                      "trace_sha256": digest(s.trace_p, s.trace_q, k),
                                      "min_gap_p": float(s.gap_p[:k].min()), "min_gap_q": float(s.gap_q[:k].min())}
            print(name, P, ent["windows"][str(P)], f"{time.time() - t0:.1f}s", flush=True)
        out[name] = ent
        with open(path, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
