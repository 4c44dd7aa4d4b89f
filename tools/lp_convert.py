#!/usr/bin/env python
"""LP text <-> MPS converter and CPU-solver harness.

Working stand-in for the reference's two GLPK side tools:
  * glpk_interface.cpp:16-104  (MPS -> solver text; unfinished there: no separator between m and n at :83,
                                row senses / objective sense ignored, no slack block appended)
  * solver_glpk.cpp:15-39      (read fixed MPS, glp_simplex, print x[i] and the optimal objective)
GLPK is not installed in this image, so the solve goes through HiGHS (scipy.optimize.linprog,
method "highs-ds" = dual simplex, one thread); the MPS written here is what `solver_glpk.cpp`
reads (`glp_read_mps(GLP_MPS_DECK)`), so the GLPK cross-check runs unchanged wherever GLPK exists.

  lp_convert.py to-mps   in.txt  out.mps     solver text (max c'x, Ax <= b, x >= 0, slack block last) -> fixed MPS
  lp_convert.py from-mps in.mps  out.txt     MPS (N/L/G/E rows, RHS, simple bounds rejected) -> solver text [A_s, I]
  lp_convert.py solve    in.txt|in.mps [--engine]   HiGHS dual simplex (or, with --engine, the B200 engine in fp64),
                                             output in solver_glpk.cpp's format, so the two can be diffed
  lp_convert.py to-bin   in.txt  out.b200lp  solver text -> binary twin (include/b200lp_io.h; parsed by the library)
  lp_convert.py from-bin in.b200lp out.txt   binary -> solver text (shortest round-trip decimals)

MPS has no portable objective-sense record in the fixed format, so the objective row is written NEGATED
(min -c'x) with a comment line saying so; `solve` and `from-mps` undo it (marker `* OBJSENSE MAX (negated)`).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NEG_MARK = "* OBJSENSE MAX (negated): objective row holds -c, the optimum of the LP is -(MPS optimum)"


def read_text(path, dtype=np.float64):
    from simplex_method_gpu_b200.solver import read_lp
    return read_lp(path, dtype=dtype)


def split_slack(A, c):
    """Structural part of [A_s, I] (v4:272-277 assumes the identity block; we check it)."""
    m, n = A.shape
    ns = n - m
    if ns >= 0 and np.array_equal(A[:, ns:], np.eye(m, dtype=A.dtype)) and not np.any(c[ns:]):
        return A[:, :ns], c[:ns]
    return A, c        # no recognisable slack block: every column is structural


def _num(v: float) -> str:
    """Most precise decimal that fits the 12-character number field of fixed MPS (~9-10 significant digits)."""
    s = repr(float(v))
    for prec in range(16, 0, -1):
        if len(s) <= 12:
            break
        s = "%.*g" % (prec, v)
    return s


def write_mps(path, A, b, c, name="B200LP"):
    """Fixed-format MPS: fields at columns 2-3, 5-12, 15-22, 25-36, 40-47, 50-61."""
    As, cs = split_slack(A, c)
    m, ns = As.shape
    with open(path, "w") as f:
        f.write(f"NAME          {name}\n{NEG_MARK}\nROWS\n N  COST\n")
        for i in range(m):
            f.write(f" L  R{i}\n")
        f.write("COLUMNS\n")
        for j in range(ns):
            ent = []
            if cs[j] != 0:
                ent.append(("COST", -cs[j]))
            ent += [(f"R{i}", As[i, j]) for i in np.nonzero(As[:, j])[0]]
            for k in range(0, len(ent), 2):
                pair = ent[k:k + 2]
                line = f"    {'X%d' % j:<8}  {pair[0][0]:<8}  {_num(pair[0][1]):>12}"
                if len(pair) == 2:
                    line += f"   {pair[1][0]:<8}  {_num(pair[1][1]):>12}"
                f.write(line + "\n")
        f.write("RHS\n")
        for i in range(m):
            if b[i] != 0:
                f.write(f"    {'RHS':<8}  {'R%d' % i:<8}  {_num(b[i]):>12}\n")
        f.write("ENDATA\n")


def read_mps(path):
    """Fixed or free MPS with N/L/G/E rows and default bounds (x >= 0).  Returns (A_s, sense, b, c, negated, col names)."""
    rows, sense, obj = {}, [], None
    cols, entries, rhs = {}, [], {}
    negated, section = False, None
    with open(path) as f:
        for raw in f:
            if raw.startswith("*"):
                negated |= raw.strip() == NEG_MARK.strip()
                continue
            if not raw.strip():
                continue
            if not raw[0].isspace():
                section = raw.split()[0].upper()
                if section == "ENDATA":
                    break
                continue
            t = raw.split()
            if section == "ROWS":
                kind, nm = t[0].upper(), t[1]
                if kind == "N":
                    obj = obj or nm
                else:
                    rows[nm] = len(sense)
                    sense.append(kind)
            elif section == "COLUMNS":
                if len(t) >= 3 and t[1] == "'MARKER'":
                    raise ValueError("integer markers are not supported (LP only)")
                cn = t[0]
                cols.setdefault(cn, len(cols))
                for k in range(1, len(t) - 1, 2):
                    entries.append((t[k], cols[cn], float(t[k + 1])))
            elif section == "RHS":
                start = 1 if len(t) % 2 == 1 else 0
                for k in range(start, len(t) - 1, 2):
                    rhs[t[k]] = float(t[k + 1])
            elif section in ("BOUNDS", "RANGES"):
                raise ValueError(f"{section} section is not supported: the solver text format is Ax <= b, x >= 0 only")
    m, ns = len(sense), len(cols)
    A = np.zeros((m, ns), order="F")
    c = np.zeros(ns)
    for rn, j, v in entries:
        if rn == obj:
            c[j] = v
        else:
            A[rows[rn], j] = v
    b = np.zeros(m)
    for rn, v in rhs.items():
        if rn in rows:
            b[rows[rn]] = v
    return A, sense, b, c, negated, list(cols)


def mps_to_standard(A, sense, b, c, negated):
    """max c'x, Ax <= b, x >= 0 with the slack identity appended: G rows are negated, E rows split in two."""
    cmax = -c if negated else c
    blocks, rhs = [], []
    for i, s in enumerate(sense):
        if s in ("L", "E"):
            blocks.append(A[i]); rhs.append(b[i])
        if s in ("G", "E"):
            blocks.append(-A[i]); rhs.append(-b[i])
    As = np.asfortranarray(np.vstack(blocks)) if blocks else np.zeros((0, A.shape[1]), order="F")
    m = As.shape[0]
    full = np.asfortranarray(np.hstack([As, np.eye(m)]))
    return full, np.asarray(rhs, dtype=np.float64), np.concatenate([cmax, np.zeros(m)])


def solve_highs(A, b, c):
    """HiGHS dual simplex on max c'x, Ax <= b, x >= 0 (structural columns only).  Returns (status, z, x)."""
    from scipy.optimize import linprog
    As, cs = split_slack(A, c)
    r = linprog(-cs, A_ub=As, b_ub=b, bounds=(0, None), method="highs-ds")
    return r.status, (-r.fun if r.status == 0 else float("nan")), (r.x if r.x is not None else np.zeros(As.shape[1]))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    for nm in ("to-mps", "from-mps", "to-bin", "from-bin"):
        sp = sub.add_parser(nm)
        sp.add_argument("src")
        sp.add_argument("dst")
    sp = sub.add_parser("solve")
    sp.add_argument("src")
    sp.add_argument("--engine", action="store_true", help="solve with libb200lp.so (needs a B200) instead of HiGHS")
    a = ap.parse_args(argv)

    if a.cmd in ("to-bin", "from-bin"):
        from simplex_method_gpu_b200.solver import read_lp_native, write_lp_native
        A, b, c = read_lp_native(a.src, dtype=np.float64)          # reads either form (detected by the magic)
        write_lp_native(a.dst, A, b, c, binary=a.cmd == "to-bin")
    elif a.cmd == "to-mps":
        A, b, c = read_text(a.src)
        write_mps(a.dst, A, b, c, name=os.path.splitext(os.path.basename(a.src))[0].upper()[:8] or "B200LP")
    elif a.cmd == "from-mps":
        from simplex_method_gpu_b200.solver import write_lp
        A, b, c = mps_to_standard(*read_mps(a.src)[:5])
        if np.any(b < 0):
            print("warning: some right-hand sides are negative; the slack basis is not feasible (no phase 1 in the solver)",
                  file=sys.stderr)
        write_lp(a.dst, A, b, c)
    else:
        if a.src.lower().endswith(".mps"):
            A, b, c = mps_to_standard(*read_mps(a.src)[:5])
        else:
            A, b, c = read_text(a.src)
        if a.engine:
            import simplex_method_gpu_b200 as lp
            try:
                sol = lp.solve(A, b, c, eps=1e-9, max_iter=1 << 40, dtype=np.float64, trace_cap=1)
            except lp.capi.B200LPError as exc:
                print(exc, file=sys.stderr)
                return 2
            ns = split_slack(A, c)[0].shape[1]
            status = {lp.SolveStatus.OptimumFound: 0, lp.SolveStatus.Unbounded: 3}.get(sol.status, 1)
            z, x = sol.z, sol.x(A.shape[1])[:ns]
        else:
            status, z, x = solve_highs(A, b, c)
        if status != 0:
            print({2: "Problem has no feasible solution", 3: "Problem unbounded"}.get(status, f"HiGHS status {status}"))
            return 1
        for i, v in enumerate(x):                      # solver_glpk.cpp:30-36 prints 1-based x[i]
            print(f"x[{i + 1}] = {v:g}")
        print(f"Optimal objective: {z:g}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
