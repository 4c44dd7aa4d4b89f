"""In-tree build of the native engine (no JIT cache: the .so travels with the repo).

nvcc cross-compiles for sm_100a without a GPU.  Outputs:
  simplex_method_gpu_b200/libb200lp.so   C-ABI engine (include/b200lp.h)
  bin/solver.out                         CLI with the reference's main() contract
  bin/solver_glpk.out                    solver_glpk.cpp-shaped CPU harness (links GLPK only where <glpk.h> exists)
"""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libb200lp.so")
CLI_PATH = os.path.join(_ROOT, "bin", "solver.out")
GLPK_PATH = os.path.join(_ROOT, "bin", "solver_glpk.out")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine has no prebuilt binary and no CPU fallback")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libb200lp.so and bin/solver.out when sources are newer than the binaries."""
    lib_src = [os.path.join(_CSRC, f) for f in ("engine.cu", "kernels.cuh", "lp_io.cpp")] + \
              [os.path.join(_ROOT, "include", h) for h in ("b200lp.h", "b200lp_io.h")]
    if force or _stale(LIB_PATH, lib_src):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB_PATH, os.path.join(_CSRC, "engine.cu"),
               os.path.join(_CSRC, "lp_io.cpp")]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, capture_output=not verbose)
    cli_src = [os.path.join(_CSRC, "solver_main.cpp"), os.path.join(_ROOT, "include", "b200lp.h"),
               os.path.join(_ROOT, "include", "b200lp_io.h")]
    if os.path.exists(cli_src[0]) and (force or _stale(CLI_PATH, cli_src + [LIB_PATH])):
        os.makedirs(os.path.dirname(CLI_PATH), exist_ok=True)
        cmd = [_nvcc(), "-O2", "-std=c++17", "-o", CLI_PATH, cli_src[0], "-I", os.path.join(_ROOT, "include"),
               "-L", _PKG, "-lb200lp", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../simplex_method_gpu_b200"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, capture_output=not verbose)
    glpk_src = os.path.join(_ROOT, "tools", "solver_glpk_harness.cpp")
    if os.path.exists(glpk_src) and (force or _stale(GLPK_PATH, [glpk_src])):
        os.makedirs(os.path.dirname(GLPK_PATH), exist_ok=True)
        cxx = shutil.which("g++") or "g++"
        have_glpk = any(os.path.exists(os.path.join(d, "glpk.h")) for d in ("/usr/include", "/usr/local/include"))
        cmd = [cxx, "-O2", "-std=c++17", glpk_src, "-o", GLPK_PATH] + (["-lglpk"] if have_glpk else [])
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, capture_output=not verbose)
    return LIB_PATH
