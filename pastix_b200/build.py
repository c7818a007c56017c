"""Builds the in-tree CUDA library pastix_b200/lib/libpastix_b200.so for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "lib", "libpastix_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("engine.cu", "csc_build.cu", "probe.cu")]
HEADERS = [os.path.join(_HERE, "csrc", f) for f in ("scalar.cuh", "symbol.cuh", "kernels_factor.cuh", "kernels_solve.cuh", "kernels_solve_dag.cuh", "kernels_solve_dag2.cuh", "kernels_solve_dag3.cuh", "kernels_small.cuh", "kernels_raff.cuh", "kernels_mma.cuh", "mma.cuh", "kernels_dist.cuh", "dist_plan.h", "csc_build.h")] + \
          [os.path.join(_HERE, "..", "include", "pastix_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--threads", "4"]


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build_cuda(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """`defines`/`out`: tuning variants (e.g. -DPB200_NBMAX_D=64 into lib/libpastix_b200_nb64.so, selected at
    run time with PB200_LIB=<path>); the default build is the product."""
    if out is None and not force and up_to_date():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + SOURCES
    print("[pastix_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None)
