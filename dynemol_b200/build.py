"""Build the sm_100a shared library IN-TREE (dynemol_b200/lib/libdynemol_b200.so).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  `python -m dynemol_b200.build [--force] [--verbose]`.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdynemol_b200.so")
SOURCES = ["propagator.cu", "legacy_abi.cu", "team.cu", "xpu_abi.cu"]
import glob
# every header the sources include: csrc/*.cuh plus the public C header
HEADERS = sorted(os.path.basename(f) for f in glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join("..", "..", "include", "dynemol_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_LIB = "/usr/local/cuda/lib64"


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    """out: alternative output path (tuning variants built with DYB_NVCC_DEFS, loaded via DYNEMOL_B200_LIB)."""
    global LIB
    if out:
        LIB = out
        force = True
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    extra = os.environ.get("DYB_NVCC_DEFS", "").split()      # e.g. "-DDYB_TILE_COLS=4 -DDYB_TMA_STAGES=3" (tuning experiments)
    cmd = [NVCC] + extra + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-ccbin", host_cxx, "-Xcompiler", "-fPIC,-O2,-fvisibility=default,-fopenmp", "-shared",
           "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + \
          ["-L" + CUDA_LIB, "-lcublas", "-lcusolver", "-lgomp", "-Xlinker", "-rpath," + CUDA_LIB]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libdynemol_b200.so")
    return LIB


if __name__ == "__main__":
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=outs[0] if outs else None))
