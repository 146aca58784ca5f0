"""In-tree build of libcoalgpu.so (nvcc, sm_100a only). The .so is git-ignored but ships to the GPU box."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("COALGPU_LIB") or os.path.join(HERE, "libcoalgpu.so")   # COALGPU_LIB: tuning variants only
SOURCES = ["coal_api.cu", "coal_tables.cpp"]
DEPS = SOURCES + ["coal_kernels.cu", "coal_multi.cu", "coal_device.cuh", "coal_math.h", "coal_tables.h", os.path.join("..", "..", "include", "coalgpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if it is missing or older than its sources; returns its path."""
    if force or _stale():
        nvcc = os.environ.get("NVCC", "nvcc")
        extra = os.environ.get("COALGPU_DEFS", "").split()
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
