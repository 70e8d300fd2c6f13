"""Build the sm_100a kernels into the in-tree C-ABI library (nvcc, no torch headers).

    python -m ood_dfq_b200.build            # -> ood_dfq_b200/csrc/liboodfq_b200.so

The library is git-ignored but travels to the GPU box with the snapshot.  nvcc
cross-compiles without a GPU; the CUDA runtime is linked statically so the .so loads
anywhere the driver is present (and, for symbol checks, even where it is not).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "liboodfq_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "oodfq_b200.h")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    tmp = f"{LIB}.tmp.{os.getpid()}"          # build aside, then rename: a concurrent loader never sees half a file
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed building liboodfq_b200.so")
    os.replace(tmp, LIB)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
