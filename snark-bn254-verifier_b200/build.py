"""Builds libbn254v.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The shared library is the product; nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("BN254V_LIB") or os.path.join(HERE, "libbn254v.so")  # BN254V_LIB: experiment builds
SOURCES = [os.path.join(CSRC, "bn254v.cu")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _deps():
    out = [os.path.join(HERE, "..", "include", "bn254v.h")]
    for name in os.listdir(CSRC):
        if name.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, name))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/bn254v.cu -> libbn254v.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libbn254v.so (there is no CPU fallback)")
    extra = os.environ.get("BN254V_NVCC_EXTRA", "").split()  # experiment builds (e.g. -DBN_NO_FP6_LAZY)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
