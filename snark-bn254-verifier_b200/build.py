"""Builds libbn254v.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The shared library is the product; nothing here falls back to a CPU implementation.  The kernel translation units
(csrc/k_*.cu) and the host side (csrc/bn254v.cu) are compiled by parallel nvcc processes and linked into one library.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("BN254V_LIB") or os.path.join(HERE, "libbn254v.so")  # BN254V_LIB: experiment builds
OBJ_DIR = os.path.join(HERE, "build", os.path.basename(LIB))
SOURCES = [os.path.join(CSRC, n) for n in ("bn254v.cu", "k_groth16.cu", "k_groth16_agg.cu", "k_plonk.cu", "k_pairing.cu", "k_aux.cu")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
STAMP = LIB + ".flags"  # the flags the library was built with (a flag change must trigger a rebuild too)


def _deps():
    """Every file the library is compiled from: all of csrc/ (the .inc bodies hold most of the arithmetic) and the
    public headers."""
    inc = os.path.join(HERE, "..", "include")
    out = [os.path.join(inc, n) for n in sorted(os.listdir(inc)) if n.endswith(".h")]
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h", ".inc")):
            out.append(os.path.join(CSRC, name))
    return out


def _flag_string() -> str:
    return " ".join(NVCC_FLAGS + os.environ.get("BN254V_NVCC_EXTRA", "").split())


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    if any(os.path.getmtime(p) > t for p in _deps()):
        return True
    try:
        return open(STAMP).read() != _flag_string()
    except OSError:
        return True


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libbn254v.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libbn254v.so (there is no CPU fallback)")
    extra = os.environ.get("BN254V_NVCC_EXTRA", "").split()  # experiment builds (e.g. -DBN_SYNC_FINE)
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (os.path.basename(src), res.stdout, res.stderr))
        return obj, res.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, err in results:
            print(err)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] +
                         [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(_flag_string())
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
