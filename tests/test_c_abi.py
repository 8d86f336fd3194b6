"""The C ABI driven from plain C (gcc): tests/c_abi/abi_driver.c links libbn254v.so directly -- no Python in the call
path.  CPU: it must see BN254V_E_NO_DEVICE from every compute entry point; GPU: it runs its own parity checks."""
import os
import subprocess

import pytest

from conftest import ROOT


def _build(pkg):
    pkg.load_library()
    src = os.path.join(ROOT, "tests", "c_abi", "abi_driver.c")
    exe = os.path.join(ROOT, "tests", "c_abi", "abi_driver")
    libdir = os.path.dirname(pkg.library_path())
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(pkg.library_path())):
        subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                        "-L", libdir, "-l:libbn254v.so", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_c_driver_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = subprocess.run([_build(pkg)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "NO-DEVICE-OK" in res.stdout, res.stdout + res.stderr


@pytest.mark.gpu
def test_c_driver_on_device(pkg):
    res = subprocess.run([_build(pkg)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "C-ABI-OK" in res.stdout, res.stdout + res.stderr
