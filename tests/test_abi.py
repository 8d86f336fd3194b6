"""CPU: the C-ABI shared library builds, loads and exports every symbol include/bn254v.h declares; without a
CUDA device every compute entry point fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _header_functions(name):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bn254v_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    declared = _header_functions("bn254v.h")
    assert len(declared) >= 16
    bench = _header_functions("bn254v_bench.h")
    assert len(bench) >= 13
    for name in declared + bench:
        assert hasattr(lib, name), name
    assert sorted(pkg.EXPORTS) == declared
    assert sorted(pkg.BENCH_EXPORTS) == bench
    # measurement / synthetic-workload helpers are not part of the verifier's header
    assert not any(n in declared for n in ("bn254v_imad_peak", "bn254v_groth16_synth", "bn254v_launch_count"))


def test_stale_check_tracks_inc_files(pkg):
    """An edit to a .inc body (most of the arithmetic) must trigger a rebuild."""
    b = pkg._build
    deps = [os.path.basename(p) for p in b._deps()]
    assert "tower_body.inc" in deps and "pairing_body.inc" in deps and "bn254v_bench.h" in deps
    assert not b.is_stale()
    p = os.path.join(b.CSRC, "tower_body.inc")
    st = os.stat(p)
    try:
        os.utime(p, (st.st_atime, os.path.getmtime(b.LIB) + 10))
        assert b.is_stale()
    finally:
        os.utime(p, (st.st_atime, st.st_mtime))
    assert not b.is_stale()
    old = os.environ.get("BN254V_NVCC_EXTRA")
    try:
        os.environ["BN254V_NVCC_EXTRA"] = "-DBN_SYNC_FINE"
        assert b.is_stale()  # other flags than the library was built with
    finally:
        if old is None:
            del os.environ["BN254V_NVCC_EXTRA"]
        else:
            os.environ["BN254V_NVCC_EXTRA"] = old


def test_status_names(pkg):
    assert pkg.status_name(pkg.OK_TRUE) == "OK_TRUE"
    assert pkg.status_name(pkg.ERR_PAIRING_CHECK_FAILED) == "ERR_PAIRING_CHECK_FAILED"
    assert pkg.status_name(pkg.PANIC_NOT_IN_SUBGROUP) == "PANIC_NOT_IN_SUBGROUP"


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.LibraryError) as e:
        pkg.groth16_synth(1, 4)
    assert e.value.code == pkg.E_NO_DEVICE
    with pytest.raises(pkg.LibraryError) as e:
        pkg.pairing_product_batch(np.zeros((1, 1, 64), np.uint8), np.zeros((1, 1, 128), np.uint8), 1)
    assert e.value.code == pkg.E_NO_DEVICE


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "snark-bn254-verifier_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import bn254_oracle" not in txt and "plonk_oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/bn254_oracle.py", "").replace("oracle/`", ""), f
