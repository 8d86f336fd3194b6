"""CPU: the C-ABI shared library builds, loads and exports every symbol include/bn254v.h declares; without a
CUDA device every compute entry point fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "bn254v.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bn254v_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    declared = _header_functions()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(pkg.EXPORTS) == declared


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
