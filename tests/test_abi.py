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


def test_aggregate_check_host_sums(pkg):
    """The host half of bn254v_groth16_batch_all_valid (no device needed): s = sum r_i, t_j = sum r_i x_ij mod r."""
    import ctypes
    from helpers import agg_batch_scalars
    import bn254_oracle as bo
    lib = pkg.load_library()
    rng = np.random.default_rng(5)
    for n, k in ((1, 2), (7, 0), (300, 2), (1000, 3)):
        rnd = rng.integers(0, 256, 16 * n, dtype=np.uint8)
        xs = [[int.from_bytes(rng.bytes(32), "big") % bo.R for _ in range(k)] for _ in range(n)]
        if n > 1 and k:
            xs[1][0] = bo.R - 1
            xs[0][k - 1] = (1 << 256) - 1  # not a field member: the device rejects it, the host sum must still not overflow
            rnd[:32] = 255
        inp = np.frombuffer(b"".join(int(x).to_bytes(32, "big") for row in xs for x in row), dtype=np.uint8)
        out = np.zeros(32 * (1 + k), dtype=np.uint8)
        lib.bn254v_agg_host_sums(rnd.ctypes.data, inp.ctypes.data if k else None, k, n, out.ctypes.data)
        assert out.tobytes() == agg_batch_scalars(rnd.tobytes(), xs), (n, k)
    with pytest.raises(pkg.LibraryError) as e:  # and the entry point itself needs a device
        import torch
        if torch.cuda.is_available():
            raise pkg.LibraryError(pkg.E_NO_DEVICE, "skipped: a GPU is present")
        pkg.Groth16Verifier.batch_all_valid([bytes(256)], bytes(10), [[1, 2]])
    assert e.value.code in (pkg.E_NO_DEVICE, pkg.E_VK_PARSE)


def test_chacha20_expansion_of_the_scalar_seed(pkg):
    """RFC 8439 2.3.2 block vector; the key stream the library expands its getrandom seed with = counter-mode blocks."""
    lib = pkg.load_library()
    key = np.arange(32, dtype=np.uint8)
    nonce = np.frombuffer(bytes.fromhex("000000090000004a00000000"), dtype=np.uint8).copy()
    out = np.zeros(64, dtype=np.uint8)
    lib.bn254v_chacha20_block(key.ctypes.data, 1, nonce.ctypes.data, out.ctypes.data)
    assert out.tobytes().hex() == ("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                                   "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")
    zero = np.zeros(12, dtype=np.uint8)
    for n in (0, 1, 63, 64, 65, 1000):
        got = np.zeros(n + 1, dtype=np.uint8)
        got[n] = 0xA5  # guard byte
        lib.bn254v_chacha20_expand(key.ctypes.data, n, got.ctypes.data)
        want = b""
        for b in range((n + 63) // 64):
            lib.bn254v_chacha20_block(key.ctypes.data, b, zero.ctypes.data, out.ctypes.data)
            want += out.tobytes()
        assert got[:n].tobytes() == want[:n] and got[n] == 0xA5
    a, b = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
    lib.bn254v_chacha20_expand(key.ctypes.data, 256, a.ctypes.data)
    key[0] ^= 1
    lib.bn254v_chacha20_expand(key.ctypes.data, 256, b.ctypes.data)
    assert (a != b).sum() > 200


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "snark-bn254-verifier_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import bn254_oracle" not in txt and "plonk_oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/bn254_oracle.py", "").replace("oracle/`", ""), f
