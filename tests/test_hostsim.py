"""CPU: the kernels' per-proof device routines, compiled for the host (tests/hostsim), against the
oracle's golden vectors.  This exercises the exact code the CUDA kernels run (PTX carry chains are
emulated, see field.cuh) -- it is a debugging aid for a GPU-less container, not a product path."""
import ctypes

import numpy as np
import pytest

import bn254_oracle as bo
from helpers import build_hostsim, groth16_malformed_suite, load_json, oracle_groth16_status, pt_bytes


@pytest.fixture(scope="module")
def hs():
    return build_hostsim()


def _limbs(v):
    return (ctypes.c_uint32 * 8)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def _val(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


def test_montgomery_mul_both_fields(hs):
    import random
    rng = random.Random(1)
    for fn, mod in ((hs.hs_fp_mul, bo.P), (hs.hs_fr_mul, bo.R)):
        rinv = pow(1 << 256, -1, mod)
        for _ in range(200):
            a, b = rng.randrange(mod), rng.randrange(mod)
            out = (ctypes.c_uint32 * 8)()
            fn(out, _limbs(a), _limbs(b))
            assert _val(out) == a * b * rinv % mod
        for a, b in ((0, 5), (mod - 1, mod - 1), (1, mod - 1)):
            out = (ctypes.c_uint32 * 8)()
            fn(out, _limbs(a), _limbs(b))
            assert _val(out) == a * b * rinv % mod


def test_pairing_products_match_golden(hs):
    for c in load_json("pairing_golden.json"):
        ml, gt = ctypes.create_string_buffer(384), ctypes.create_string_buffer(384)
        one = hs.hs_pairing_product(c["k"], bytes.fromhex(c["g1"]), bytes.fromhex(c["g2"]), ml, gt)
        assert ml.raw.hex() == c["miller"] and gt.raw.hex() == c["gt"] and bool(one) == c["is_one"]


def _vk_points(case):
    vk = bo.load_groth16_verifying_key_from_bytes(bytes.fromhex(case["vk"]))
    if case["sign_mode"] == 0:
        beta, gamma = vk["beta2"], vk["gamma2"]
    else:
        beta, gamma = bo.g2_neg(vk["beta2"]), bo.g2_neg(vk["gamma2"])
    delta = bo.g2_neg(vk["delta2"])
    blob = bo.g1_to_bytes(vk["alpha"]) + bo.g2_to_bytes(beta) + bo.g2_to_bytes(gamma) + bo.g2_to_bytes(delta)
    blob += b"".join(bo.g1_to_bytes(k) for k in vk["k"])
    return blob, len(vk["k"])


def test_groth16_matches_golden(hs):
    hs.hs_groth16_vk_new.restype = ctypes.c_void_p
    hs.hs_groth16_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                     ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    hs.hs_groth16_vk_target.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    hs.hs_groth16_vk_free.argtypes = [ctypes.c_void_p]
    for case in load_json("groth16_golden.json")["cases"]:
        blob, n_ic = _vk_points(case)
        vk = hs.hs_groth16_vk_new(blob, n_ic)
        tgt = ctypes.create_string_buffer(384)
        hs.hs_groth16_vk_target(vk, tgt)
        assert tgt.raw.hex() == case["alpha_beta"]
        for pr in case["proofs"][:4]:
            L, ml, gt = (ctypes.create_string_buffer(n) for n in (64, 384, 384))
            inputs = b"".join(int(x).to_bytes(32, "big") for x in pr["inputs"])
            st = hs.hs_groth16_verify(vk, bytes.fromhex(pr["proof"]), 256, inputs, 2, L, ml, gt)
            assert st == (0 if pr["valid"] else 1)
            assert L.raw == pt_bytes(pr["L"]) and ml.raw.hex() == pr["miller"] and gt.raw.hex() == pr["gt"]
        hs.hs_groth16_vk_free(vk)


def test_groth16_malformed_classes(hs, pkg):
    case = load_json("groth16_golden.json")["cases"][0]
    td = bo.Groth16Trapdoor(case["seed"], 2, 0)
    blob, n_ic = _vk_points(case)
    hs.hs_groth16_vk_new.restype = ctypes.c_void_p
    hs.hs_groth16_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                     ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    vk = hs.hs_groth16_vk_new(blob, n_ic)
    names = {"OK_TRUE": 0, "OK_FALSE": 1, "ERR_PREPARE_INPUTS": 2, "PANIC_FIELD_NOT_MEMBER": 16, "PANIC_NOT_ON_CURVE": 17,
             "PANIC_NOT_IN_SUBGROUP": 18, "PANIC_IDENTITY": 19, "PANIC_SHORT_BUFFER": 20}
    for name, pb, xs, want in groth16_malformed_suite(td):
        if "ABI only" not in name:
            assert oracle_groth16_status(pb, bytes.fromhex(case["vk"]), xs) == want, name
        inputs = b"".join(int(x).to_bytes(32, "big") for x in xs)
        st = hs.hs_groth16_verify(vk, pb, len(pb), inputs, len(xs), None, None, None)
        assert st == names[want], name
    pb, xs, _ = td.proof(0, corrupt=False)
    st = hs.hs_groth16_verify(vk, pb, 256, int(xs[0]).to_bytes(32, "big"), 1, None, None, None)
    assert st == names["ERR_PREPARE_INPUTS"]
