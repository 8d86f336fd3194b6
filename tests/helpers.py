"""Shared test helpers (oracle-side data builders)."""
import ctypes
import json
import os
import subprocess

import numpy as np

import bn254_oracle as bo
from conftest import GOLDEN, ROOT


def g2_point_outside_subgroup(seed=5):
    """A point of E'(Fq2) that is not in the r-torsion (the cofactor is ~2^254, so any random point works)."""
    x = (seed, 1)
    while True:
        rhs = bo.fp2_add(bo.fp2_mul(bo.fp2_sqr(x), x), bo.B2)
        y = bo.fp2_sqrt(rhs)
        if y is not None:
            pt = (x, y)
            assert bo.g2_is_on_curve(pt) and not bo.g2_in_subgroup(pt)
            return pt
        x = (x[0] + 1, 1)


def groth16_malformed_suite(td, index=0):
    """(name, proof bytes, inputs, expected status name) covering every PANIC_/ERR_ class of the Groth16 path."""
    pb, xs, _ = td.proof(index, corrupt=False)
    out = [("valid", pb, xs, "OK_TRUE")]
    b = bytearray(pb); b[0:32] = (bo.P + 1).to_bytes(32, "big"); out.append(("A.x>=p", bytes(b), xs, "PANIC_FIELD_NOT_MEMBER"))
    b = bytearray(pb); b[96:128] = bo.P.to_bytes(32, "big"); out.append(("B.x0==p", bytes(b), xs, "PANIC_FIELD_NOT_MEMBER"))
    b = bytearray(pb); b[63] ^= 1; out.append(("A off curve", bytes(b), xs, "PANIC_NOT_ON_CURVE"))
    b = bytearray(pb); b[255] ^= 2; out.append(("C off curve", bytes(b), xs, "PANIC_NOT_ON_CURVE"))
    b = bytearray(pb); b[191] ^= 1; out.append(("B off curve", bytes(b), xs, "PANIC_NOT_ON_CURVE"))
    b = bytearray(pb); b[64:192] = bo.g2_to_bytes(g2_point_outside_subgroup()); out.append(("B not in G2", bytes(b), xs, "PANIC_NOT_IN_SUBGROUP"))
    out.append(("short", pb[:255], xs, "PANIC_SHORT_BUFFER"))
    out.append(("x0 == 0", pb, [0, xs[1]], "PANIC_IDENTITY"))
    # x >= r cannot be expressed as a bn::Fr: the reference's caller fails in Fr::from_slice before verify().
    # The byte-level ABI reports it per proof; the oracle (which takes ints mod r) has no such case.
    out.append(("x1 >= r (ABI only)", pb, [xs[0], bo.R], "PANIC_FIELD_NOT_MEMBER"))
    out.append(("gnark tail", pb + bytes(68), xs, "OK_TRUE"))  # 324-byte gnark proof: trailing bytes ignored
    return out


def oracle_groth16_status(pb, vk, xs):
    try:
        return "OK_TRUE" if bo.groth16_verifier_verify(pb, vk, xs) else "OK_FALSE"
    except bo.Groth16Error:
        return "ERR_PREPARE_INPUTS"
    except bo.PanicError as e:
        return "PANIC_" + e.kind


def load_json(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def pt_bytes(pt_hex):
    return int(pt_hex[0], 16).to_bytes(32, "big") + int(pt_hex[1], 16).to_bytes(32, "big")


def build_hostsim():
    """g++ build of the kernels' __host__ __device__ code (test-only helper library)."""
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    so = os.path.join(ROOT, "tests", "hostsim", "_hostsim.so")
    csrc = os.path.join(ROOT, "snark-bn254-verifier_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)
