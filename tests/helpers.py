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
    # B is parsed before C and before prepare_inputs: its subgroup failure is the one reported
    b = bytearray(pb); b[64:192] = bo.g2_to_bytes(g2_point_outside_subgroup(6)); b[255] ^= 2
    out.append(("B not in G2 + C off curve", bytes(b), xs, "PANIC_NOT_IN_SUBGROUP"))
    b[255] ^= 2
    out.append(("B not in G2 + x0 == 0", bytes(b), [0, xs[1]], "PANIC_NOT_IN_SUBGROUP"))
    b = bytearray(pb); b[63] ^= 1; b[64:192] = bo.g2_to_bytes(g2_point_outside_subgroup(7))
    out.append(("A off curve + B not in G2", bytes(b), xs, "PANIC_NOT_ON_CURVE"))
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
        subprocess.run(["g++", "-O2", "-std=c++17", "-DBN254_COUNT_MULS", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)


PLONK_STATUS = {"OK_TRUE": 0, "ERR_BSB22_MISMATCH": 3, "ERR_INVALID_WITNESS": 4, "ERR_INVERSE_NOT_FOUND": 5,
                "ERR_OPENING_POLY_MISMATCH": 6, "ERR_INVALID_NUMBER_OF_DIGESTS": 7, "ERR_PAIRING_CHECK_FAILED": 8,
                "PANIC_FIELD_NOT_MEMBER": 16, "PANIC_NOT_ON_CURVE": 17, "PANIC_IDENTITY": 19, "PANIC_SHORT_BUFFER": 20,
                "PANIC_DIV_BY_ZERO": 21, "PANIC_INDEX": 22}
# oracle exception kinds -> status names of include/bn254v.h
_PLONK_ERR = {"BSB22_COMMITMENT_MISMATCH": "ERR_BSB22_MISMATCH", "INVALID_WITNESS": "ERR_INVALID_WITNESS",
              "INVERSE_NOT_FOUND": "ERR_INVERSE_NOT_FOUND", "OPENING_POLY_MISMATCH": "ERR_OPENING_POLY_MISMATCH",
              "INVALID_NUMBER_OF_DIGESTS": "ERR_INVALID_NUMBER_OF_DIGESTS", "PAIRING_CHECK_FAILED": "ERR_PAIRING_CHECK_FAILED"}


def oracle_plonk_status(pb, vk, xs, rnd=0x1234567):
    import plonk_oracle as po
    try:
        po.plonk_verifier_verify(pb, vk, xs, rnd=rnd)
        return "OK_TRUE"
    except po.PlonkError as e:
        return _PLONK_ERR[e.kind]
    except bo.PanicError as e:
        return "PANIC_" + e.kind


def plonk_fixture(prog):
    fx = load_json("fixtures.json")[f"{prog}_plonk"]
    return bytes.fromhex(fx["raw_proof"]), [int(s) for s in fx["inputs"]]


def plonk_vk_bytes():
    return open(os.path.join(GOLDEN, "plonk_vk.bin"), "rb").read()


def plonk_structural_suite(prog="fibonacci"):
    """Framing-level edge cases of load_plonk_proof_from_bytes / verify_plonk shape checks:
    (name, proof bytes, inputs)."""
    raw, xs = plonk_fixture(prog)
    out = [("valid", raw, xs)]
    out.append(("short<516", raw[:515], xs))
    out.append(("short claimed", raw[:516 + 40], xs))
    out.append(("short tail", raw[:-1], xs))
    b = bytearray(raw); b[512:516] = (5).to_bytes(4, "big")
    out.append(("5 claimed then garbage", bytes(b), xs))
    out.append(("one public input", raw, xs[:1]))
    out.append(("three public inputs", raw, xs + [7]))
    # no BSB22 commitment: nb = 0 and the 64 commitment bytes removed
    off = 516 + 32 * 7 + 96
    b = bytearray(raw[:off]) + (0).to_bytes(4, "big")
    out.append(("no bsb22 commitment", bytes(b), xs))
    b = bytearray(raw); b[off + 4 + 63] ^= 1
    out.append(("bsb22 off curve", bytes(b), xs))
    b = bytearray(raw); b[516 + 32 * 7 + 64:516 + 32 * 7 + 96] = bo.R.to_bytes(32, "big")
    out.append(("zu == r", bytes(b), xs))
    out.append(("trailing bytes", raw + bytes(40), xs))
    return out


def plonk_shape_variant(nqcp, nb_public, prog="fibonacci"):
    """A PlonK (vk bytes, proof bytes, inputs) of another circuit shape than the one bundled VK (nQcp = 1, nPub = 2),
    derived from the bundled fixture: `nqcp` BSB22 commitments / Qcp points / commitment indexes and `nb_public` public
    inputs.  No prover exists for these shapes, so claimed[0] is SOLVED to equal the constant term of the linearised
    polynomial (it enters no earlier challenge): the proof then passes the OpeningPolyMismatch check, runs both MSM
    rounds and the pairing, and fails there -- every intermediate (challenges, PI, digests, pairing inputs, Fq12
    values) is defined and can be compared with the oracle."""
    import plonk_oracle as po
    raw, xs = plonk_fixture(prog)
    vk = plonk_vk_bytes()
    head, nq0 = vk[:368], int.from_bytes(vk[368:372], "big")
    assert nq0 == 1
    qcp0 = vk[372:404]
    rest = vk[404:404 + 160 + 33788]           # g1, g2[0], g2[1], zero-filled lines
    tail = vk[404 + 160 + 33788:]
    assert int.from_bytes(tail[:8], "big") == 1
    cci0 = int.from_bytes(tail[8:16], "big")
    extra_qcp = [vk[112 + 32 * 3:112 + 32 * 4], vk[112 + 32 * 4:112 + 32 * 5], vk[112 + 32 * 5:112 + 32 * 6]]  # Ql Qr Qm
    qcps = ([qcp0] + extra_qcp)[:nqcp]
    ccis = [cci0 + 3 * i for i in range(nqcp)]
    head = head[:72] + nb_public.to_bytes(8, "big") + head[80:]
    new_vk = head + nqcp.to_bytes(4, "big") + b"".join(qcps) + rest + len(ccis).to_bytes(8, "big") + \
        b"".join(c.to_bytes(8, "big") for c in ccis)
    # proof: 8 points | ncl | claimed | zsH zu | nbsb | bsb..
    pts = raw[:512]
    claimed = [raw[516 + 32 * i:548 + 32 * i] for i in range(7)]
    off = 516 + 32 * 7
    zsh_zu = raw[off:off + 96]
    bsb0 = raw[off + 100:off + 164]
    bsbs = ([bsb0] + [raw[64 * i:64 * i + 64] for i in range(3)])[:nqcp]      # further commitments: L R O (valid points)
    cl = claimed[:6] + [claimed[6]] * nqcp
    inputs = (list(xs) + [12345, 67890, 13579])[:nb_public]

    def build(c0):
        c = [c0] + cl[1:]
        return pts + len(c).to_bytes(4, "big") + b"".join(c) + zsh_zu + len(bsbs).to_bytes(4, "big") + b"".join(bsbs)

    d = {}
    try:
        po.plonk_verifier_verify(build(cl[0]), new_vk, inputs, rnd=99, debug=d)
    except po.PlonkError as e:
        assert e.kind == "OPENING_POLY_MISMATCH", e.kind
    proof = build(d["const_lin"].to_bytes(32, "big"))
    return new_vk, proof, inputs


TWIST_COFACTOR_SMALL_PRIMES = (10069, 5864401, 1875725156269)


def random_twist_point(rng):
    """A random point of E'(Fq2) (almost surely outside G2: the cofactor is ~2^254)."""
    while True:
        x = (int(rng.integers(0, 1 << 62)) * int(rng.integers(1, 1 << 62)) % bo.P, int(rng.integers(0, 1 << 62)) % bo.P)
        y = bo.fp2_sqrt(bo.fp2_add(bo.fp2_mul(bo.fp2_sqr(x), x), bo.B2))
        if y is not None:
            if int(rng.integers(0, 2)):
                y = bo.fp2_neg(y)
            return (x, y)


def twist_point_of_order(ell, rng):
    """A point of E'(Fq2) of exact prime order ell (ell divides the cofactor 2p - r)."""
    n_twist = bo.R * (2 * bo.P - bo.R)
    while True:
        pt = bo.g2_mul_raw(random_twist_point(rng), n_twist // ell)
        if pt is not None:
            assert bo.g2_mul_raw(pt, ell) is None
            return pt


# ---- opt-in aggregate Groth16 check (csrc/groth16_agg.cuh)
GLV_LAMBDA = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd  # (x, y) -> (beta x, y) = [lambda](x, y) on G1


def agg_scalar(rnd16: bytes) -> int:
    """r_i of one proof from its 16 scalar bytes: a (LE u64, made odd) + b (LE u64) * lambda mod r."""
    a = int.from_bytes(rnd16[:8], "little") | 1
    b = int.from_bytes(rnd16[8:16], "little")
    return (a + b * GLV_LAMBDA) % bo.R


def agg_batch_scalars(rnd: bytes, inputs) -> bytes:
    """s = sum r_i and t_j = sum r_i x_ij (mod r) as 32-byte big-endian scalars: what the library's host side computes."""
    rs = [agg_scalar(rnd[16 * i:16 * i + 16]) for i in range(len(inputs))]
    n_in = len(inputs[0]) if inputs else 0
    out = [sum(rs) % bo.R] + [sum(r * int(xs[j]) for r, xs in zip(rs, inputs)) % bo.R for j in range(n_in)]
    return b"".join(v.to_bytes(32, "big") for v in out)
