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


def test_dedicated_squaring_both_fields(hs):
    """fe_sqr_short (36 + 72 multiply-adds: off-diagonal products once, doubled, plus the squares) equals a * a / 2^256
    mod m for random values, for every value below 2 m that the callers may pass (operand sums), and for limb patterns
    that drive every carry of the accumulation (all-ones limbs, single limbs, alternating)."""
    import random
    rng = random.Random(7)
    for which, mod in ((0, bo.P), (1, bo.R)):
        rinv = pow(1 << 256, -1, mod)
        vals = [rng.randrange(mod) for _ in range(300)] + [rng.randrange(mod, 2 * mod) for _ in range(100)]
        vals += [0, 1, mod - 1, mod, mod + 1, 2 * mod - 1, (1 << 255) - 1, (1 << 254) - 1]
        vals += [0xFFFFFFFF << (32 * i) for i in range(8) if (0xFFFFFFFF << (32 * i)) < 2 * mod]
        vals += [v for v in (int("ffffffff00000000" * 4, 16) >> 2, int("00000000ffffffff" * 4, 16), int("f" * 63, 16) >> 1)
                 if v < 2 * mod]
        for a in (v for v in vals if v < 2 * mod):
            out = (ctypes.c_uint32 * 8)()
            hs.hs_fe_sqr_short(which, out, _limbs(a))
            assert _val(out) == a * a * rinv % mod, hex(a)


def test_wide_square_is_exact_for_every_limb_pattern(hs):
    """fe_sqr_wide is a plain integer square: a * a over 16 words for ANY 256-bit a -- all-ones (every product and every
    carry at its maximum), single limbs, alternating limbs, random values."""
    import random
    rng = random.Random(3)
    vals = [(1 << 256) - 1, 0, 1, 1 << 255, (1 << 255) - 1, int("ffffffff00000000" * 4, 16), int("00000000ffffffff" * 4, 16)]
    vals += [0xFFFFFFFF << (32 * i) for i in range(8)] + [((1 << 256) - 1) ^ (0xFFFFFFFF << (32 * i)) for i in range(8)]
    vals += [rng.getrandbits(256) for _ in range(300)]
    for a in vals:
        out = (ctypes.c_uint32 * 16)()
        hs.hs_fe_sqr_wide(out, _limbs(a))
        assert sum(int(out[i]) << (32 * i) for i in range(16)) == a * a, hex(a)


def test_divsteps_inversion_equals_fermat_and_python(hs):
    """fe_inv (Bernstein-Yang divsteps, 20 x 30) against the Fermat chain and against Python's pow, both fields: random
    values, small and near-modulus values, sparse limbs, powers of two, zero."""
    rng = np.random.default_rng(2)
    for which, mod in ((0, bo.P), (1, bo.R)):
        R = (1 << 256) % mod
        vals = [0, 1, 2, 3, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 1 << 30, (1 << 30) - 1, 1 << 60, 1 << 240,
                (1 << 253) % mod, (1 << 255) % mod, mod >> 1, 0x3fffffff << 30, int("55" * 31, 16) % mod, int("aa" * 31, 16) % mod]
        vals += [pow(2, k, mod) for k in range(0, 256, 17)] + [mod - pow(2, k, mod) for k in range(1, 256, 23)]
        vals += [int.from_bytes(rng.bytes(32), "big") % mod for _ in range(400)]
        vals += [int.from_bytes(rng.bytes(4), "big") for _ in range(20)]
        vals += [(int.from_bytes(rng.bytes(32), "big") & ~((1 << 120) - 1 << 60)) % mod for _ in range(20)]  # a hole of zero bits
        for a in vals:
            am = a * R % mod  # Montgomery form
            o1, o2 = (ctypes.c_uint32 * 8)(), (ctypes.c_uint32 * 8)()
            hs.hs_fe_inv(which, _limbs(am), o1, o2)
            want = pow(a, -1, mod) * R % mod if a else 0
            assert _val(o1) == want, (which, hex(a))
            assert _val(o2) == want, (which, hex(a))


def test_fp2_mul_sqr_lazy_reduction_edge_cases(hs):
    """The lazy-reduction Fp2 multiplier (3 wide products + 2 wide reductions) on random and extreme operands."""
    import random
    rng = random.Random(7)
    Rm = 1 << 256
    rinv = pow(Rm, -1, bo.P)
    edge = [0, 1, 2, bo.P - 1, bo.P - 2, (bo.P - 1) // 2, (1 << 253), Rm % bo.P, (Rm * Rm) % bo.P, 0xFFFFFFFF, (1 << 224) - 1]

    def words(c0, c1):
        return (ctypes.c_uint32 * 16)(*([(c0 >> (32 * i)) & 0xFFFFFFFF for i in range(8)] +
                                        [(c1 >> (32 * i)) & 0xFFFFFFFF for i in range(8)]))

    def val(a):
        return sum(int(a[i]) << (32 * i) for i in range(8)), sum(int(a[8 + i]) << (32 * i) for i in range(8))

    cases = [(a0, a1, b0, b1) for a0 in edge[:6] for a1 in edge[:6] for b0 in (0, bo.P - 1, 5) for b1 in (0, bo.P - 1, 7)]
    cases += [tuple(rng.choice(edge) for _ in range(4)) for _ in range(300)]
    cases += [tuple(rng.randrange(bo.P) for _ in range(4)) for _ in range(1500)]
    out = (ctypes.c_uint32 * 16)()
    for a0, a1, b0, b1 in cases:
        hs.hs_fp2_mul(out, words(a0, a1), words(b0, b1))
        # Montgomery: inputs are xR, result is (x y) R  => result = a b R^-1
        c0 = (a0 * b0 - a1 * b1) * rinv % bo.P
        c1 = (a0 * b1 + a1 * b0) * rinv % bo.P
        assert val(out) == (c0, c1), (a0, a1, b0, b1)
        hs.hs_fp2_sqr(out, words(a0, a1))
        assert val(out) == ((a0 * a0 - a1 * a1) * rinv % bo.P, 2 * a0 * a1 * rinv % bo.P)
        # (9 + u)(a0 + a1 u) through the multiply-by-9 / estimated-quotient path (linear: representation-agnostic)
        hs.hs_fp2_mul_xi(out, words(a0, a1))
        assert val(out) == ((9 * a0 - a1) % bo.P, (9 * a1 + a0) % bo.P), (a0, a1)
        h = (ctypes.c_uint32 * 8)()
        hs.hs_fp_halve(h, (ctypes.c_uint32 * 8)(*[(a0 >> (32 * i)) & 0xFFFFFFFF for i in range(8)]))
        assert sum(int(h[i]) << (32 * i) for i in range(8)) == a0 * pow(2, -1, bo.P) % bo.P


def test_fp6_mul_lazy_reduction_edge_cases(hs):
    """The lazily reduced Fq6 multiplication (6 unreduced Fq2 products, signed 512-bit sums, multiplication by xi with
    a reduction modulo p 2^256, 6 Montgomery reductions) against big-integer schoolbook arithmetic, on random operands
    and on the operands that drive the intermediate sums to the ends of their ranges (all coefficients 0 / p-1)."""
    import itertools
    import random
    rng = random.Random(13)
    rinv = pow(1 << 256, -1, bo.P)
    P = bo.P

    def words(cs):
        return (ctypes.c_uint32 * 48)(*[(c >> (32 * i)) & 0xFFFFFFFF for c in cs for i in range(8)])

    def f2mul(x, y):
        return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)

    def f2add(x, y):
        return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)

    def xi(x):
        return ((9 * x[0] - x[1]) % P, (9 * x[1] + x[0]) % P)

    def ref(a, b):
        A = [(a[0], a[1]), (a[2], a[3]), (a[4], a[5])]
        B = [(b[0], b[1]), (b[2], b[3]), (b[4], b[5])]
        c0 = f2add(f2mul(A[0], B[0]), xi(f2add(f2mul(A[1], B[2]), f2mul(A[2], B[1]))))
        c1 = f2add(f2add(f2mul(A[0], B[1]), f2mul(A[1], B[0])), xi(f2mul(A[2], B[2])))
        c2 = f2add(f2add(f2mul(A[0], B[2]), f2mul(A[1], B[1])), f2mul(A[2], B[0]))
        return [v * rinv % P for c in (c0, c1, c2) for v in c]

    hi, lo = P - 1, 0
    cases = []
    for bits in itertools.product((lo, hi), repeat=12):  # every 0 / p-1 pattern of both operands
        cases.append((list(bits[:6]), list(bits[6:])))
    edge = [0, 1, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 1 << 253, (1 << 256) % P]
    cases += [([rng.choice(edge) for _ in range(6)], [rng.choice(edge) for _ in range(6)]) for _ in range(1500)]
    cases += [([rng.randrange(P) for _ in range(6)], [rng.randrange(P) for _ in range(6)]) for _ in range(3000)]
    out = (ctypes.c_uint32 * 48)()
    for a, b in cases:
        hs.hs_fp6_mul(out, words(a), words(b))
        got = [sum(int(out[8 * k + i]) << (32 * i) for i in range(8)) for k in range(6)]
        assert got == ref(a, b), (a, b)


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
    hs.hs_groth16_vk_new_tables.restype = ctypes.c_void_p
    for ci, case in enumerate(load_json("groth16_golden.json")["cases"]):
        blob, n_ic = _vk_points(case)
        # case 0: prepare_inputs through the fixed-base tables (what the CUDA vk_load builds); case 1: double-and-add
        vk = (hs.hs_groth16_vk_new_tables if ci == 0 else hs.hs_groth16_vk_new)(blob, n_ic)
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
    hs.hs_groth16_vk_new_tables.restype = ctypes.c_void_p
    vk = hs.hs_groth16_vk_new_tables(blob, n_ic)
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


# ------------------------------------------------------------------------------------------------ PlonK
def _hs_plonk(hs):
    hs.hs_plonk_vk_new.restype = ctypes.c_void_p
    hs.hs_plonk_vk_new.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    hs.hs_plonk_vk_free.argtypes = [ctypes.c_void_p]
    hs.hs_plonk_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                   ctypes.c_char_p] + [ctypes.c_char_p] * 4
    from helpers import plonk_vk_bytes
    vkb = plonk_vk_bytes()
    return hs.hs_plonk_vk_new(vkb, len(vkb))


def test_sha256_known_answers(hs):
    import hashlib
    out = ctypes.create_string_buffer(32)
    for m in (b"", b"abc", b"a" * 55, b"a" * 56, b"a" * 64, bytes(range(256)) * 5):
        hs.hs_sha256(m, len(m), out)
        assert out.raw == hashlib.sha256(m).digest()


def test_plonk_fixtures_bit_exact(hs):
    """The 4 bundled proofs: verdict Ok(true) (what the reference's test_programs pins) and every intermediate."""
    from helpers import plonk_fixture
    vk = _hs_plonk(hs)
    gold = load_json("plonk_golden.json")
    for prog, g in gold.items():
        pr, xs = plonk_fixture(prog)
        inputs = b"".join(x.to_bytes(32, "big") for x in xs)
        g1, fr, ml, gt = (ctypes.create_string_buffer(n) for n in (256, 256, 384, 384))
        st = hs.hs_plonk_verify(vk, pr, len(pr), inputs, 2, int(g["rnd"], 16).to_bytes(32, "big"), g1, fr, ml, gt)
        assert st == 0
        for i, nm in enumerate(["gamma", "beta", "alpha", "zeta", "kzg_gamma", "pi", "const_lin"]):
            assert fr.raw[32 * i:32 * i + 32].hex() == g[nm][2:], nm
        assert fr.raw[224:256].hex() == g["hashed_bsb22"][0][2:]
        assert g1.raw[0:64] == pt_bytes(g["lin_digest"]) and g1.raw[64:128] == pt_bytes(g["folded_digest"])
        assert g1.raw[128:192] == pt_bytes(g["pair_g1"][0]) and g1.raw[192:256] == pt_bytes(g["pair_g1"][1])
        assert ml.raw.hex() == g["miller"] and gt.raw.hex() == g["gt"]
    hs.hs_plonk_vk_free(vk)


def test_plonk_fixed_base_tables_give_the_same_points(hs):
    """Same fixture through the VK-constant window tables (what the CUDA vk_load builds): identical digests / GT."""
    from helpers import plonk_fixture
    vk = _hs_plonk(hs)
    hs.hs_plonk_vk_add_tables.argtypes = [ctypes.c_void_p]
    hs.hs_plonk_vk_add_tables(vk)
    g = load_json("plonk_golden.json")["sha2"]
    pr, xs = plonk_fixture("sha2")
    inputs = b"".join(x.to_bytes(32, "big") for x in xs)
    g1, fr, ml, gt = (ctypes.create_string_buffer(n) for n in (256, 256, 384, 384))
    st = hs.hs_plonk_verify(vk, pr, len(pr), inputs, 2, int(g["rnd"], 16).to_bytes(32, "big"), g1, fr, ml, gt)
    assert st == 0
    assert g1.raw[0:64] == pt_bytes(g["lin_digest"]) and g1.raw[64:128] == pt_bytes(g["folded_digest"])
    assert g1.raw[128:192] == pt_bytes(g["pair_g1"][0]) and g1.raw[192:256] == pt_bytes(g["pair_g1"][1])
    assert ml.raw.hex() == g["miller"] and gt.raw.hex() == g["gt"]
    hs.hs_plonk_vk_free(vk)


def test_plonk_mutations_and_structural_cases(hs):
    from helpers import PLONK_STATUS, oracle_plonk_status, plonk_structural_suite, plonk_vk_bytes
    vk = _hs_plonk(hs)
    rnd = (0x1234567).to_bytes(32, "big")
    for m in load_json("plonk_mutations.json"):
        pr = bytes.fromhex(m["raw_proof"])
        inputs = b"".join(int(s).to_bytes(32, "big") for s in m["inputs"])
        st = hs.hs_plonk_verify(vk, pr, len(pr), inputs, 2, rnd, None, None, None, None)
        assert st == PLONK_STATUS[m["status"]], (m["program"], m["mutation"])
    vkb = plonk_vk_bytes()
    for name, pr, xs in plonk_structural_suite():
        want = oracle_plonk_status(pr, vkb, xs)
        inputs = b"".join(int(x).to_bytes(32, "big") for x in xs)
        st = hs.hs_plonk_verify(vk, pr, len(pr), inputs, len(xs), rnd, None, None, None, None)
        assert st == PLONK_STATUS[want], (name, want, st)
    hs.hs_plonk_vk_free(vk)


def test_g2_subgroup_tests_agree(hs):
    """The 63-bit subgroup test, the 127-bit one and the Miller-loop end-point test agree with the oracle's [r]P == 0
    on subgroup points, on random points of E'(Fq2) outside the subgroup, on points of the form (subgroup point +
    cofactor-torsion point) and on points of the small order 10069 (where the step formulas can hit their
    exceptional cases)."""
    from helpers import g2_point_outside_subgroup
    for k in (1, 2, 12345, bo.R - 1):
        pt = bo.g2_mul(bo.G2_GEN, k)
        assert hs.hs_g2_subgroup_both(bo.g2_to_bytes(pt)) == 7
    h2 = 2 * bo.P - bo.R
    assert h2 % 10069 == 0
    for seed in (5, 6, 7):
        small = bo.g2_mul_raw(g2_point_outside_subgroup(seed), bo.R * (h2 // 10069))
        if small is None:
            continue
        assert bo.g2_mul_raw(small, 10069) is None
        for k in (1, 2, 3, 5034, 5035, 10068):
            pt = bo.g2_mul_raw(small, k)
            assert hs.hs_g2_subgroup_both(bo.g2_to_bytes(pt)) == 0
            mixed = bo.g2_add(pt, bo.g2_mul(bo.G2_GEN, 77 + seed))
            assert hs.hs_g2_subgroup_both(bo.g2_to_bytes(mixed)) == 0
    for seed in (5, 6, 7, 100, 1000):
        pt = g2_point_outside_subgroup(seed)
        assert not bo.g2_in_subgroup(pt)
        assert hs.hs_g2_subgroup_both(bo.g2_to_bytes(pt)) == 0
        mixed = bo.g2_add(pt, bo.g2_mul(bo.G2_GEN, seed))
        assert hs.hs_g2_subgroup_both(bo.g2_to_bytes(mixed)) == 0
        # [r]P of an outside point is a non-trivial point killed by the cofactor: still outside
        tors = bo.g2_mul_raw(pt, bo.R)
        if tors is not None:
            assert hs.hs_g2_subgroup_both(bo.g2_to_bytes(tors)) == 0


def test_glv_windowed_scalar_multiplication(hs):
    """The GLV + 4-bit-window scalar multiplication of the PlonK term kernels against the oracle's double-and-add."""
    import random
    rng = random.Random(11)
    pts = [bo.G1_GEN, bo.g1_mul(bo.G1_GEN, 0xABCDEF1234567), bo.g1_mul(bo.G1_GEN, bo.R - 5)]
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    ks = [0, 1, 2, 15, 16, bo.R - 1, bo.R - 2, lam, lam - 1, lam + 1, (lam * lam) % bo.R, bo.R // 2, 1 << 127, (1 << 128) - 1,
          1 << 128, (1 << 253)] + [rng.randrange(bo.R) for _ in range(60)]
    out = ctypes.create_string_buffer(64)
    for pt in pts:
        for k in ks:
            ok = hs.hs_g1_mul_w4(out, bo.g1_to_bytes(pt), k.to_bytes(32, "big"))
            want = bo.g1_mul(pt, k)
            if want is None:
                assert ok == 0
            else:
                assert ok == 1 and out.raw == bo.g1_to_bytes(want), hex(k)


def _plonk_debug_matches_oracle(d, g1, fr, ml, gt):
    """canonical debug bytes (bn254v_debug layout) against the oracle's debug dict"""
    for j, nm in enumerate(["gamma", "beta", "alpha", "zeta", "kzg_gamma", "pi", "const_lin"]):
        assert fr[32 * j:32 * j + 32] == d[nm].to_bytes(32, "big"), nm
    if d["hashed_bsb22"]:
        assert fr[224:256] == d["hashed_bsb22"][0].to_bytes(32, "big")
    assert g1[0:64] == bo.g1_to_bytes(d["lin_digest"]) and g1[64:128] == bo.g1_to_bytes(d["folded_digest"])
    assert g1[128:192] == bo.g1_to_bytes(d["pair_g1"][0]) and g1[192:256] == bo.g1_to_bytes(d["pair_g1"][1])
    assert ml == bo.fp12_to_bytes(d["miller"]) and gt == bo.fp12_to_bytes(d["gt"])


def test_plonk_other_circuit_shapes(hs):
    """VK shapes other than the bundled one (nQcp = 1, nPub = 2): no BSB22 commitment / one public input, two
    commitments / three public inputs, three commitments.  The proofs are derived from a bundled one with claimed[0]
    solved (helpers.plonk_shape_variant), so the whole path runs and every intermediate is compared with the oracle."""
    import plonk_oracle as po
    from helpers import plonk_shape_variant
    _hs_plonk(hs)
    for nq, npub in ((0, 1), (2, 3), (3, 2)):
        vkb, pr, xs = plonk_shape_variant(nq, npub)
        d = {}
        with pytest.raises(po.PlonkError) as e:
            po.plonk_verifier_verify(pr, vkb, xs, rnd=4242, debug=d)
        assert e.value.kind == "PAIRING_CHECK_FAILED"
        vk = hs.hs_plonk_vk_new(vkb, len(vkb))
        assert vk
        inputs = b"".join(x.to_bytes(32, "big") for x in xs)
        g1, fr, ml, gt = (ctypes.create_string_buffer(n) for n in (256, 256, 384, 384))
        st = hs.hs_plonk_verify(vk, pr, len(pr), inputs, npub, (4242).to_bytes(32, "big"), g1, fr, ml, gt)
        assert st == 8, (nq, npub, st)  # ERR_PAIRING_CHECK_FAILED
        _plonk_debug_matches_oracle(d, g1.raw, fr.raw, ml.raw, gt.raw)
        hs.hs_plonk_vk_free(vk)


def test_pairing_product_skips_identity_pairs(hs):
    """bn::pairing_batch skips a pair with an identity member (all-zero bytes here): the Miller and GT values are
    those of the remaining pairs; a set of identities only gives 1."""
    c = [x for x in load_json("pairing_golden.json") if x["k"] == 3][0]
    g1, g2 = bytearray(bytes.fromhex(c["g1"])), bytearray(bytes.fromhex(c["g2"]))
    pts = [(bo.uncompressed_bytes_to_g1_point(bytes(g1[64 * j:64 * j + 64])),
            bo.uncompressed_bytes_to_g2_point(bytes(g2[128 * j:128 * j + 128]))) for j in range(3)]
    for zero_g1, zero_g2 in (({1}, set()), (set(), {0}), ({0}, {2}), ({0, 1, 2}, set()), ({0, 1}, {2})):
        a, b = bytearray(g1), bytearray(g2)
        pairs = []
        for j in range(3):
            if j in zero_g1:
                a[64 * j:64 * j + 64] = bytes(64)
            if j in zero_g2:
                b[128 * j:128 * j + 128] = bytes(128)
            pairs.append((None if j in zero_g1 else pts[j][0], None if j in zero_g2 else pts[j][1]))
        m = bo.miller_product(pairs)
        m = bo.FP12_ONE if m is None else m
        want_gt = bo.final_exponentiation(m)
        ml, gt = ctypes.create_string_buffer(384), ctypes.create_string_buffer(384)
        one = hs.hs_pairing_product(3, bytes(a), bytes(b), ml, gt)
        assert ml.raw == bo.fp12_to_bytes(m) and gt.raw == bo.fp12_to_bytes(want_gt), (zero_g1, zero_g2)
        assert bool(one) == (want_gt == bo.FP12_ONE)


# ------------------------------------------------------------------------------------------------ three lanes per proof
def _fp12_words(rng):
    """a random Fq12 element as 96 Montgomery words (12 coefficients, little-endian limbs)"""
    out = b""
    for _ in range(12):
        v = (int.from_bytes(rng.bytes(40), "big") % bo.P) * (1 << 256) % bo.P
        out += v.to_bytes(32, "little")
    return out


def test_trio_fp12_operations_equal_the_sequential_ones(hs):
    """csrc/trio.cuh (three lanes per proof, sliced at Fq2 granularity), run in lock-step on the host: every Fq12
    operation gives the same limbs as tower_body.inc, also with edge operands (zero / one coefficients)."""
    rng = np.random.default_rng(7)
    one = (1 << 256) % bo.P
    cases = [(_fp12_words(rng), _fp12_words(rng)) for _ in range(12)]
    sparse = bytearray(_fp12_words(rng))
    sparse[64:128] = bytes(64)
    sparse[320:384] = bytes(64)
    cases.append((bytes(sparse), _fp12_words(rng)))
    unit = one.to_bytes(32, "little") + bytes(352)
    cases.append((_fp12_words(rng), unit))
    pm1 = ((bo.P - 1) * (1 << 256) % bo.P).to_bytes(32, "little")
    cases.append((pm1 * 12, pm1 * 12))
    for a, b in cases:
        assert hs.hs_trio_fp12_ops(a, b) == 0


def test_trio_final_exponentiation_and_plonk_miller(hs):
    """The sliced final exponentiation on the golden Miller values, and the sliced two-pair Miller loop of the PlonK
    KZG check (pair table) on the golden pairing inputs: canonical bytes as committed."""
    out = ctypes.create_string_buffer(384)
    for c in load_json("pairing_golden.json"):
        hs.hs_trio_final_exp(bytes.fromhex(c["miller"]), out)
        assert out.raw.hex() == c["gt"]
    vk = _hs_plonk(hs)
    hs.hs_trio_plonk_miller.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    for prog, g in load_json("plonk_golden.json").items():
        pf = pt_bytes(g["pair_g1"][0]) + pt_bytes(g["pair_g1"][1])
        assert hs.hs_trio_plonk_miller(vk, pf, out) == 0
        assert out.raw.hex() == g["miller"]


def test_trio_g2_steps_and_miller_loops(hs):
    """The sliced doubling / addition steps, line products and the two Miller-loop shapes (Groth16: one variable pair
    plus the VK pair table, with the end-point G2 verdict; raw products: k variable pairs with identity masks) give the
    sequential code's limbs -- on points of G2 and on points of the twist outside G2."""
    from helpers import g2_point_outside_subgroup
    g = load_json("pairing_golden.json")
    c = [x for x in g if x["k"] == 3][0]
    g1, g2 = bytes.fromhex(c["g1"]), bytes.fromhex(c["g2"])
    outside = bo.g2_to_bytes(g2_point_outside_subgroup(9))
    for q in (g2[:128], g2[128:256], outside):
        assert hs.hs_trio_g2_steps(q, g1[:64]) == 0
    for x in g:
        k = x["k"]
        for skip in (0, 1, (1 << k) - 1) if k > 1 else (0,):
            assert hs.hs_trio_pairing_miller(k, bytes.fromhex(x["g1"]), bytes.fromhex(x["g2"]), skip) == 0
    case = load_json("groth16_golden.json")["cases"][0]
    blob, n_ic = _vk_points(case)
    hs.hs_groth16_vk_new.restype = ctypes.c_void_p
    vk = hs.hs_groth16_vk_new(blob, n_ic)
    hs.hs_trio_groth16_miller.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p,
                                          ctypes.POINTER(ctypes.c_int)]
    for pr in case["proofs"][:3]:
        raw = bytes.fromhex(pr["proof"])
        lc = pt_bytes(pr["L"]) + raw[192:256]
        ing2 = ctypes.c_int(-1)
        assert hs.hs_trio_groth16_miller(vk, raw[:64], raw[64:192], lc, ctypes.byref(ing2)) == 0
        assert ing2.value == 1
    ing2 = ctypes.c_int(-1)
    assert hs.hs_trio_groth16_miller(vk, raw[:64], outside, lc, ctypes.byref(ing2)) == 0 and ing2.value == 0
    hs.hs_groth16_vk_free(ctypes.c_void_p(vk))


def test_plonk_joint_msm_form_gives_the_same_values(hs):
    """The large-batch form of the PlonK MSM rounds (terms of a sum evaluated jointly with shared doublings, VK terms of
    a sum in one thread): bundled fixtures, other circuit shapes and mutated proofs give the same statuses and the same
    canonical intermediates as the oracle (and hence as the per-term form)."""
    import plonk_oracle as po
    from helpers import PLONK_STATUS, plonk_fixture, plonk_shape_variant, plonk_vk_bytes
    _hs_plonk(hs)
    hs.hs_plonk_verify_joint.argtypes = hs.hs_plonk_verify.argtypes
    vkb = plonk_vk_bytes()
    vk = hs.hs_plonk_vk_new(vkb, len(vkb))
    hs.hs_plonk_vk_add_tables.argtypes = [ctypes.c_void_p]
    hs.hs_plonk_vk_add_tables(vk)  # the VK terms through the fixed-base tables, as on the device
    gold = load_json("plonk_golden.json")
    for prog in ("fibonacci", "tendermint"):
        g = gold[prog]
        pr, xs = plonk_fixture(prog)
        inputs = b"".join(x.to_bytes(32, "big") for x in xs)
        g1, fr, ml, gt = (ctypes.create_string_buffer(n) for n in (256, 256, 384, 384))
        st = hs.hs_plonk_verify_joint(vk, pr, len(pr), inputs, 2, int(g["rnd"], 16).to_bytes(32, "big"), g1, fr, ml, gt)
        assert st == 0
        assert g1.raw[0:64] == pt_bytes(g["lin_digest"]) and g1.raw[64:128] == pt_bytes(g["folded_digest"])
        assert g1.raw[128:192] == pt_bytes(g["pair_g1"][0]) and g1.raw[192:256] == pt_bytes(g["pair_g1"][1])
        assert ml.raw.hex() == g["miller"] and gt.raw.hex() == g["gt"]
    muts = [m for m in load_json("plonk_mutations.json") if m["program"] == "sha2"]
    for m in muts:
        pr = bytes.fromhex(m["raw_proof"])
        inputs = b"".join(int(x).to_bytes(32, "big") for x in m["inputs"])
        st = hs.hs_plonk_verify_joint(vk, pr, len(pr), inputs, 2, (77).to_bytes(32, "big"), None, None, None, None)
        assert st == PLONK_STATUS[m["status"]], m["mutation"]
    hs.hs_plonk_vk_free(vk)
    for nq, npub in ((0, 1), (2, 3), (3, 2)):  # no fixed-base tables here: the variable-base fallback of the VK terms
        vkb2, pr, xs = plonk_shape_variant(nq, npub)
        d = {}
        with pytest.raises(po.PlonkError):
            po.plonk_verifier_verify(pr, vkb2, xs, rnd=4242, debug=d)
        vk2 = hs.hs_plonk_vk_new(vkb2, len(vkb2))
        inputs = b"".join(x.to_bytes(32, "big") for x in xs)
        g1, fr, ml, gt = (ctypes.create_string_buffer(n) for n in (256, 256, 384, 384))
        st = hs.hs_plonk_verify_joint(vk2, pr, len(pr), inputs, npub, (4242).to_bytes(32, "big"), g1, fr, ml, gt)
        assert st == 8
        _plonk_debug_matches_oracle(d, g1.raw, fr.raw, ml.raw, gt.raw)
        hs.hs_plonk_vk_free(vk2)


# ------------------------------------------------------------------------------------------------ aggregate Groth16 check
def test_signed_window_digit_patterns(hs):
    """The signed 4-bit windows (digit = nibble of (k + 0x88..8) - 8) on scalars whose nibbles sit on the edges of the
    digit range -- all 8 (every window carries into the next), all 7 (the largest digit without a carry), all 0xf,
    alternating, a lone top bit -- through the aggregate check's double-scalar routine, whose 64-bit halves are windowed as they come:
    [a + b lambda] P for two points sharing one table inversion, against the oracle."""
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    pats = [0x8888888888888888, 0x7777777777777777, 0xffffffffffffffff, 0x8000000000000000, 0x0000000000000001,
            0x0f0f0f0f0f0f0f0f, 0xf0f0f0f0f0f0f0f1, 0x789abcdef0123457, 0x8888888877777777, 0x1]
    p0, p1 = bo.g1_mul(bo.G1_GEN, 0x1234567), bo.g1_mul(bo.G1_GEN, bo.R - 77)
    pts = bo.g1_to_bytes(p0) + bo.g1_to_bytes(p1)
    out = ctypes.create_string_buffer(128)
    for a in pats:
        for b in pats + [0]:
            k = (a + b * lam) % bo.R
            assert hs.hs_g1_mul_glv64_2(out, pts, a.to_bytes(8, "little"), b.to_bytes(8, "little")) == 1
            assert out.raw[:64] == bo.g1_to_bytes(bo.g1_mul(p0, k)), (hex(a), hex(b))
            assert out.raw[64:] == bo.g1_to_bytes(bo.g1_mul(p1, k)), (hex(a), hex(b))


def test_groth16_aggregate_check(hs):
    """csrc/groth16_agg.cuh on the host: the per-proof Miller values are ML(r_i A_i, B_i) of the oracle, an all-valid batch
    passes (with and without the window tables, for two fold widths), one invalid proof anywhere fails it, malformed
    proofs keep the per-proof statuses."""
    from helpers import agg_scalar, agg_batch_scalars
    case = load_json("groth16_golden.json")["cases"][0]
    td = bo.Groth16Trapdoor(case["seed"], 2, 0)
    blob, n_ic = _vk_points(case)
    hs.hs_groth16_vk_new.restype = ctypes.c_void_p
    hs.hs_groth16_vk_new_tables.restype = ctypes.c_void_p
    hs.hs_groth16_vk_add_agg_tables.argtypes = [ctypes.c_void_p]
    hs.hs_groth16_agg.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p,
                                  ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]
    vk_plain = hs.hs_groth16_vk_new(blob, n_ic)
    vk_tab = hs.hs_groth16_vk_new_tables(blob, n_ic)
    hs.hs_groth16_vk_add_agg_tables(vk_tab)
    rng = np.random.default_rng(11)

    def run(vk, recs, per=8, want_f=False):
        n = len(recs)
        stride = max(len(pb) for pb, _ in recs)
        proofs = b"".join(pb.ljust(stride, b"\0") for pb, _ in recs)
        lens = (ctypes.c_uint32 * n)(*[len(pb) for pb, _ in recs])
        inputs = b"".join(int(x).to_bytes(32, "big") for _, xs in recs for x in xs)
        rnd = rng.integers(0, 256, 16 * n, dtype=np.uint8).tobytes()
        scal = agg_batch_scalars(rnd, [xs for _, xs in recs])
        st = ctypes.create_string_buffer(n)
        fo = ctypes.create_string_buffer(384 * n) if want_f else None
        verdict = hs.hs_groth16_agg(vk, proofs, stride, lens, n, inputs, 2, rnd, scal, per, st, fo)
        return verdict, list(st.raw), rnd, (fo.raw if want_f else None)

    valid = []
    for i in range(5):
        pb, xs, _ = td.proof(i, corrupt=False)
        valid.append((pb, xs))
    verdict, st, rnd, f = run(vk_tab, valid, per=2, want_f=True)
    assert verdict == 1 and st == [0] * 5
    for i in (0, 3):  # per-proof Miller value = the oracle's Miller loop of (r_i A_i, B_i)
        A = bo.uncompressed_bytes_to_g1_point(valid[i][0][:64])
        B = bo.uncompressed_bytes_to_g2_point(valid[i][0][64:192])
        want = bo.miller_product([(bo.g1_mul(A, agg_scalar(rnd[16 * i:16 * i + 16])), B)])
        assert f[384 * i:384 * i + 384] == bo.fp12_to_bytes(want)
    assert run(vk_plain, valid[:3], per=8)[0] == 1  # generic scalar multiplications instead of the tables
    assert run(vk_tab, valid[:1])[0] == 1           # a batch of one
    # one invalid proof (every corruption class keeps the points valid: per-proof status would be OK_FALSE)
    for idx in range(0, 10):
        pb, xs, ok = td.proof(idx, corrupt=True)
        if ok:
            continue
        for pos in (0, 2):
            recs = list(valid[:3])
            recs.insert(pos, (pb, xs))
            verdict, st, _, _ = run(vk_tab, recs, per=2)
            assert verdict == 0 and st == [0] * 4, (idx, pos)
    # two invalid proofs whose errors cancel without the random scalars: C_0 + D and C_1 - D
    (p0, x0), (p1, x1) = valid[0], valid[1]
    D = bo.g1_mul(bo.G1_GEN, 12345)
    c0 = bo.g1_add(bo.uncompressed_bytes_to_g1_point(p0[192:256]), D)
    c1 = bo.g1_add(bo.uncompressed_bytes_to_g1_point(p1[192:256]), bo.g1_neg(D))
    recs = [(p0[:192] + bo.g1_to_bytes(c0), x0), (p1[:192] + bo.g1_to_bytes(c1), x1), valid[2]]
    assert run(vk_tab, recs)[0] == 0
    # malformed proofs: the per-proof statuses of groth16_verify_one, batch verdict 0
    names = {"OK_TRUE": 0, "OK_FALSE": 1, "ERR_PREPARE_INPUTS": 2, "PANIC_FIELD_NOT_MEMBER": 16, "PANIC_NOT_ON_CURVE": 17,
             "PANIC_NOT_IN_SUBGROUP": 18, "PANIC_IDENTITY": 19, "PANIC_SHORT_BUFFER": 20}
    suite = [(name, pb, xs, want) for name, pb, xs, want in groth16_malformed_suite(td)]
    recs = [(pb, xs) for _, pb, xs, _ in suite]
    verdict, st, _, _ = run(vk_tab, recs)
    assert verdict == 0
    assert st == [names[w] for _, _, _, w in suite], [n for (n, _, _, w), s_ in zip(suite, st) if names[w] != s_]
    # ... and the well-formed ones among them alone pass
    good = [(pb, xs) for _, pb, xs, w in suite if w == "OK_TRUE"]
    assert run(vk_tab, good)[0] == 1
