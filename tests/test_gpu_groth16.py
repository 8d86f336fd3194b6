"""GPU parity: the CUDA Groth16 path (through the C ABI) against the oracle and the golden vectors.
Bit-exact: verdicts, prepare_inputs output L, canonical Fq12 Miller value and GT value."""
import numpy as np
import pytest

import bn254_oracle as bo
from helpers import groth16_malformed_suite, load_json, oracle_groth16_status, pt_bytes

pytestmark = pytest.mark.gpu


def test_golden_cases_bit_exact(gpu):
    for case in load_json("groth16_golden.json")["cases"]:
        ver = type("V", (gpu.Groth16Verifier,), {"sign_mode": case["sign_mode"]})
        proofs = [bytes.fromhex(p["proof"]) for p in case["proofs"]]
        inputs = [[int(x) for x in p["inputs"]] for p in case["proofs"]]
        status, dbg = ver.verify_batch(proofs, bytes.fromhex(case["vk"]), inputs, debug=True)
        for i, p in enumerate(case["proofs"]):
            assert status[i] == (gpu.OK_TRUE if p["valid"] else gpu.OK_FALSE)
            assert dbg.g1[i, 0].tobytes() == pt_bytes(p["L"])
            assert dbg.miller[i].tobytes().hex() == p["miller"]
            assert dbg.gt[i].tobytes().hex() == p["gt"]


def test_single_verify_semantics(gpu):
    """Groth16Verifier::verify keeps the reference's outcomes: Ok(true) / Ok(false) / Err / panic."""
    td = bo.Groth16Trapdoor(7, 2, 0)
    vk = td.vk_bytes()
    pb, xs, _ = td.proof(0, corrupt=False)
    assert gpu.Groth16Verifier.verify(pb, vk, xs) is True
    assert gpu.Groth16Verifier.verify(pb, vk, [xs[0] + 1, xs[1]]) is False
    with pytest.raises(gpu.Groth16Error):
        gpu.Groth16Verifier.verify(pb, vk, xs[:1])
    with pytest.raises(gpu.VerifierPanic):
        gpu.Groth16Verifier.verify(pb[:100], vk, xs)
    with pytest.raises(gpu.VerifierPanic):
        gpu.Groth16Verifier.verify(pb, vk[:100], xs)


def test_malformed_and_ragged_batch(gpu):
    """Every PANIC_/ERR_ class in one ragged batch (different record lengths), statuses as the oracle's."""
    td = bo.Groth16Trapdoor(7, 2, 0)
    vk = td.vk_bytes()
    suite = [c for c in groth16_malformed_suite(td) if len(c[2]) == 2]
    status = gpu.Groth16Verifier.verify_batch([c[1] for c in suite], vk, [c[2] for c in suite])
    for (name, pb, xs, want), st in zip(suite, status):
        assert gpu.status_name(st) == want, name
        if "ABI only" not in name:
            assert oracle_groth16_status(pb, vk, xs) == want, name


def test_empty_batch(gpu):
    td = bo.Groth16Trapdoor(7, 2, 0)
    st = gpu.Groth16Verifier.verify_batch(np.zeros((0, 256), np.uint8), td.vk_bytes(), np.zeros((0, 2, 32), np.uint8))
    assert st.shape == (0,)


@pytest.mark.parametrize("sign_mode", [0, 1])
def test_synth_generator_matches_oracle_generator(gpu, sign_mode):
    """The on-device workload generator and the oracle's generator emit identical bytes."""
    seed = 99 + sign_mode
    vk, proofs, inputs, expected = gpu.groth16_synth(seed, 12, sign_mode=sign_mode, first_index=1000)
    td = bo.Groth16Trapdoor(seed, 2, sign_mode)
    assert vk == td.vk_bytes()
    for i in range(12):
        pb, xs, valid = td.proof(1000 + i)
        assert proofs[i].tobytes() == pb
        assert [int.from_bytes(inputs[i, j].tobytes(), "big") for j in range(2)] == xs
        assert expected[i] == (gpu.OK_TRUE if valid else gpu.OK_FALSE)


def test_synth_batch_vs_oracle_all_values(gpu):
    """Seeded 48-proof batch: verdict, L, Miller and GT compared with the oracle for every proof."""
    vk, proofs, inputs, expected = gpu.groth16_synth(31337, 48)
    status, dbg = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs, debug=True)
    assert (status == expected).all()
    vkp = bo.load_groth16_verifying_key_from_bytes(vk)
    for i in range(48):
        d = {}
        xs = [int.from_bytes(inputs[i, j].tobytes(), "big") for j in range(2)]
        ok = bo.verify_groth16(vkp, bo.load_groth16_proof_from_bytes(proofs[i].tobytes()), xs, d)
        assert ok == (status[i] == gpu.OK_TRUE)
        assert bo.g1_to_bytes(d["L"]) == dbg.g1[i, 0].tobytes()
        assert bo.fp12_to_bytes(d["miller"]) == dbg.miller[i].tobytes()
        assert bo.fp12_to_bytes(d["gt"]) == dbg.gt[i].tobytes()


def test_full_size_batch_properties(gpu):
    """BASELINE config 2 at full size (2^16, 50 % corrupted): every verdict equals the generator's expected
    verdict, exactly half are rejected, and verdicts are invariant under a permutation of the batch."""
    n = 1 << 16
    vk, proofs, inputs, expected = gpu.groth16_synth(2024, n)
    status = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs)
    assert (status == expected).all()
    assert int((status == gpu.OK_FALSE).sum()) == n // 2
    rng = np.random.default_rng(1)
    perm = rng.permutation(n)[:4096]
    st2 = gpu.Groth16Verifier.verify_batch(proofs[perm], vk, inputs[perm])
    assert (st2 == status[perm]).all()
    # device-resident path gives the same answers
    batch = gpu.Groth16DeviceBatch(vk, proofs[:8192], inputs[:8192])
    st3, ms = batch.verify()
    assert (st3 == status[:8192]).all() and ms > 0
    batch.free()


def test_multi_wave_batch_with_malformed_proofs(gpu):
    """240 000 proofs (the 384-thread / 168-register two-launch shape used from 148 x 384 x 4 proofs up), with malformed
    records scattered through the batch: every verdict equals the generator's, every malformed record gets its own
    status, and rejected threads leaving early do not disturb their blocks' barriers."""
    n = 240000
    vk, proofs, inputs, expected = gpu.groth16_synth(77, n)
    proofs = proofs.copy(); expected = expected.copy()
    rng = np.random.default_rng(3)
    # off-curve A, off-curve C, B.x0 == p, on a few hundred random positions
    pos = rng.choice(n, size=600, replace=False)
    for j, i in enumerate(pos):
        kind = j % 3
        if kind == 0:
            proofs[i, 63] ^= 1; expected[i] = gpu.PANIC_NOT_ON_CURVE
        elif kind == 1:
            proofs[i, 255] ^= 2; expected[i] = gpu.PANIC_NOT_ON_CURVE
        else:
            proofs[i, 96:128] = np.frombuffer(bo.P.to_bytes(32, "big"), dtype=np.uint8)
            expected[i] = gpu.PANIC_FIELD_NOT_MEMBER
    status = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs)
    assert (status == expected).all()


def test_more_public_inputs(gpu):
    """|IC| other than 3 (prepare_inputs loops over n_public)."""
    for n_public in (1, 4):
        vk, proofs, inputs, expected = gpu.groth16_synth(5, 6, n_public=n_public)
        status, dbg = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs, debug=True)
        assert (status == expected).all()
        vkp = bo.load_groth16_verifying_key_from_bytes(vk)
        xs = [int.from_bytes(inputs[0, j].tobytes(), "big") for j in range(n_public)]
        assert bo.g1_to_bytes(bo.prepare_inputs(vkp, xs)) == dbg.g1[0, 0].tobytes()


def test_full_size_sample_against_cpp_oracle(gpu):
    """2^16 batch: L, Miller and GT of a 2^10 strided sample bit-exact against the C++ restatement of the reference
    (which recomputes everything per call the way the crate does), and its verdicts on the same sample."""
    import os
    import ref_cpu
    n = 1 << 16
    vk, proofs, inputs, expected = gpu.groth16_synth(4242, n)
    idx = np.arange(0, n, n >> 10)
    status, dbg = gpu.Groth16Verifier.verify_batch(proofs[idx], vk, inputs[idx], debug=True)
    _, st_c, l_c, ml_c, gt_c = ref_cpu.groth16_verify_batch(vk, proofs[idx], inputs[idx], threads=os.cpu_count() or 1,
                                                            debug=True)
    assert (status == st_c).all() and (status == expected[idx]).all()
    assert (dbg.g1[:, 0] == l_c).all() and (dbg.miller == ml_c).all() and (dbg.gt == gt_c).all()


def test_g2_membership_hardening(gpu):
    """The end-point subgroup test (csrc/pairing_body.inc, ate_endpoint_in_g2) on the device against the reference's
    predicate [r]Q == O (verifier/src/converter.rs:135-153 -> AffineG2::new): >= 4096 points of E'(Fq2) as the proof's
    B -- random points outside G2, points of each small prime order dividing the cofactor, products of those, sums of
    an order-r point with a small-order point -- next to points of G2.  Every status equals the C++ oracle's and, on a
    sample and on all small-order classes, the Python oracle's."""
    import os
    import ref_cpu
    from helpers import TWIST_COFACTOR_SMALL_PRIMES, random_twist_point, twist_point_of_order
    rng = np.random.default_rng(20260118)
    td = bo.Groth16Trapdoor(31, 2, 0)
    vk = td.vk_bytes()
    pb, xs, _ = td.proof(0, corrupt=False)
    points, klass = [], []
    for _ in range(3840):
        points.append(random_twist_point(rng)), klass.append("random")
    small = {ell: [twist_point_of_order(ell, rng) for _ in range(24)] for ell in TWIST_COFACTOR_SMALL_PRIMES}
    for ell, pts in small.items():
        for p in pts:
            points.append(p), klass.append("order %d" % ell)
    l0, l1, l2 = TWIST_COFACTOR_SMALL_PRIMES
    for i in range(24):  # composite small order
        points.append(bo.g2_add(small[l0][i], small[l1][i])), klass.append("order l0*l1")
        points.append(bo.g2_add(small[l1][i], small[l2][i])), klass.append("order l1*l2")
    in_g2 = [bo.g2_mul(bo.G2_GEN, int(rng.integers(1, 1 << 62))) for _ in range(96)]
    for i in range(72):  # an order-r point plus a small-order point
        ell = TWIST_COFACTOR_SMALL_PRIMES[i % 3]
        points.append(bo.g2_add(in_g2[i], small[ell][i % 24])), klass.append("G2 + order %d" % ell)
    for p in in_g2:
        points.append(p), klass.append("in G2")
    n = len(points)
    assert n >= 4096
    proofs = np.tile(np.frombuffer(pb, np.uint8), (n, 1)).copy()
    for i, p in enumerate(points):
        proofs[i, 64:192] = np.frombuffer(bo.g2_to_bytes(p), np.uint8)
    inputs = np.tile(gpu.fr_to_be(xs)[None], (n, 1, 1))
    status = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs)
    _, st_c = ref_cpu.groth16_verify_batch(vk, proofs, inputs, threads=os.cpu_count() or 1)
    assert (status == st_c).all(), [(klass[i], int(status[i]), int(st_c[i])) for i in np.nonzero(status != st_c)[0][:8]]
    for i, k in enumerate(klass):
        want_in = k == "in G2"
        assert (status[i] != gpu.PANIC_NOT_IN_SUBGROUP) == want_in, (k, int(status[i]))
        if want_in:
            assert status[i] == gpu.OK_FALSE  # a valid proof whose B was replaced
    check = [i for i, k in enumerate(klass) if k != "random"] + list(range(0, 3840, 19))
    for i in check:
        assert bo.g2_in_subgroup(points[i]) == (klass[i] == "in G2"), klass[i]
    # the same points in a batch small enough for the fused small-batch kernel (another launch shape)
    sub = np.array([i for i, k in enumerate(klass) if k != "random"][:96] + list(range(32)))
    assert (gpu.Groth16Verifier.verify_batch(proofs[sub], vk, inputs[sub]) == status[sub]).all()


def test_bundled_groth16_raw_proofs_through_the_decoder(gpu):
    """The reference repo holds 4 Groth16 raw proofs (324 bytes) but not their VK: under a trapdoor VK
    the device decoder must accept A, B, C (on curve, B in G2) and the verdict must be the oracle's Ok(false)."""
    fx = load_json("fixtures.json")
    td = bo.Groth16Trapdoor(3, 2, 0)
    vk = td.vk_bytes()
    proofs, inputs = [], []
    for prog in ("fibonacci", "is-prime", "sha2", "tendermint"):
        f = fx[f"{prog}_groth16"]
        raw = bytes.fromhex(f["raw_proof"])
        assert len(raw) == 324
        proofs.append(raw), inputs.append([int(s) for s in f["inputs"]])
    status = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs)
    for pr, xs, st in zip(proofs, inputs, status):
        assert gpu.status_name(st) == oracle_groth16_status(pr, vk, xs) == "OK_FALSE"


@pytest.mark.parametrize("n", [100, 30000, 60000, 240000])
def test_failed_proofs_inside_warps_every_launch_shape(gpu, n):
    """Malformed and rejected proofs scattered inside warps of well-formed ones, for every launch shape (32-thread
    blocks, 128 x 2, 448 two-launch, 384 two-launch): the pairing kernels contain block-wide barriers, so a failed proof
    must neither hang its block nor disturb its neighbours.  Every status as expected."""
    td = bo.Groth16Trapdoor(2024, 2, 0)
    vk, proofs, inputs, expected = gpu.groth16_synth(2024, n)
    assert vk == td.vk_bytes()
    suite = [c for c in groth16_malformed_suite(td) if len(c[2]) == 2 and len(c[1]) == 256]
    rng = np.random.default_rng(n)
    pos = rng.choice(n, size=min(n // 3, 40 * len(suite)), replace=False)
    proofs, inputs, expected = proofs.copy(), inputs.copy(), expected.copy()
    names = {gpu.status_name(s): s for s in range(0, 24)}
    for j, p in enumerate(pos):
        name, pb, xs, want = suite[j % len(suite)]
        proofs[p] = np.frombuffer(pb, np.uint8)
        inputs[p] = gpu.fr_to_be(xs)
        expected[p] = names[want]
    status = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs)
    bad = np.nonzero(status != expected)[0]
    assert bad.size == 0, [(int(i), gpu.status_name(status[i]), gpu.status_name(expected[i])) for i in bad[:8]]


@pytest.mark.parametrize("stride", [256, 260, 257, 324])
def test_record_alignment_paths(gpu, stride):
    """Field elements are read with 128-bit loads from 16-byte aligned records (stride 256), 32-bit loads from 4-byte
    aligned ones (260, 324 = gnark's raw proof length) and byte loads otherwise (257): same statuses, same values."""
    n = 300
    vk, proofs, inputs, expected = gpu.groth16_synth(99, n)
    wide = np.zeros((n, stride), np.uint8)
    wide[:, :256] = proofs
    wide[:, 256:] = 0xAB  # trailing bytes of a record are ignored
    status, dbg = gpu.Groth16Verifier.verify_batch(wide, vk, inputs, debug=True)
    ref_status, ref_dbg = gpu.Groth16Verifier.verify_batch(proofs, vk, inputs, debug=True)
    assert (status == expected).all() and (ref_status == expected).all()
    assert (dbg.miller == ref_dbg.miller).all() and (dbg.gt == ref_dbg.gt).all() and (dbg.g1 == ref_dbg.g1).all()
