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
