"""GPU: the opt-in aggregate Groth16 check (bn254v_groth16_batch_all_valid) against the per-proof CUDA path and the
oracle.  The answer must be True exactly for batches whose per-proof statuses are all OK_TRUE."""
import numpy as np
import pytest

import bn254_oracle as bo
from helpers import groth16_malformed_suite

pytestmark = pytest.mark.gpu


def _valid_part(gpu, seed, n_total, n_public=2):
    vk, proofs, inputs, expected = gpu.groth16_synth(seed, n_total, n_public=n_public)
    keep = expected == gpu.OK_TRUE
    return vk, np.ascontiguousarray(proofs[keep]), np.ascontiguousarray(inputs[keep]), proofs, inputs, expected


def test_all_valid_batches_pass_and_any_invalid_member_fails(gpu):
    vk, vp, vi, proofs, inputs, expected = _valid_part(gpu, 21, 4000)
    n = vp.shape[0]
    assert n == 2000
    ver = gpu.Groth16Verifier
    ok, st = ver.batch_all_valid(vp, vk, vi, want_status=True)  # scalars drawn by the library
    assert ok is True and (st == gpu.OK_TRUE).all()
    rng = np.random.default_rng(3)
    rnd = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    assert ver.batch_all_valid(vp, vk, vi, rnd16=rnd) is True
    rnd[:] = 0  # degenerate caller scalars: a = 1, b = 0 for every proof (r_i = 1) -- still a correct "yes"
    assert ver.batch_all_valid(vp, vk, vi, rnd16=rnd) is True
    # the 50 % corrupted batch: every record is well-formed (status OK_TRUE = "in the aggregate"), the answer is no
    ok, st = ver.batch_all_valid(proofs, vk, inputs, want_status=True)
    assert ok is False and (st == gpu.OK_TRUE).all()
    # one invalid member of each corruption class, at several positions
    bad_idx = np.flatnonzero(expected != gpu.OK_TRUE)
    for c, pos in zip(range(5), (0, 1, 777, n - 1, n)):
        j = [i for i in bad_idx if (i >> 1) % 5 == c][0]
        p2 = np.insert(vp, pos, proofs[j], axis=0)
        i2 = np.insert(vi, pos, inputs[j], axis=0)
        assert ver.batch_all_valid(p2, vk, i2) is False, (c, pos)
        assert ver.verify_batch(p2, vk, i2)[pos] == gpu.OK_FALSE
    # small batches
    assert ver.batch_all_valid(vp[:1], vk, vi[:1]) is True
    assert ver.batch_all_valid(proofs[bad_idx[:1]], vk, inputs[bad_idx[:1]]) is False
    assert ver.batch_all_valid(vp[:9], vk, vi[:9]) is True
    assert ver.batch_all_valid(np.zeros((0, 256), np.uint8), vk, np.zeros((0, 2, 32), np.uint8)) is True


def test_errors_that_cancel_without_the_scalars(gpu):
    """C_0 + D and C_1 - D: the product of the two pairing equations still holds, each one alone does not."""
    td = bo.Groth16Trapdoor(7, 2, 0)
    vk = td.vk_bytes()
    recs = [td.proof(i, corrupt=False) for i in range(4)]
    D = bo.g1_mul(bo.G1_GEN, 987654321)
    p0, p1 = recs[0][0], recs[1][0]
    c0 = bo.g1_add(bo.uncompressed_bytes_to_g1_point(p0[192:256]), D)
    c1 = bo.g1_add(bo.uncompressed_bytes_to_g1_point(p1[192:256]), bo.g1_neg(D))
    proofs = [p0[:192] + bo.g1_to_bytes(c0), p1[:192] + bo.g1_to_bytes(c1), recs[2][0], recs[3][0]]
    inputs = [r[1] for r in recs]
    ver = gpu.Groth16Verifier
    assert list(ver.verify_batch(proofs, vk, inputs)) == [gpu.OK_FALSE, gpu.OK_FALSE, gpu.OK_TRUE, gpu.OK_TRUE]
    for _ in range(3):
        assert ver.batch_all_valid(proofs, vk, inputs) is False
    # with r_0 = r_1 (caller-chosen scalars: what the header warns against) the errors cancel: the check is only as good
    # as the scalars are unpredictable
    assert ver.batch_all_valid(proofs, vk, inputs, rnd16=np.zeros((4, 16), np.uint8)) is True
    assert ver.batch_all_valid([r[0] for r in recs], vk, inputs) is True


def test_malformed_records_keep_their_per_proof_status(gpu):
    td = bo.Groth16Trapdoor(7, 2, 0)
    vk = td.vk_bytes()
    suite = [c for c in groth16_malformed_suite(td) if len(c[2]) == 2]
    ver = gpu.Groth16Verifier
    want = ver.verify_batch([c[1] for c in suite], vk, [c[2] for c in suite])
    ok, st = ver.batch_all_valid([c[1] for c in suite], vk, [c[2] for c in suite], want_status=True)
    assert ok is False
    assert list(st) == list(want), [c[0] for c, a, b in zip(suite, st, want) if a != b]
    good = [c for c in suite if c[3] == "OK_TRUE"]
    assert len(good) == 2 and ver.batch_all_valid([c[1] for c in good], vk, [c[2] for c in good]) is True
    # a batch in which one record is malformed and the rest valid
    recs = [td.proof(i, corrupt=False) for i in range(40)]
    for name, pb, xs, w in suite:
        if w == "OK_TRUE":
            continue
        proofs = [r[0] for r in recs[:17]] + [pb] + [r[0] for r in recs[17:]]
        inputs = [r[1] for r in recs[:17]] + [xs] + [r[1] for r in recs[17:]]
        ok, st = ver.batch_all_valid(proofs, vk, inputs, want_status=True)
        assert ok is False and gpu.status_name(st[17]) == w and (np.delete(st, 17) == gpu.OK_TRUE).all(), name
    # wrong number of public inputs for the VK
    ok, st = ver.batch_all_valid([r[0] for r in recs[:3]], vk, [r[1][:1] for r in recs[:3]], want_status=True)
    assert ok is False and (st == gpu.ERR_PREPARE_INPUTS).all()


@pytest.mark.parametrize("n", [60000, 240000])
def test_every_launch_shape_of_the_per_proof_half(gpu, n):
    """448-thread blocks (one wave) and 384-thread blocks (several waves); the last block is partly filled."""
    vk, vp, vi, proofs, inputs, expected = _valid_part(gpu, 31, 2 * n)
    assert vp.shape[0] == n
    ver = gpu.Groth16Verifier
    assert ver.batch_all_valid(vp, vk, vi) is True
    bad = np.flatnonzero(expected != gpu.OK_TRUE)[-1]
    vp[n - 1] = proofs[bad]
    vi[n - 1] = inputs[bad]
    assert ver.batch_all_valid(vp, vk, vi) is False


def test_other_numbers_of_public_inputs_and_sign_mode(gpu):
    for n_public, sign_mode in ((1, 0), (4, 1)):
        vk, proofs, inputs, expected = gpu.groth16_synth(5, 600, n_public=n_public, sign_mode=sign_mode)
        ver = type("V", (gpu.Groth16Verifier,), {"sign_mode": sign_mode})
        keep = expected == gpu.OK_TRUE
        vp, vi = np.ascontiguousarray(proofs[keep]), np.ascontiguousarray(inputs[keep])
        assert (ver.verify_batch(vp, vk, vi) == gpu.OK_TRUE).all()
        assert ver.batch_all_valid(vp, vk, vi) is True
        assert ver.batch_all_valid(proofs[:20], vk, inputs[:20]) is False
