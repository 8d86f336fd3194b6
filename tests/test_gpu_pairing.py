"""GPU parity: raw k-pair pairing products (bn::pairing_batch) -- canonical Fq12 Miller and GT values."""
import numpy as np
import pytest

import bn254_oracle as bo
from helpers import load_json

pytestmark = pytest.mark.gpu


def test_pairing_golden_bit_exact(gpu):
    for c in load_json("pairing_golden.json"):
        k = c["k"]
        g1 = np.frombuffer(bytes.fromhex(c["g1"]), dtype=np.uint8).reshape(1, k, 64)
        g2 = np.frombuffer(bytes.fromhex(c["g2"]), dtype=np.uint8).reshape(1, k, 128)
        is_one, ml, gt = gpu.pairing_product_batch(g1, g2, k, want_values=True)
        assert ml[0].tobytes().hex() == c["miller"]
        assert gt[0].tobytes().hex() == c["gt"]
        assert bool(is_one[0]) == c["is_one"]


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_pairing_synth_vs_oracle(gpu, k):
    g1, g2, expected = gpu.pairing_synth(4242, 8, k=k)
    is_one, ml, gt = gpu.pairing_product_batch(g1, g2, k, want_values=True)
    assert (is_one == expected).all()
    for i in (0, 1, 5):
        pairs = [(bo.uncompressed_bytes_to_g1_point(g1[i, j].tobytes()),
                  bo.uncompressed_bytes_to_g2_point(g2[i, j].tobytes())) for j in range(k)]
        m = bo.miller_product(pairs)
        assert bo.fp12_to_bytes(m) == ml[i].tobytes()
        assert bo.fp12_to_bytes(bo.final_exponentiation(m)) == gt[i].tobytes()


def test_pairing_batch_large_properties(gpu):
    """2^14 4-pair sets: is_one equals the generator's construction (odd indices solved to 1); the GT value
    of a set is invariant under permuting its pairs (the product is commutative)."""
    n = 1 << 14
    g1, g2, expected = gpu.pairing_synth(11, n, k=4)
    is_one = gpu.pairing_product_batch(g1, g2, 4)
    assert (is_one == expected).all() and int(is_one.sum()) == n // 2
    sub = slice(0, 64)
    _, _, gt_a = gpu.pairing_product_batch(g1[sub], g2[sub], 4, want_values=True)
    _, _, gt_b = gpu.pairing_product_batch(g1[sub][:, ::-1], g2[sub][:, ::-1], 4, want_values=True)
    assert (gt_a == gt_b).all()


def test_full_size_config4_against_cpp_oracle(gpu):
    """BASELINE configs[3] at full size: 2^20 random 4-pair sets, every is_one bit as constructed (odd indices solved
    to 1); canonical Miller and GT values of a 2^12 strided sample bit-exact against the C++ oracle."""
    import os
    import ref_cpu
    n = 1 << 20
    g1, g2, expected = gpu.pairing_synth(2025, n, k=4)
    is_one = gpu.pairing_product_batch(g1, g2, 4)
    assert (is_one == expected).all() and int(is_one.sum()) == n // 2
    idx = np.arange(0, n, n >> 12)
    _, ml, gt = gpu.pairing_product_batch(g1[idx], g2[idx], 4, want_values=True)
    _, one_c, ml_c, gt_c = ref_cpu.pairing_product_batch(g1[idx], g2[idx], 4, threads=os.cpu_count() or 1)
    assert (ml == ml_c).all() and (gt == gt_c).all() and (one_c == expected[idx]).all()


@pytest.mark.parametrize("n,k", [(60000, 2), (240000, 1)])
def test_two_launch_shapes_write_the_same_values(gpu, n, k):
    """Batches of more than one wave run as k_pairing_miller | k_pairing_finish (448 x 128 registers at 60 000 sets,
    384 x 168 at 240 000): every is_one bit as constructed, and the canonical Miller / GT values those kernels write
    are bit-exact against the C++ oracle on a strided sample."""
    import os
    import ref_cpu
    g1, g2, expected = gpu.pairing_synth(77 + k, n, k=k)
    is_one, ml, gt = gpu.pairing_product_batch(g1, g2, k, want_values=True)
    assert (is_one == expected).all()
    idx = np.arange(0, n, n // 96)
    _, one_c, ml_c, gt_c = ref_cpu.pairing_product_batch(g1[idx], g2[idx], k, threads=os.cpu_count() or 1)
    assert (ml[idx] == ml_c).all() and (gt[idx] == gt_c).all() and (one_c == expected[idx]).all()


def test_identity_pairs_are_skipped(gpu):
    """A pair with an identity member (all-zero bytes) is skipped as in substrate-bn's pairing_batch: Miller and GT
    values are those of the remaining pairs (bit-exact against the oracle); a set of identities only gives 1."""
    n, k = 64, 4
    g1, g2, _ = gpu.pairing_synth(99, n, k=k)
    g1, g2 = g1.copy(), g2.copy()
    rng = np.random.default_rng(5)
    masks = []
    for i in range(n):
        z1 = set(np.nonzero(rng.integers(0, 3, k) == 0)[0].tolist()) if i % 4 else set()
        z2 = set(np.nonzero(rng.integers(0, 4, k) == 0)[0].tolist()) if i % 4 else set()
        if i == 5:
            z1, z2 = {0, 1, 2, 3}, set()
        if i == 6:
            z1, z2 = {0, 2}, {1, 3}
        for j in z1:
            g1[i, j] = 0
        for j in z2:
            g2[i, j] = 0
        masks.append((z1, z2))
    is_one, ml, gt = gpu.pairing_product_batch(g1, g2, k, want_values=True)
    for i in list(range(12)) + [20, 33, 47, 63]:
        z1, z2 = masks[i]
        pairs = [(None if j in z1 else bo.uncompressed_bytes_to_g1_point(g1[i, j].tobytes()),
                  None if j in z2 else bo.uncompressed_bytes_to_g2_point(g2[i, j].tobytes())) for j in range(k)]
        m = bo.miller_product(pairs)
        m = bo.FP12_ONE if m is None else m
        e = bo.final_exponentiation(m)
        assert bo.fp12_to_bytes(m) == ml[i].tobytes() and bo.fp12_to_bytes(e) == gt[i].tobytes(), (i, z1, z2)
        assert bool(is_one[i]) == (e == bo.FP12_ONE)
    assert is_one[5] == 1 and is_one[6] == 1
