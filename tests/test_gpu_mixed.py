"""GPU parity: mixed Groth16 + PlonK batches through bn254v_verify_many (BASELINE.json configs[4] shape) and the
library's VK cache."""
import os

import numpy as np
import pytest

import bn254_oracle as bo
from helpers import plonk_vk_bytes

pytestmark = pytest.mark.gpu


def test_vk_cache_is_keyed_by_vk_hash(gpu):
    gpu.vk_cache_clear()
    assert gpu.vk_cache_size() == 0
    tds = [bo.Groth16Trapdoor(s, 2, 0) for s in (41, 42)]
    for rep in range(3):
        for td in tds:
            pb, xs, valid = td.proof(1)
            assert gpu.Groth16Verifier.verify(pb, td.vk_bytes(), xs) is valid
    assert gpu.vk_cache_size() == 2
    gnark = type("V", (gpu.Groth16Verifier,), {"sign_mode": 1})
    pb, xs, _ = tds[0].proof(0, corrupt=False)
    assert gnark.verify(pb, tds[0].vk_bytes(), xs) is False  # same bytes, other sign convention: its own handle
    assert gpu.vk_cache_size() == 3
    gpu.PlonkVerifier.verify_batch(np.zeros((1, 904), np.uint8), plonk_vk_bytes(), np.zeros((1, 2, 32), np.uint8))
    assert gpu.vk_cache_size() == 4
    gpu.vk_cache_clear()
    assert gpu.vk_cache_size() == 0
    pb, xs, valid = tds[1].proof(2)
    assert gpu.Groth16Verifier.verify(pb, tds[1].vk_bytes(), xs) is valid  # rebuilt after the clear


def test_unparsable_vk_in_a_mixed_batch(gpu):
    td = bo.Groth16Trapdoor(9, 2, 0)
    pb, xs, valid = td.proof(0, corrupt=False)
    items = [("groth16", pb, td.vk_bytes(), xs), ("groth16", pb, td.vk_bytes()[:100], xs),
             ("plonk", pb, td.vk_bytes(), xs), ("groth16", pb, td.vk_bytes(), [xs[0]])]
    st = gpu.verify_many(items)
    assert [gpu.status_name(s) for s in st] == ["OK_TRUE", "PANIC_VK_PARSE", "PANIC_VK_PARSE", "ERR_PREPARE_INPUTS"]


def test_mixed_batch_config5_parity(gpu):
    """2^16 items: 2^15 trapdoor Groth16 proofs (50 % corrupted) interleaved one-to-one with 2^15 PlonK proofs (bundled
    fixtures replicated, 50 % mutated), through ONE bn254v_verify_many call.  Every status equals the generators'
    expectation, and a strided sample of 512 items equals the C++ oracle's verdict for the same item."""
    import ref_cpu
    import workloads
    half = 1 << 15
    vk_g, pr_g, in_g, exp_g = gpu.groth16_synth(777, half)
    pr_p, in_p, rnd_p, exp_p = workloads.plonk_workload(half, seed=9)
    vk_p = plonk_vk_bytes()
    status = gpu.verify_mixed_arrays(vk_g, pr_g, in_g, vk_p, pr_p, in_p, rnd_p)
    assert (status[0::2] == exp_g).all() and (status[1::2] == exp_p).all()
    idx = np.arange(0, half, half // 256)
    _, st_c = ref_cpu.groth16_verify_batch(vk_g, pr_g[idx], in_g[idx], threads=os.cpu_count() or 1)
    assert (status[0::2][idx] == st_c).all()
    _, st_c = ref_cpu.plonk_verify_batch(vk_p, pr_p[idx], in_p[idx], rnd_p[idx], threads=os.cpu_count() or 1)
    assert (status[1::2][idx] == st_c).all()


def test_mixed_batch_of_several_chunks_per_group(gpu):
    """600 000 items: both groups are cut into several chunks inside bn254v_verify_many (a small first chunk, then 2^20
    Groth16 / 2^18 PlonK proofs), gathered into alternating staging sets while the previous chunk runs; every status
    equals the generators' expectation and comes back at its item's position.  Verified twice over one prebuilt item
    array (MixedItems), the second time into a caller-supplied buffer."""
    import workloads
    half = 300000
    vk_g, pr_g, in_g, exp_g = gpu.groth16_synth(4242, half)
    pr_p, in_p, rnd_p, exp_p = workloads.plonk_workload(half, seed=21)
    items = gpu.MixedItems(vk_g, pr_g, in_g, plonk_vk_bytes(), pr_p, in_p, rnd_p)
    status = items.verify()
    assert (status[0::2] == exp_g).all() and (status[1::2] == exp_p).all()
    out = np.full(2 * half, 255, np.uint8)
    assert items.verify(out=out) is out
    assert (out == status).all()


def test_calls_from_several_host_threads(gpu):
    """The batch entry points share per-device streams and scratch pools: concurrent callers are serialised inside the
    library and every caller gets its own batch's answers."""
    import threading
    import workloads
    vk_p = workloads.plonk_vk_bytes()
    jobs = []
    for t in range(4):
        vk, proofs, inputs, expected = gpu.groth16_synth(100 + t, 3000 + 500 * t)
        jobs.append(("groth16", vk, proofs, inputs, None, expected))
    pp, pi, pr, pe = workloads.plonk_workload(512, seed=9)
    jobs.append(("plonk", vk_p, pp, pi, pr, pe))
    g1, g2, e1 = gpu.pairing_synth(5, 2000, k=2)
    jobs.append(("pairing", None, g1, g2, None, e1))
    results, errors = [None] * len(jobs), []

    def work(j):
        try:
            kind, vk, a, b, rnd, _ = jobs[j]
            for _ in range(3):
                if kind == "groth16":
                    results[j] = gpu.Groth16Verifier.verify_batch(a, vk, b)
                    keep = results[j] == gpu.OK_TRUE
                    assert gpu.Groth16Verifier.batch_all_valid(np.ascontiguousarray(a[keep]), vk, np.ascontiguousarray(b[keep])) is True
                elif kind == "plonk":
                    results[j] = gpu.PlonkVerifier.verify_batch(a, vk, b, rnd=rnd)
                else:
                    results[j] = gpu.pairing_product_batch(a, b, 2)
        except Exception as e:  # surfaced in the main thread
            errors.append((j, repr(e)))

    threads = [threading.Thread(target=work, args=(j,)) for j in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for (kind, *_rest, expected), got in zip(jobs, results):
        assert (np.asarray(got) == expected).all(), kind
