"""GPU parity: the CUDA PlonK path (through the C ABI) against the reference's bundled fixtures, the oracle's
golden intermediates (challenges, PI, linearised digest, folded digest, pairing inputs, Fq12 values) and the
mutation -> status map."""
import numpy as np
import pytest

import bn254_oracle as bo
from helpers import (PLONK_STATUS, load_json, oracle_plonk_status, plonk_fixture, plonk_structural_suite, plonk_vk_bytes,
                     pt_bytes)

pytestmark = pytest.mark.gpu
PROGS = ["fibonacci", "is-prime", "sha2", "tendermint"]


def test_bundled_fixtures_verify_true(gpu):
    """What the reference's own test pins (examples/script/src/main.rs:182-245): Ok(true) for every bundled proof."""
    vk = plonk_vk_bytes()
    for prog in PROGS:
        pr, xs = plonk_fixture(prog)
        assert gpu.PlonkVerifier.verify(pr, vk, xs) is True


def test_fixture_intermediates_bit_exact(gpu):
    vk = plonk_vk_bytes()
    gold = load_json("plonk_golden.json")
    proofs, inputs, rnds = [], [], []
    for prog in PROGS:
        pr, xs = plonk_fixture(prog)
        proofs.append(pr), inputs.append(xs), rnds.append(int(gold[prog]["rnd"], 16))
    status, dbg = gpu.PlonkVerifier.verify_batch(proofs, vk, inputs, rnd=rnds, debug=True)
    assert (status == gpu.OK_TRUE).all()
    for i, prog in enumerate(PROGS):
        g = gold[prog]
        for j, nm in enumerate(["gamma", "beta", "alpha", "zeta", "kzg_gamma", "pi", "const_lin"]):
            assert dbg.fr[i, j].tobytes().hex() == g[nm][2:], (prog, nm)
        assert dbg.fr[i, 7].tobytes().hex() == g["hashed_bsb22"][0][2:]
        assert dbg.g1[i, 0].tobytes() == pt_bytes(g["lin_digest"])
        assert dbg.g1[i, 1].tobytes() == pt_bytes(g["folded_digest"])
        assert dbg.g1[i, 2].tobytes() == pt_bytes(g["pair_g1"][0])
        assert dbg.g1[i, 3].tobytes() == pt_bytes(g["pair_g1"][1])
        assert dbg.miller[i].tobytes().hex() == g["miller"] and dbg.gt[i].tobytes().hex() == g["gt"]


def test_mutation_status_map(gpu):
    """All 96 committed mutated proofs (4 programs x 24 classes) in one batch."""
    vk = plonk_vk_bytes()
    muts = load_json("plonk_mutations.json")
    status = gpu.PlonkVerifier.verify_batch([bytes.fromhex(m["raw_proof"]) for m in muts], vk,
                                            [[int(s) for s in m["inputs"]] for m in muts], rnd=[77] * len(muts))
    for m, st in zip(muts, status):
        assert st == PLONK_STATUS[m["status"]], (m["program"], m["mutation"], gpu.status_name(st))


def test_error_semantics_single_verify(gpu):
    """PlonkVerifier::verify never returns Ok(false): it is True or an error (verifier/src/plonk/verify.rs:316)."""
    vk = plonk_vk_bytes()
    by = {(m["program"], m["mutation"]): m for m in load_json("plonk_mutations.json")}
    m = by[("sha2", "claimed0+1")]
    with pytest.raises(gpu.PlonkError) as e:
        gpu.PlonkVerifier.verify(bytes.fromhex(m["raw_proof"]), vk, [int(s) for s in m["inputs"]])
    assert e.value.kind == "OpeningPolyMismatch"
    m = by[("sha2", "batchedH*2")]
    with pytest.raises(gpu.PlonkError) as e:
        gpu.PlonkVerifier.verify(bytes.fromhex(m["raw_proof"]), vk, [int(s) for s in m["inputs"]])
    assert e.value.kind == "PairingCheckFailed"
    m = by[("sha2", "L-offcurve")]
    with pytest.raises(gpu.VerifierPanic):
        gpu.PlonkVerifier.verify(bytes.fromhex(m["raw_proof"]), vk, [int(s) for s in m["inputs"]])


def test_structural_edge_cases_ragged_batch(gpu):
    vk = plonk_vk_bytes()
    suite = [c for c in plonk_structural_suite() if len(c[2]) == 2]
    status = gpu.PlonkVerifier.verify_batch([c[1] for c in suite], vk, [c[2] for c in suite], rnd=[5] * len(suite))
    for (name, pr, xs), st in zip(suite, status):
        assert gpu.status_name(st).replace("_OUT_OF_RANGE", "") == oracle_plonk_status(pr, vk, xs), name
    # wrong number of public inputs -> Err(InvalidWitness)
    pr, xs = plonk_fixture("fibonacci")
    for bad_inputs in (xs[:1], xs + [7]):
        with pytest.raises(gpu.PlonkError) as e:
            gpu.PlonkVerifier.verify(pr, vk, bad_inputs)
        assert e.value.kind == "InvalidWitness"


def test_verdict_independent_of_rnd_but_points_are_not(gpu):
    vk = plonk_vk_bytes()
    pr, xs = plonk_fixture("tendermint")
    rnds = [1, 2, bo.R - 1, 0xDEADBEEF, 0x1234567, (1 << 256) - 1]
    status, dbg = gpu.PlonkVerifier.verify_batch([pr] * len(rnds), vk, [xs] * len(rnds), rnd=rnds, debug=True)
    assert (status == gpu.OK_TRUE).all()
    assert len({dbg.g1[i, 2].tobytes() for i in range(len(rnds))}) == len(rnds)
    # same values as the oracle for one of them
    import plonk_oracle as po
    d = {}
    po.plonk_verifier_verify(pr, vk, xs, rnd=0x1234567, debug=d)
    assert bo.g1_to_bytes(d["pair_g1"][0]) == dbg.g1[4, 2].tobytes()
    assert bo.fp12_to_bytes(d["gt"]) == dbg.gt[4].tobytes()


def test_full_size_batch_config3(gpu):
    """BASELINE configs[2]: 2^14 proofs = bundled fixtures replicated, 50 % mutated; every status as expected."""
    import workloads
    n = 1 << 14
    proofs, inputs, rnd, expected = workloads.plonk_workload(n, seed=3)
    status = gpu.PlonkVerifier.verify_batch(proofs, plonk_vk_bytes(), inputs, rnd=rnd)
    assert (status == expected).all()
    assert int((status == gpu.OK_TRUE).sum()) == n // 2
    assert {int(s) for s in np.unique(status)} == {0, 6, 8}


def test_big_batch_several_chunks(gpu):
    """600 000 proofs: three workspace chunks (2^18 proofs each at most), survivors listed per chunk; every status as
    expected."""
    import workloads
    n = 600000
    proofs, inputs, rnd, expected = workloads.plonk_workload(n, seed=5)
    status = gpu.PlonkVerifier.verify_batch(proofs, plonk_vk_bytes(), inputs, rnd=rnd)
    assert (status == expected).all()
    assert {int(s) for s in np.unique(status)} == {0, 6, 8}


def test_mixed_batch_over_several_vks(gpu):
    """verify_many: Groth16 proofs under two different VKs (reference and gnark sign conventions are the same wire
    format; here two trapdoor VKs) interleaved with PlonK proofs; statuses return in input order."""
    vk_p = plonk_vk_bytes()
    items, want = [], []
    tds = [bo.Groth16Trapdoor(21, 2, 0), bo.Groth16Trapdoor(22, 2, 0)]
    vks = [td.vk_bytes() for td in tds]
    muts = [m for m in load_json("plonk_mutations.json") if m["program"] == "tendermint"][:6]
    for i in range(6):
        for t, td in enumerate(tds):
            pb, xs, valid = td.proof(i)
            items.append(("groth16", pb, vks[t], xs))
            want.append(0 if valid else 1)
        m = muts[i]
        items.append(("plonk", bytes.fromhex(m["raw_proof"]), vk_p, [int(s) for s in m["inputs"]]))
        want.append(PLONK_STATUS[m["status"]])
    # a proof checked under the other VK is rejected
    pb, xs, _ = tds[0].proof(0, corrupt=False)
    items.append(("groth16", pb, vks[1], xs))
    want.append(1)
    st = gpu.verify_many(items)
    assert st.tolist() == want


def test_other_circuit_shapes(gpu):
    """VK shapes other than the bundled one (nQcp = 1, nPub = 2; BN_MAX_QCP / BN_MAX_PLONK_PUBLIC in plonk.cuh): no
    BSB22 commitment / one public input, two commitments / three inputs, three commitments.  Proofs derived from a
    bundled one with claimed[0] solved (helpers.plonk_shape_variant): the whole path runs (both MSM rounds with the other
    term counts, the KZG transcript over more digests) and every intermediate equals the oracle's."""
    import plonk_oracle as po
    from helpers import plonk_shape_variant
    for nq, npub in ((0, 1), (2, 3), (3, 2)):
        vkb, pr, xs = plonk_shape_variant(nq, npub)
        d = {}
        with pytest.raises(po.PlonkError) as e:
            po.plonk_verifier_verify(pr, vkb, xs, rnd=4242, debug=d)
        assert e.value.kind == "PAIRING_CHECK_FAILED"
        status, dbg = gpu.PlonkVerifier.verify_batch([pr] * 3, vkb, [xs] * 3, rnd=[4242] * 3, debug=True)
        assert (status == gpu.ERR_PAIRING_CHECK_FAILED).all()
        for j, nm in enumerate(["gamma", "beta", "alpha", "zeta", "kzg_gamma", "pi", "const_lin"]):
            assert dbg.fr[1, j].tobytes() == d[nm].to_bytes(32, "big"), (nq, npub, nm)
        if nq:
            assert dbg.fr[1, 7].tobytes() == d["hashed_bsb22"][0].to_bytes(32, "big")
        assert dbg.g1[1, 0].tobytes() == bo.g1_to_bytes(d["lin_digest"])
        assert dbg.g1[1, 1].tobytes() == bo.g1_to_bytes(d["folded_digest"])
        assert dbg.g1[1, 2].tobytes() == bo.g1_to_bytes(d["pair_g1"][0])
        assert dbg.g1[1, 3].tobytes() == bo.g1_to_bytes(d["pair_g1"][1])
        assert dbg.miller[1].tobytes() == bo.fp12_to_bytes(d["miller"]) and dbg.gt[1].tobytes() == bo.fp12_to_bytes(d["gt"])


def test_library_drawn_batch_opening_scalars(gpu):
    """rnd = None is the production path: the library draws the scalars of kzg::batch_verify_multi_points from the OS
    CSPRNG (kzg.rs:149-154).  Verdicts do not depend on them; the pairing inputs do, so two calls differ."""
    vk = plonk_vk_bytes()
    pr, xs = plonk_fixture("sha2")
    st1, d1 = gpu.PlonkVerifier.verify_batch([pr] * 4, vk, [xs] * 4, debug=True)
    st2, d2 = gpu.PlonkVerifier.verify_batch([pr] * 4, vk, [xs] * 4, debug=True)
    assert (st1 == gpu.OK_TRUE).all() and (st2 == gpu.OK_TRUE).all()
    seen = {d.g1[i, 2].tobytes() for d in (d1, d2) for i in range(4)}
    assert len(seen) == 8


def test_joint_msm_form_and_thread_per_proof_pairing_on_small_inputs():
    """Large batches evaluate the MSM terms of each sum jointly (shared doublings) and run stage E with one proof per
    thread; small ones use one thread per term and three lanes per proof.  Forcing the large-batch forms on the bundled
    fixtures (environment switches, read once per process -> a subprocess) must reproduce every golden intermediate and
    the whole mutation -> status map."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests"); sys.path.insert(0, %r + "/oracle")
import __graft_entry__ as ge
from helpers import PLONK_STATUS, load_json, plonk_fixture, plonk_vk_bytes, pt_bytes
pkg = ge.load_package(); pkg.init(None)
vk = plonk_vk_bytes()
gold = load_json("plonk_golden.json")
progs = ["fibonacci", "is-prime", "sha2", "tendermint"]
proofs, inputs, rnds = [], [], []
for prog in progs:
    pr, xs = plonk_fixture(prog)
    proofs.append(pr), inputs.append(xs), rnds.append(int(gold[prog]["rnd"], 16))
status, dbg = pkg.PlonkVerifier.verify_batch(proofs, vk, inputs, rnd=rnds, debug=True)
assert (status == 0).all()
for i, prog in enumerate(progs):
    g = gold[prog]
    assert dbg.g1[i, 0].tobytes() == pt_bytes(g["lin_digest"]) and dbg.g1[i, 1].tobytes() == pt_bytes(g["folded_digest"])
    assert dbg.g1[i, 2].tobytes() == pt_bytes(g["pair_g1"][0]) and dbg.g1[i, 3].tobytes() == pt_bytes(g["pair_g1"][1])
    assert dbg.miller[i].tobytes().hex() == g["miller"] and dbg.gt[i].tobytes().hex() == g["gt"]
muts = load_json("plonk_mutations.json")
st = pkg.PlonkVerifier.verify_batch([bytes.fromhex(m["raw_proof"]) for m in muts], vk,
                                    [[int(s) for s in m["inputs"]] for m in muts], rnd=[77] * len(muts))
assert [int(s) for s in st] == [PLONK_STATUS[m["status"]] for m in muts]
print("JOINT-OK")
''' % (ROOT, ROOT, ROOT)
    env = dict(os.environ, BN254V_PLONK_JOINT_MIN="0", BN254V_TRIO_MAX="0")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=280)
    assert "JOINT-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
