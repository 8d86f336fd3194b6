"""CPU: pins the C++ restatement (oracle/bn254_ref.cpp, the fast oracle and CPU baseline) on the Python oracle's
golden vectors, and the two workload generators on each other."""
import numpy as np

import bn254_oracle as bo
import ref_cpu
from helpers import load_json, pt_bytes


def test_pairing_golden():
    for c in load_json("pairing_golden.json"):
        k = c["k"]
        _, one, ml, gt = ref_cpu.pairing_product_batch(np.frombuffer(bytes.fromhex(c["g1"]), np.uint8),
                                                       np.frombuffer(bytes.fromhex(c["g2"]), np.uint8), k)
        assert ml[0].tobytes().hex() == c["miller"] and gt[0].tobytes().hex() == c["gt"]
        assert bool(one[0]) == c["is_one"]


def test_groth16_golden_reference_equation():
    case = load_json("groth16_golden.json")["cases"][0]
    proofs = np.array([np.frombuffer(bytes.fromhex(p["proof"]), np.uint8) for p in case["proofs"]])
    inputs = np.array([[np.frombuffer(int(x).to_bytes(32, "big"), np.uint8) for x in p["inputs"]] for p in case["proofs"]])
    _, st, dl, dm, dg = ref_cpu.groth16_verify_batch(bytes.fromhex(case["vk"]), proofs, inputs, threads=2, debug=True)
    for i, p in enumerate(case["proofs"]):
        assert st[i] == (0 if p["valid"] else 1)
        assert dl[i].tobytes() == pt_bytes(p["L"])
        assert dm[i].tobytes().hex() == p["miller"] and dg[i].tobytes().hex() == p["gt"]


def test_generators_agree_and_verdicts():
    vk, proofs, inputs, expected = ref_cpu.groth16_synth(424242, 10, first_index=77)
    td = bo.Groth16Trapdoor(424242, 2, 0)
    assert vk == td.vk_bytes()
    for i in range(10):
        pb, xs, valid = td.proof(77 + i)
        assert pb == proofs[i].tobytes() and valid == (expected[i] == 0)
        assert [int.from_bytes(inputs[i, j].tobytes(), "big") for j in range(2)] == xs
    _, st = ref_cpu.groth16_verify_batch(vk, proofs, inputs, threads=2)
    assert (st == expected).all()


def test_malformed_classes():
    from helpers import groth16_malformed_suite
    td = bo.Groth16Trapdoor(7, 2, 0)
    vk = td.vk_bytes()
    names = {"OK_TRUE": 0, "OK_FALSE": 1, "PANIC_FIELD_NOT_MEMBER": 16, "PANIC_NOT_ON_CURVE": 17,
             "PANIC_NOT_IN_SUBGROUP": 18, "PANIC_IDENTITY": 19}
    for name, pb, xs, want in groth16_malformed_suite(td):
        if len(pb) != 256 or len(xs) != 2:
            continue
        proofs = np.frombuffer(pb, np.uint8).reshape(1, 256)
        inputs = np.array([[np.frombuffer(int(x).to_bytes(32, "big"), np.uint8) for x in xs]])
        _, st = ref_cpu.groth16_verify_batch(vk, proofs, inputs)
        assert st[0] == names[want], name


def test_plonk_port_fixtures_and_mutation_map():
    """The C++ PlonK restatement: bundled fixtures Ok(true) with the golden GT value, and the status of every committed
    mutated proof (4 programs x 24 classes), i.e. the same map the Python oracle produced."""
    import workloads
    from helpers import PLONK_STATUS, plonk_fixture
    vk = workloads.plonk_vk_bytes()
    gold = load_json("plonk_golden.json")
    for prog, g in gold.items():
        pr, xs = plonk_fixture(prog)
        proofs = np.frombuffer(pr, np.uint8).reshape(1, -1)
        inputs = np.array([[np.frombuffer(x.to_bytes(32, "big"), np.uint8) for x in xs]])
        rnd = np.frombuffer(int(g["rnd"], 16).to_bytes(32, "big"), np.uint8).reshape(1, 32)
        _, st, gt = ref_cpu.plonk_verify_batch(vk, proofs, inputs, rnd, want_gt=True)
        assert st[0] == 0 and gt[0].tobytes().hex() == g["gt"]
    muts = load_json("plonk_mutations.json")
    stride = max(len(m["raw_proof"]) // 2 for m in muts)
    proofs = np.zeros((len(muts), stride), np.uint8)
    lens = np.zeros(len(muts), np.uint32)
    for i, m in enumerate(muts):
        b = bytes.fromhex(m["raw_proof"])
        proofs[i, :len(b)] = np.frombuffer(b, np.uint8)
        lens[i] = len(b)
    inputs = np.array([[np.frombuffer(int(s).to_bytes(32, "big"), np.uint8) for s in m["inputs"]] for m in muts])
    rnd = np.tile(np.frombuffer((77).to_bytes(32, "big"), np.uint8), (len(muts), 1))
    _, st = ref_cpu.plonk_verify_batch(vk, proofs, inputs, rnd, threads=4, lens=lens)
    for m, s in zip(muts, st):
        assert s == PLONK_STATUS[m["status"]], (m["program"], m["mutation"])
