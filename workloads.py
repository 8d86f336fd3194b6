"""Synthetic workloads shared by bench.py and the parity tests (data only; no oracle code is executed here).

PlonK (BASELINE.json configs[2]): the reference's 4 bundled SP1 PlonK proofs, replicated, with 50 % of the records
replaced by a mutated copy.  The mutated proofs and their expected statuses are the committed fixtures in
tests/golden/plonk_mutations.json (generated once by oracle/make_fixtures.py from the reference's bundled data).
"""
from __future__ import annotations

import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")
STATUS = {"OK_TRUE": 0, "ERR_OPENING_POLY_MISMATCH": 6, "ERR_PAIRING_CHECK_FAILED": 8}
PROGRAMS = ["fibonacci", "is-prime", "sha2", "tendermint"]


def plonk_vk_bytes() -> bytes:
    return open(os.path.join(GOLDEN, "plonk_vk.bin"), "rb").read()


def plonk_workload(n, seed=1, late_reject_only=False):
    """Returns (proofs[n, 904] u8, inputs[n, 2, 32] u8, rnd[n, 32] u8, expected[n] u8).
    Even records are the bundled proofs (round-robin over the 4 programs), odd records are mutated copies
    (round-robin over the mutation classes whose outcome is Err(..), i.e. the malformed/panic classes are left
    to the edge-case tests).  `late_reject_only` keeps only the classes that fail in the final pairing check."""
    muts = json.load(open(os.path.join(GOLDEN, "plonk_mutations.json")))
    valid = {m["program"]: m for m in muts if m["mutation"] == "valid"}
    bad = [m for m in muts if m["status"] in ("ERR_OPENING_POLY_MISMATCH", "ERR_PAIRING_CHECK_FAILED")]
    if late_reject_only:
        bad = [m for m in bad if m["status"] == "ERR_PAIRING_CHECK_FAILED"]

    def rec(m):
        p = np.frombuffer(bytes.fromhex(m["raw_proof"]), dtype=np.uint8)
        x = np.stack([np.frombuffer(int(s).to_bytes(32, "big"), dtype=np.uint8) for s in m["inputs"]])
        return p, x, STATUS[m["status"]]

    vrec = [rec(valid[p]) for p in PROGRAMS]
    brec = [rec(m) for m in bad]
    stride = 904
    proofs = np.zeros((n, stride), dtype=np.uint8)
    inputs = np.zeros((n, 2, 32), dtype=np.uint8)
    expected = np.zeros(n, dtype=np.uint8)
    # build by tiling the two pools
    even = np.arange(0, n, 2)
    odd = np.arange(1, n, 2)
    vp = np.stack([r[0] for r in vrec]); vx = np.stack([r[1] for r in vrec]); vs = np.array([r[2] for r in vrec], np.uint8)
    bp = np.stack([r[0] for r in brec]); bx = np.stack([r[1] for r in brec]); bs = np.array([r[2] for r in brec], np.uint8)
    vi = (even // 2) % len(vrec)
    bi = (odd // 2) % len(brec)
    proofs[even], inputs[even], expected[even] = vp[vi], vx[vi], vs[vi]
    proofs[odd], inputs[odd], expected[odd] = bp[bi], bx[bi], bs[bi]
    rng = np.random.default_rng(seed)
    rnd = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    return proofs, inputs, rnd, expected
