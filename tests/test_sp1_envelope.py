"""CPU: the SP1 envelope loader (host framing) on two of the reference's bundled fixtures."""
import hashlib
import os

import pytest

from conftest import GOLDEN
from helpers import load_json, plonk_vk_bytes


def _mod(pkg):
    from importlib import import_module
    return import_module("snark_bn254_verifier_b200.sp1_envelope")


def test_envelopes_match_extracted_fixtures(pkg):
    env = _mod(pkg)
    fx = load_json("fixtures.json")
    for name, kind, n in (("sha2_plonk", "plonk", 904), ("sha2_groth16", "groth16", 324)):
        e = env.load(os.path.join(GOLDEN, "envelopes", name + "_proof.bin"))
        assert e["kind"] == kind and len(e["raw_proof"]) == n
        assert e["raw_proof"].hex() == fx[name]["raw_proof"]
        assert [str(v) for v in e["public_inputs"]] == fx[name]["inputs"]
        assert e["vkey_hash"].hex() == fx[name]["vkey_hash"]
    assert env.load(os.path.join(GOLDEN, "envelopes", "sha2_plonk_proof.bin"))["vkey_hash"] == \
        hashlib.sha256(plonk_vk_bytes()).digest()


def test_envelope_errors(pkg):
    env = _mod(pkg)
    data = open(os.path.join(GOLDEN, "envelopes", "sha2_plonk_proof.bin"), "rb").read()
    with pytest.raises(env.EnvelopeError):
        env.parse(data[:100])
    with pytest.raises(env.EnvelopeError):
        env.parse((0).to_bytes(4, "little") + data[4:])


@pytest.mark.gpu
def test_envelope_feeds_verify(gpu):
    """Fixture file -> envelope -> PlonkVerifier.verify == Ok(true), as the reference's test_programs expects."""
    env = _mod(gpu)
    e = env.load(os.path.join(GOLDEN, "envelopes", "sha2_plonk_proof.bin"))
    assert gpu.PlonkVerifier.verify(e["raw_proof"], plonk_vk_bytes(), e["public_inputs"]) is True
