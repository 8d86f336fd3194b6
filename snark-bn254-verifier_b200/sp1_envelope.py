"""SP1 `SP1ProofWithPublicValues` v2.0.0 envelope (bincode) -> (kind, raw gnark proof, public inputs).

Host-side framing only: this is what the reference's example script does before it calls the verifier
(`SP1ProofWithPublicValues::load` + `try_as_plonk()/try_as_groth_16()`, `hex::decode(raw_proof)`, decimal public
inputs -> `Fr`; examples/script/src/main.rs:115-138, :193-213), so that the bundled `examples/binaries/*.bin`
fixtures feed `verify` / `verify_batch` directly.  Layout (little-endian lengths): u32 variant
(2 = Plonk, 3 = Groth16) | String public_inputs[0] | String public_inputs[1] | String encoded_proof |
String raw_proof (hex) | [u8; 32] vkey hash | ...
"""
from __future__ import annotations

PLONK, GROTH16 = 2, 3


class EnvelopeError(ValueError):
    pass


def parse(data: bytes):
    """Returns {"kind": "plonk"|"groth16", "raw_proof": bytes, "public_inputs": [int, int], "vkey_hash": bytes}."""
    off = 0

    def need(n):
        if off + n > len(data):
            raise EnvelopeError("truncated SP1 proof envelope")

    need(4)
    variant = int.from_bytes(data[0:4], "little")
    off = 4
    if variant not in (PLONK, GROTH16):
        raise EnvelopeError("unsupported SP1 proof variant %d (only Plonk = 2 and Groth16 = 3 carry a gnark proof)" % variant)

    def rd_str():
        nonlocal off
        need(8)
        n = int.from_bytes(data[off:off + 8], "little")
        off += 8
        need(n)
        s = data[off:off + n]
        off += n
        return s.decode("ascii")

    in0, in1 = rd_str(), rd_str()
    rd_str()  # encoded_proof (the Solidity-verifier encoding; unused here)
    raw_hex = rd_str()
    need(32)
    vkey_hash = data[off:off + 32]
    return {"kind": "plonk" if variant == PLONK else "groth16", "raw_proof": bytes.fromhex(raw_hex),
            "public_inputs": [int(in0), int(in1)], "vkey_hash": bytes(vkey_hash)}


def load(path):
    with open(path, "rb") as f:
        return parse(f.read())
