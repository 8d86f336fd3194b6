"""Regenerates tests/golden/ from the reference's bundled data (run in the build container,
where /root/reference exists; the GPU box only sees the committed outputs).

Inputs (data files only, no reference source is copied):
  /root/reference/examples/binaries/*_{plonk,groth16}_proof.bin   bincode SP1ProofWithPublicValues v2.0.0
  /root/reference/examples/program/elf/plonk                      guest ELF embedding plonk_vk.bin

Outputs:
  tests/golden/plonk_vk.bin                 34 368-byte gnark PlonK VK (sha256 = fixtures' plonk_vkey_hash)
  tests/golden/fixtures.json                raw proofs (hex) + public inputs (decimal) for the 8 envelopes
  tests/golden/plonk_golden.json            oracle intermediates for the 4 PlonK fixtures (challenges, digests, GT)
  tests/golden/plonk_mutations.json         mutated PlonK proofs + expected status
  tests/golden/groth16_golden.json          trapdoor Groth16 instances: VK, proofs, inputs, L, Miller, GT, verdicts
  tests/golden/pairing_golden.json          k-pair products: inputs, canonical Miller and GT values

Envelope layout follows what the reference's test unwraps at examples/script/src/main.rs:115-138.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import bn254_oracle as bo  # noqa: E402
import plonk_oracle as po  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(HERE, "..", "tests", "golden")
PROGRAMS = ["fibonacci", "is-prime", "sha2", "tendermint"]
PLONK_VK_OFFSET, PLONK_VK_LEN = 315128, 34368
PLONK_VK_SHA256 = "4aca240a3e5296e6a565f98dc728c6f48f8de4792a8fa365038c3b86952176f5"


def parse_envelope(data: bytes):
    """bincode: u32 variant (2 = Plonk, 3 = Groth16); Vec<String> public_inputs is a fixed [String;2]
    (u64 len + bytes each); String encoded_proof; String raw_proof; [u8;32] vkey hash; ..."""
    off = 0
    variant = int.from_bytes(data[off:off + 4], "little"); off += 4

    def rd_str():
        nonlocal off
        n = int.from_bytes(data[off:off + 8], "little"); off += 8
        s = data[off:off + n]; off += n
        return s.decode()

    in0, in1 = rd_str(), rd_str()
    encoded = rd_str()
    raw = rd_str()
    vkey_hash = data[off:off + 32]
    return {"variant": variant, "inputs": [in0, in1], "encoded_proof": encoded, "raw_proof": raw,
            "vkey_hash": vkey_hash.hex()}


def main():
    os.makedirs(OUT, exist_ok=True)
    elf = open(os.path.join(REF, "examples/program/elf/plonk"), "rb").read()
    vk = elf[PLONK_VK_OFFSET:PLONK_VK_OFFSET + PLONK_VK_LEN]
    assert hashlib.sha256(vk).hexdigest() == PLONK_VK_SHA256
    open(os.path.join(OUT, "plonk_vk.bin"), "wb").write(vk)

    fixtures = {}
    for prog in PROGRAMS:
        for mode in ("plonk", "groth16"):
            env = parse_envelope(open(os.path.join(REF, f"examples/binaries/{prog}_{mode}_proof.bin"), "rb").read())
            assert env["variant"] == (2 if mode == "plonk" else 3)
            if mode == "plonk":
                assert env["vkey_hash"] == PLONK_VK_SHA256
            fixtures[f"{prog}_{mode}"] = {"raw_proof": env["raw_proof"], "inputs": env["inputs"],
                                           "vkey_hash": env["vkey_hash"]}
    json.dump(fixtures, open(os.path.join(OUT, "fixtures.json"), "w"), indent=1)

    # ---- PlonK goldens
    hx = lambda v: "0x%064x" % v
    pt = lambda p: [hx(p[0]), hx(p[1])]
    golden = {}
    for prog in PROGRAMS:
        fx = fixtures[f"{prog}_plonk"]
        dbg = {}
        ok = po.plonk_verifier_verify(bytes.fromhex(fx["raw_proof"]), vk, [int(s) for s in fx["inputs"]],
                                      rnd=0xDEADBEEF, debug=dbg)
        assert ok is True
        golden[prog] = {
            "verdict": True, "rnd": hx(0xDEADBEEF),
            **{k: hx(dbg[k]) for k in ("gamma", "beta", "alpha", "zeta", "kzg_gamma", "pi", "const_lin", "folded_eval")},
            "hashed_bsb22": [hx(v) for v in dbg["hashed_bsb22"]],
            "lin_digest": pt(dbg["lin_digest"]), "folded_digest": pt(dbg["folded_digest"]),
            "pair_g1": [pt(p) for p in dbg["pair_g1"]],
            "miller": bo.fp12_to_bytes(dbg["miller"]).hex(), "gt": bo.fp12_to_bytes(dbg["gt"]).hex(),
        }
    json.dump(golden, open(os.path.join(OUT, "plonk_golden.json"), "w"), indent=1)

    # ---- PlonK mutations (SURVEY.md 8(d).3 / Appendix C status map)
    muts = []
    for prog in PROGRAMS:
        fx = fixtures[f"{prog}_plonk"]
        raw = bytes.fromhex(fx["raw_proof"])
        inputs = [int(s) for s in fx["inputs"]]
        for name, (mraw, minputs) in plonk_mutations(raw, inputs).items():
            try:
                po.plonk_verifier_verify(mraw, vk, minputs, rnd=0x1234567)
                status = "OK_TRUE"
            except po.PlonkError as e:
                status = "ERR_" + e.kind
            except bo.PanicError as e:
                status = "PANIC_" + e.kind
            muts.append({"program": prog, "mutation": name, "raw_proof": mraw.hex(),
                         "inputs": [str(v) for v in minputs], "status": status})
    json.dump(muts, open(os.path.join(OUT, "plonk_mutations.json"), "w"), indent=1)

    # ---- Groth16 trapdoor goldens
    g16 = {"cases": []}
    for seed, sign_mode in ((7, 0), (11, 1)):
        td = bo.Groth16Trapdoor(seed, 2, sign_mode)
        vkb = td.vk_bytes()
        case = {"seed": seed, "sign_mode": sign_mode, "vk": vkb.hex(), "proofs": []}
        for i in range(8):
            pb, xs, valid = td.proof(i)
            dbg = {}
            vkp = bo.load_groth16_verifying_key_from_bytes(vkb)
            if sign_mode == 1:
                # standard gnark check e(A,B) e(L,-gamma) e(C,-delta) == e(alpha, beta)
                vkp = dict(vkp, beta2=bo.g2_neg(vkp["beta2"]), gamma2=bo.g2_neg(vkp["gamma2"]))
            r = bo.verify_groth16(vkp, bo.load_groth16_proof_from_bytes(pb), xs, dbg)
            assert r == valid, (seed, i)
            case["proofs"].append({"proof": pb.hex(), "inputs": [str(x) for x in xs], "valid": valid,
                                   "L": pt(dbg["L"]), "miller": bo.fp12_to_bytes(dbg["miller"]).hex(),
                                   "gt": bo.fp12_to_bytes(dbg["gt"]).hex()})
        case["alpha_beta"] = bo.fp12_to_bytes(dbg["alpha_beta"]).hex()
        g16["cases"].append(case)
    json.dump(g16, open(os.path.join(OUT, "groth16_golden.json"), "w"), indent=1)

    # ---- raw pairing products
    pg = []
    for k in (1, 2, 3, 4):
        for trial in range(3):
            ss = [bo.synth_scalar(1000 + k, trial, 2 * j) for j in range(k)]
            ts = [bo.synth_scalar(1000 + k, trial, 2 * j + 1) for j in range(k)]
            if trial == 1 and k > 1:  # force product == 1: sum s_j t_j == 0
                acc = sum(s * t for s, t in zip(ss[:-1], ts[:-1])) % bo.R
                ss[-1] = (-acc) * pow(ts[-1], bo.R - 2, bo.R) % bo.R
            g1s = [bo.g1_mul(bo.G1_GEN, s) for s in ss]
            g2s = [bo.g2_mul(bo.G2_GEN, t) for t in ts]
            ml = bo.miller_product(list(zip(g1s, g2s)))
            gt = bo.final_exponentiation(ml)
            pg.append({"k": k, "g1": b"".join(bo.g1_to_bytes(p) for p in g1s).hex(),
                       "g2": b"".join(bo.g2_to_bytes(q) for q in g2s).hex(),
                       "miller": bo.fp12_to_bytes(ml).hex(), "gt": bo.fp12_to_bytes(gt).hex(),
                       "is_one": gt == bo.FP12_ONE})
    json.dump(pg, open(os.path.join(OUT, "pairing_golden.json"), "w"), indent=1)
    print("fixtures written to", os.path.abspath(OUT))


def plonk_mutations(raw: bytes, inputs):
    """Mutation classes of SURVEY.md 8(d).3.  Offsets follow the gnark proof layout
    (verifier/src/plonk/converter.rs:121-178) for the fixture shape (7 claimed, 1 BSB22)."""
    out = {}
    R = bo.R

    def fr_add1(b: bytearray, off):
        v = (int.from_bytes(b[off:off + 32], "big") + 1) % R
        b[off:off + 32] = v.to_bytes(32, "big")

    def pt_scale(b: bytearray, off, k):
        p = bo.uncompressed_bytes_to_g1_point(bytes(b[off:off + 64]))
        b[off:off + 64] = bo.g1_to_bytes(bo.g1_mul(p, k))

    out["valid"] = (raw, list(inputs))
    out["input0+1"] = (raw, [(inputs[0] + 1) % R, inputs[1]])
    out["input1+1"] = (raw, [inputs[0], (inputs[1] + 1) % R])
    ncl = int.from_bytes(raw[512:516], "big")
    off_zs = 516 + 32 * ncl
    for ci in range(ncl):
        b = bytearray(raw); fr_add1(b, 516 + 32 * ci); out[f"claimed{ci}+1"] = (bytes(b), list(inputs))
    b = bytearray(raw); fr_add1(b, off_zs + 64); out["zu+1"] = (bytes(b), list(inputs))
    for name, off, k in (("L*2", 0, 2), ("R*3", 64, 3), ("O*2", 128, 2), ("Z*2", 192, 2), ("H0*2", 256, 2),
                         ("H1*5", 320, 5), ("H2*2", 384, 2), ("batchedH*2", 448, 2), ("zshiftedH*3", off_zs, 3),
                         ("BSB22*2", off_zs + 100, 2)):
        b = bytearray(raw); pt_scale(b, off, k); out[name] = (bytes(b), list(inputs))
    # malformed classes (reference panics)
    b = bytearray(raw); b[0:32] = (bo.P + 5).to_bytes(32, "big"); out["Lx>=p"] = (bytes(b), list(inputs))
    b = bytearray(raw); b[63] ^= 1; out["L-offcurve"] = (bytes(b), list(inputs))
    b = bytearray(raw); b[516:548] = (bo.R + 1).to_bytes(32, "big"); out["claimed0>=r"] = (bytes(b), list(inputs))
    return out


if __name__ == "__main__":
    main()
