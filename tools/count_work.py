"""Counts the algorithmic work of the GPU-shaped algorithms: field multiplications executed by the kernels'
own per-proof routines, run on the host build (tests/hostsim, -DBN254_COUNT_MULS).  One Fp/Fr multiplication =
136 limb multiply-adds with 8 x 32-bit CIOS Montgomery (64 product + 72 reduction).  Writes profiles/workcount.json,
which bench.py uses for `roofline.achieved`.
Run: python tools/count_work.py
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main(write=True):
    import bn254_oracle as bo
    from helpers import build_hostsim, load_json, plonk_fixture, plonk_vk_bytes
    hs = build_hostsim()
    hs.hs_mul_count.restype = ctypes.c_ulonglong
    hs.hs_groth16_vk_new_tables.restype = ctypes.c_void_p
    hs.hs_groth16_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                     ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    hs.hs_plonk_vk_new.restype = ctypes.c_void_p
    hs.hs_plonk_vk_new.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    hs.hs_plonk_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                   ctypes.c_char_p] + [ctypes.c_char_p] * 4
    out = {"macs_per_mul": 136, "unit": "limb multiply-adds (32x32+64) per proof, counted on the host build of the kernels' "
                                        "per-proof routines: 136 per Montgomery multiplication, 64 per wide product, "
                                        "72 per wide reduction; *_fp_mul = macs / 136 (multiplication equivalents)"}
    # Groth16: mean over 8 trapdoor proofs (valid and corrupted cost the same)
    case = load_json("groth16_golden.json")["cases"][0]
    vk = bo.load_groth16_verifying_key_from_bytes(bytes.fromhex(case["vk"]))
    blob = bo.g1_to_bytes(vk["alpha"]) + bo.g2_to_bytes(vk["beta2"]) + bo.g2_to_bytes(vk["gamma2"]) + \
        bo.g2_to_bytes(bo.g2_neg(vk["delta2"])) + b"".join(bo.g1_to_bytes(k) for k in vk["k"])
    h = hs.hs_groth16_vk_new_tables(blob, len(vk["k"]))
    hs.hs_mul_count(1)
    n = 0
    for pr in case["proofs"]:
        inputs = b"".join(int(x).to_bytes(32, "big") for x in pr["inputs"])
        hs.hs_groth16_verify(h, bytes.fromhex(pr["proof"]), 256, inputs, 2, None, None, None)
        n += 1
    out["groth16_macs"] = hs.hs_mul_count(1) // n
    out["groth16_fp_mul"] = out["groth16_macs"] // 136
    # split at the kernel boundary: the final exponentiation alone (same Fq12 chain for every input)
    c = load_json("pairing_golden.json")[0]
    hs.hs_mul_count(1)
    hs.hs_final_exp_only(bytes.fromhex(c["miller"]))
    out["groth16_finish_macs"] = hs.hs_mul_count(1)
    out["groth16_miller_macs"] = out["groth16_macs"] - out["groth16_finish_macs"]
    # ... of which k_groth16_prepare (decode, validation, prepare_inputs)
    hs.hs_groth16_prepare_macs.restype = ctypes.c_ulonglong
    hs.hs_groth16_prepare_macs.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int]
    tot = 0
    for pr in case["proofs"]:
        inputs = b"".join(int(x).to_bytes(32, "big") for x in pr["inputs"])
        tot += hs.hs_groth16_prepare_macs(h, bytes.fromhex(pr["proof"]), 256, inputs, 2)
    out["groth16_prepare_macs"] = tot // n
    # opt-in aggregate check (csrc/groth16_agg.cuh): one proof's share = validation, [r] A, [r] C | single-pair Miller loop
    hs.hs_groth16_agg_proof_macs.restype = ctypes.c_ulonglong
    hs.hs_groth16_agg_proof_macs.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                             ctypes.c_char_p, ctypes.POINTER(ctypes.c_ulonglong)]
    tot = cpart = 0
    for i, pr in enumerate(case["proofs"]):
        inputs = b"".join(int(x).to_bytes(32, "big") for x in pr["inputs"])
        cp = ctypes.c_ulonglong(0)
        rnd = bytes((37 * i + 11 * k + 5) & 255 for k in range(16))
        tot += hs.hs_groth16_agg_proof_macs(h, bytes.fromhex(pr["proof"]), 256, inputs, 2, rnd, ctypes.byref(cp))
        cpart += cp.value
    out["groth16_agg_macs"] = tot // n
    out["groth16_agg_prepare_macs"] = cpart // n
    # raw pairing products
    for c in load_json("pairing_golden.json"):
        if c["is_one"]:
            continue
        hs.hs_mul_count(1)
        hs.hs_pairing_product(c["k"], bytes.fromhex(c["g1"]), bytes.fromhex(c["g2"]), None, None)
        out["pairing_product_k%d_macs" % c["k"]] = hs.hs_mul_count(1)
    # PlonK: full path (valid proof) per stage of the staged GPU path, and early reject.  The VK-constant bases go
    # through the fixed-base window tables, exactly as on the device (hs_plonk_vk_add_tables) -- without them the
    # 8 + nQcp VK terms would be counted as variable-base multiplications the GPU never executes.
    vkb = plonk_vk_bytes()
    pv = hs.hs_plonk_vk_new(vkb, len(vkb))
    hs.hs_plonk_vk_add_tables.argtypes = [ctypes.c_void_p]
    hs.hs_plonk_vk_add_tables(pv)
    hs.hs_plonk_stage_macs.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_int,
                                       ctypes.c_char_p, ctypes.POINTER(ctypes.c_ulonglong)]
    pr, xs = plonk_fixture("fibonacci")
    inputs = b"".join(x.to_bytes(32, "big") for x in xs)
    hs.hs_mul_count(1)
    assert hs.hs_plonk_verify(pv, pr, len(pr), inputs, 2, (77).to_bytes(32, "big"), None, None, None, None) == 0
    out["plonk_full_path_macs"] = hs.hs_mul_count(1)
    out["plonk_full_path_mul"] = out["plonk_full_path_macs"] // 136
    stages = (ctypes.c_ulonglong * 5)()
    assert hs.hs_plonk_stage_macs(pv, pr, len(pr), inputs, 2, (77).to_bytes(32, "big"), stages) == 0
    out["plonk_stage_macs"] = {k: int(v) for k, v in zip(("stage_a", "terms0", "stage_c", "terms1", "stage_e"), stages)}
    assert sum(out["plonk_stage_macs"].values()) == out["plonk_full_path_macs"]
    # the large-batch form: MSM terms of a sum evaluated jointly (shared doublings), chunks of >= 2^15 proofs
    hs.hs_plonk_stage_macs_form.argtypes = hs.hs_plonk_stage_macs.argtypes + [ctypes.c_int]
    assert hs.hs_plonk_stage_macs_form(pv, pr, len(pr), inputs, 2, (77).to_bytes(32, "big"), stages, 1) == 0
    out["plonk_joint_stage_macs"] = {k: int(v) for k, v in zip(("stage_a", "terms0", "stage_c", "terms1", "stage_e"), stages)}
    out["plonk_joint_full_path_macs"] = sum(out["plonk_joint_stage_macs"].values())
    bad = [m for m in load_json("plonk_mutations.json") if m["program"] == "fibonacci" and m["mutation"] == "claimed0+1"][0]
    hs.hs_mul_count(1)
    hs.hs_plonk_verify(pv, bytes.fromhex(bad["raw_proof"]), 904, inputs, 2, (77).to_bytes(32, "big"), None, None, None, None)
    out["plonk_early_reject_macs"] = hs.hs_mul_count(1)
    out["plonk_early_reject_mul"] = out["plonk_early_reject_macs"] // 136
    if write:
        json.dump(out, open(os.path.join(ROOT, "profiles", "workcount.json"), "w"), indent=1)
    return out


if __name__ == "__main__":
    print(json.dumps(main(), indent=1))
