"""CPU ORACLE (test infrastructure only) -- PlonK half.

Line-by-line restatement of the reference's gnark-PlonK verifier:

  * verifier/src/plonk/verify.rs:46-317    verify_plonk
  * verifier/src/plonk/verify.rs:319-396   bind_public_data / derive_randomness / batch_invert
  * verifier/src/plonk/kzg.rs:46-190       derive_gamma / fold / fold_proof / batch_verify_multi_points
  * verifier/src/transcript.rs:14-107      Fiat-Shamir transcript (SHA-256)
  * verifier/src/hash_to_field.rs:30-97    RFC 9380 expand_message_xmd
  * verifier/src/plonk/converter.rs:18-185 gnark VK / proof framing, g1_to_bytes

SHA-256 is taken from hashlib (FIPS 180-4; the reference uses the `sha2` crate 0.10.8).
Pinned by tests/test_oracle.py against the four bundled PlonK fixtures and the golden
challenge / digest values of SURVEY.md Appendix C.
"""
from __future__ import annotations

import hashlib

from bn254_oracle import (P, R, PanicError, g1_add, g1_mul, g1_neg, g1_msm, g1_to_bytes, fr_from_slice,
                          uncompressed_bytes_to_g1_point, compressed_x_to_g1_point,
                          compressed_x_to_g2_point, miller_product, final_exponentiation, FP12_ONE)


class PlonkError(Exception):
    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


# ------------------------------------------------------------------ transcript.rs
class Transcript:
    def __init__(self, ids):
        self.order = list(ids)
        self.bindings = {i: [] for i in ids}
        self.value = {}
        self.prev = None  # (position, digest)

    def bind(self, cid, data: bytes):
        if cid not in self.bindings:
            raise PlonkError("CHALLENGE_NOT_FOUND")
        if cid in self.value:
            raise PlonkError("CHALLENGE_ALREADY_COMPUTED")
        self.bindings[cid].append(bytes(data))

    def compute_challenge(self, cid) -> bytes:
        if cid not in self.bindings:
            raise PlonkError("CHALLENGE_NOT_FOUND")
        if cid in self.value:
            return self.value[cid]
        pos = self.order.index(cid)
        h = hashlib.sha256()
        h.update(cid.encode())
        if pos != 0:
            if self.prev is None or self.prev[0] != pos - 1:
                raise PlonkError("PREVIOUS_CHALLENGE_NOT_COMPUTED")
            h.update(self.prev[1])
        for b in self.bindings[cid]:
            h.update(b)
        d = h.digest()
        self.value[cid] = d
        self.prev = (pos, d)
        return d


# ------------------------------------------------------------------ hash_to_field.rs
def expand_msg_xmd(msg: bytes, dst: bytes, length: int) -> bytes:
    ell = (length + 31) // 32
    assert ell <= 255 and len(dst) <= 255
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(b"\x00" * 64 + msg + bytes([(length >> 8) & 0xFF, length & 0xFF, 0]) + dst_prime).digest()
    b1 = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bytearray(b1)
    bi = b1
    for i in range(2, ell + 1):
        x = bytes(p ^ q for p, q in zip(b0, bi))
        bi = hashlib.sha256(x + bytes([i]) + dst_prime).digest()
        out += bi
    return bytes(out[:length])


def hash_to_field_bsb22(msg: bytes) -> int:
    """WrappedHashToField::new(b"BSB22-Plonk"); write(msg); sum() -> 48 bytes -> mod r."""
    return int.from_bytes(expand_msg_xmd(msg, b"BSB22-Plonk", 48), "big") % R


def fr_bytes(v: int) -> bytes:
    return int(v % R).to_bytes(32, "big")


# ------------------------------------------------------------------ plonk/converter.rs
def load_plonk_verifying_key_from_bytes(buf: bytes):
    try:
        size = int.from_bytes(buf[0:8], "big")
        size_inv = fr_from_slice(buf[8:40])
        generator = fr_from_slice(buf[40:72])
        nb_public = int.from_bytes(buf[72:80], "big")
        coset_shift = fr_from_slice(buf[80:112])
        names = ["s0", "s1", "s2", "ql", "qr", "qm", "qo", "qk"]
        pts = {n: compressed_x_to_g1_point(buf[112 + 32 * i:144 + 32 * i]) for i, n in enumerate(names)}
        if len(buf) < 372:
            raise PanicError("SHORT_BUFFER")
        nqcp = int.from_bytes(buf[368:372], "big")
        off = 372
        qcp = []
        for _ in range(nqcp):
            qcp.append(compressed_x_to_g1_point(buf[off:off + 32]))
            off += 32
        g1 = compressed_x_to_g1_point(buf[off:off + 32])
        g2_0 = compressed_x_to_g2_point(buf[off + 32:off + 96])
        g2_1 = compressed_x_to_g2_point(buf[off + 96:off + 160])
        off += 160 + 33788
        if len(buf) < off + 8:
            raise PanicError("SHORT_BUFFER")
        nidx = int.from_bytes(buf[off:off + 8], "big")
        off += 8
        if len(buf) < off + 8 * nidx:
            raise PanicError("SHORT_BUFFER")
        idx = [int.from_bytes(buf[off + 8 * i:off + 8 * i + 8], "big") for i in range(nidx)]
    except IndexError:
        raise PanicError("SHORT_BUFFER")
    vk = {"size": size, "size_inv": size_inv, "generator": generator, "nb_public": nb_public,
          "coset_shift": coset_shift, "s": [pts["s0"], pts["s1"], pts["s2"]], "qcp": qcp,
          "g1": g1, "g2": [g2_0, g2_1], "cci": idx}
    for n in ("ql", "qr", "qm", "qo", "qk"):
        vk[n] = pts[n]
    return vk


def _need(buf, n):
    if len(buf) < n:
        raise PanicError("SHORT_BUFFER")


def load_plonk_proof_from_bytes(buf: bytes):
    _need(buf, 516)
    pts = [uncompressed_bytes_to_g1_point(buf[64 * i:64 * i + 64]) for i in range(8)]
    ncl = int.from_bytes(buf[512:516], "big")
    off = 516
    claimed = []
    for _ in range(ncl):
        _need(buf, off + 32)
        claimed.append(fr_from_slice(buf[off:off + 32]))
        off += 32
    _need(buf, off + 100)
    zs_h = uncompressed_bytes_to_g1_point(buf[off:off + 64])
    zs_v = fr_from_slice(buf[off + 64:off + 96])
    nb = int.from_bytes(buf[off + 96:off + 100], "big")
    off += 100
    bsb = []
    for _ in range(nb):
        _need(buf, off + 64)
        bsb.append(uncompressed_bytes_to_g1_point(buf[off:off + 64]))
        off += 64
    return {"lro": pts[0:3], "z": pts[3], "h": pts[4:7], "batched_h": pts[7], "claimed": claimed,
            "zs_h": zs_h, "zs_value": zs_v, "bsb22": bsb}


# ------------------------------------------------------------------ plonk/verify.rs helpers
def _bind_public_data(fs, vk, public_inputs):
    for pt in vk["s"]:
        fs.bind("gamma", g1_to_bytes(pt))
    for n in ("ql", "qr", "qm", "qo", "qk"):
        fs.bind("gamma", g1_to_bytes(vk[n]))
    for pt in vk["qcp"]:
        fs.bind("gamma", g1_to_bytes(pt))
    for x in public_inputs:
        fs.bind("gamma", fr_bytes(x))


def _derive_randomness(fs, cid, points):
    for pt in points or []:
        fs.bind(cid, g1_to_bytes(pt))
    return int.from_bytes(fs.compute_challenge(cid), "big") % R


def _fr_inv(a):
    return pow(a, R - 2, R)


def batch_invert(v):
    """verifier/src/plonk/verify.rs:364-396 (zeros are skipped)."""
    nz = [f for f in v if f % R]
    if not nz:
        return list(v)
    prod = []
    tmp = 1
    for f in nz:
        tmp = tmp * f % R
        prod.append(tmp)
    tmp = _fr_inv(tmp)
    out = list(v)
    idxs = [i for i, f in enumerate(v) if f % R]
    for k in range(len(idxs) - 1, -1, -1):
        i = idxs[k]
        s = prod[k - 1] if k > 0 else 1
        new_tmp = tmp * v[i] % R
        out[i] = tmp * s % R
        tmp = new_tmp
    return out


# ------------------------------------------------------------------ plonk/kzg.rs
def _derive_gamma(point, digests, claimed, data_transcript):
    t = Transcript(["gamma"])
    t.bind("gamma", fr_bytes(point))
    for d in digests:
        t.bind("gamma", g1_to_bytes(d))
    for c in claimed:
        t.bind("gamma", fr_bytes(c))
    if data_transcript is not None:
        t.bind("gamma", data_transcript)
    return int.from_bytes(t.compute_challenge("gamma"), "big") % R


def _msm(points, scalars):
    r = g1_msm(points, scalars)
    if r is None:
        raise PanicError("IDENTITY")
    return r


def _fold(di, fai, ci):
    ev = 0
    for f, c in zip(fai, ci):
        ev = (ev + f * c) % R
    return _msm(di, ci), ev


def fold_proof(digests, batched_h, claimed, point, data_transcript, debug=None):
    if len(digests) != len(claimed):
        raise PlonkError("INVALID_NUMBER_OF_DIGESTS")
    gamma = _derive_gamma(point, digests, claimed, data_transcript)
    gi = [1]
    for _ in range(1, len(digests)):
        gi.append(gi[-1] * gamma % R)
    folded_digest, folded_eval = _fold(digests, claimed, gi)
    if debug is not None:
        debug["kzg_gamma"] = gamma
    return (batched_h, folded_eval), folded_digest


def batch_verify_multi_points(digests, proofs, points, vk, rnd, debug=None):
    """`rnd` replaces the reference's OsRng draw: random_numbers = [1, rnd]."""
    n = len(digests)
    if n != len(proofs) or n != len(points):
        raise PlonkError("INVALID_NUMBER_OF_DIGESTS")
    assert n == 2, "the reference has todo!() for a single digest"
    rn = [1, rnd % R]
    quotients = [pr[0] for pr in proofs]
    folded_quotients = _msm(quotients, rn)
    evals = [pr[1] for pr in proofs]
    folded_digests, folded_evals = _fold(digests, evals, rn)
    fec = g1_mul(vk["g1"], folded_evals)
    if fec is None:
        raise PanicError("IDENTITY")
    folded_digests = g1_add(folded_digests, g1_neg(fec))
    if folded_digests is None:
        raise PanicError("IDENTITY")
    rn2 = [rn[i] * points[i] % R for i in range(n)]
    fpq = _msm(quotients, rn2)
    folded_digests = g1_add(folded_digests, fpq)
    if folded_digests is None:
        raise PanicError("IDENTITY")
    folded_quotients = g1_neg(folded_quotients)
    ml = miller_product([(folded_digests, vk["g2"][0]), (folded_quotients, vk["g2"][1])])
    gt = final_exponentiation(ml)
    if debug is not None:
        debug.update({"pair_g1": [folded_digests, folded_quotients], "miller": ml, "gt": gt})
    if gt != FP12_ONE:
        raise PlonkError("PAIRING_CHECK_FAILED")


# ------------------------------------------------------------------ verify_plonk
def verify_plonk(vk, proof, public_inputs, rnd=0xDEADBEEF, debug=None):
    """Returns True or raises PlonkError / PanicError (the reference never returns Ok(false))."""
    if len(proof["bsb22"]) != len(vk["qcp"]):
        raise PlonkError("BSB22_COMMITMENT_MISMATCH")
    if len(public_inputs) != vk["nb_public"]:
        raise PlonkError("INVALID_WITNESS")
    fs = Transcript(["gamma", "beta", "alpha", "zeta"])
    _bind_public_data(fs, vk, public_inputs)
    gamma = _derive_randomness(fs, "gamma", proof["lro"])
    beta = _derive_randomness(fs, "beta", None)
    alpha = _derive_randomness(fs, "alpha", list(proof["bsb22"]) + [proof["z"]])
    zeta = _derive_randomness(fs, "zeta", proof["h"])

    n = vk["size"]
    zeta_power_n = pow(zeta, n, R)
    zh_zeta = (zeta_power_n - 1) % R
    if (zeta - 1) % R == 0:
        raise PlonkError("INVERSE_NOT_FOUND")
    lagrange_one = _fr_inv((zeta - 1) % R) * zh_zeta % R * vk["size_inv"] % R

    pi = 0
    accw = 1
    dens = []
    for _ in public_inputs:
        dens.append((zeta - accw) % R)
        accw = accw * vk["generator"] % R
    inv_dens = batch_invert(dens)
    accw = 1
    for i, w in enumerate(public_inputs):
        xi_li = zh_zeta * inv_dens[i] % R * vk["size_inv"] % R * accw % R * w % R
        accw = accw * vk["generator"] % R
        pi = (pi + xi_li) % R

    hashed = []
    for i, cci in enumerate(vk["cci"]):
        # proof.bsb22_commitments[i] -- index panic if the proof has fewer commitments
        if i >= len(proof["bsb22"]):
            raise PanicError("INDEX")
        hashed_cmt = hash_to_field_bsb22(g1_to_bytes(proof["bsb22"][i]))
        hashed.append(hashed_cmt)
        w_pow_i = pow(vk["generator"], vk["nb_public"] + cci, R)
        den = (zeta - w_pow_i) % R
        if den == 0:
            raise PanicError("DIV_BY_ZERO")
        lagrange = zh_zeta * w_pow_i % R * _fr_inv(den) % R * vk["size_inv"] % R
        pi = (pi + lagrange * hashed_cmt) % R

    cl = proof["claimed"]
    if len(cl) < 6:
        raise PanicError("INDEX")
    l, r, o, s1, s2 = cl[1], cl[2], cl[3], cl[4], cl[5]
    zu = proof["zs_value"]
    a2l1 = lagrange_one * alpha % R * alpha % R
    const_lin = (beta * s1 + gamma + l) % R
    const_lin = const_lin * ((beta * s2 + gamma + r) % R) % R
    const_lin = const_lin * ((o + gamma) % R) % R
    const_lin = const_lin * alpha % R * zu % R
    const_lin = (const_lin - a2l1 + pi) % R
    const_lin = (-const_lin) % R
    if debug is not None:
        debug.update({"gamma": gamma, "beta": beta, "alpha": alpha, "zeta": zeta, "pi": pi,
                      "hashed_bsb22": hashed, "const_lin": const_lin})
    if const_lin != cl[0]:
        raise PlonkError("OPENING_POLY_MISMATCH")

    _s1 = (beta * s1 + l + gamma) % R
    tmp = (beta * s2 + r + gamma) % R
    _s1 = _s1 * tmp % R * beta % R * alpha % R * zu % R
    u = vk["coset_shift"]
    _s2 = (beta * zeta + gamma + l) % R
    _s2 = _s2 * ((beta * u % R * zeta + gamma + r) % R) % R
    _s2 = _s2 * ((beta * u % R * u % R * zeta + gamma + o) % R) % R
    _s2 = (-(_s2 * alpha)) % R
    coeff_z = (a2l1 + _s2) % R
    rl = l * r % R
    zn2 = pow(zeta, n + 2, R)
    zn2sq = zn2 * zn2 % R
    zeta_n_plus_two_zh = (-(zn2 * zh_zeta)) % R
    zeta_n_plus_two_square_zh = (-(zn2sq * zh_zeta)) % R
    zh = (-zh_zeta) % R

    points = list(proof["bsb22"]) + [vk["ql"], vk["qr"], vk["qm"], vk["qo"], vk["qk"], vk["s"][2],
                                     proof["z"], proof["h"][0], proof["h"][1], proof["h"][2]]
    scalars = list(cl[6:]) + [l, r, rl, o, 1, _s1, coeff_z, zh, zeta_n_plus_two_zh, zeta_n_plus_two_square_zh]
    # AffineG1::msm zips points with scalars (shorter length wins)
    lin_digest = _msm(points[:len(scalars)], scalars[:len(points)])

    digests = [lin_digest, proof["lro"][0], proof["lro"][1], proof["lro"][2], vk["s"][0], vk["s"][1]] + list(vk["qcp"])
    (folded_h, folded_eval), folded_digest = fold_proof(digests, proof["batched_h"], cl, zeta, fr_bytes(zu), debug)
    if debug is not None:
        debug.update({"lin_digest": lin_digest, "folded_digest": folded_digest, "folded_eval": folded_eval})
    shifted_zeta = zeta * vk["generator"] % R
    batch_verify_multi_points([folded_digest, proof["z"]],
                              [(folded_h, folded_eval), (proof["zs_h"], zu)],
                              [zeta, shifted_zeta], vk, rnd, debug)
    return True


def plonk_verifier_verify(proof_bytes, vk_bytes, public_inputs, rnd=0xDEADBEEF, debug=None):
    """PlonkVerifier::verify (verifier/src/lib.rs:69-74)."""
    proof = load_plonk_proof_from_bytes(proof_bytes)
    vk = load_plonk_verifying_key_from_bytes(vk_bytes)
    return verify_plonk(vk, proof, public_inputs, rnd, debug)
