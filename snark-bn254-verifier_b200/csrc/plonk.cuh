// Per-proof gnark PlonK verification (BSB22 commitments, linearised polynomial, KZG batch opening).
// Replaces, per proof:
//   load_plonk_proof_from_bytes     verifier/src/plonk/converter.rs:121-178  (framing + point/scalar validation)
//   verify_plonk                    verifier/src/plonk/verify.rs:46-317
//   bind_public_data / derive_randomness / batch_invert   verifier/src/plonk/verify.rs:319-396
//   Transcript::compute_challenge   verifier/src/transcript.rs:68-107        (SHA-256 on device)
//   WrappedHashToField              verifier/src/hash_to_field.rs:30-97       (expand_message_xmd)
//   kzg::derive_gamma / fold / fold_proof / batch_verify_multi_points         verifier/src/plonk/kzg.rs:46-190
// VK-constant work is hoisted to vk_load: the SHA-256 state after the VK prefix of the gamma transcript,
// omega^(nPub + cci), the canonical bytes of S1/S2/Qcp hashed by the KZG transcript and the two G2 line tables.
#pragma once
#include "io.cuh"
#include "pairing.cuh"
#include "sha256.cuh"

namespace bn254 {

#define BN_MAX_QCP 4
#define BN_MAX_PLONK_PUBLIC 16
#define BN_MAX_CLAIMED (6 + BN_MAX_QCP)

struct PlonkVkDev {
  uint64_t size;
  int n_public, n_qcp;
  Fr size_inv, generator, coset_shift;       // Montgomery
  Fr w_pow_cci[BN_MAX_QCP];                  // omega^(n_public + cci[i]), Montgomery
  G1Aff s[3], ql, qr, qm, qo, qk, g1;
  G1Aff qcp[BN_MAX_QCP];
  G2Aff g2[2];
  Sha256 gamma_prefix;                       // state after "gamma" | S1 S2 S3 Ql Qr Qm Qo Qk | Qcp..
  uint8_t kzg_vk_bytes[64 * (2 + BN_MAX_QCP)];  // canonical S1 | S2 | Qcp..  (KZG transcript)
  Line g2_lines[2][BN_N_LINES];
};

HD void plonk_vk_prepare(PlonkVkDev& vk) {
  g2_precompute(vk.g2_lines[0], vk.g2[0]);
  g2_precompute(vk.g2_lines[1], vk.g2[1]);
}

// ---- Fr helpers (Montgomery unless said otherwise)
HD Fr fr_one() { return fe_one<FrCfg>(); }
HD Fr fr_mul(const Fr& a, const Fr& b) { return fe_mul(a, b); }
HD Fr fr_add(const Fr& a, const Fr& b) { return fe_add(a, b); }
HD Fr fr_sub(const Fr& a, const Fr& b) { return fe_sub(a, b); }
HD Fr fr_neg(const Fr& a) { return fe_neg(a); }
// 32 big-endian bytes -> value mod r (Fr::from_bytes_be_mod_order), Montgomery
HD Fr fr_from_be_mod_order(const uint8_t* b) {
  Fr t;
  fe_from_be_bytes(t, b);
  fe_reduce_full(t);
  return fe_to_mont(t);
}
// 32 big-endian bytes, must be < r (Fr::from_slice)
HD bool fr_load_be(Fr& out, const uint8_t* b) {
  Fr t;
  bool ok = fe_from_be_bytes(t, b);
  out = fe_to_mont(t);
  return ok;
}
HD void fr_store_be(uint8_t* b, const Fr& a) { fe_to_be_bytes(b, fe_from_mont(a)); }
HDN Fr fr_pow_u64(const Fr& a, uint64_t e) {
  Fr r = fr_one();
  bool started = false;
  for (int i = 63; i >= 0; i--) {
    if (started) r = fr_mul(r, r);
    if ((e >> i) & 1) {
      r = started ? fr_mul(r, a) : a;
      started = true;
    }
  }
  return r;
}
HD uint32_t be32_at(const uint8_t* b) {
  return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
}

// acc += [k] P   (k Montgomery); AffineG1::msm is a plain sum of scalar multiples
HDN void msm_acc(G1Jac& acc, const G1Aff& p, const Fr& k_mont) {
  Fr k = fe_from_mont(k_mont);
  G1Jac t = scalar_mul(p, k.v);
  acc = jac_add(acc, t);
}

// SHA-256 transcript pieces
HD void sha_bytes(Sha256& s, const char* str, int n) {
  for (int i = 0; i < n; i++) sha256_put(s, (uint8_t)str[i]);
}

// RFC 9380 expand_message_xmd(SHA-256, DST = "BSB22-Plonk", 48 bytes) of a 64-byte point, then mod r.
HDN Fr hash_to_field_bsb22(const uint8_t* pt64) {
  const char dst[12] = {'B', 'S', 'B', '2', '2', '-', 'P', 'l', 'o', 'n', 'k', 11};  // DST || len(DST)
  uint8_t b0[32], b1[32], b2[32];
  Sha256 s;
  sha256_init(s);
  for (int i = 0; i < 64; i++) sha256_put(s, 0);
  sha256_update(s, pt64, 64);
  sha256_put(s, 0), sha256_put(s, 48), sha256_put(s, 0);
  sha_bytes(s, dst, 12);
  sha256_final(s, b0);
  sha256_init(s);
  sha256_update(s, b0, 32);
  sha256_put(s, 1);
  sha_bytes(s, dst, 12);
  sha256_final(s, b1);
  sha256_init(s);
  for (int i = 0; i < 32; i++) sha256_put(s, b0[i] ^ b1[i]);
  sha256_put(s, 2);
  sha_bytes(s, dst, 12);
  sha256_final(s, b2);
  // value = BE(b1 | b2[0..16]) = b1 * 2^128 + b2_hi
  Fr hi = fr_from_be_mod_order(b1);
  Fr two128 = fe_zero<FrCfg>();
  two128.v[4] = 1;
  two128 = fe_to_mont(two128);
  Fr lo = fe_zero<FrCfg>();
  for (int i = 0; i < 4; i++) lo.v[3 - i] = be32_at(b2 + 4 * i);
  lo = fe_to_mont(lo);
  return fr_add(fr_mul(hi, two128), lo);
}

struct PlonkDebug {
  uint8_t* g1;      // 4 x 64: lin digest, folded digest, pairing G1 #0, #1   (or null)
  uint8_t* fr;      // 8 x 32: gamma, beta, alpha, zeta, kzg gamma, PI, const_lin, hashed BSB22[0]
  uint8_t* miller;  // 384
  uint8_t* gt;      // 384
};

HD int plonk_verify_one(const PlonkVkDev& vk, const uint8_t* pr, uint32_t len, const uint8_t* inputs_be, int n_inputs,
                        const uint8_t* rnd_be, const PlonkDebug& dbg) {
  // ---- public inputs are bn::Fr in the reference's API: a value >= r cannot be constructed by the caller
  for (int i = 0; i < n_inputs; i++) {
    Fr t;
    if (!fe_from_be_bytes(t, inputs_be + 32 * i)) return BN254V_PANIC_FIELD_NOT_MEMBER;
  }
  // ---- load_plonk_proof_from_bytes (plonk/converter.rs:121-178)
  if (len < 516) return BN254V_PANIC_SHORT_BUFFER;
  G1Aff P8[8];  // L R O Z H0 H1 H2 batchedH
  for (int i = 0; i < 8; i++) {
    int st = load_g1_checked(P8[i], pr + 64 * i);
    if (st != BN254V_OK_TRUE) return st;
  }
  uint32_t ncl = be32_at(pr + 512);
  uint32_t off = 516;
  Fr cl[BN_MAX_CLAIMED];
  for (uint32_t i = 0; i < ncl; i++) {
    if ((uint64_t)off + 32 > len) return BN254V_PANIC_SHORT_BUFFER;
    Fr t;
    if (!fr_load_be(t, pr + off)) return BN254V_PANIC_FIELD_NOT_MEMBER;
    if (i < BN_MAX_CLAIMED) cl[i] = t;
    off += 32;
  }
  if ((uint64_t)off + 100 > len) return BN254V_PANIC_SHORT_BUFFER;
  const uint32_t off_zsh = off;
  G1Aff zs_h;
  {
    int st = load_g1_checked(zs_h, pr + off);
    if (st != BN254V_OK_TRUE) return st;
  }
  Fr zu;
  if (!fr_load_be(zu, pr + off + 64)) return BN254V_PANIC_FIELD_NOT_MEMBER;
  uint32_t nbsb = be32_at(pr + off + 96);
  off += 100;
  const uint32_t off_bsb = off;
  G1Aff bsb[BN_MAX_QCP];
  for (uint32_t i = 0; i < nbsb; i++) {
    if ((uint64_t)off + 64 > len) return BN254V_PANIC_SHORT_BUFFER;
    G1Aff t;
    int st = load_g1_checked(t, pr + off);
    if (st != BN254V_OK_TRUE) return st;
    if (i < BN_MAX_QCP) bsb[i] = t;
    off += 64;
  }
  // ---- verify_plonk shape checks (plonk/verify.rs:52-59)
  if ((int)nbsb != vk.n_qcp || nbsb > BN_MAX_QCP) return BN254V_ERR_BSB22_MISMATCH;
  if (n_inputs != vk.n_public) return BN254V_ERR_INVALID_WITNESS;

  // ---- Fiat-Shamir challenges (plonk/verify.rs:62-95)
  uint8_t dg[32];
  Sha256 sh = vk.gamma_prefix;
  sha256_update(sh, inputs_be, 32u * n_inputs);
  sha256_update(sh, pr, 192);  // L R O
  sha256_final(sh, dg);
  Fr gamma = fr_from_be_mod_order(dg);
  sha256_init(sh);
  sha_bytes(sh, "beta", 4);
  sha256_update(sh, dg, 32);
  sha256_final(sh, dg);
  Fr beta = fr_from_be_mod_order(dg);
  sha256_init(sh);
  sha_bytes(sh, "alpha", 5);
  sha256_update(sh, dg, 32);
  sha256_update(sh, pr + off_bsb, 64 * nbsb);
  sha256_update(sh, pr + 192, 64);  // Z
  sha256_final(sh, dg);
  Fr alpha = fr_from_be_mod_order(dg);
  sha256_init(sh);
  sha_bytes(sh, "zeta", 4);
  sha256_update(sh, dg, 32);
  sha256_update(sh, pr + 256, 192);  // H0 H1 H2
  sha256_final(sh, dg);
  Fr zeta = fr_from_be_mod_order(dg);

  // ---- zeta^n - 1, L_1(zeta), PI(zeta) (plonk/verify.rs:98-163); one shared inversion (Montgomery's trick) for
  //      (zeta-1), the public-input denominators and the BSB22 denominators: the inverses are the same field elements.
  const Fr one = fr_one();
  Fr zeta_n = fr_pow_u64(zeta, vk.size);
  Fr zh_zeta = fr_sub(zeta_n, one);
  Fr zm1 = fr_sub(zeta, one);
  if (fe_is_zero(zm1)) return BN254V_ERR_INVERSE_NOT_FOUND;
  Fr dens[1 + BN_MAX_PLONK_PUBLIC + BN_MAX_QCP], pref[1 + BN_MAX_PLONK_PUBLIC + BN_MAX_QCP];
  int nd = 0;
  dens[nd++] = zm1;
  {
    Fr accw = one;
    for (int i = 0; i < n_inputs; i++) {
      dens[nd++] = fr_sub(zeta, accw);
      accw = fr_mul(accw, vk.generator);
    }
  }
  for (int i = 0; i < vk.n_qcp; i++) {
    Fr d = fr_sub(zeta, vk.w_pow_cci[i]);
    if (fe_is_zero(d)) return BN254V_PANIC_DIV_BY_ZERO;  // Fr `/=` by zero (plonk/verify.rs:157)
    dens[nd++] = d;
  }
  {
    Fr run = one;
    for (int i = 0; i < nd; i++) {  // batch_invert skips zeros (plonk/verify.rs:364-396)
      pref[i] = run;
      if (!fe_is_zero(dens[i])) run = fr_mul(run, dens[i]);
    }
    Fr inv = fe_inv(run);
    for (int i = nd - 1; i >= 0; i--) {
      if (fe_is_zero(dens[i])) continue;
      Fr t = fr_mul(inv, pref[i]);
      inv = fr_mul(inv, dens[i]);
      dens[i] = t;
    }
  }
  Fr lagrange_one = fr_mul(fr_mul(dens[0], zh_zeta), vk.size_inv);
  Fr pi = fe_zero<FrCfg>();
  {
    Fr accw = one;
    for (int i = 0; i < n_inputs; i++) {
      Fr w;
      fr_load_be(w, inputs_be + 32 * i);
      Fr li = fr_mul(fr_mul(fr_mul(fr_mul(zh_zeta, dens[1 + i]), vk.size_inv), accw), w);
      accw = fr_mul(accw, vk.generator);
      pi = fr_add(pi, li);
    }
  }
  Fr hashed0 = fe_zero<FrCfg>();
  for (int i = 0; i < vk.n_qcp; i++) {
    Fr hc = hash_to_field_bsb22(pr + off_bsb + 64 * i);
    if (i == 0) hashed0 = hc;
    Fr lag = fr_mul(fr_mul(fr_mul(zh_zeta, vk.w_pow_cci[i]), dens[1 + n_inputs + i]), vk.size_inv);
    pi = fr_add(pi, fr_mul(lag, hc));
  }

  // ---- linearised polynomial constant term (plonk/verify.rs:166-214)
  if (ncl < 6) return BN254V_PANIC_INDEX_OUT_OF_RANGE;
  const Fr &l = cl[1], &r = cl[2], &o = cl[3], &s1 = cl[4], &s2 = cl[5];
  Fr a2l1 = fr_mul(fr_mul(lagrange_one, alpha), alpha);
  Fr t1 = fr_add(fr_add(fr_mul(beta, s1), gamma), l);
  Fr t2 = fr_add(fr_add(fr_mul(beta, s2), gamma), r);
  Fr const_lin = fr_mul(fr_mul(fr_mul(fr_mul(t1, t2), fr_add(o, gamma)), alpha), zu);
  const_lin = fr_neg(fr_add(fr_sub(const_lin, a2l1), pi));
  if (dbg.fr) {
    fr_store_be(dbg.fr, gamma), fr_store_be(dbg.fr + 32, beta), fr_store_be(dbg.fr + 64, alpha);
    fr_store_be(dbg.fr + 96, zeta), fr_store_be(dbg.fr + 160, pi), fr_store_be(dbg.fr + 192, const_lin);
    fr_store_be(dbg.fr + 224, hashed0);
  }
  if (!fe_eq(const_lin, cl[0])) return BN254V_ERR_OPENING_POLY_MISMATCH;
  // A claimed-value count other than 6 + nQcp reaches fold_proof's length check after the linearisation MSM.
  if ((int)ncl != 6 + vk.n_qcp) return BN254V_ERR_INVALID_NUMBER_OF_DIGESTS;

  // ---- linearised polynomial digest (plonk/verify.rs:218-284)
  Fr s1c = fr_mul(fr_mul(fr_mul(fr_mul(t1, t2), beta), alpha), zu);  // (b s1 + l + g)(b s2 + r + g) b a zu
  const Fr& u = vk.coset_shift;
  Fr bz = fr_mul(beta, zeta);
  Fr buz = fr_mul(bz, u);
  Fr s2c = fr_mul(fr_mul(fr_add(fr_add(bz, gamma), l), fr_add(fr_add(buz, gamma), r)),
                  fr_add(fr_add(fr_mul(buz, u), gamma), o));
  s2c = fr_neg(fr_mul(s2c, alpha));
  Fr coeff_z = fr_add(a2l1, s2c);
  Fr rl = fr_mul(l, r);
  Fr zn2 = fr_mul(fr_mul(zeta_n, zeta), zeta);  // zeta^(n+2)
  Fr zn2zh = fr_neg(fr_mul(zn2, zh_zeta));
  Fr zn2sqzh = fr_neg(fr_mul(fr_mul(zn2, zn2), zh_zeta));
  Fr zh = fr_neg(zh_zeta);

  G1Jac acc = jac_identity<Fp>();
  for (int i = 0; i < vk.n_qcp; i++) msm_acc(acc, bsb[i], cl[6 + i]);
  msm_acc(acc, vk.ql, l);
  msm_acc(acc, vk.qr, r);
  msm_acc(acc, vk.qm, rl);
  msm_acc(acc, vk.qo, o);
  acc = jac_add_mixed(acc, vk.qk);  // scalar 1
  msm_acc(acc, vk.s[2], s1c);
  msm_acc(acc, P8[3], coeff_z);
  msm_acc(acc, P8[4], zh);
  msm_acc(acc, P8[5], zn2zh);
  msm_acc(acc, P8[6], zn2sqzh);
  G1Aff lin;
  if (!to_affine(lin, acc)) return BN254V_PANIC_IDENTITY;
  uint8_t lin_bytes[64];
  store_g1(lin_bytes, lin);
  if (dbg.g1) memcpy(dbg.g1, lin_bytes, 64);

  // ---- kzg::fold_proof (plonk/kzg.rs:87-126): gamma_kzg = H("gamma" | zeta | digests | claimed | zu)
  sha256_init(sh);
  sha_bytes(sh, "gamma", 5);
  uint8_t tmp32[32];
  fr_store_be(tmp32, zeta);
  sha256_update(sh, tmp32, 32);
  sha256_update(sh, lin_bytes, 64);
  sha256_update(sh, pr, 192);  // L R O
  sha256_update(sh, vk.kzg_vk_bytes, 64 * (2 + vk.n_qcp));
  sha256_update(sh, pr + 516, 32 * ncl);
  sha256_update(sh, pr + off_zsh + 64, 32);  // zu
  sha256_final(sh, dg);
  Fr kg = fr_from_be_mod_order(dg);
  if (dbg.fr) fr_store_be(dbg.fr + 128, kg);
  // folded digest = sum gamma^i D_i, folded eval = sum gamma^i claimed_i;  D = lin, L, R, O, S1, S2, Qcp..
  Fr gi = kg;
  Fr folded_eval = cl[0];
  G1Jac fd = to_jac(lin);
  for (int i = 1; i < 6 + vk.n_qcp; i++) {
    const G1Aff& D = i <= 3 ? P8[i - 1] : (i <= 5 ? vk.s[i - 4] : vk.qcp[i - 6]);
    msm_acc(fd, D, gi);
    folded_eval = fr_add(folded_eval, fr_mul(cl[i], gi));
    gi = fr_mul(gi, kg);
  }
  if (is_identity(fd)) return BN254V_PANIC_IDENTITY;
  if (dbg.g1) {
    G1Aff t;
    to_affine(t, fd);
    store_g1(dbg.g1 + 64, t);
  }

  // ---- kzg::batch_verify_multi_points (plonk/kzg.rs:128-190): digests [folded, Z], proofs [(batchedH, folded_eval),
  //      (zsH, zu)], points [zeta, omega zeta], random numbers [1, rnd]
  Fr rnd = fr_from_be_mod_order(rnd_be);
  G1Jac fq = to_jac(P8[7]);  // folded quotients = batchedH + rnd zsH
  msm_acc(fq, zs_h, rnd);
  if (is_identity(fq)) return BN254V_PANIC_IDENTITY;
  G1Jac fdg = fd;  // folded digests = folded + rnd Z
  msm_acc(fdg, P8[3], rnd);
  if (is_identity(fdg)) return BN254V_PANIC_IDENTITY;
  Fr fev = fr_add(folded_eval, fr_mul(zu, rnd));
  {
    Fr k = fe_from_mont(fev);
    G1Jac fec = scalar_mul(vk.g1, k.v);  // vk.g1 * folded_evals, .into() AffineG1
    if (is_identity(fec)) return BN254V_PANIC_IDENTITY;
    fec.y = neg(fec.y);
    fdg = jac_add(fdg, fec);
    if (is_identity(fdg)) return BN254V_PANIC_IDENTITY;
  }
  {
    Fr shifted = fr_mul(zeta, vk.generator);
    G1Jac fpq = jac_identity<Fp>();
    msm_acc(fpq, P8[7], zeta);
    msm_acc(fpq, zs_h, fr_mul(rnd, shifted));
    if (is_identity(fpq)) return BN254V_PANIC_IDENTITY;
    fdg = jac_add(fdg, fpq);
    if (is_identity(fdg)) return BN254V_PANIC_IDENTITY;
  }
  fq.y = neg(fq.y);
  G1Aff pf[2];
  to_affine(pf[0], fdg);
  to_affine(pf[1], fq);
  if (dbg.g1) {
    store_g1(dbg.g1 + 128, pf[0]);
    store_g1(dbg.g1 + 192, pf[1]);
  }
  const Line* tabs[2] = {vk.g2_lines[0], vk.g2_lines[1]};
  Fp12 f;
  miller_loop<0, 2>(f, nullptr, nullptr, pf, tabs);
  if (dbg.miller) fp12_to_bytes(dbg.miller, f);
  final_exponentiation(f, f);
  if (dbg.gt) fp12_to_bytes(dbg.gt, f);
  return eq(f, fp12_one()) ? BN254V_OK_TRUE : BN254V_ERR_PAIRING_CHECK_FAILED;
}

}  // namespace bn254
