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
#include "groth16.cuh"
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
  const G1Aff* fixed_tables;                 // [7 + n_qcp + 1][32][255] window tables of the VK bases, or null
  Line g2_lines[2][BN_N_LINES];
  LinePairKF g2_pairs[BN_N_LINES];           // product coefficients of the two KZG G2 line tables
};

// VK-constant MSM bases, in table order: Ql Qr Qm Qo S3 S1 S2 Qcp.. g1(KZG)
#define BN_PLONK_N_FIXED(nq) (8 + (nq))
HD const G1Aff& plonk_fixed_base(const PlonkVkDev& vk, int idx) {
  switch (idx) {
    case 0: return vk.ql;
    case 1: return vk.qr;
    case 2: return vk.qm;
    case 3: return vk.qo;
    case 4: return vk.s[2];
    case 5: return vk.s[0];
    case 6: return vk.s[1];
  }
  if (idx < 7 + vk.n_qcp) return vk.qcp[idx - 7];
  return vk.g1;
}

HD void plonk_vk_prepare(PlonkVkDev& vk) {
  g2_precompute(vk.g2_lines[0], vk.g2[0]);
  g2_precompute(vk.g2_lines[1], vk.g2[1]);
  line_pair_table(vk.g2_pairs, vk.g2_lines[0], vk.g2_lines[1]);
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

// ---- GLV: the curve y^2 = x^3 + 3 has the endomorphism phi(x, y) = (beta x, y) = [lambda](x, y).  A scalar splits as
// k = k1 + k2 lambda (mod r) with |k1|, |k2| < 2^128, so [k]P = [k1]P + [k2]phi(P) needs half the doublings.
// c1 = (k g1) >> 256, c2 = (k g2) >> 256 (rounded-down lattice coordinates; any integers give an exact identity, these
// keep k1, k2 short), k1 = k - c1 a1 - c2 a2, k2 = -c1 b1 - c2 b2, all modulo 2^256 in two's complement.
HD void glv_mul_hi(uint32_t* c /*5*/, const uint32_t* k /*8*/, const uint32_t* g, int g_limbs) {
  uint32_t t[14];
  for (int i = 0; i < 14; i++) t[i] = 0;
  for (int j = 0; j < g_limbs; j++) {
    uint64_t carry = 0;
    const uint32_t gj = g[j];
    for (int i = 0; i < 8; i++) {
      uint64_t x = (uint64_t)k[i] * gj + t[i + j] + carry;
      t[i + j] = (uint32_t)x;
      carry = x >> 32;
    }
    t[8 + j] = (uint32_t)carry;
  }
  for (int i = 0; i < 5; i++) c[i] = t[8 + i];
}
// acc (8 limbs) += c (5 limbs) * m (8 limbs)  mod 2^256
HD void glv_mac_lo(uint32_t* acc, const uint32_t* c, const uint32_t* m) {
  for (int j = 0; j < 5; j++) {
    uint64_t carry = 0;
    for (int i = 0; i + j < 8; i++) {
      uint64_t x = (uint64_t)m[i] * c[j] + acc[i + j] + carry;
      acc[i + j] = (uint32_t)x;
      carry = x >> 32;
    }
  }
}
HD bool glv_abs(uint32_t* v) {  // two's complement -> magnitude; returns the sign
  const bool neg = (v[7] >> 31) != 0;
  if (neg) {
    uint64_t carry = 1;
    for (int i = 0; i < 8; i++) {
      uint64_t x = (uint64_t)(~v[i]) + carry;
      v[i] = (uint32_t)x;
      carry = x >> 32;
    }
  }
  return neg;
}
#define BN_GLV_CONST(name, fn)          \
  uint32_t name[8];                     \
  _Pragma("unroll") for (int _i = 0; _i < 8; _i++) name[_i] = fn(_i);

// [k] P for a proof-supplied point: GLV split, then SIGNED 4-bit windows for both halves over one 8-entry AFFINE
// table (phi of a table entry is one multiplication by beta) with shared doublings: 132 doublings + at most 66 mixed
// additions instead of 252 + 64; uniform control flow across a warp apart from zero digits.  k: plain 8 x u32 LE (< r).
// Same group element as the reference's AffineG1 * Fr; AffineG1::msm is a plain sum of such terms.
// k -> (|k1|, |k2|, signs)
HD void glv_split(uint32_t* k1, uint32_t* k2, bool& n1, bool& n2, const uint32_t* k) {
  uint32_t c1[5], c2[5];
  {
    BN_GLV_CONST(g1, K::glv_g1)
    BN_GLV_CONST(g2, K::glv_g2)
    glv_mul_hi(c1, k, g1, 3);
    glv_mul_hi(c2, k, g2, 5);
  }
  for (int i = 0; i < 8; i++) k1[i] = k[i], k2[i] = 0;
  {
    BN_GLV_CONST(na1, K::glv_neg_a1)
    BN_GLV_CONST(na2, K::glv_neg_a2)
    BN_GLV_CONST(nb1, K::glv_neg_b1)
    BN_GLV_CONST(nb2, K::glv_neg_b2)
    glv_mac_lo(k1, c1, na1);
    glv_mac_lo(k1, c2, na2);
    glv_mac_lo(k2, c1, nb1);
    glv_mac_lo(k2, c2, nb2);
  }
  n1 = glv_abs(k1), n2 = glv_abs(k2);
}
// Signed digits without carries: v + 0x888..8 (33 nibbles of 8) has nibbles n_w with sum (n_w - 8) 16^w = v, so the
// digit of window w is n_w - 8 in [-8, 7] and can be read most significant window first.  Needs v < 7 * 2^128 (the GLV
// halves are < 2^128: tools/gen_constants.py glv(), tests/test_hostsim.py).
HD void w4_offset(uint32_t* v) {
  v[0] = cc::add_cc(v[0], 0x88888888u);
  v[1] = cc::addc_cc(v[1], 0x88888888u);
  v[2] = cc::addc_cc(v[2], 0x88888888u);
  v[3] = cc::addc_cc(v[3], 0x88888888u);
  v[4] = cc::addc(v[4], 0x8u);
}
HD int w4_digit(const uint32_t* v, int w) { return (int)((v[w >> 3] >> (4 * (w & 7))) & 15) - 8; }
// tab[8 j + d - 1] = d P_j, d = 1..8, in AFFINE coordinates for each of n points, with one inversion for all of them
// (Montgomery's trick over the 7 n Jacobian z coordinates).  The points are on the curve (validated in stage A; VK
// points by the VK loader), hence of order r: no multiple below r is the identity and no z is zero.  An affine entry
// makes every window addition a mixed one (11 multiplications instead of 16) and takes 64 bytes instead of 96; the
// affine coordinates of a multiple are unique, so the sums are the same group elements as with Jacobian entries.
#ifndef BN_MSM_MAX
#define BN_MSM_MAX 5
#endif
HD void g1_w4_tables(G1Aff* tab, const G1Aff* p, int n) {
  Fp z[7 * BN_MSM_MAX], c[7 * BN_MSM_MAX];
  Fp run = fe_one<FpCfg>();
  for (int j = 0; j < n; j++) {
    tab[8 * j] = p[j];
    G1Jac a = jac_double(to_jac(p[j]));
    for (int d = 0; d < 7; d++) {
      if (d) a = jac_add_mixed(a, p[j]);
      tab[8 * j + 1 + d] = G1Aff{a.x, a.y};  // Jacobian X, Y until the pass below
      z[7 * j + d] = a.z;
      c[7 * j + d] = run;  // product of all earlier z
      run = fe_mul(run, a.z);
    }
  }
  Fp iv = fe_inv(run);
  for (int i = 7 * n - 1; i >= 0; i--) {
    const Fp zi = fe_mul(iv, c[i]);
    iv = fe_mul(iv, z[i]);
    const Fp zi2 = fe_sqr_short(zi);
    G1Aff& e = tab[8 * (i / 7) + 1 + (i % 7)];
    e.x = fe_mul(e.x, zi2);
    e.y = fe_mul(e.y, fe_mul(zi2, zi));
  }
}
// acc += [digit of k1 at window w] P + [digit of k2 at window w] phi(P); k1, k2 offset by w4_offset
HD void g1_w4_add_digits(G1Jac& acc, const G1Aff* tab, const uint32_t* k1, const uint32_t* k2, bool n1, bool n2,
                         const Fp& beta, int w) {
  const int d1 = w4_digit(k1, w), d2 = w4_digit(k2, w);
  if (d1) {
    G1Aff t = tab[(d1 < 0 ? -d1 : d1) - 1];
    if (n1 != (d1 < 0)) t.y = fe_neg(t.y);
    acc = jac_add_mixed(acc, t);
  }
  if (d2) {
    G1Aff t = tab[(d2 < 0 ? -d2 : d2) - 1];
    t.x = fe_mul(t.x, beta);
    if (n2 != (d2 < 0)) t.y = fe_neg(t.y);
    acc = jac_add_mixed(acc, t);
  }
}
HDN G1Jac g1_mul_w4(const G1Aff& p, const uint32_t* k) {
  uint32_t k1[8], k2[8];
  bool n1, n2;
  glv_split(k1, k2, n1, n2, k);
  w4_offset(k1), w4_offset(k2);
  Fp beta;
  BN_LOAD_FP(beta, K::glv_beta, 0);
  G1Aff tab[8];
  g1_w4_tables(tab, &p, 1);
  G1Jac acc = jac_identity<Fp>();
  for (int w = 32; w >= 0; w--) {  // 33 windows: the offset form of a 128-bit value carries into the 33rd
    if (w != 32)
      for (int j = 0; j < 4; j++) acc = jac_double(acc);
    g1_w4_add_digits(acc, tab, k1, k2, n1, n2, beta, w);
  }
  return acc;
}
// sum_j [k_j] P_j for up to BN_MSM_MAX proof-supplied points, evaluated jointly (Straus): the 132 doublings are paid once
// for the group instead of once per point, and the tables of the group share one inversion.  Same group element as the
// sum of the separate products (AffineG1::msm sums its terms; only the total is converted to affine coordinates).
HDN G1Jac g1_msm_w4(const G1Aff* p, const uint32_t* const* k, int n) {
  uint32_t k1[BN_MSM_MAX][8], k2[BN_MSM_MAX][8];
  bool n1[BN_MSM_MAX], n2[BN_MSM_MAX];
  G1Aff tab[BN_MSM_MAX * 8];
  for (int j = 0; j < n; j++) {
    glv_split(k1[j], k2[j], n1[j], n2[j], k[j]);
    w4_offset(k1[j]), w4_offset(k2[j]);
  }
  g1_w4_tables(tab, p, n);
  Fp beta;
  BN_LOAD_FP(beta, K::glv_beta, 0);
  G1Jac acc = jac_identity<Fp>();
  for (int w = 32; w >= 0; w--) {
    if (w != 32)
      for (int j = 0; j < 4; j++) acc = jac_double(acc);
    for (int j = 0; j < n; j++) g1_w4_add_digits(acc, tab + 8 * j, k1[j], k2[j], n1[j], n2[j], beta, w);
  }
  return acc;
}
// [k] B for a VK-constant base with a window table T[w][d-1] = d * 2^(8w) * B (see groth16.cuh)
HDN G1Jac g1_mul_fixed(const G1Aff* tab, const uint32_t* k) {
  G1Jac acc = jac_identity<Fp>();
  for (int w = 0; w < BN_IC_WINDOWS; w++) {
    uint32_t d = (k[w >> 2] >> (8 * (w & 3))) & 0xff;
    if (d) acc = jac_add_mixed(acc, tab[w * BN_IC_ENTRIES + (d - 1)]);
  }
  return acc;
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

// ---------------------------------------------------------------------------------------------------------------
// The verification is cut into stages so that the MSM terms -- where a PlonK verification spends its time -- can
// run one term per thread (k_plonk_terms) while the sequential parts run one proof per thread:
//   stage A  parse + validate, Fiat-Shamir challenges, Fr pipeline, const_lin check    -> scalars of the 10+nQcp
//            terms of the linearised-polynomial MSM                                   (early rejects stop here)
//   terms 0  one scalar multiplication per (proof, term)
//   stage C  sum -> linearised digest (affine, hashed), KZG gamma, folded evaluation    -> scalars of the 5+nQcp
//            fold terms and the 5 terms of batch_verify_multi_points
//   terms 1  one scalar multiplication per (proof, term)
//   stage E  sums with the reference's identity checks, 2-pair Miller loop, final exponentiation, verdict
// plonk_verify_one() runs the same stage functions back to back in one thread (host simulation, tiny batches).
// ---------------------------------------------------------------------------------------------------------------
#define BN_PLONK_MAX_T (BN_MAX_QCP + 10)

struct PlonkWork {
  uint32_t sc[BN_PLONK_MAX_T][8];  // plain scalars of the current stage's terms
  G1Jac part[BN_PLONK_MAX_T];      // term results of the current stage
  Fr zeta;                         // Montgomery
  G1Aff lin;                       // linearised polynomial digest
  G1Aff pair[2];                   // G1 inputs of the final pairing check (stage D -> three-lane stage E)
};
HD int plonk_n_terms(const PlonkVkDev& vk, int stage) { return vk.n_qcp + 10; }  // both stages: nQcp + 10

HD void plonk_put_scalar(PlonkWork& w, int t, const Fr& k_mont) {
  Fr k = fe_from_mont(k_mont);
#pragma unroll
  for (int i = 0; i < 8; i++) w.sc[t][i] = k.v[i];
}

HD int plonk_stage_a(PlonkWork& w, const PlonkVkDev& vk, const uint8_t* pr, uint32_t len, const uint8_t* inputs_be,
                     int n_inputs, const PlonkDebug& dbg) {
  // ---- public inputs are bn::Fr in the reference's API: a value >= r cannot be constructed by the caller
  for (int i = 0; i < n_inputs; i++) {
    Fr t;
    if (!fe_from_be_bytes(t, inputs_be + 32 * i)) return BN254V_PANIC_FIELD_NOT_MEMBER;
  }
  // ---- load_plonk_proof_from_bytes (plonk/converter.rs:121-178)
  if (len < 516) return BN254V_PANIC_SHORT_BUFFER;
  for (int i = 0; i < 8; i++) {  // L R O Z H0 H1 H2 batchedH
    G1Aff t;
    int st = load_g1_checked(t, pr + 64 * i);
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
  {
    G1Aff t;
    int st = load_g1_checked(t, pr + off);  // z-shifted opening proof H
    if (st != BN254V_OK_TRUE) return st;
  }
  Fr zu;
  if (!fr_load_be(zu, pr + off + 64)) return BN254V_PANIC_FIELD_NOT_MEMBER;
  uint32_t nbsb = be32_at(pr + off + 96);
  off += 100;
  const uint32_t off_bsb = off;
  for (uint32_t i = 0; i < nbsb; i++) {
    if ((uint64_t)off + 64 > len) return BN254V_PANIC_SHORT_BUFFER;
    G1Aff t;
    int st = load_g1_checked(t, pr + off);
    if (st != BN254V_OK_TRUE) return st;
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
  w.zeta = zeta;

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
      Fr x;
      fr_load_be(x, inputs_be + 32 * i);
      Fr li = fr_mul(fr_mul(fr_mul(fr_mul(zh_zeta, dens[1 + i]), vk.size_inv), accw), x);
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

  // ---- scalars of the linearised polynomial digest (plonk/verify.rs:218-284)
  Fr s1c = fr_mul(fr_mul(fr_mul(fr_mul(t1, t2), beta), alpha), zu);  // (b s1 + l + g)(b s2 + r + g) b a zu
  const Fr& u = vk.coset_shift;
  Fr bz = fr_mul(beta, zeta);
  Fr buz = fr_mul(bz, u);
  Fr s2c = fr_mul(fr_mul(fr_add(fr_add(bz, gamma), l), fr_add(fr_add(buz, gamma), r)),
                  fr_add(fr_add(fr_mul(buz, u), gamma), o));
  s2c = fr_neg(fr_mul(s2c, alpha));
  Fr zn2 = fr_mul(fr_mul(zeta_n, zeta), zeta);  // zeta^(n+2)
  // term order: BSB22[i] * claimed[6+i] | Ql*l Qr*r Qm*rl Qo*o Qk*1 S3*s1c | Z*coeff_z H0*zh H1*zn2zh H2*zn2sqzh
  int t = 0;
  for (int i = 0; i < vk.n_qcp; i++) plonk_put_scalar(w, t++, cl[6 + i]);
  plonk_put_scalar(w, t++, l);
  plonk_put_scalar(w, t++, r);
  plonk_put_scalar(w, t++, fr_mul(l, r));
  plonk_put_scalar(w, t++, o);
  plonk_put_scalar(w, t++, one);
  plonk_put_scalar(w, t++, s1c);
  plonk_put_scalar(w, t++, fr_add(a2l1, s2c));
  plonk_put_scalar(w, t++, fr_neg(zh_zeta));
  plonk_put_scalar(w, t++, fr_neg(fr_mul(zn2, zh_zeta)));
  plonk_put_scalar(w, t++, fr_neg(fr_mul(fr_mul(zn2, zn2), zh_zeta)));
  return BN254V_OK_TRUE;
}

// offsets of the tail of a well-formed proof (6 + nQcp claimed values; survivors of stage A only)
HD uint32_t plonk_off_zsh(const PlonkVkDev& vk) { return 516 + 32 * (6 + vk.n_qcp); }
HD uint32_t plonk_off_bsb(const PlonkVkDev& vk) { return plonk_off_zsh(vk) + 100; }

HD G1Jac plonk_fixed_or_var(const PlonkVkDev& vk, int base, const uint32_t* k) {
  if (vk.fixed_tables) return g1_mul_fixed(vk.fixed_tables + (size_t)base * BN_IC_WINDOWS * BN_IC_ENTRIES, k);
  return g1_mul_w4(plonk_fixed_base(vk, base), k);
}
HD G1Jac plonk_var(const uint8_t* pt_bytes, const uint32_t* k) {
  G1Aff p;
  load_g1_unchecked(p, pt_bytes);  // validated in stage A
  return g1_mul_w4(p, k);
}

// One scalar multiplication: term `t` of stage 0 (linearised digest) or stage 1 (fold + batch opening).
HD void plonk_term(PlonkWork& w, const PlonkVkDev& vk, const uint8_t* pr, int stage, int t) {
  const uint32_t* k = w.sc[t];
  const int nq = vk.n_qcp;
  G1Jac res;
  if (stage == 0) {
    if (t < nq) res = plonk_var(pr + plonk_off_bsb(vk) + 64 * t, k);
    else {
      int j = t - nq;
      if (j < 4) res = plonk_fixed_or_var(vk, j, k);           // Ql Qr Qm Qo
      else if (j == 4) res = to_jac(vk.qk);                    // Qk * 1
      else if (j == 5) res = plonk_fixed_or_var(vk, 4, k);     // S3
      else res = plonk_var(pr + 192 + 64 * (j - 6), k);        // Z H0 H1 H2
    }
  } else {
    if (t < 5 + nq) {
      int i = t + 1;  // digest index: lin(0) L R O S1 S2 Qcp..
      if (i <= 3) res = plonk_var(pr + 64 * (i - 1), k);
      else res = plonk_fixed_or_var(vk, 5 + (i - 4), k);       // S1 S2 Qcp..
    } else {
      int j = t - (5 + nq);
      if (j == 0) res = plonk_var(pr + plonk_off_zsh(vk), k);        // zsH * rnd
      else if (j == 1) res = plonk_var(pr + 192, k);                 // Z * rnd
      else if (j == 2) res = plonk_fixed_or_var(vk, 7 + nq, k);      // g1 * folded evals
      else if (j == 3) res = plonk_var(pr + 448, k);                 // batchedH * zeta
      else res = plonk_var(pr + plonk_off_zsh(vk), k);               // zsH * (rnd omega zeta)
    }
  }
  w.part[t] = res;
}

// ---- joint evaluation (large batches): the terms of a stage that the verifier only ever sums are evaluated in groups
// with shared doublings (g1_msm_w4, at most BN_MSM_MAX proof points per group) and the VK terms of a sum in one thread.
//   stage 0 (every term goes into the linearised digest): groups of the nq + 4 proof points BSB22.. Z H0 H1 H2, then one
//           item for Ql Qr Qm Qo Qk S3;
//   stage 1: 0 = L R O, 1 = S1 S2 Qcp.. (both into the folded digest), 2 = zeta batchedH + (rnd omega zeta) zsH,
//           3 = rnd zsH, 4 = rnd Z, 5 = (folded evaluations) g1.
// The per-term form above stays for small batches, where the number of threads matters more than their work.
HD int plonk_n_items(int nq, int stage, bool joint) {
  if (!joint) return nq + 10;
  return stage == 0 ? (nq + 4 + BN_MSM_MAX - 1) / BN_MSM_MAX + 1 : 6;
}
HD G1Jac plonk_fixed_sum(const PlonkVkDev& vk, G1Jac acc, const int* bases, const uint32_t* const* ks, int n) {
  for (int j = 0; j < n; j++) {
    if (vk.fixed_tables) {
      const G1Aff* tab = vk.fixed_tables + (size_t)bases[j] * BN_IC_WINDOWS * BN_IC_ENTRIES;
      for (int w = 0; w < BN_IC_WINDOWS; w++) {
        const uint32_t d = (ks[j][w >> 2] >> (8 * (w & 3))) & 0xff;
        if (d) acc = jac_add_mixed(acc, tab[w * BN_IC_ENTRIES + (d - 1)]);
      }
    } else {
      acc = jac_add(acc, g1_mul_w4(plonk_fixed_base(vk, bases[j]), ks[j]));
    }
  }
  return acc;
}
HD void plonk_item_joint(PlonkWork& w, const PlonkVkDev& vk, const uint8_t* pr, int stage, int item) {
  const int nq = vk.n_qcp, b = 5 + nq;
  G1Aff pts[BN_MSM_MAX];
  const uint32_t* ks[BN_MAX_QCP + 6];
  int bases[BN_MAX_QCP + 6];
  G1Jac res;
  if (stage == 0) {
    const int nv = nq + 4, ng = (nv + BN_MSM_MAX - 1) / BN_MSM_MAX;
    if (item < ng) {
      int n = 0;
      for (int v = item * BN_MSM_MAX; v < nv && n < BN_MSM_MAX; v++, n++) {
        // proof point v: BSB22 commitment v (term v) or Z H0 H1 H2 (terms nq + 6 ..)
        const int t = v < nq ? v : nq + 6 + (v - nq);
        load_g1_unchecked(pts[n], v < nq ? pr + plonk_off_bsb(vk) + 64 * v : pr + 192 + 64 * (v - nq));
        ks[n] = w.sc[t];
      }
      res = g1_msm_w4(pts, ks, n);
    } else {
      for (int j = 0; j < 4; j++) bases[j] = j, ks[j] = w.sc[nq + j];  // Ql Qr Qm Qo
      bases[4] = 4, ks[4] = w.sc[nq + 5];                              // S3
      res = plonk_fixed_sum(vk, to_jac(vk.qk), bases, ks, 5);           // + Qk * 1
    }
  } else {
    if (item == 0) {
      for (int j = 0; j < 3; j++) load_g1_unchecked(pts[j], pr + 64 * j), ks[j] = w.sc[j];  // L R O
      res = g1_msm_w4(pts, ks, 3);
    } else if (item == 1) {
      for (int j = 0; j < 2 + nq; j++) bases[j] = 5 + j, ks[j] = w.sc[3 + j];  // S1 S2 Qcp..
      res = plonk_fixed_sum(vk, jac_identity<Fp>(), bases, ks, 2 + nq);
    } else if (item == 2) {
      load_g1_unchecked(pts[0], pr + 448), ks[0] = w.sc[b + 3];                 // batchedH * zeta
      load_g1_unchecked(pts[1], pr + plonk_off_zsh(vk)), ks[1] = w.sc[b + 4];   // zsH * (rnd omega zeta)
      res = g1_msm_w4(pts, ks, 2);
    } else if (item == 3) {
      res = plonk_var(pr + plonk_off_zsh(vk), w.sc[b + 0]);                      // zsH * rnd
    } else if (item == 4) {
      res = plonk_var(pr + 192, w.sc[b + 1]);                                    // Z * rnd
    } else {
      res = plonk_fixed_or_var(vk, 7 + nq, w.sc[b + 2]);                         // g1 * folded evaluations
    }
  }
  w.part[item] = res;
}

// Launch order of the terms of a stage: position -> term index, variable-base terms first (see k_plonk_terms).
//   stage 0 terms: [0, nq) BSB22 (variable) | nq + {0..3} Ql Qr Qm Qo, +4 Qk, +5 S3 (fixed) | nq + {6..9} Z H0 H1 H2 (variable)
//   stage 1 terms: 0..2 L R O (variable) | 3 .. 4 + nq S1 S2 Qcp (fixed) | b + {0, 1, 3, 4} (variable), b + 2 g1 (fixed); b = 5 + nq
HD int plonk_term_order(int nq, int stage, int pos) {
  if (stage == 0) {
    if (pos < nq) return pos;                   // BSB22
    if (pos < nq + 4) return nq + 6 + (pos - nq);  // Z H0 H1 H2
    return nq + (pos - nq - 4);                 // the six VK terms
  }
  const int b = 5 + nq;
  if (pos < 3) return pos;                      // L R O
  if (pos < 7) {
    const int j = pos - 3;                      // zsH rnd, Z rnd, batchedH zeta, zsH rnd omega zeta
    return b + (j < 2 ? j : j + 1);
  }
  if (pos < 7 + 2 + nq) return 3 + (pos - 7);   // S1 S2 Qcp..
  return b + 2;                                 // g1
}

HD int plonk_stage_c(PlonkWork& w, const PlonkVkDev& vk, const uint8_t* pr, const uint8_t* rnd_be,
                     const PlonkDebug& dbg, bool joint = false) {
  const int nq = vk.n_qcp;
  G1Jac acc = w.part[0];
  for (int t = 1; t < plonk_n_items(nq, 0, joint); t++) acc = jac_add(acc, w.part[t]);
  if (!to_affine(w.lin, acc)) return BN254V_PANIC_IDENTITY;
  uint8_t lin_bytes[64];
  store_g1(lin_bytes, w.lin);
  if (dbg.g1) memcpy(dbg.g1, lin_bytes, 64);
  // ---- kzg::fold_proof (plonk/kzg.rs:87-126): gamma_kzg = H("gamma" | zeta | digests | claimed | zu)
  const uint32_t ncl = 6 + nq, off_zsh = plonk_off_zsh(vk);
  Sha256 sh;
  uint8_t dg[32], tmp32[32];
  sha256_init(sh);
  sha_bytes(sh, "gamma", 5);
  fr_store_be(tmp32, w.zeta);
  sha256_update(sh, tmp32, 32);
  sha256_update(sh, lin_bytes, 64);
  sha256_update(sh, pr, 192);  // L R O
  sha256_update(sh, vk.kzg_vk_bytes, 64 * (2 + nq));
  sha256_update(sh, pr + 516, 32 * ncl);
  sha256_update(sh, pr + off_zsh + 64, 32);  // zu
  sha256_final(sh, dg);
  Fr kg = fr_from_be_mod_order(dg);
  if (dbg.fr) fr_store_be(dbg.fr + 128, kg);
  // fold scalars gamma^i and folded evaluation sum gamma^i claimed_i
  Fr gi = kg, folded_eval, c;
  fr_load_be(folded_eval, pr + 516);
  for (int i = 1; i < 6 + nq; i++) {
    plonk_put_scalar(w, i - 1, gi);
    fr_load_be(c, pr + 516 + 32 * i);
    folded_eval = fr_add(folded_eval, fr_mul(c, gi));
    gi = fr_mul(gi, kg);
  }
  // batch_verify_multi_points scalars: random numbers [1, rnd], points [zeta, omega zeta]
  Fr rnd = fr_from_be_mod_order(rnd_be), zu;
  fr_load_be(zu, pr + off_zsh + 64);
  const int b = 5 + nq;
  plonk_put_scalar(w, b + 0, rnd);
  plonk_put_scalar(w, b + 1, rnd);
  plonk_put_scalar(w, b + 2, fr_add(folded_eval, fr_mul(zu, rnd)));
  plonk_put_scalar(w, b + 3, w.zeta);
  plonk_put_scalar(w, b + 4, fr_mul(rnd, fr_mul(w.zeta, vk.generator)));
  return BN254V_OK_TRUE;
}

// The G1 side of kzg::fold and kzg::batch_verify_multi_points (plonk/kzg.rs:74-85, 128-178): sums of the term results
// with the reference's AffineG1 conversions (an identity intermediate panics there) -> the two G1 inputs of the pairing.
HD int plonk_stage_d(G1Aff* pf, PlonkWork& w, const PlonkVkDev& vk, const uint8_t* pr, const PlonkDebug& dbg,
                     bool joint = false) {
  const int nq = vk.n_qcp;
  // where the stage-1 results are: per-term form b terms of the fold, then rnd zsH, rnd Z, evals g1, and the two terms
  // of the points-quotients sum; joint form items 0, 1 (fold), 3, 4, 5, and 2 (already summed)
  const int n_fold = joint ? 2 : 5 + nq, b = joint ? 3 : 5 + nq;
  int st = BN254V_OK_TRUE;
  // folded digest = lin + sum gamma^i D_i (kzg::fold)
  G1Jac fd = to_jac(w.lin);
  for (int t = 0; t < n_fold; t++) fd = jac_add(fd, w.part[t]);
  if (is_identity(fd)) st = BN254V_PANIC_IDENTITY;
  if (st == BN254V_OK_TRUE && dbg.g1) {
    G1Aff t;
    to_affine(t, fd);
    store_g1(dbg.g1 + 64, t);
  }
  G1Aff bh;
  load_g1_unchecked(bh, pr + 448);
  G1Jac fq = jac_add_mixed(w.part[b + 0], bh);  // folded quotients = batchedH + rnd zsH
  if (is_identity(fq)) st = BN254V_PANIC_IDENTITY;
  G1Jac fdg = jac_add(fd, w.part[b + 1]);       // folded digests = folded + rnd Z
  if (is_identity(fdg)) st = BN254V_PANIC_IDENTITY;
  G1Jac fec = w.part[b + 2];                    // vk.g1 * folded evals
  if (is_identity(fec)) st = BN254V_PANIC_IDENTITY;
  fec.y = neg(fec.y);
  fdg = jac_add(fdg, fec);
  if (is_identity(fdg)) st = BN254V_PANIC_IDENTITY;
  G1Jac fpq = joint ? w.part[2] : jac_add(w.part[b + 3], w.part[b + 4]);  // zeta batchedH + rnd omega zeta zsH
  if (is_identity(fpq)) st = BN254V_PANIC_IDENTITY;
  fdg = jac_add(fdg, fpq);
  if (is_identity(fdg)) st = BN254V_PANIC_IDENTITY;
  fq.y = neg(fq.y);
  if (st == BN254V_OK_TRUE) {
    to_affine2(pf[0], fdg, pf[1], fq);  // both checked non-identity above; one shared inversion
    if (dbg.g1) {
      store_g1(dbg.g1 + 128, pf[0]);
      store_g1(dbg.g1 + 192, pf[1]);
    }
  }
  return st;
}

// `live == false`: a spare thread, or a proof that stage C already ended -- the pairing below contains block-wide
// phase barriers, so the thread still runs it (on VK points) and its result is discarded.  For the same reason the
// identity panics of the reference's AffineG1 conversions are recorded and the pairing runs on substitute points.
HD int plonk_stage_e(PlonkWork& w, const PlonkVkDev& vk, const uint8_t* pr, const PlonkDebug& dbg, bool live = true,
                     bool joint = false) {
  int st = live ? BN254V_OK_TRUE : BN254V_STATUS_UNSET;
  G1Aff pf[2] = {vk.g1, vk.g1};
  if (live) st = plonk_stage_d(pf, w, vk, pr, dbg, joint);
  const bool ok = st == BN254V_OK_TRUE;
  if (!ok) pf[0] = pf[1] = vk.g1;
  Fp12 f;
  miller_loop_pairtab<0>(f, nullptr, nullptr, pf, vk.g2_pairs);
  if (ok && dbg.miller) fp12_to_bytes(dbg.miller, f);
  final_exponentiation(f, f);
  if (!ok) return st;
  if (dbg.gt) fp12_to_bytes(dbg.gt, f);
  return eq(f, fp12_one()) ? BN254V_OK_TRUE : BN254V_ERR_PAIRING_CHECK_FAILED;
}

// All stages in one thread.
HD int plonk_verify_one(const PlonkVkDev& vk, const uint8_t* pr, uint32_t len, const uint8_t* inputs_be, int n_inputs,
                        const uint8_t* rnd_be, const PlonkDebug& dbg, bool joint = false) {
  PlonkWork w;
  int st = plonk_stage_a(w, vk, pr, len, inputs_be, n_inputs, dbg);
  if (st != BN254V_OK_TRUE) return st;
  for (int t = 0; t < plonk_n_items(vk.n_qcp, 0, joint); t++) {
    if (joint) plonk_item_joint(w, vk, pr, 0, t);
    else plonk_term(w, vk, pr, 0, t);
  }
  st = plonk_stage_c(w, vk, pr, rnd_be, dbg, joint);
  if (st != BN254V_OK_TRUE) return st;
  for (int t = 0; t < plonk_n_items(vk.n_qcp, 1, joint); t++) {
    if (joint) plonk_item_joint(w, vk, pr, 1, t);
    else plonk_term(w, vk, pr, 1, t);
  }
  return plonk_stage_e(w, vk, pr, dbg, true, joint);
}

}  // namespace bn254
