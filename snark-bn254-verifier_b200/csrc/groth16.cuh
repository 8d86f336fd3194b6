// Per-proof Groth16 verification (one proof per thread in the v1 kernels).
// Replaces verify_groth16 / prepare_inputs (reference verifier/src/groth16/verify.rs:53-78) and the
// per-proof half of load_groth16_proof_from_bytes (verifier/src/groth16/converter.rs:14-26).
//
// GPU shape (same mathematical values, VK-constant work hoisted to vk_load):
//   target  = e(alpha, beta')             computed once per VK  (reference recomputes it per call, :70)
//   lines   = G2::precompute(gamma'), G2::precompute(delta')    once per VK
//   verdict = FE( ML[(A,B), (L,gamma'), (C,delta')] ) == target
// where (beta', gamma', delta') = (-beta_file, gamma, -delta) for the reference equation and
// (beta_file, -gamma, -delta) for gnark's (sign_mode 1).
#pragma once
#include "io.cuh"
#include "pairing.cuh"

namespace bn254 {

#define BN_MAX_IC 9  // up to 8 public inputs

// Fixed-base tables for prepare_inputs: the IC bases are VK-constant, so [x] IC_i is a sum of 32 table entries
// T[i][w][d-1] = d * 2^(8w) * IC_{i+1}, d = 1..255 (affine), one per non-zero byte of x.  Same group element as
// the reference's 254-bit double-and-add, hence the same affine L.
#define BN_IC_WINDOWS 32
#define BN_IC_ENTRIES 255

struct Groth16VkDev {
  int n_ic;
  const G1Aff* ic_table;     // [n_ic - 1][BN_IC_WINDOWS][BN_IC_ENTRIES] in device memory, or null
  const G1Aff* agg_table;    // [2][BN_IC_WINDOWS][BN_IC_ENTRIES]: the same for IC_0 and alpha (groth16_agg.cuh), or null
  G1Aff alpha;
  G2Aff beta, gamma, delta;  // already sign-adjusted (beta', gamma', delta')
  G1Aff ic[BN_MAX_IC];
  Fp12 target;               // e(alpha, beta')
  Line gamma_lines[BN_N_LINES];
  Line delta_lines[BN_N_LINES];
  LinePairKF gd_pairs[BN_N_LINES];  // product coefficients of the gamma and delta lines (pairing_body.inc)
};

// VK-constant precomputation (runs once per VK, single thread).
HD void groth16_vk_prepare(Groth16VkDev& vk) {
  g2_precompute(vk.gamma_lines, vk.gamma);
  g2_precompute(vk.delta_lines, vk.delta);
  line_pair_table(vk.gd_pairs, vk.gamma_lines, vk.delta_lines);
  Fp12 f;
  miller_loop<1, 0>(f, &vk.alpha, &vk.beta, nullptr, nullptr);
  final_exponentiation(vk.target, f);
}

// One (base, window) slice of the fixed-base table.
HD void groth16_ic_table_slice(G1Aff* out, const G1Aff& base, int w) {
  G1Jac b = to_jac(base);
  for (int i = 0; i < 8 * w; i++) b = jac_double(b);
  G1Aff step;
  to_affine(step, b);  // a VK point has order r: never the identity
  G1Jac cur = b;
  out[0] = step;
  for (int d = 2; d <= BN_IC_ENTRIES; d++) {
    cur = jac_add_mixed(cur, step);
    to_affine(out[d - 1], cur);
  }
}

// L = IC_0 + sum x_i IC_{i+1}; affine accumulation semantics of the reference: an identity term
// (x_i == 0) or an identity partial sum panics inside substrate-bn.
HD int groth16_prepare_inputs(G1Aff& L, const Groth16VkDev& vk, const uint8_t* inputs_be, int n_inputs) {
  if (n_inputs + 1 != vk.n_ic) return BN254V_ERR_PREPARE_INPUTS;
  G1Jac acc = to_jac(vk.ic[0]);
  for (int i = 0; i < n_inputs; i++) {
    Fr x;
    if (!fr_load_be_plain(x, inputs_be + 32 * i)) return BN254V_PANIC_FIELD_NOT_MEMBER;
    G1Jac term;
    if (vk.ic_table) {
      term = jac_identity<Fp>();
      const G1Aff* tab = vk.ic_table + (size_t)i * BN_IC_WINDOWS * BN_IC_ENTRIES;
      for (int w = 0; w < BN_IC_WINDOWS; w++) {
        uint32_t d = (x.v[w >> 2] >> (8 * (w & 3))) & 0xff;
        if (d) term = jac_add_mixed(term, tab[w * BN_IC_ENTRIES + (d - 1)]);
      }
    } else {
      term = scalar_mul(vk.ic[i + 1], x.v);
    }
    if (is_identity(term)) return BN254V_PANIC_IDENTITY;
    acc = jac_add(acc, term);
    if (is_identity(acc)) return BN254V_PANIC_IDENTITY;
  }
  to_affine(L, acc);
  return BN254V_OK_TRUE;
}

struct Groth16Debug {
  uint8_t* L;       // 64 B or null
  uint8_t* miller;  // 384 B or null
  uint8_t* gt;      // 384 B or null
};

// proof: >= 256 bytes (A | B | C), proof_len: valid bytes.
// The verification in pieces, so that the batch kernels can run them as separate launches (each with a fraction of the
// code and stack): groth16_parse_one (decode, validate, prepare_inputs), the Miller loop -> the Fq12 Miller value
// (groth16_miller_one = both, for the fused small-batch kernel), groth16_finish_one (final exponentiation and
// comparison with e(alpha, beta')).
// Barrier discipline: miller_loop_pairtab / final_exponentiation contain block-wide phase barriers, so every thread of a
// block must reach them the same number of times.  A proof that fails before the Miller loop is therefore NOT ended
// early: its status is recorded and the thread runs the loop on substitute VK points (always valid, order r), whose
// result is discarded.  `live == false` (a spare thread of the last block) does the same and writes nothing.
// Decode, validate and prepare_inputs: everything of a proof that precedes the Miller loop.  On BN254V_OK_TRUE the
// points A, B (on the curve; its G2 membership is read off the Miller loop's end point), C and L are set.
HD int groth16_parse_one(G1Aff& A, G2Aff& B, G1Aff& C, G1Aff& L, const Groth16VkDev& vk, const uint8_t* proof,
                         uint32_t proof_len, const uint8_t* inputs_be, int n_inputs, bool live = true) {
  int st = BN254V_OK_TRUE;
  if (!live) st = BN254V_STATUS_UNSET;
  else if (proof_len < 256) st = BN254V_PANIC_SHORT_BUFFER;
  if (st == BN254V_OK_TRUE) st = load_g1_checked(A, proof);
  if (st == BN254V_OK_TRUE) {
    st = load_g2_on_curve(B, proof + 64);
    if (st == BN254V_OK_TRUE) {
      // B's subgroup test (the last check of AffineG2::new) is read off the end point of the Miller loop.  The
      // reference parses B before C and before the public inputs, so on a later failure B's verdict still comes first
      // (rare path, separate scalar multiplication, no barriers).
      st = load_g1_checked(C, proof + 192);
      if (st == BN254V_OK_TRUE) st = groth16_prepare_inputs(L, vk, inputs_be, n_inputs);
      if (st != BN254V_OK_TRUE && !g2_in_subgroup<false>(B)) st = BN254V_PANIC_NOT_IN_SUBGROUP;
    }
  }
  return st;
}
HD int groth16_miller_one(Fp12& f, const Groth16VkDev& vk, const uint8_t* proof, uint32_t proof_len,
                          const uint8_t* inputs_be, int n_inputs, const Groth16Debug& dbg, bool live = true) {
  G1Aff A, C, L;
  G2Aff B;
  const int st = groth16_parse_one(A, B, C, L, vk, proof, proof_len, inputs_be, n_inputs, live);
  const bool ok = st == BN254V_OK_TRUE;
  if (!ok) A = vk.alpha, B = vk.beta, L = vk.ic[0], C = vk.ic[0];  // substitute inputs; the result is discarded

  G1Aff pf[2] = {L, C};
  bool in_g2;
  miller_loop_pairtab<1>(f, &A, &B, pf, vk.gd_pairs, &in_g2);
  if (!ok) return st;
  if (!in_g2) return BN254V_PANIC_NOT_IN_SUBGROUP;
  if (dbg.L) store_g1(dbg.L, L);
  if (dbg.miller) fp12_to_bytes(dbg.miller, f);
  return BN254V_OK_TRUE;
}
// `decide == false`: run the exponentiation for its barriers only (the proof already has its status).
HD int groth16_finish_one(Fp12& f, const Groth16VkDev& vk, const Groth16Debug& dbg, bool decide = true) {
  final_exponentiation(f, f);
  if (!decide) return BN254V_STATUS_UNSET;
  if (dbg.gt) fp12_to_bytes(dbg.gt, f);
  return eq(f, vk.target) ? BN254V_OK_TRUE : BN254V_OK_FALSE;
}

// proof: >= 256 bytes (A | B | C), proof_len: valid bytes.  Returns BN254V_STATUS_UNSET for a spare thread.
HD int groth16_verify_one(const Groth16VkDev& vk, const uint8_t* proof, uint32_t proof_len,
                          const uint8_t* inputs_be, int n_inputs, const Groth16Debug& dbg, bool live = true) {
  Fp12 f;
  int st = groth16_miller_one(f, vk, proof, proof_len, inputs_be, n_inputs, dbg, live);
  const bool ok = st == BN254V_OK_TRUE;
  if (!ok) f = vk.target;  // any non-zero value: the exponentiation runs for its barriers only
  int st2 = groth16_finish_one(f, vk, dbg, ok);
  return ok ? st2 : st;
}

// Raw k-pair product (bn::pairing_batch): all G2 variable.  A pair with an identity member -- encoded as all-zero
// bytes: 64 for G1, 128 for G2 -- is skipped, as substrate-bn's pairing_batch skips it: the thread still walks the
// loop on generator points (barriers, uniform control flow) with that pair masked out, i.e. its lines replaced by the
// sparse element 1, so the Miller value is exactly the product over the remaining pairs.  Returns is_one.
HD bool all_zero_bytes(const uint8_t* b, int n) {
  uint32_t t = 0;
  for (int i = 0; i < n; i++) t |= b[i];
  return t == 0;
}
// the points of one k-pair set; bit j of `skip`: pair j has an identity member (substituted by the generators)
template <int KP>
HD uint32_t pairing_product_load(G1Aff* p, G2Aff* q, const uint8_t* g1, const uint8_t* g2) {
  uint32_t skip = 0;
  for (int j = 0; j < KP; j++) {
    if (all_zero_bytes(g1 + 64 * j, 64) || all_zero_bytes(g2 + 128 * j, 128)) {
      skip |= 1u << j;
      p[j] = g1_generator();
      q[j] = g2_generator_dev();
    } else {
      load_g1_unchecked(p[j], g1 + 64 * j);
      load_g2_unchecked(q[j], g2 + 128 * j);
    }
  }
  return skip;
}
template <int KP>
HD bool pairing_product_one(const uint8_t* g1, const uint8_t* g2, uint8_t* miller_out, uint8_t* gt_out) {
  G1Aff p[KP];
  G2Aff q[KP];
  const uint32_t skip = pairing_product_load<KP>(p, q, g1, g2);
  Fp12 f;
  miller_loop<KP, 0>(f, p, q, nullptr, nullptr, skip);
  if (miller_out) fp12_to_bytes(miller_out, f);
  final_exponentiation(f, f);
  if (gt_out) fp12_to_bytes(gt_out, f);
  return eq(f, fp12_one());
}

}  // namespace bn254
