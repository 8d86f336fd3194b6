// Opt-in aggregate Groth16 check: "is every proof of this batch valid?" with ONE final exponentiation per batch.
// SURVEY.md 8(f).4; no counterpart in the reference, which verifies proof by proof (verifier/src/groth16/verify.rs:65-78).
//
// Per proof i the reference equation is  e(A_i, B_i) e(L_i, gamma') e(C_i, delta') == e(alpha, beta')  (groth16.cuh).
// With scalars r_i drawn after the proofs are fixed, all n equations hold, except with probability n / 2^128-ish, iff
//   prod_i ML(r_i A_i, B_i) * ML(sum r_i L_i, gamma') * ML(sum r_i C_i, delta') * ML(-(sum r_i) alpha, beta')
// final-exponentiates to 1, where  sum r_i L_i = (sum r_i) IC_0 + sum_j (sum_i r_i x_ij) IC_{j+1}.
// So a proof costs two short scalar multiplications and ONE single-pair Miller loop (instead of a three-pair loop and
// a final exponentiation), and the batch costs a product tree, a three-pair Miller loop and one final exponentiation.
// The [r_i] C_i come first, in their own cheap pass: then the sum tree, the batch's three points and its Miller loop
// (a handful of threads) run on a second stream underneath the proofs' Miller loops instead of after them.
//
// r_i = a_i + b_i lambda with a_i (odd), b_i uniform 64-bit: (a, b) -> a + b lambda is injective on that box (the
// shortest vector of the GLV lattice is ~2^127), so r_i is uniform over 2^127 distinct non-zero values of Fr, and
// [r_i] P = [a_i] P + [b_i] phi(P) needs 64 doublings instead of 128.
//
// What is checked per proof before the aggregate -- and reported in status[i] exactly as the per-proof kernels do: record
// length, coordinates < p, A, B, C on their curves, B in G2 (read off the end point of its Miller loop), the number
// of public inputs, inputs < r and != 0.  NOT reproduced: substrate-bn's panic on an identity PARTIAL sum inside
// prepare_inputs (it needs a discrete-log relation between the IC points of the VK).
#pragma once
#include "groth16.cuh"
#include "plonk.cuh"  // g1_w4_tables / g1_w4_add_digits / g1_mul_fixed

namespace bn254 {

// [a + b lambda] P_0 and [a + b lambda] P_1, a and b 64-bit (2 LE words each): signed 4-bit windows over two 8-entry
// affine tables that share one inversion (plonk.cuh g1_w4_tables); 17 windows, 64 doublings per point.
HDN void g1_mul_glv64_2(G1Jac* out, const G1Aff* p, const uint32_t* a, const uint32_t* b) {
  Fp beta;
  BN_LOAD_FP(beta, K::glv_beta, 0);
  uint32_t ka[5] = {a[0], a[1], 0, 0, 0}, kb[5] = {b[0], b[1], 0, 0, 0};
  w4_offset(ka), w4_offset(kb);  // the windows above the 17th hold the digit 0
  G1Aff tab[16];
  g1_w4_tables(tab, p, 2);
  for (int v = 0; v < 2; v++) {
    G1Jac acc = jac_identity<Fp>();
    for (int w = 16; w >= 0; w--) {
      if (w != 16)
        for (int j = 0; j < 4; j++) acc = jac_double(acc);
      g1_w4_add_digits(acc, tab + 8 * v, ka, kb, false, false, beta, w);
    }
    out[v] = acc;
  }
}

// the scalar halves of proof i: 16 bytes = a (LE u64) | b (LE u64); a is made odd, so r_i != 0
HD void groth16_agg_scalar(uint32_t* a, uint32_t* b, const uint8_t* rnd16) {
  for (int k = 0; k < 2; k++) {
    a[k] = (uint32_t)rnd16[4 * k] | ((uint32_t)rnd16[4 * k + 1] << 8) | ((uint32_t)rnd16[4 * k + 2] << 16) | ((uint32_t)rnd16[4 * k + 3] << 24);
    b[k] = (uint32_t)rnd16[8 + 4 * k] | ((uint32_t)rnd16[8 + 4 * k + 1] << 8) | ((uint32_t)rnd16[8 + 4 * k + 2] << 16) |
           ((uint32_t)rnd16[8 + 4 * k + 3] << 24);
  }
  a[0] |= 1;
}

// Everything of a proof that precedes its Miller loop.  Same statuses, in the same order, as groth16_parse_one.
HD int groth16_agg_parse_one(G1Aff& A, G2Aff& B, G1Aff& C, const Groth16VkDev& vk, const uint8_t* proof,
                             uint32_t proof_len, const uint8_t* inputs_be, int n_inputs, bool live) {
  int st = BN254V_OK_TRUE;
  if (!live) st = BN254V_STATUS_UNSET;
  else if (proof_len < 256) st = BN254V_PANIC_SHORT_BUFFER;
  if (st == BN254V_OK_TRUE) st = load_g1_checked(A, proof);
  if (st == BN254V_OK_TRUE) {
    st = load_g2_on_curve(B, proof + 64);
    if (st == BN254V_OK_TRUE) {
      st = load_g1_checked(C, proof + 192);
      if (st == BN254V_OK_TRUE && n_inputs + 1 != vk.n_ic) st = BN254V_ERR_PREPARE_INPUTS;
      for (int i = 0; i < n_inputs && st == BN254V_OK_TRUE; i++) {
        Fr x;
        if (!fr_load_be_plain(x, inputs_be + 32 * i)) st = BN254V_PANIC_FIELD_NOT_MEMBER;
        else if (fe_is_zero(x)) st = BN254V_PANIC_IDENTITY;  // [0] IC_i: the reference's affine sum panics
      }
      if (st != BN254V_OK_TRUE && !g2_in_subgroup<false>(B)) st = BN254V_PANIC_NOT_IN_SUBGROUP;
    }
  }
  return st;
}

// One proof's shares.  (1) Everything before its Miller loop (k_groth16_agg_prepare: one thread per proof in small
// blocks, no barriers): validation, rA = [r_i] A_i in affine coordinates, rc = [r_i] C_i.  A proof that fails validation
// contributes rc = O.  First and on its own so that the batch's own pairing, which needs sum rc, can run while the
// Miller loops of the proofs are still going.
HD int groth16_agg_prepare_one(G1Aff& rA, G2Aff& B, G1Jac& rc, const Groth16VkDev& vk, const uint8_t* proof,
                               uint32_t proof_len, const uint8_t* inputs_be, int n_inputs, const uint8_t* rnd16) {
  G1Aff A, C;
  const int st = groth16_agg_parse_one(A, B, C, vk, proof, proof_len, inputs_be, n_inputs, true);
  rc = jac_identity<Fp>();
  if (st != BN254V_OK_TRUE) return st;
  uint32_t a[2], b[2];
  groth16_agg_scalar(a, b, rnd16);
  const G1Aff pts[2] = {A, C};
  G1Jac res[2];
  g1_mul_glv64_2(res, pts, a, b);
  rc = res[1];
  to_affine(rA, res[0]);  // A has order r and r_i != 0: never the identity
  return st;
}
// (2) f = ML(r_i A_i, B_i) and B's membership in G2, read off the end point of the loop.  `ok == false` (the proof failed
// validation, or a spare thread of the last block): the loop runs on substitute VK points for its block-wide barriers.
HD bool groth16_agg_miller_one(Fp12& f, const Groth16VkDev& vk, G1Aff rA, G2Aff B, bool ok) {
  if (!ok) rA = vk.alpha, B = vk.beta;
  bool in_g2;
  miller_loop<1, 0>(f, &rA, &B, nullptr, nullptr, 0, &in_g2);
  if (!ok || !in_g2) f = fp12_one();
  return in_g2;
}

// The batch's own three pairs: (-(s) alpha, beta'), (s IC_0 + sum t_j IC_j, gamma'), (sum r_i C_i, delta').
// scal_be: s | t_1 | .. | t_n as 32-byte big-endian scalars (< r), computed by the host from the r_i and the public
// inputs.  False when one of the three G1 points is the identity (probability ~2^-127 for honest input; the caller
// then reports "not all valid" and the per-proof path decides).
HD bool groth16_agg_points(G1Aff& nsa, G1Aff& sl, G1Aff& sc, const Groth16VkDev& vk, const uint8_t* scal_be,
                           const G1Jac& sum_rc) {
  const int n_inputs = vk.n_ic - 1;
  Fr s;
  if (!fr_load_be_plain(s, scal_be)) return false;
  G1Jac l, a;
  if (vk.agg_table) {
    l = g1_mul_fixed(vk.agg_table, s.v);
    a = g1_mul_fixed(vk.agg_table + (size_t)BN_IC_WINDOWS * BN_IC_ENTRIES, s.v);
  } else {
    l = scalar_mul(vk.ic[0], s.v);
    a = scalar_mul(vk.alpha, s.v);
  }
  for (int j = 0; j < n_inputs; j++) {
    Fr t;
    if (!fr_load_be_plain(t, scal_be + 32 * (j + 1))) return false;
    l = jac_add(l, vk.ic_table ? g1_mul_fixed(vk.ic_table + (size_t)j * BN_IC_WINDOWS * BN_IC_ENTRIES, t.v)
                               : scalar_mul(vk.ic[j + 1], t.v));
  }
  if (is_identity(l) || is_identity(a) || is_identity(sum_rc)) return false;
  to_affine(sl, l);
  to_affine(nsa, a);
  nsa.y = fe_neg(nsa.y);
  to_affine(sc, sum_rc);
  return true;
}

// The batch's own three-pair Miller value f' (false: degenerate batch, see groth16_agg_points), and the verdict from it
// and the folded product F of the proofs' Miller values.  (One thread; the CUDA path runs the same steps on three lanes:
// k_groth16_agg_fprime3 / k_groth16_agg_final3.)
HD bool groth16_agg_fprime(Fp12& fp, const G1Jac& sum_rc, const Groth16VkDev& vk, const uint8_t* scal_be) {
  G1Aff nsa, pf[2];
  if (!groth16_agg_points(nsa, pf[0], pf[1], vk, scal_be, sum_rc)) return false;
  miller_loop_pairtab<1>(fp, &nsa, &vk.beta, pf, vk.gd_pairs, nullptr);
  return true;
}
HD bool groth16_agg_final(const Fp12& F, const Fp12& fp) {
  Fp12 f;
  mul(f, F, fp);
  final_exponentiation(f, f);
  return eq(f, fp12_one());
}

}  // namespace bn254
