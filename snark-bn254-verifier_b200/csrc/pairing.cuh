// Optimal-ate multi-Miller loop and final exponentiation -- the GPU replacement for
// bn::pairing / bn::pairing_batch (reference call sites verifier/src/groth16/verify.rs:70,73 and
// verifier/src/plonk/kzg.rs:180-187).  Formulas follow substrate-bn's flipped Miller loop
//  so that the canonical Fq12 Miller value is bit-identical:
//   * homogeneous projective doubling / mixed-addition steps with line (ell_0, ell_vw, ell_vv),
//   * 64-digit ATE_LOOP_COUNT_NAF followed by the two Frobenius additions,
//   * one shared Fq12 accumulator for all pairs of a check,
//   * final exponentiation = easy part + Fuentes-Castaneda hard part with cyclotomic squarings.
#pragma once
#include "curve.cuh"

namespace bn254 {

struct Line {
  Fp2 ell_0, ell_vw, ell_vv;
};

// R <- 2R, returns the tangent line coefficients.
HDN Line doubling_step(G2Jac& r) {
  Fp2 a = fp2_halve(mul(r.x, r.y));
  Fp2 b = sqr(r.y);
  Fp2 c = sqr(r.z);
  Fp2 d = add(dbl(c), c);
  Fp2 e = mul(fp2_b2(), d);
  Fp2 f = add(dbl(e), e);
  Fp2 g = fp2_halve(add(b, f));
  Fp2 h = sub(sqr(add(r.y, r.z)), add(b, c));
  Fp2 i = sub(e, b);
  Fp2 j = sqr(r.x);
  Fp2 e2 = sqr(e);
  r.x = mul(a, sub(b, f));
  r.y = sub(sqr(g), add(dbl(e2), e2));
  r.z = mul(b, h);
  return Line{mul_xi(i), neg(h), add(dbl(j), j)};
}

// R <- R + Q (Q affine), returns the chord line coefficients.
HDN Line addition_step(G2Jac& r, const G2Aff& q) {
  Fp2 d = sub(r.x, mul(r.z, q.x));
  Fp2 e = sub(r.y, mul(r.z, q.y));
  Fp2 f = sqr(d);
  Fp2 g = sqr(e);
  Fp2 h = mul(d, f);
  Fp2 i = mul(r.x, f);
  Fp2 j = sub(add(mul(r.z, g), h), dbl(i));
  Fp2 ell0 = mul_xi(sub(mul(e, q.x), mul(d, q.y)));
  r.x = mul(d, j);
  r.y = sub(mul(e, sub(i, j)), mul(h, r.y));
  r.z = mul(r.z, h);
  return Line{ell0, d, neg(e)};
}

HD G2Aff g2_mul_by_q(const G2Aff& q) { return g2_psi(q); }

// f <- f * line(P): mul_by_024(ell_0, ell_vw * P.y, ell_vv * P.x)
HDN void apply_line(Fp12& f, const Line& l, const G1Aff& p) {
  mul_by_024(f, l.ell_0, scale(l.ell_vw, p.y), scale(l.ell_vv, p.x));
}

#define BN_N_LINES 87

// G2::precompute -> 87 line triples for a fixed (VK-constant) G2 point.
HDN void g2_precompute(Line* out, const G2Aff& q) {
  G2Jac r = to_jac(q);
  G2Aff nq = neg(q);
  int idx = 0;
  for (int k = 0; k < 64; k++) {
    out[idx++] = doubling_step(r);
    int d = K::ate_digit(k);
    if (d == 1) out[idx++] = addition_step(r, q);
    else if (d == 3) out[idx++] = addition_step(r, nq);
  }
  G2Aff q1 = g2_mul_by_q(q);
  G2Aff q2 = neg(g2_mul_by_q(q1));
  out[idx++] = addition_step(r, q1);
  out[idx++] = addition_step(r, q2);
}

// Shared-accumulator Miller loop over NV pairs with a variable G2 point (lines computed on the
// fly) and NF pairs whose G2 point has a precomputed line table.
// `active_v` / `active_f`: pairs with an identity member are skipped (substrate-bn pairing_batch).
template <int NV, int NF>
HD void miller_loop(Fp12& f, const G1Aff* pv, const G2Aff* qv, const G1Aff* pf, const Line* const* tables) {
  f = fp12_one();
  G2Jac r[NV > 0 ? NV : 1];
  G2Aff nq[NV > 0 ? NV : 1];
  for (int v = 0; v < NV; v++) {
    r[v] = to_jac(qv[v]);
    nq[v] = neg(qv[v]);
  }
  int idx = 0;
  for (int k = 0; k < 64; k++) {
    BN_PHASE_SYNC();
    sqr(f, f);
    for (int v = 0; v < NV; v++) {
      BN_PHASE_SYNC_FINE();
      Line l = doubling_step(r[v]);
      BN_PHASE_SYNC_FINE();
      apply_line(f, l, pv[v]);
    }
    BN_PHASE_SYNC_FINE();
    for (int t = 0; t < NF; t++) apply_line(f, tables[t][idx], pf[t]);
    idx++;
    int d = K::ate_digit(k);
    if (d != 0) {
      for (int v = 0; v < NV; v++) {
        BN_PHASE_SYNC_FINE();
        Line l = addition_step(r[v], d == 1 ? qv[v] : nq[v]);
        BN_PHASE_SYNC_FINE();
        apply_line(f, l, pv[v]);
      }
      BN_PHASE_SYNC_FINE();
      for (int t = 0; t < NF; t++) apply_line(f, tables[t][idx], pf[t]);
      idx++;
    }
  }
  BN_PHASE_SYNC();
  // Frobenius additions: all pairs consume coefficient idx (Q1) then idx+1 (Q2), as
  // substrate-bn's miller_loop_batch shares the coefficient index across pairs.
  G2Aff q1[NV > 0 ? NV : 1], q2[NV > 0 ? NV : 1];
  for (int v = 0; v < NV; v++) {
    q1[v] = g2_mul_by_q(qv[v]);
    q2[v] = neg(g2_mul_by_q(q1[v]));
  }
  for (int v = 0; v < NV; v++) {
    Line l = addition_step(r[v], q1[v]);
    apply_line(f, l, pv[v]);
  }
  for (int t = 0; t < NF; t++) apply_line(f, tables[t][idx], pf[t]);
  idx++;
  for (int v = 0; v < NV; v++) {
    Line l = addition_step(r[v], q2[v]);
    apply_line(f, l, pv[v]);
  }
  for (int t = 0; t < NF; t++) apply_line(f, tables[t][idx], pf[t]);
}

// r = conj(a^x), x = BN parameter, for a in the cyclotomic subgroup (r must not alias a).  Left-to-right over the
// non-adjacent form of x (24 non-zero digits instead of 28 bits; a^-1 = conj(a) is free there): the same element
// a^x as substrate-bn's plain square-and-multiply, with 4 fewer Fq12 multiplications.
HDN void exp_by_neg_z(Fp12& r, const Fp12& a) {
  const uint64_t pos = 0x450a14044a890a01ull, neg = 0x20815000200010ull;  // x = pos - neg, top digit (bit 62) positive
  Fp12 ai;
  conj(ai, a);
  r = a;
  for (int i = 61; i >= 0; i--) {
    BN_PHASE_SYNC();
    cyclotomic_sqr(r, r);
    if ((pos >> i) & 1) {
      BN_PHASE_SYNC_FINE();
      mul(r, r, a);
    } else if ((neg >> i) & 1) {
      BN_PHASE_SYNC_FINE();
      mul(r, r, ai);
    }
  }
  BN_PHASE_SYNC();
  conj(r, r);
}

// Fq12::final_exponentiation (r may alias f).  `f` must be non-zero (a Miller value always is).
// Same chain as substrate-bn, written over 8 reused buffers.
HDN void final_exponentiation(Fp12& r, const Fp12& f) {
  Fp12 T, A, B, D, E, Kk, L, X;
  inv(A, f);
  conj(X, f);
  mul(T, X, A);  // f^(p^6 - 1)
  frobenius<2>(A, T);
  mul(T, A, T);  // t = f^((p^6-1)(p^2+1))
  exp_by_neg_z(A, T);      // a
  cyclotomic_sqr(B, A);    // b
  cyclotomic_sqr(X, B);    // c
  mul(D, X, B);            // d
  exp_by_neg_z(E, D);      // e
  cyclotomic_sqr(X, E);    // f
  exp_by_neg_z(A, X);      // g
  conj(A, A);              // i = conj(g)
  mul(Kk, A, E);           // j = i e
  conj(X, D);              // h
  mul(Kk, Kk, X);          // k = j h
  mul(L, Kk, B);           // l = k b
  mul(X, Kk, E);           // m = k e
  mul(X, T, X);            // n = t m
  frobenius<1>(A, L);      // o
  mul(X, A, X);            // p = o n
  frobenius<2>(A, Kk);     // q
  mul(X, A, X);            // r = q p
  conj(A, T);              // s
  mul(A, A, L);            // t' = s l
  frobenius<3>(B, A);      // u
  mul(r, B, X);
}

}  // namespace bn254
