// Extension tower Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3-xi), xi = 9+u, Fq12 = Fq6[w]/(w^2-v)
// in substrate-bn's basis, all coefficients in Montgomery form.
// Replaces bn::{Fq2, Fq6, Fq12} used through bn::pairing / pairing_batch
// (reference call sites verifier/src/groth16/verify.rs:70,73; verifier/src/plonk/kzg.rs:180).
#pragma once
#include "constants.cuh"
#include "field.cuh"

namespace bn254 {

// ------------------------------------------------------------------------------------------ Fp
HD Fp add(const Fp& a, const Fp& b) { return fe_add(a, b); }
HD Fp sub(const Fp& a, const Fp& b) { return fe_sub(a, b); }
HD Fp neg(const Fp& a) { return fe_neg(a); }
HD Fp dbl(const Fp& a) { return fe_dbl(a); }
HD Fp mul(const Fp& a, const Fp& b) { return fe_mul(a, b); }
HD Fp sqr(const Fp& a) { return fe_sqr_short(a); }  // (G1 arithmetic; the Fq2 code squares with fe_mul)
HD bool is_zero(const Fp& a) { return fe_is_zero(a); }
HD bool eq(const Fp& a, const Fp& b) { return fe_eq(a, b); }
HD Fp inv(const Fp& a) { return fe_inv(a); }

#define BN_LOAD_FP(dst, fn, off)                                   \
  {                                                                \
    _Pragma("unroll") for (int _i = 0; _i < 8; _i++)(dst).v[_i] = fn((off) * 8 + _i); \
  }

HD Fp fp_two_inv() {
  Fp r;
  BN_LOAD_FP(r, K::two_inv, 0);
  return r;
}
HD Fp fp_three() {
  Fp r;
  BN_LOAD_FP(r, K::three, 0);
  return r;
}
HD Fp fp_halve(const Fp& a) { return fe_halve(a); }

// ------------------------------------------------------------------------------------------ Fp2
struct Fp2 {
  Fp c0, c1;
};

HD Fp2 fp2_zero() { return Fp2{fe_zero<FpCfg>(), fe_zero<FpCfg>()}; }
HD Fp2 fp2_one() { return Fp2{fe_one<FpCfg>(), fe_zero<FpCfg>()}; }
HD Fp2 add(const Fp2& a, const Fp2& b) { return Fp2{fe_add(a.c0, b.c0), fe_add(a.c1, b.c1)}; }
HD Fp2 sub(const Fp2& a, const Fp2& b) { return Fp2{fe_sub(a.c0, b.c0), fe_sub(a.c1, b.c1)}; }
// a + b without the conditional subtractions (components < 2p).  Only valid as ONE operand of mul(Fp2, Fp2) whose other
// operand is fully reduced: the lazy product then needs a0' b1 + a1' b0 < 4 p^2 < p 2^256 (a' < 2p, b < p), and the
// internal operand sum a0' + a1' < 4p still fits 256 bits.  The product is the same fully reduced element.
HD Fp2 add_nr(const Fp2& a, const Fp2& b) { return Fp2{fe_add_nr(a.c0, b.c0), fe_add_nr(a.c1, b.c1)}; }
HD Fp2 neg(const Fp2& a) { return Fp2{fe_neg(a.c0), fe_neg(a.c1)}; }
HD Fp2 dbl(const Fp2& a) { return Fp2{fe_dbl(a.c0), fe_dbl(a.c1)}; }
HD Fp2 conj(const Fp2& a) { return Fp2{a.c0, fe_neg(a.c1)}; }
HD bool is_zero(const Fp2& a) { return fe_is_zero(a.c0) && fe_is_zero(a.c1); }
HD bool eq(const Fp2& a, const Fp2& b) { return fe_eq(a.c0, b.c0) && fe_eq(a.c1, b.c1); }

// Karatsuba with lazy reduction: 3 full 512-bit products and 2 Montgomery reductions (336 multiply-adds instead of
// the 408 of three Montgomery multiplications).  Out of line, operands by value (registers): one copy in the binary.
//   c0 = a0 b0 - a1 b1           (+ p 2^256 when negative: still < p 2^256 and congruent)
//   c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1      (sums unreduced, < 2p; the difference is a0 b1 + a1 b0 >= 0)
HD Fp2 mul_inl(const Fp2& a, const Fp2& b) {
  uint32_t T0[16], T1[16], S[16];
  fe_mul_wide(T0, a.c0, b.c0);
  fe_mul_wide(T1, a.c1, b.c1);
  fe_mul_wide(S, fe_add_nr(a.c0, a.c1), fe_add_nr(b.c0, b.c1));
  wide_sub(S, T0);
  wide_sub(S, T1);
  Fp2 r;
  r.c1 = fe_redc_wide<FpCfg>(S);
  if (wide_sub(T0, T1)) wide_add_mod_hi<FpCfg>(T0);  // predicated by ptxas
  r.c0 = fe_redc_wide<FpCfg>(T0);
  return r;
}
HDN Fp2 mul(Fp2 a, Fp2 b) { return mul_inl(a, b); }
// (a0+a1)(a0-a1), 2 a0 a1
HDN Fp2 sqr(Fp2 a) {
  Fp c1 = fe_mul(fe_add_nr(a.c0, a.c0), a.c1);  // 2 a0 a1 (first operand < 2p)
  Fp c0 = fe_mul(fe_add_nr(a.c0, a.c1), fe_sub(a.c0, a.c1));
  return Fp2{c0, c1};
}
HD Fp2 scale(const Fp2& a, const Fp& k) { return Fp2{fe_mul(a.c0, k), fe_mul(a.c1, k)}; }
// a j + b k for Fq scalars j, k (one reduction per component)
HD Fp2 scale2_add(const Fp2& a, const Fp& j, const Fp2& b, const Fp& k) {
  return Fp2{fe_mul2_add(a.c0, j, b.c0, k), fe_mul2_add(a.c1, j, b.c1, k)};
}
// multiply by xi = 9 + u: (9 a0 - a1) + (a0 + 9 a1) u
HDN Fp2 mul_xi(Fp2 a) {
  return Fp2{fe_mul9_add(a.c0, fe_mod_minus(a.c1)), fe_mul9_add(a.c1, a.c0)};
}
HD Fp2 inv(const Fp2& a) {
  Fp n = fe_inv(fe_add(fe_sqr(a.c0), fe_sqr(a.c1)));
  return Fp2{fe_mul(a.c0, n), fe_neg(fe_mul(a.c1, n))};
}
HD Fp2 fp2_halve(const Fp2& a) { return Fp2{fe_halve(a.c0), fe_halve(a.c1)}; }

// ---- unreduced Fq2 values for the lazily reduced Fq6 multiplication (tower_body.inc, mul_lazy; an experiment that is
// bit-exact and has 21 % fewer multiply-adds per Fq6 multiplication but measured 11 % slower per Groth16 batch on B200
// than the plain Karatsuba below -- see DESIGN.md section 6 -- so it is opt-in: -DBN_FP6_MUL_LAZY).
// A Wide2 holds re + im u as two signed 512-bit integers (Montgomery scale 2^256 not yet divided out).  In units of
// p^2 (p 2^256 = 5.29 p^2, signed 512-bit range +-13.9 p^2): a product of reduced operands has re in (-1, 1) and
// im in [0, 2).
struct alignas(16) Wide {
  uint32_t v[16];
};
struct Wide2 {
  Wide re, im;
};
// Karatsuba, 3 full products, no reduction: re = a0 b0 - a1 b1, im = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1
HD void mul_wide_inl(Wide2& r, const Fp2& a, const Fp2& b) {
  uint32_t T1[16];
  fe_mul_wide(r.re.v, a.c0, b.c0);
  fe_mul_wide(T1, a.c1, b.c1);
  fe_mul_wide(r.im.v, fe_add_nr(a.c0, a.c1), fe_add_nr(b.c0, b.c1));
  wide_sub(r.im.v, r.re.v);
  wide_sub(r.im.v, T1);
  wide_sub(r.re.v, T1);
}
HDN void mul_wide(Wide2& r, Fp2 a, Fp2 b) {
  Wide2 t;
  mul_wide_inl(t, a, b);
  r = t;  // whole-struct copy: 128-bit local stores
}
// x +- y with y copied to registers half by half as a whole struct (128-bit local loads; element-wise reads between
// the carry-chain asm statements are not merged by the compiler)
HD void wadd(Wide2& x, const Wide2& y) {
  {
    const Wide t = y.re;
    wide_add(x.re.v, t.v);
  }
  const Wide t = y.im;
  wide_add(x.im.v, t.v);
}
HD void wsub(Wide2& x, const Wide2& y) {
  {
    const Wide t = y.re;
    wide_sub(x.re.v, t.v);
  }
  const Wide t = y.im;
  wide_sub(x.im.v, t.v);
}
// x <- xi x = (9 re - im) + (9 im + re) u, each component brought back to |.| <= 2.65 p^2 modulo p 2^256.
// Valid for re in (-2, 2), im in [0, 4) (p^2 units): 9 re - im in (-22, 18), 9 im + re in (-2, 38).
HD void mul_xi_wide_inl(Wide2& x) {
  Wide t;
  wide_mul9_addsub<FpCfg, true>(t.v, x.re.v, x.im.v);
  wide_mul9_addsub<FpCfg, false>(x.im.v, x.im.v, x.re.v);
  x.re = t;
}
HDN void mul_xi_wide(Wide2& x) {
  Wide2 t = x;
  mul_xi_wide_inl(t);
  x = t;
}
// One output coefficient of the lazily reduced Fq6 multiplication: (XI ? xi : 1)(a b - X - Y) + Z, reduced.  The running
// value stays in registers from the products to the reductions; X, Y, Z stream in from local memory once.
// Ranges (p^2 units): a b - X - Y has re in (-2, 2), im in [0, 4) when XI; the final value has re in (-5.29, 5.29) and
// im in (-5.29, 5.29), or im in (-5.29, 10.58) with TWICE.
template <bool XI, bool TWICE>
HDN Fp2 mul_sub2_add(Fp2 a, Fp2 b, const Wide2& X, const Wide2& Y, const Wide2& Z) {
  Wide2 P;
  mul_wide_inl(P, a, b);
  wsub(P, X);
  wsub(P, Y);
  if (XI) mul_xi_wide_inl(P);
  wadd(P, Z);
  Fp2 r;
  r.c0 = fe_redc_wide_signed<FpCfg, false>(P.re.v);
  r.c1 = fe_redc_wide_signed<FpCfg, TWICE>(P.im.v);
  return r;
}
#define BN_HAVE_FP6_MUL_LAZY  // tower_body.inc: mul_lazy(Fp6&, ..) is compiled; -DBN_FP6_MUL_LAZY makes it the Fq6 multiplier

#define BN_LOAD_FP2(dst, fn, idx) \
  {                               \
    BN_LOAD_FP((dst).c0, fn, 2 * (idx)); \
    BN_LOAD_FP((dst).c1, fn, 2 * (idx) + 1); \
  }

HD Fp2 fp2_b2() {
  Fp2 r;
  BN_LOAD_FP2(r, K::b2, 0);
  return r;
}

// Frobenius coefficients xi^(i(p^k-1)/6)
template <int KK>
HD Fp2 frob_coeff(int i) {  // i = 1..5
  Fp2 r;
  if (KK == 1) { BN_LOAD_FP2(r, K::frob1, i - 1); }
  else if (KK == 2) { BN_LOAD_FP2(r, K::frob2, i - 1); }
  else { BN_LOAD_FP2(r, K::frob3, i - 1); }
  return r;
}
template <int KK>
HD Fp frob_coeff_fp(int i) { return frob_coeff<KK>(i).c0; }

#include "tower_body.inc"

// canonical serialisation: 12 x 32-byte BE, order c0.c0.c0, c0.c0.c1, c0.c1.c0, ..., c1.c2.c1
HDN void fp12_to_bytes(uint8_t* out, const Fp12& a) {
  const Fp2* cs[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
  for (int i = 0; i < 6; i++) {
    fe_to_be_bytes(out + 64 * i, fe_from_mont(cs[i]->c0));
    fe_to_be_bytes(out + 64 * i + 32, fe_from_mont(cs[i]->c1));
  }
}

}  // namespace bn254
