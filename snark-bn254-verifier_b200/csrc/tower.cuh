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
HD Fp sqr(const Fp& a) { return fe_sqr(a); }
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
HD Fp fp_halve(const Fp& a) { return fe_mul(a, fp_two_inv()); }

// ------------------------------------------------------------------------------------------ Fp2
struct Fp2 {
  Fp c0, c1;
};

HD Fp2 fp2_zero() { return Fp2{fe_zero<FpCfg>(), fe_zero<FpCfg>()}; }
HD Fp2 fp2_one() { return Fp2{fe_one<FpCfg>(), fe_zero<FpCfg>()}; }
HD Fp2 add(const Fp2& a, const Fp2& b) { return Fp2{fe_add(a.c0, b.c0), fe_add(a.c1, b.c1)}; }
HD Fp2 sub(const Fp2& a, const Fp2& b) { return Fp2{fe_sub(a.c0, b.c0), fe_sub(a.c1, b.c1)}; }
HD Fp2 neg(const Fp2& a) { return Fp2{fe_neg(a.c0), fe_neg(a.c1)}; }
HD Fp2 dbl(const Fp2& a) { return Fp2{fe_dbl(a.c0), fe_dbl(a.c1)}; }
HD Fp2 conj(const Fp2& a) { return Fp2{a.c0, fe_neg(a.c1)}; }
HD bool is_zero(const Fp2& a) { return fe_is_zero(a.c0) && fe_is_zero(a.c1); }
HD bool eq(const Fp2& a, const Fp2& b) { return fe_eq(a.c0, b.c0) && fe_eq(a.c1, b.c1); }

// Karatsuba with lazy reduction: 3 full 512-bit products and 2 Montgomery reductions (336 multiply-adds instead of
// the 408 of three Montgomery multiplications).  Out of line, operands by value (registers): one copy in the binary.
//   c0 = a0 b0 - a1 b1           (+ p 2^256 when negative: still < p 2^256 and congruent)
//   c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1      (sums unreduced, < 2p; the difference is a0 b1 + a1 b0 >= 0)
HDN Fp2 mul(Fp2 a, Fp2 b) {
  uint32_t T0[16], T1[16], S[16];
  fe_mul_wide(T0, a.c0, b.c0);
  fe_mul_wide(T1, a.c1, b.c1);
  fe_mul_wide(S, fe_add_nr(a.c0, a.c1), fe_add_nr(b.c0, b.c1));
  wide_sub(S, T0);
  wide_sub(S, T1);
  Fp2 r;
  r.c1 = fe_redc_wide<FpCfg>(S);
  uint32_t bw = wide_sub(T0, T1);
  wide_add_mod_hi<FpCfg>(T0, bw);
  r.c0 = fe_redc_wide<FpCfg>(T0);
  return r;
}
// (a0+a1)(a0-a1), 2 a0 a1
HDN Fp2 sqr(Fp2 a) {
  Fp c1 = fe_mul(fe_add_nr(a.c0, a.c0), a.c1);  // 2 a0 a1 (first operand < 2p)
  Fp c0 = fe_mul(fe_add_nr(a.c0, a.c1), fe_sub(a.c0, a.c1));
  return Fp2{c0, c1};
}
HD Fp2 scale(const Fp2& a, const Fp& k) { return Fp2{fe_mul(a.c0, k), fe_mul(a.c1, k)}; }
// multiply by xi = 9 + u: (9 a0 - a1) + (a0 + 9 a1) u
HDN Fp2 mul_xi(Fp2 a) {
  Fp t0 = fe_dbl(fe_dbl(fe_dbl(a.c0)));  // 8 a0
  Fp t1 = fe_dbl(fe_dbl(fe_dbl(a.c1)));  // 8 a1
  return Fp2{fe_sub(fe_add(t0, a.c0), a.c1), fe_add(fe_add(t1, a.c1), a.c0)};
}
HD Fp2 inv(const Fp2& a) {
  Fp n = fe_inv(fe_add(fe_sqr(a.c0), fe_sqr(a.c1)));
  return Fp2{fe_mul(a.c0, n), fe_neg(fe_mul(a.c1, n))};
}
HD Fp2 fp2_halve(const Fp2& a) { return scale(a, fp_two_inv()); }

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

// ------------------------------------------------------------------------------------------ Fp6
// Fp6 / Fp12 values live in (local) memory; every operation below is destination-passing -- `r` may alias any
// input -- so no 192 / 384-byte temporaries are created for return values and the per-thread stack stays small.
struct Fp6 {
  Fp2 c0, c1, c2;
};
HD Fp6 fp6_zero() { return Fp6{fp2_zero(), fp2_zero(), fp2_zero()}; }
HD Fp6 fp6_one() { return Fp6{fp2_one(), fp2_zero(), fp2_zero()}; }
HD bool eq(const Fp6& a, const Fp6& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1) && eq(a.c2, b.c2); }

// element-wise helpers over n consecutive Fp (an Fp6 is 6, an Fp12 is 12): rolled loops, one copy in the binary
HDN void fpn_add(Fp* r, const Fp* a, const Fp* b, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) r[i] = fe_add(a[i], b[i]);
}
HDN void fpn_sub(Fp* r, const Fp* a, const Fp* b, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) r[i] = fe_sub(a[i], b[i]);
}
HDN void fpn_neg(Fp* r, const Fp* a, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) r[i] = fe_neg(a[i]);
}
HD void add(Fp6& r, const Fp6& a, const Fp6& b) { fpn_add(&r.c0.c0, &a.c0.c0, &b.c0.c0, 6); }
HD void sub(Fp6& r, const Fp6& a, const Fp6& b) { fpn_sub(&r.c0.c0, &a.c0.c0, &b.c0.c0, 6); }
HD void neg(Fp6& r, const Fp6& a) { fpn_neg(&r.c0.c0, &a.c0.c0, 6); }
HD void dbl(Fp6& r, const Fp6& a) { fpn_add(&r.c0.c0, &a.c0.c0, &a.c0.c0, 6); }
// multiply by v
HD void mul_v(Fp6& r, const Fp6& a) {
  Fp2 t = mul_xi(a.c2);
  Fp2 a0 = a.c0, a1 = a.c1;
  r.c0 = t;
  r.c1 = a0;
  r.c2 = a1;
}

// Karatsuba: 6 Fp2 multiplications
HDN void mul(Fp6& r, const Fp6& a, const Fp6& b) {
  Fp2 v0 = mul(a.c0, b.c0);
  Fp2 v1 = mul(a.c1, b.c1);
  Fp2 v2 = mul(a.c2, b.c2);
  Fp2 t0 = sub(sub(mul(add(a.c1, a.c2), add(b.c1, b.c2)), v1), v2);
  Fp2 t1 = sub(sub(mul(add(a.c0, a.c1), add(b.c0, b.c1)), v0), v1);
  Fp2 t2 = sub(sub(mul(add(a.c0, a.c2), add(b.c0, b.c2)), v0), v2);
  r.c0 = add(v0, mul_xi(t0));
  r.c1 = add(t1, mul_xi(v2));
  r.c2 = add(t2, v1);
}
// CH-SQR2: 2 mul + 3 sqr in Fp2
HDN void sqr(Fp6& r, const Fp6& a) {
  Fp2 s0 = sqr(a.c0);
  Fp2 s1 = dbl(mul(a.c0, a.c1));
  Fp2 s2 = sqr(add(sub(a.c0, a.c1), a.c2));
  Fp2 s3 = dbl(mul(a.c1, a.c2));
  Fp2 s4 = sqr(a.c2);
  r.c0 = add(s0, mul_xi(s3));
  r.c1 = add(s1, mul_xi(s4));
  r.c2 = sub(add(add(s1, s2), s3), add(s0, s4));
}
HDN void inv(Fp6& r, const Fp6& a) {
  Fp2 c0 = sub(sqr(a.c0), mul_xi(mul(a.c1, a.c2)));
  Fp2 c1 = sub(mul_xi(sqr(a.c2)), mul(a.c0, a.c1));
  Fp2 c2 = sub(sqr(a.c1), mul(a.c0, a.c2));
  Fp2 t = add(mul(a.c0, c0), mul_xi(add(mul(a.c2, c1), mul(a.c1, c2))));
  Fp2 ti = inv(t);
  r.c0 = mul(c0, ti);
  r.c1 = mul(c1, ti);
  r.c2 = mul(c2, ti);
}

// ------------------------------------------------------------------------------------------ Fp12
struct Fp12 {
  Fp6 c0, c1;
};
HD Fp12 fp12_one() { return Fp12{fp6_one(), fp6_zero()}; }
HD bool eq(const Fp12& a, const Fp12& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
// unitary inverse (conjugate)
HD void conj(Fp12& r, const Fp12& a) {
  if (&r != &a) r.c0 = a.c0;
  neg(r.c1, a.c1);
}

// Karatsuba: 3 Fp6 multiplications
HDN void mul(Fp12& r, const Fp12& a, const Fp12& b) {
  Fp6 v0, v1, s, t;
  mul(v0, a.c0, b.c0);
  mul(v1, a.c1, b.c1);
  add(s, a.c0, a.c1);
  add(t, b.c0, b.c1);
  mul(s, s, t);
  sub(s, s, v0);
  sub(r.c1, s, v1);
  mul_v(t, v1);
  add(r.c0, v0, t);
}
// complex squaring: 2 Fp6 multiplications
HDN void sqr(Fp12& r, const Fp12& a) {
  Fp6 ab, s, t;
  mul(ab, a.c0, a.c1);
  add(s, a.c0, a.c1);
  mul_v(t, a.c1);
  add(t, t, a.c0);
  mul(s, s, t);       // (a0 + a1)(a0 + v a1)
  sub(s, s, ab);
  mul_v(t, ab);
  sub(r.c0, s, t);
  dbl(r.c1, ab);
}
HDN void inv(Fp12& r, const Fp12& a) {
  Fp6 t, u;
  sqr(t, a.c0);
  sqr(u, a.c1);
  mul_v(u, u);
  sub(t, t, u);
  inv(t, t);
  mul(r.c0, a.c0, t);
  mul(u, a.c1, t);
  neg(r.c1, u);
}

// f <- f * (x0 + x2 v^2 + x4 v w)  -- substrate-bn's mul_by_024(ell_0 = x0, ell_vw = x4, ell_vv = x2):
// the sparse operand is Fq12{c0: (x0, 0, x2), c1: (0, x4, 0)}.  14 Fp2 multiplications.
HDN void mul_by_024(Fp12& f, const Fp2& x0, const Fp2& x4, const Fp2& x2) {
  Fp6 AS0, BS1, T, S;
  {
    const Fp6& A = f.c0;
    // A * (x0, 0, x2): 5 mul
    Fp2 a0x0 = mul(A.c0, x0);
    Fp2 a2x2 = mul(A.c2, x2);
    Fp2 a1x0 = mul(A.c1, x0);
    Fp2 a1x2 = mul(A.c1, x2);
    Fp2 cross = sub(sub(mul(add(A.c0, A.c2), add(x0, x2)), a0x0), a2x2);  // a0x2 + a2x0
    AS0.c0 = add(a0x0, mul_xi(a1x2));
    AS0.c1 = add(a1x0, mul_xi(a2x2));
    AS0.c2 = cross;
  }
  {
    const Fp6& B = f.c1;
    // B * (0, x4, 0): 3 mul -> (xi b2x4, b0x4, b1x4)
    BS1.c0 = mul_xi(mul(B.c2, x4));
    BS1.c1 = mul(B.c0, x4);
    BS1.c2 = mul(B.c1, x4);
  }
  // (A+B) * (x0, x4, x2): 6 mul
  S.c0 = x0, S.c1 = x4, S.c2 = x2;
  add(T, f.c0, f.c1);
  mul(T, T, S);
  sub(T, T, AS0);
  sub(f.c1, T, BS1);
  mul_v(BS1, BS1);
  add(f.c0, AS0, BS1);
}

// Granger-Scott squaring for cyclotomic-subgroup elements: 6 Fp2 mul-equivalents (18 m).
HD void fp4_sqr(Fp2& t0, Fp2& t1, const Fp2& z0, const Fp2& z1) {
  Fp2 tmp = mul(z0, z1);
  t0 = sub(sub(mul(add(z0, z1), add(z0, mul_xi(z1))), tmp), mul_xi(tmp));
  t1 = dbl(tmp);
}
HDN void cyclotomic_sqr(Fp12& r, const Fp12& a) {
  // z0=c0.c0 z4=c0.c1 z3=c0.c2 z2=c1.c0 z1=c1.c1 z5=c1.c2; each output pair depends on one input pair only
  Fp2 t0, t1;
  {
    Fp2 z0 = a.c0.c0, z1 = a.c1.c1;
    fp4_sqr(t0, t1, z0, z1);
    r.c0.c0 = add(dbl(sub(t0, z0)), t0);  // 3 t0 - 2 z0
    r.c1.c1 = add(dbl(add(t1, z1)), t1);  // 3 t1 + 2 z1
  }
  Fp2 z2 = a.c1.c0, z3 = a.c0.c2, z4 = a.c0.c1, z5 = a.c1.c2;
  Fp2 t2, t3;
  fp4_sqr(t2, t3, z2, z3);
  fp4_sqr(t0, t1, z4, z5);  // t4, t5
  Fp2 tmp = mul_xi(t1);
  r.c1.c0 = add(dbl(add(tmp, z2)), tmp);
  r.c0.c2 = add(dbl(sub(t0, z3)), t0);
  r.c0.c1 = add(dbl(sub(t2, z4)), t2);
  r.c1.c2 = add(dbl(add(t3, z5)), t3);
}

// Frobenius^k, k in {1,2,3}: coefficient a_i of w^i -> conj^k(a_i) * xi^(i(p^k-1)/6)
template <int KK>
HD Fp2 frob_coeff(int i) {  // i = 1..5
  Fp2 r;
  if (KK == 1) { BN_LOAD_FP2(r, K::frob1, i - 1); }
  else if (KK == 2) { BN_LOAD_FP2(r, K::frob2, i - 1); }
  else { BN_LOAD_FP2(r, K::frob3, i - 1); }
  return r;
}
template <int KK>
HDN void frobenius(Fp12& r, const Fp12& a) {
  // w-power order: a0=c0.c0, a1=c1.c0, a2=c0.c1, a3=c1.c1, a4=c0.c2, a5=c1.c2 (element-wise: r may alias a)
  if (KK & 1) {
    r.c0.c0 = conj(a.c0.c0);
    r.c1.c0 = mul(conj(a.c1.c0), frob_coeff<KK>(1));
    r.c0.c1 = mul(conj(a.c0.c1), frob_coeff<KK>(2));
    r.c1.c1 = mul(conj(a.c1.c1), frob_coeff<KK>(3));
    r.c0.c2 = mul(conj(a.c0.c2), frob_coeff<KK>(4));
    r.c1.c2 = mul(conj(a.c1.c2), frob_coeff<KK>(5));
  } else {
    // p^2: coefficients lie in Fp (c1 == 0)
    r.c0.c0 = a.c0.c0;
    r.c1.c0 = scale(a.c1.c0, frob_coeff<KK>(1).c0);
    r.c0.c1 = scale(a.c0.c1, frob_coeff<KK>(2).c0);
    r.c1.c1 = scale(a.c1.c1, frob_coeff<KK>(3).c0);
    r.c0.c2 = scale(a.c0.c2, frob_coeff<KK>(4).c0);
    r.c1.c2 = scale(a.c1.c2, frob_coeff<KK>(5).c0);
  }
}

// canonical serialisation: 12 x 32-byte BE, order c0.c0.c0, c0.c0.c1, c0.c1.c0, ..., c1.c2.c1
HDN void fp12_to_bytes(uint8_t* out, const Fp12& a) {
  const Fp2* cs[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
  for (int i = 0; i < 6; i++) {
    fe_to_be_bytes(out + 64 * i, fe_from_mont(cs[i]->c0));
    fe_to_be_bytes(out + 64 * i + 32, fe_from_mont(cs[i]->c1));
  }
}

}  // namespace bn254
