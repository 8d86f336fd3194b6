// G1 (y^2 = x^3 + 3 over Fq) and G2 (y^2 = x^3 + 3/xi over Fq2) group operations.
// Replaces bn::{AffineG1, AffineG2, G1, G2}: `AffineG::new` validation (reference
// verifier/src/converter.rs:78-88,135-153), `AffineG1 * Fr`, `+` (verifier/src/groth16/verify.rs:61),
// `G1 * Fr` (verifier/src/plonk/kzg.rs:169) and `AffineG1::msm` (verifier/src/plonk/verify.rs:284).
#pragma once
#include "tower.cuh"

namespace bn254 {

template <class F>
struct Aff {
  F x, y;
};
template <class F>
struct Jac {
  F x, y, z;  // z == 0 <=> identity
};
typedef Aff<Fp> G1Aff;
typedef Aff<Fp2> G2Aff;
typedef Jac<Fp> G1Jac;
typedef Jac<Fp2> G2Jac;

template <class F> HD F f_one();
template <> HD Fp f_one<Fp>() { return fe_one<FpCfg>(); }
template <> HD Fp2 f_one<Fp2>() { return fp2_one(); }
template <class F> HD F f_zero();
template <> HD Fp f_zero<Fp>() { return fe_zero<FpCfg>(); }
template <> HD Fp2 f_zero<Fp2>() { return fp2_zero(); }
template <class F> HD F curve_b();
template <> HD Fp curve_b<Fp>() { return fp_three(); }
template <> HD Fp2 curve_b<Fp2>() { return fp2_b2(); }

template <class F>
HD Jac<F> jac_identity() {
  return Jac<F>{f_one<F>(), f_one<F>(), f_zero<F>()};
}
template <class F>
HD Jac<F> to_jac(const Aff<F>& p) {
  return Jac<F>{p.x, p.y, f_one<F>()};
}
template <class F>
HD bool is_identity(const Jac<F>& p) {
  return is_zero(p.z);
}
template <class F>
HD bool on_curve(const Aff<F>& p) {
  return eq(sqr(p.y), add(mul(sqr(p.x), p.x), curve_b<F>()));
}
template <class F>
HD Aff<F> neg(const Aff<F>& p) {
  return Aff<F>{p.x, neg(p.y)};
}

// dbl-2009-l (a = 0): 2M + 5S
template <class F>
HDN Jac<F> jac_double(const Jac<F>& p) {
  F A = sqr(p.x);
  F B = sqr(p.y);
  F C = sqr(B);
  F D = dbl(sub(sub(sqr(add(p.x, B)), A), C));
  F E = add(dbl(A), A);
  F X3 = sub(sqr(E), dbl(D));
  F C8 = dbl(dbl(dbl(C)));
  F Y3 = sub(mul(E, sub(D, X3)), C8);
  F Z3 = dbl(mul(p.y, p.z));
  return Jac<F>{X3, Y3, Z3};  // z == 0 stays 0
}

// madd-2007-bl mixed addition (q affine, not identity), complete w.r.t. p == identity, p == +-q.
template <class F>
HDN Jac<F> jac_add_mixed(const Jac<F>& p, const Aff<F>& q) {
  if (is_identity(p)) return to_jac(q);
  F Z1Z1 = sqr(p.z);
  F U2 = mul(q.x, Z1Z1);
  F S2 = mul(mul(q.y, p.z), Z1Z1);
  F H = sub(U2, p.x);
  F rr = sub(S2, p.y);
  if (is_zero(H)) {
    if (is_zero(rr)) return jac_double(p);
    return jac_identity<F>();
  }
  F HH = sqr(H);
  F I = dbl(dbl(HH));
  F J = mul(H, I);
  F r2 = dbl(rr);
  F V = mul(p.x, I);
  F X3 = sub(sub(sqr(r2), J), dbl(V));
  F Y3 = sub(mul(r2, sub(V, X3)), dbl(mul(p.y, J)));
  F Z3 = sub(sub(sqr(add(p.z, H)), Z1Z1), HH);
  return Jac<F>{X3, Y3, Z3};
}

// add-2007-bl full Jacobian addition
template <class F>
HDN Jac<F> jac_add(const Jac<F>& p, const Jac<F>& q) {
  if (is_identity(p)) return q;
  if (is_identity(q)) return p;
  F Z1Z1 = sqr(p.z);
  F Z2Z2 = sqr(q.z);
  F U1 = mul(p.x, Z2Z2);
  F U2 = mul(q.x, Z1Z1);
  F S1 = mul(mul(p.y, q.z), Z2Z2);
  F S2 = mul(mul(q.y, p.z), Z1Z1);
  F H = sub(U2, U1);
  F rr = sub(S2, S1);
  if (is_zero(H)) {
    if (is_zero(rr)) return jac_double(p);
    return jac_identity<F>();
  }
  F I = sqr(dbl(H));
  F J = mul(H, I);
  F r2 = dbl(rr);
  F V = mul(U1, I);
  F X3 = sub(sub(sqr(r2), J), dbl(V));
  F Y3 = sub(mul(r2, sub(V, X3)), dbl(mul(S1, J)));
  F Z3 = mul(sub(sub(sqr(add(p.z, q.z)), Z1Z1), Z2Z2), H);
  return Jac<F>{X3, Y3, Z3};
}

// Returns false for the identity (substrate-bn: "Unable to convert G1 to AffineG1").
template <class F>
HDN bool to_affine(Aff<F>& out, const Jac<F>& p) {
  if (is_identity(p)) return false;
  F zi = inv(p.z);
  F zi2 = sqr(zi);
  out.x = mul(p.x, zi2);
  out.y = mul(p.y, mul(zi2, zi));
  return true;
}

// MSB-first double-and-add over a 256-bit plain scalar (8 LE words).
template <class F, bool SYNC = false>
HDN Jac<F> scalar_mul(const Aff<F>& p, const uint32_t* k) {
  Jac<F> acc = jac_identity<F>();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (SYNC && (i & 3) == 3) BN_PHASE_SYNC();
    if (started) acc = jac_double(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) {
      acc = jac_add_mixed(acc, p);
      started = true;
    }
  }
  return acc;
}

// ---- G2 subgroup membership.  substrate-bn checks [r-1]P + P == 0 by a 254-bit scalar
// multiplication; for BN curves the identical predicate is  psi(P) == [6x^2]P  where psi is the
// untwist-Frobenius-twist endomorphism (the eigenvalue of psi on G2 is p = 6x^2 mod r, and no other
// point of E'(Fq2) satisfies it).  6x^2 is 127 bits: half the doublings.
HD G2Aff g2_psi(const G2Aff& q) {
  Fp2 cx, cy;
  BN_LOAD_FP2(cx, K::frob1, 1);  // xi^((p-1)/3)
  BN_LOAD_FP2(cy, K::frob1, 2);  // xi^((p-1)/2)
  return G2Aff{mul(conj(q.x), cx), mul(conj(q.y), cy)};
}
// Reference predicate, 127-bit scalar: psi(P) == [6x^2]P (kept for cross-checking the faster test below).
HDN bool g2_in_subgroup_6x2(const G2Aff& q) {
  // 6x^2 = 0x6f4d8248eeb859fbf83e9682e87cfd46 (127 bits).  Exactness: psi satisfies
  // psi^2 - t psi + p = 0 and gcd((6x^2)^2 - t 6x^2 + p, #E'(Fq2)/r) = 1, so the test forces ord(P) | r.
  const uint32_t k[8] = {0xe87cfd46u, 0xf83e9682u, 0xeeb859fbu, 0x6f4d8248u, 0, 0, 0, 0};
  G2Jac lhs = scalar_mul<Fp2, false>(q, k);
  G2Aff ps = g2_psi(q);
  if (is_identity(lhs)) return false;
  Fp2 z2 = sqr(lhs.z);
  return eq(lhs.x, mul(ps.x, z2)) && eq(lhs.y, mul(ps.y, mul(z2, lhs.z)));
}

HD G2Jac g2_psi_jac(const G2Jac& p) {  // psi in Jacobian coordinates: conjugation commutes with X/Z^2, Y/Z^3
  Fp2 cx, cy;
  BN_LOAD_FP2(cx, K::frob1, 1);
  BN_LOAD_FP2(cy, K::frob1, 2);
  return G2Jac{mul(conj(p.x), cx), mul(conj(p.y), cy), conj(p.z)};
}
template <class F>
HD bool jac_eq(const Jac<F>& a, const Jac<F>& b) {
  bool ia = is_identity(a), ib = is_identity(b);
  if (ia || ib) return ia && ib;
  F za2 = sqr(a.z), zb2 = sqr(b.z);
  return eq(mul(a.x, zb2), mul(b.x, za2)) && eq(mul(a.y, mul(zb2, b.z)), mul(b.y, mul(za2, a.z)));
}
// BN subgroup test with a 63-bit scalar (El Housni-Guillevic-Piellard, "Co-factor clearing and subgroup membership
// testing on pairing-friendly curves", the test gnark-crypto uses for BN254):
//   P in G2  <=>  [x+1]P + psi([x]P) + psi^2([x]P) == psi^3([2x]P).
// Same predicate as substrate-bn's [r]P == 0 on every point of E'(Fq2), at a quarter of the doublings.
HDN bool g2_in_subgroup(const G2Aff& q) {
  const uint32_t k[8] = {0x4a6909f1u, 0x44e992b4u, 0, 0, 0, 0, 0, 0};  // x = 0x44e992b44a6909f1
  G2Jac a = scalar_mul<Fp2, true>(q, k);  // [x]P
  G2Jac b = g2_psi_jac(a);                 // psi([x]P)
  G2Jac c = g2_psi_jac(b);                 // psi^2([x]P)
  G2Jac d = g2_psi_jac(c);                 // psi^3([x]P)
  G2Jac lhs = jac_add(jac_add(jac_add_mixed(a, q), b), c);
  return jac_eq(lhs, jac_double(d));
}

}  // namespace bn254
