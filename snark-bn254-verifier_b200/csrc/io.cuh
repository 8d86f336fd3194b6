// Wire-format <-> Montgomery conversion and point validation.
// Device-side restatement of the per-proof part of the reference's parsers:
//   uncompressed_bytes_to_g1_point  verifier/src/converter.rs:78-88   (Fq::from_slice, AffineG1::new)
//   uncompressed_bytes_to_g2_point  verifier/src/converter.rs:135-153 (x1|x0|y1|y0, AffineG2::new)
// Status codes are those of include/bn254v.h.
#pragma once
#include "../../include/bn254v.h"
#include "curve.cuh"

namespace bn254 {

// 32-byte BE -> Montgomery Fp; false when >= p
HD bool fp_load_be(Fp& out, const uint8_t* b) {
  Fp t;
  bool ok = fe_from_be_bytes(t, b);
  out = fe_to_mont(t);
  return ok;
}
HD void fp_store_be(uint8_t* b, const Fp& a) { fe_to_be_bytes(b, fe_from_mont(a)); }

// 32-byte BE -> plain Fr limbs (not Montgomery); false when >= r
HD bool fr_load_be_plain(Fr& out, const uint8_t* b) { return fe_from_be_bytes(out, b); }

HD int load_g1_checked(G1Aff& p, const uint8_t* b) {
  bool ok = fp_load_be(p.x, b);
  ok = fp_load_be(p.y, b + 32) && ok;
  if (!ok) return BN254V_PANIC_FIELD_NOT_MEMBER;
  if (!on_curve(p)) return BN254V_PANIC_NOT_ON_CURVE;
  return BN254V_OK_TRUE;
}
HD void load_g1_unchecked(G1Aff& p, const uint8_t* b) {
  fp_load_be(p.x, b);
  fp_load_be(p.y, b + 32);
}
HD void load_g2_unchecked(G2Aff& q, const uint8_t* b) {
  fp_load_be(q.x.c1, b);
  fp_load_be(q.x.c0, b + 32);
  fp_load_be(q.y.c1, b + 64);
  fp_load_be(q.y.c0, b + 96);
}
// AffineG2::new without its subgroup test: field membership and the curve equation only.  The Groth16 path gets the
// subgroup verdict from the end point of the Miller loop (pairing_body.inc, ate_endpoint_in_g2).
HD int load_g2_on_curve(G2Aff& q, const uint8_t* b) {
  bool ok = fp_load_be(q.x.c1, b);
  ok = fp_load_be(q.x.c0, b + 32) && ok;
  ok = fp_load_be(q.y.c1, b + 64) && ok;
  ok = fp_load_be(q.y.c0, b + 96) && ok;
  if (!ok) return BN254V_PANIC_FIELD_NOT_MEMBER;
  if (!on_curve(q)) return BN254V_PANIC_NOT_ON_CURVE;
  return BN254V_OK_TRUE;
}
HD int load_g2_checked(G2Aff& q, const uint8_t* b) {
  int st = load_g2_on_curve(q, b);
  if (st != BN254V_OK_TRUE) return st;
  if (!g2_in_subgroup<false>(q)) return BN254V_PANIC_NOT_IN_SUBGROUP;
  return BN254V_OK_TRUE;
}
HD void store_g1(uint8_t* b, const G1Aff& p) {
  fp_store_be(b, p.x);
  fp_store_be(b + 32, p.y);
}
HD void store_g2(uint8_t* b, const G2Aff& q) {
  fp_store_be(b, q.x.c1);
  fp_store_be(b + 32, q.x.c0);
  fp_store_be(b + 64, q.y.c1);
  fp_store_be(b + 96, q.y.c0);
}

}  // namespace bn254
