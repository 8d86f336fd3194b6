// G1 (y^2 = x^3 + 3 over Fq) and G2 (y^2 = x^3 + 3/xi over Fq2) group operations.
// Replaces bn::{AffineG1, AffineG2, G1, G2}: `AffineG::new` validation (reference
// verifier/src/converter.rs:78-88,135-153), `AffineG1 * Fr`, `+` (verifier/src/groth16/verify.rs:61),
// `G1 * Fr` (verifier/src/plonk/kzg.rs:169) and `AffineG1::msm` (verifier/src/plonk/verify.rs:284).
#pragma once
#include "tower.cuh"

namespace bn254 {

#include "curve_body.inc"

// group generators (substitute inputs for masked / failed proofs, synthetic workloads)
HD G1Aff g1_generator() {
  G1Aff g;
  BN_LOAD_FP(g.x, K::g1_gen, 0);
  BN_LOAD_FP(g.y, K::g1_gen, 1);
  return g;
}
HD G2Aff g2_generator_dev() {
  G2Aff g;
  BN_LOAD_FP(g.x.c0, K::g2_gen, 0);
  BN_LOAD_FP(g.x.c1, K::g2_gen, 1);
  BN_LOAD_FP(g.y.c0, K::g2_gen, 2);
  BN_LOAD_FP(g.y.c1, K::g2_gen, 3);
  return g;
}


}  // namespace bn254
