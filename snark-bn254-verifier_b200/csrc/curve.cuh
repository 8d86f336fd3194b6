// G1 (y^2 = x^3 + 3 over Fq) and G2 (y^2 = x^3 + 3/xi over Fq2) group operations.
// Replaces bn::{AffineG1, AffineG2, G1, G2}: `AffineG::new` validation (reference
// verifier/src/converter.rs:78-88,135-153), `AffineG1 * Fr`, `+` (verifier/src/groth16/verify.rs:61),
// `G1 * Fr` (verifier/src/plonk/kzg.rs:169) and `AffineG1::msm` (verifier/src/plonk/verify.rs:284).
#pragma once
#include "tower.cuh"

namespace bn254 {

#include "curve_body.inc"

}  // namespace bn254
