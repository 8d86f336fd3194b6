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
typedef Line LineF;  // stored form of a line (VK tables); identical to the compute form in the scalar build
#define BN_LD_LINE(l) (l)
#include "pairing_body.inc"
#undef BN_LD_LINE

}  // namespace bn254
