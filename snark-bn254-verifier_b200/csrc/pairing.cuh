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
struct LinePairKF {  // stored coefficients K0..K8 of the product of two VK-constant lines (pairing_body.inc)
  Fp2 k[9];
};
#define BN_LD_FP2(x) (x)
#define BN_LD_LINE(l) (l)
#include "pairing_body.inc"
#undef BN_LD_LINE
#undef BN_LD_FP2

// Pair table of two VK-constant G2 points from their line tables (once per VK).
HDN void line_pair_table(LinePairKF* out, const Line* t1, const Line* t2) {
  for (int i = 0; i < BN_N_LINES; i++) {
    const Fp2 &x0 = t1[i].ell_0, &a = t1[i].ell_vv, &c = t1[i].ell_vw;
    const Fp2 &y0 = t2[i].ell_0, &b = t2[i].ell_vv, &d = t2[i].ell_vw;
    out[i].k[0] = mul(x0, y0);
    out[i].k[1] = mul_xi(mul(c, d));
    out[i].k[2] = mul_xi(mul(a, b));
    out[i].k[3] = mul(x0, b);
    out[i].k[4] = mul(a, y0);
    out[i].k[5] = mul_xi(mul(a, d));
    out[i].k[6] = mul_xi(mul(c, b));
    out[i].k[7] = mul(x0, d);
    out[i].k[8] = mul(c, y0);
  }
}

}  // namespace bn254
