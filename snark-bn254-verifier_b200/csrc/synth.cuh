// Deterministic synthetic workloads (BASELINE.json configs 2 and 4), generated on device.
// Same PRNG / scalar definitions as the oracle's generator (oracle/bn254_oracle.py: splitmix64,
// synth_scalar, Groth16Trapdoor) so the two can be compared byte for byte in the parity tests.
// Trapdoor Groth16 proofs solve the reference's equation (verifier/src/groth16/verify.rs:70-77):
//   c = (a b + l gamma + alpha beta) / delta        (sign_mode 0)
//   c = (a b - l gamma - alpha beta) / delta        (sign_mode 1, gnark)
#pragma once
#include "io.cuh"

namespace bn254 {

HD uint64_t splitmix64_next(uint64_t& state) {
  state += 0x9E3779B97F4A7C15ull;
  uint64_t z = state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// plain (non-Montgomery) scalar in [1, r)
HD Fr synth_scalar(uint64_t seed, uint64_t index, uint64_t slot) {
  uint64_t st = seed * 0xD1342543DE82EF95ull + index * 0x2545F4914F6CDD1Dull + slot * 0x9E3779B97F4A7C15ull +
                0x632BE59BD9B4E019ull;
  Fr v;
  for (int w = 0; w < 4; w++) {
    uint64_t z = splitmix64_next(st);
    v.v[2 * w] = (uint32_t)z;
    v.v[2 * w + 1] = (uint32_t)(z >> 32);
  }
  fe_reduce_full(v);
  if (fe_is_zero(v)) v.v[0] = 1;
  return v;
}

#define BN_SYNTH_VK_INDEX (1ull << 40)

#define BN_MAX_IC_SYNTH 9

struct Groth16Trapdoor {  // Montgomery-form scalars
  Fr alpha, beta, gamma, delta, delta_inv;
  Fr ic[BN_MAX_IC_SYNTH];
};

HD void trapdoor_init(Groth16Trapdoor& td, uint64_t seed, int n_public) {
  td.alpha = fe_to_mont(synth_scalar(seed, BN_SYNTH_VK_INDEX, 0));
  td.beta = fe_to_mont(synth_scalar(seed, BN_SYNTH_VK_INDEX, 1));
  td.gamma = fe_to_mont(synth_scalar(seed, BN_SYNTH_VK_INDEX, 2));
  td.delta = fe_to_mont(synth_scalar(seed, BN_SYNTH_VK_INDEX, 3));
  td.delta_inv = fe_inv(td.delta);
  for (int i = 0; i <= n_public; i++) td.ic[i] = fe_to_mont(synth_scalar(seed, BN_SYNTH_VK_INDEX, 4 + i));
}

HD void synth_corruption(bool& bad, int& klass, uint64_t seed, uint64_t index) {
  uint64_t j = index >> 1;
  uint64_t st = seed ^ (j * 0xA24BAED4963EE407ull) ^ 0x9FB21C651E98DF25ull;
  uint64_t z = splitmix64_next(st);
  bad = (index & 1) == (z & 1);
  klass = (int)(j % 5);
}

// xs: plain limbs; a, b, c: Montgomery
HD void synth_base_scalars(Fr* xs, Fr& a, Fr& b, Fr& c, const Groth16Trapdoor& td, uint64_t seed, uint64_t index,
                           int n_public, int sign_mode) {
  for (int i = 0; i < n_public; i++) xs[i] = synth_scalar(seed, index, 8 + i);
  xs[0].v[7] &= 0x00ffffffu;  // SP1's vkey hash is 31 bytes
  if (fe_is_zero(xs[0])) xs[0].v[0] = 1;
  a = fe_to_mont(synth_scalar(seed, index, 0));
  b = fe_to_mont(synth_scalar(seed, index, 1));
  Fr ell = td.ic[0];
  for (int i = 0; i < n_public; i++) ell = fe_add(ell, fe_mul(fe_to_mont(xs[i]), td.ic[i + 1]));
  Fr ab = fe_mul(a, b), lg = fe_mul(ell, td.gamma), al = fe_mul(td.alpha, td.beta);
  Fr num = sign_mode == 0 ? fe_add(fe_add(ab, lg), al) : fe_sub(fe_sub(ab, lg), al);
  c = fe_mul(num, td.delta_inv);
}

HD void store_g1_mul_gen(uint8_t* out, const Fr& k_mont) {
  Fr k = fe_from_mont(k_mont);
  G1Aff p;
  to_affine(p, scalar_mul(g1_generator(), k.v));
  store_g1(out, p);
}
HD void store_g2_mul_gen(uint8_t* out, const Fr& k_mont) {
  Fr k = fe_from_mont(k_mont);
  G2Aff q;
  to_affine(q, scalar_mul(g2_generator_dev(), k.v));
  store_g2(out, q);
}

// One synthetic proof: 256 proof bytes, n_public*32 input bytes, expected status.
HD void groth16_synth_one(uint8_t* proof, uint8_t* inputs, uint8_t* expected, const Groth16Trapdoor& td,
                          uint64_t seed, uint64_t index, int n_public, int sign_mode) {
  Fr xs[BN_MAX_IC_SYNTH], a, b, c;
  synth_base_scalars(xs, a, b, c, td, seed, index, n_public, sign_mode);
  bool bad;
  int klass;
  synth_corruption(bad, klass, seed, index);
  if (bad) {
    if (klass == 0) {
      Fr one = fe_zero<FrCfg>();
      one.v[0] = 1;
      xs[0] = fe_add(xs[0], one);  // plain limbs: fe_add is representation-agnostic
    } else if (klass == 1) {
      a = fe_dbl(a);
    } else if (klass == 2) {
      c = fe_neg(c);
    } else if (klass == 3) {
      b = fe_dbl(b);
    } else {
      Fr xs2[BN_MAX_IC_SYNTH], a2, b2;
      synth_base_scalars(xs2, a2, b2, c, td, seed, index ^ 1, n_public, sign_mode);
    }
  }
  store_g1_mul_gen(proof, a);
  store_g2_mul_gen(proof + 64, b);
  store_g1_mul_gen(proof + 192, c);
  for (int i = 0; i < n_public; i++) fe_to_be_bytes(inputs + 32 * i, xs[i]);
  *expected = bad ? BN254V_OK_FALSE : BN254V_OK_TRUE;
}

// k-pair set: P_j = s_j G1, Q_j = t_j G2; odd indices (k > 1) are solved so that sum s_j t_j = 0.
HD void pairing_synth_one(uint8_t* g1, uint8_t* g2, uint8_t* expected_is_one, uint64_t seed, uint64_t index, int k) {
  Fr s[4], t[4];
  for (int j = 0; j < k; j++) {
    s[j] = fe_to_mont(synth_scalar(seed, index, 2 * j));
    t[j] = fe_to_mont(synth_scalar(seed, index, 2 * j + 1));
  }
  bool one = (index & 1) && k > 1;
  if (one) {
    Fr acc = fe_zero<FrCfg>();
    for (int j = 0; j < k - 1; j++) acc = fe_add(acc, fe_mul(s[j], t[j]));
    s[k - 1] = fe_mul(fe_neg(acc), fe_inv(t[k - 1]));
  }
  for (int j = 0; j < k; j++) {
    store_g1_mul_gen(g1 + 64 * j, s[j]);
    store_g2_mul_gen(g2 + 128 * j, t[j]);
  }
  *expected_is_one = one ? 1 : 0;
}

}  // namespace bn254
