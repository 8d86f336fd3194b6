// Host build of the kernels' __host__ __device__ code (carry flag emulated, see field.cuh).
// TEST-ONLY: lets the CPU test-suite exercise the exact per-proof routines the CUDA kernels run,
// against the oracle, in a container without a GPU.  Not part of the product library.
#include <string.h>

#include "../../snark-bn254-verifier_b200/csrc/groth16.cuh"

using namespace bn254;

extern "C" {

void hs_fp_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  Fp x, y;
  memcpy(x.v, a, 32);
  memcpy(y.v, b, 32);
  Fp z = fe_mul(x, y);
  memcpy(r, z.v, 32);
}
void hs_fr_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  Fr x, y;
  memcpy(x.v, a, 32);
  memcpy(y.v, b, 32);
  Fr z = fe_mul(x, y);
  memcpy(r, z.v, 32);
}

int hs_pairing_product(int k, const uint8_t* g1, const uint8_t* g2, uint8_t* miller_out, uint8_t* gt_out) {
  switch (k) {
    case 1: return pairing_product_one<1>(g1, g2, miller_out, gt_out);
    case 2: return pairing_product_one<2>(g1, g2, miller_out, gt_out);
    case 3: return pairing_product_one<3>(g1, g2, miller_out, gt_out);
    case 4: return pairing_product_one<4>(g1, g2, miller_out, gt_out);
  }
  return -1;
}

int hs_g2_check(const uint8_t* g2) {
  G2Aff q;
  return load_g2_checked(q, g2);
}
int hs_g1_check(const uint8_t* g1) {
  G1Aff p;
  return load_g1_checked(p, g1);
}

// vk_points: alpha(64) | beta'(128) | gamma'(128) | delta'(128) | ic[n_ic](64 each), uncompressed BE
void* hs_groth16_vk_new(const uint8_t* vk_points, int n_ic) {
  Groth16VkDev* vk = new Groth16VkDev();
  vk->n_ic = n_ic;
  load_g1_unchecked(vk->alpha, vk_points);
  load_g2_unchecked(vk->beta, vk_points + 64);
  load_g2_unchecked(vk->gamma, vk_points + 192);
  load_g2_unchecked(vk->delta, vk_points + 320);
  for (int i = 0; i < n_ic; i++) load_g1_unchecked(vk->ic[i], vk_points + 448 + 64 * i);
  groth16_vk_prepare(*vk);
  return vk;
}
void hs_groth16_vk_free(void* vk) { delete (Groth16VkDev*)vk; }
void hs_groth16_vk_target(void* vk, uint8_t* out) { fp12_to_bytes(out, ((Groth16VkDev*)vk)->target); }

int hs_groth16_verify(void* vk, const uint8_t* proof, uint32_t proof_len, const uint8_t* inputs, int n_inputs,
                      uint8_t* L, uint8_t* miller, uint8_t* gt) {
  Groth16Debug dbg{L, miller, gt};
  return groth16_verify_one(*(Groth16VkDev*)vk, proof, proof_len, inputs, n_inputs, dbg);
}
}
