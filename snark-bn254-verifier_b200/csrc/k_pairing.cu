// Raw k-pair pairing-product kernels (bn::pairing_batch; reference call sites verifier/src/groth16/verify.rs:70-77,
// verifier/src/plonk/kzg.rs:180-187), one set per thread.  sm_100a only.
#include "kernels.h"
#include "trio.cuh"

namespace bn254 {
namespace {

template <int KP, int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_pairing_product(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, size_t n,
                      uint8_t* __restrict__ is_one, uint8_t* miller_out, uint8_t* gt_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;  // spare threads walk the last set (phase barriers inside) and write nothing
  bool one = pairing_product_one<KP>(g1 + (size_t)64 * KP * i, g2 + (size_t)128 * KP * i,
                                     live && miller_out ? miller_out + 384 * i : nullptr,
                                     live && gt_out ? gt_out + 384 * i : nullptr);
  if (live) is_one[i] = one ? 1 : 0;
}

// three lanes per set (trio.cuh), for batches that cannot fill the GPU with one set per thread
template <int KP, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_pairing_product3(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, size_t n,
                       uint8_t* __restrict__ is_one, uint8_t* miller_out, uint8_t* gt_out) {
  size_t i = trio::trio_slot();
  const bool live = trio::trio_lane_valid() && i < n;
  if (!live) i = 0;
  G1Aff p[KP];
  G2Aff q[KP];
  uint32_t skip = 0;
  for (int j = 0; j < KP; j++) {
    const uint8_t *b1 = g1 + (size_t)64 * (KP * i + j), *b2 = g2 + (size_t)128 * (KP * i + j);
    if (all_zero_bytes(b1, 64) || all_zero_bytes(b2, 128)) {
      skip |= 1u << j;
      p[j] = g1_generator();
      q[j] = g2_generator_dev();
    } else {
      load_g1_unchecked(p[j], b1);
      load_g2_unchecked(q[j], b2);
    }
  }
  trio::S12 f;
  trio::miller_loop_var_s<KP>(f, p, q, skip);
  if (live && miller_out) trio::fp12s_to_bytes(miller_out + 384 * i, f);
  trio::fp12s_final_exponentiation(f, f);
  if (live && gt_out) trio::fp12s_to_bytes(gt_out + 384 * i, f);
  const bool one = trio::fp12s_eq(f, trio::fp12s_one());
  if (live && trio::lane_j() == 0) is_one[i] = one ? 1 : 0;
}

template <int KP>
void launch_k(cudaStream_t st, const uint8_t* g1, const uint8_t* g2, size_t m, uint8_t* is_one, uint8_t* ml, uint8_t* gt,
              int shape) {
#define PP(TPB) k_pairing_product<KP, TPB><<<(unsigned)((m + TPB - 1) / TPB), TPB, 0, st>>>(g1, g2, m, is_one, ml, gt)
  if (shape < 0) {  // three lanes per set
    constexpr int TPB = 128;
    const unsigned per_block = (TPB / 32) * BN_TRIOS_PER_WARP;
    k_pairing_product3<KP, TPB, 3><<<(unsigned)((m + per_block - 1) / per_block), TPB, trio::trio_smem_bytes(TPB), st>>>(
        g1, g2, m, is_one, ml, gt);
    return;
  }
  switch (shape) {
    case launch::SHAPE_448: case launch::SHAPE_384: PP(448); break;
    case launch::SHAPE_32: PP(32); break;
    default: PP(128); break;
  }
#undef PP
}

}  // namespace

namespace launch {

int pairing_product(cudaStream_t st, int k, const uint8_t* g1, const uint8_t* g2, size_t m, uint8_t* is_one,
                    uint8_t* miller_out, uint8_t* gt_out, int sm_count) {
  const int shape = m <= trio_max_items(sm_count) ? -1 : pick_shape(m, sm_count);
  switch (k) {
    case 1: launch_k<1>(st, g1, g2, m, is_one, miller_out, gt_out, shape); break;
    case 2: launch_k<2>(st, g1, g2, m, is_one, miller_out, gt_out, shape); break;
    case 3: launch_k<3>(st, g1, g2, m, is_one, miller_out, gt_out, shape); break;
    default: launch_k<4>(st, g1, g2, m, is_one, miller_out, gt_out, shape); break;
  }
  return 1;
}

}  // namespace launch
}  // namespace bn254
