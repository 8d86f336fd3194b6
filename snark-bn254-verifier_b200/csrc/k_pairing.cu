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

// Big batches as two launches (Miller loops -> fbuf | final exponentiations -> verdicts), as the Groth16 kernels do:
// each kernel has half the code and stack of the fused one, and both fit the 384-thread x 168-register shape.
template <int KP, int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_pairing_miller(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, size_t n, Fp12* __restrict__ fbuf,
                     uint8_t* miller_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;  // spare threads walk the last set (phase barriers inside) and write nothing
  G1Aff p[KP];
  G2Aff q[KP];
  const uint32_t skip = pairing_product_load<KP>(p, q, g1 + (size_t)64 * KP * i, g2 + (size_t)128 * KP * i);
  Fp12 f;
  miller_loop<KP, 0>(f, p, q, nullptr, nullptr, skip);
  if (!live) return;
  if (miller_out) fp12_to_bytes(miller_out + 384 * i, f);
  fbuf[i] = f;
}
template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_pairing_finish(size_t n, const Fp12* __restrict__ fbuf, uint8_t* __restrict__ is_one, uint8_t* gt_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;
  Fp12 f = fbuf[i];
  final_exponentiation(f, f);
  if (!live) return;
  if (gt_out) fp12_to_bytes(gt_out + 384 * i, f);
  is_one[i] = eq(f, fp12_one()) ? 1 : 0;
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
int launch_k(cudaStream_t st, const uint8_t* g1, const uint8_t* g2, size_t m, uint8_t* is_one, uint8_t* ml, uint8_t* gt,
             int shape, Fp12* fbuf, cudaEvent_t mid) {
#define PP(TPB) k_pairing_product<KP, TPB><<<(unsigned)((m + TPB - 1) / TPB), TPB, 0, st>>>(g1, g2, m, is_one, ml, gt)
  if (shape < 0) {  // three lanes per set
    constexpr int TPB = 128;
    const unsigned per_block = (TPB / 32) * BN_TRIOS_PER_WARP;
    k_pairing_product3<KP, TPB, 3><<<(unsigned)((m + per_block - 1) / per_block), TPB, trio::trio_smem_bytes(TPB), st>>>(
        g1, g2, m, is_one, ml, gt);
    return 1;
  }
  if (fbuf && (shape == launch::SHAPE_448 || shape == launch::SHAPE_384)) {
    if (shape == launch::SHAPE_448) {
      k_pairing_miller<KP, 448><<<(unsigned)((m + 447) / 448), 448, 0, st>>>(g1, g2, m, fbuf, ml);
      if (mid) cudaEventRecord(mid, st);
      k_pairing_finish<448><<<(unsigned)((m + 447) / 448), 448, 0, st>>>(m, fbuf, is_one, gt);
    } else {
      k_pairing_miller<KP, 384><<<(unsigned)((m + 383) / 384), 384, 0, st>>>(g1, g2, m, fbuf, ml);
      if (mid) cudaEventRecord(mid, st);
      k_pairing_finish<384><<<(unsigned)((m + 383) / 384), 384, 0, st>>>(m, fbuf, is_one, gt);
    }
    return 2;
  }
  switch (shape) {
    case launch::SHAPE_448: case launch::SHAPE_384: PP(448); break;
    case launch::SHAPE_32: PP(32); break;
    default: PP(128); break;
  }
#undef PP
  return 1;
}

}  // namespace

namespace launch {

// fbuf (m Fp12 of scratch, or null: one fused launch) and mid (an event recorded between the two launches, or null)
int pairing_product(cudaStream_t st, int k, const uint8_t* g1, const uint8_t* g2, size_t m, uint8_t* is_one,
                    uint8_t* miller_out, uint8_t* gt_out, int sm_count, Fp12* fbuf, cudaEvent_t mid) {
  const int shape = m <= trio_max_items(sm_count) ? -1 : pick_shape(m, sm_count);
  switch (k) {
    case 1: return launch_k<1>(st, g1, g2, m, is_one, miller_out, gt_out, shape, fbuf, mid);
    case 2: return launch_k<2>(st, g1, g2, m, is_one, miller_out, gt_out, shape, fbuf, mid);
    case 3: return launch_k<3>(st, g1, g2, m, is_one, miller_out, gt_out, shape, fbuf, mid);
    default: return launch_k<4>(st, g1, g2, m, is_one, miller_out, gt_out, shape, fbuf, mid);
  }
}

}  // namespace launch
}  // namespace bn254
