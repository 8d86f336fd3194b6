// Synthetic-workload generators (BASELINE.json configs 2 and 4) and the integer multiply-add issue-rate probe that is
// the roofline denominator.  Measurement / test support: declared in include/bn254v_bench.h, not part of the verifier.
#include "kernels.h"

namespace bn254 {
namespace {

#define BN_TPB 128

__global__ void __launch_bounds__(BN_TPB)
    k_groth16_synth(Groth16Trapdoor td, uint64_t seed, size_t first, size_t n, int n_public, int sign_mode,
                    uint8_t* proofs, uint8_t* inputs, uint8_t* expected) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  groth16_synth_one(proofs + 256 * i, inputs + (size_t)32 * n_public * i, expected + i, td, seed, first + i,
                    n_public, sign_mode);
}

__global__ void __launch_bounds__(BN_TPB)
    k_pairing_synth(uint64_t seed, size_t first, size_t n, int k, uint8_t* g1, uint8_t* g2, uint8_t* expected) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pairing_synth_one(g1 + (size_t)64 * k * i, g2 + (size_t)128 * k * i, expected + i, seed, first + i, k);
}

// Integer multiply-add issue-rate probe (the roofline denominator): 8 independent accumulator chains per thread,
// 8 warps per SMSP.  Each step is one IMAD.WIDE.U32 with a 64-bit accumulate -- written as the mad.lo.cc / madc.hi
// pair that ptxas fuses, exactly as in fe_mul -- or one 32-bit IMAD.  The multiplier a[u] changes every iteration
// (one IADD per 8 MACs) so that ptxas can neither hoist the products nor strength-reduce the loop; SASS checked:
// 32 IMAD.WIDE.U32 (or IMAD) + 4 IADD3 per unrolled iteration.  Measured on B200 at 1965 MHz: 8.69 T wide MAC/s
// (one warp-wide IMAD.WIDE per 4.3 cycles per SMSP) and 18.5 T 32-bit IMAD/s (one per 2.0 cycles).
template <bool WIDE>
__global__ void __launch_bounds__(256) k_imad_peak(int iters, uint32_t a0, uint32_t b0, uint64_t* sink) {
  uint32_t lo[8], hi[8], x[8], a[4];
#pragma unroll
  for (int u = 0; u < 4; u++) a[u] = a0 * (u + 1) + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    lo[j] = blockIdx.x;
    hi[j] = b0 + j;
    x[j] = (j + 1) * b0 + threadIdx.x;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (WIDE)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                       : "+r"(lo[j]), "+r"(hi[j])
                       : "r"(x[j]), "r"(a[u]));
        else
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(x[j]), "r"(a[u]));
      }
      a[u] += 0x9e3779b9u;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j];
  if (s == 0x12345678u) sink[0] = s;
}

}  // namespace

namespace launch {

int groth16_synth(cudaStream_t st, const Groth16Trapdoor& td, uint64_t seed, size_t first, size_t n, int n_public,
                  int sign_mode, uint8_t* proofs, uint8_t* inputs, uint8_t* expected) {
  unsigned grid = (unsigned)((n + BN_TPB - 1) / BN_TPB);
  k_groth16_synth<<<grid, BN_TPB, 0, st>>>(td, seed, first, n, n_public, sign_mode, proofs, inputs, expected);
  return 1;
}
int pairing_synth(cudaStream_t st, uint64_t seed, size_t first, size_t n, int k, uint8_t* g1, uint8_t* g2,
                  uint8_t* expected) {
  unsigned grid = (unsigned)((n + BN_TPB - 1) / BN_TPB);
  k_pairing_synth<<<grid, BN_TPB, 0, st>>>(seed, first, n, k, g1, g2, expected);
  return 1;
}
int imad_peak(cudaStream_t st, bool wide, int blocks, int threads, int iters, uint64_t* sink) {
  if (wide) k_imad_peak<true><<<blocks, threads, 0, st>>>(iters, 12345u, 6789u, sink);
  else k_imad_peak<false><<<blocks, threads, 0, st>>>(iters, 12345u, 6789u, sink);
  return 1;
}

}  // namespace launch
}  // namespace bn254
