// Kernels of the opt-in aggregate Groth16 check (groth16_agg.cuh).  sm_100a only.
//   main stream   k_groth16_agg_prepare one proof per thread, small blocks: validate, [r_i] A_i (affine), [r_i] C_i
//                 k_groth16_agg_miller  one proof per thread: the single-pair Miller loop of (r_i A_i, B_i), B in G2
//                 k_groth16_agg_fold_f  product tree over the Miller values
//                 k_groth16_agg_final3  three lanes (trio.cuh): F * f', final exponentiation, == 1
//   side stream   k_groth16_agg_fold_g  sum tree over the [r_i] C_i                             } a handful of threads,
//   (after        k_groth16_agg_points  the batch's three G1 points from the host's scalar sums  } underneath the
//    _prepare)    k_groth16_agg_fprime3 three lanes: the batch's own three-pair Miller loop f'   } proofs' Miller loops
// Same barrier discipline as k_groth16.cu: nothing returns before the last block-wide barrier.
#include "kernels.h"
#include "groth16_agg.cuh"
#include "trio.cuh"

namespace bn254 {
namespace {

__global__ void __launch_bounds__(64, 4)
    k_groth16_agg_prepare(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                          const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs,
                          const uint8_t* __restrict__ rnd16, size_t n, uint8_t* __restrict__ status,
                          Fp12* __restrict__ fbuf, G1Jac* __restrict__ gbuf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;  // (no barriers in this kernel)
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  G1Aff rA;
  G2Aff B;
  G1Jac rc;
  const int st = groth16_agg_prepare_one(rA, B, rc, *vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i,
                                         n_inputs, rnd16 + 16 * i);
  status[i] = (uint8_t)st;
  gbuf[i] = rc;
  if (st == BN254V_OK_TRUE) *(G1Aff*)&fbuf[i] = rA;  // parked for k_groth16_agg_miller
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_groth16_agg_miller(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride, size_t n,
                         uint8_t* __restrict__ status, Fp12* __restrict__ fbuf) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;  // spare threads of the last block walk substitutes and write nothing
  const bool ok = live && status[i] == BN254V_OK_TRUE;
  G1Aff rA = vk->alpha;
  G2Aff B = vk->beta;
  if (ok) {
    rA = *(const G1Aff*)&fbuf[i];
    load_g2_unchecked(B, proofs + stride * i + 64);  // validated by k_groth16_agg_prepare
  }
  Fp12 f;
  const bool in_g2 = groth16_agg_miller_one(f, *vk, rA, B, ok);
  if (!live) return;
  if (ok && !in_g2) status[i] = BN254V_PANIC_NOT_IN_SUBGROUP;
  fbuf[i] = f;
}

// one pass of a tree: out[t] = in[t per] * .. * in[t per + per - 1]
__global__ void __launch_bounds__(64)
    k_groth16_agg_fold_f(const Fp12* __restrict__ in, size_t n, Fp12* __restrict__ out, int per) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t lo = t * per;
  if (lo >= n) return;
  Fp12 f = in[lo];
  for (size_t k = lo + 1; k < n && k < lo + per; k++) mul(f, f, in[k]);
  out[t] = f;
}
__global__ void __launch_bounds__(64)
    k_groth16_agg_fold_g(const G1Jac* __restrict__ in, size_t n, G1Jac* __restrict__ out, int per) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t lo = t * per;
  if (lo >= n) return;
  G1Jac g = in[lo];
  for (size_t k = lo + 1; k < n && k < lo + per; k++) g = jac_add(g, in[k]);
  out[t] = g;
}

struct AggScratch {
  G1Aff nsa, sl, sc;
  int ok;
  Fp12 fprime;
};

__global__ void k_groth16_agg_points(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ scal_be,
                                     const G1Jac* __restrict__ sum_rc, AggScratch* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  out->ok = groth16_agg_points(out->nsa, out->sl, out->sc, *vk, scal_be, *sum_rc) ? 1 : 0;
}

template <int TPB>
__global__ void __launch_bounds__(TPB, 1) k_groth16_agg_fprime3(const Groth16VkDev* __restrict__ vk, AggScratch* sc) {
  const bool live = trio::trio_lane_valid() && trio::trio_slot() == 0;
  G1Aff A = vk->alpha, pf[2] = {vk->ic[0], vk->ic[0]};  // idle trios (and a degenerate batch) walk substitutes
  const G2Aff B = vk->beta;
  if (sc->ok) A = sc->nsa, pf[0] = sc->sl, pf[1] = sc->sc;
  trio::S12 f;
  bool in_g2;
  trio::miller_loop_pairtab1_s(f, A, B, pf, vk->gd_pairs, &in_g2);
  if (live) trio::fp12s_store(sc->fprime, f);
}

template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_groth16_agg_final3(const AggScratch* __restrict__ sc, const Fp12* __restrict__ prod, uint8_t* __restrict__ verdict) {
  const bool live = trio::trio_lane_valid() && trio::trio_slot() == 0;
  trio::S12 f = trio::fp12s_load(sc->fprime);
  const trio::S12 p = trio::fp12s_load(*prod);
  trio::fp12s_mul(f, f, p);
  trio::fp12s_final_exponentiation(f, f);
  const bool one = trio::fp12s_eq(f, trio::fp12s_one());
  if (live && trio::lane_j() == 0) *verdict = (sc->ok && one) ? 1 : 0;
}

// `per` to one per pass, ping-ponging between buf[0, m) and buf[m, ..); returns where the single result lands
template <class T, class K>
T* fold_passes(cudaStream_t st, T* buf, size_t m, K kernel, int* launches) {
  T *a = buf, *b = buf + m;
  size_t cur = m;
  int per = 4;  // the first pass has threads to spare; afterwards depth is what counts
  while (cur > 1) {
    const size_t nxt = (cur + per - 1) / per;
    kernel<<<(unsigned)((nxt + 63) / 64), 64, 0, st>>>(a, cur, b, per);
    T* t = a; a = b; b = t;
    cur = nxt;
    per = 2;
    (*launches)++;
  }
  return a;
}

}  // namespace

namespace launch {

size_t groth16_agg_scratch_bytes() { return sizeof(AggScratch); }
size_t groth16_agg_slots(size_t m) { return m + (m + 3) / 4; }

// main stream, first: validation, [r_i] A_i, [r_i] C_i
int groth16_agg_prepare(cudaStream_t st, const Groth16AggArgs& a) {
  k_groth16_agg_prepare<<<(unsigned)((a.m + 63) / 64), 64, 0, st>>>(a.vk, a.proofs, a.stride, a.lens, a.inputs, a.n_inputs,
                                                                   a.rnd16, a.m, a.status, a.fbuf, a.gbuf);
  return 1;
}

// side stream, once groth16_agg_prepare is done and a.scal_be is on the device: sum tree, points, the batch's Miller value
int groth16_agg_side(cudaStream_t st, const Groth16AggArgs& a) {
  int launches = 0;
  const G1Jac* sum = fold_passes(st, a.gbuf, a.m, k_groth16_agg_fold_g, &launches);
  AggScratch* sc = (AggScratch*)a.scratch;
  k_groth16_agg_points<<<1, 1, 0, st>>>(a.vk, a.scal_be, sum, sc);
  k_groth16_agg_fprime3<32><<<1, 32, trio::trio_smem_bytes(32), st>>>(a.vk, sc);
  return launches + 2;
}

// main stream: the proofs' Miller loops and their product tree
int groth16_agg_miller(cudaStream_t st, const Groth16AggArgs& a, int sm_count, Fp12** product) {
  const size_t m = a.m;
  const int shape = pick_shape(m, sm_count);
#define LA(TPB, MINB)                                                                                                     \
  k_groth16_agg_miller<TPB, MINB><<<(unsigned)((m + TPB - 1) / TPB), TPB, 0, st>>>(a.vk, a.proofs, a.stride, m, a.status, a.fbuf)
  switch (shape) {
    case SHAPE_448: LA(448, 1); break;
    case SHAPE_384: LA(384, 1); break;
    default: LA(128, 2); break;
  }
#undef LA
  int launches = 1;
  *product = fold_passes(st, a.fbuf, m, k_groth16_agg_fold_f, &launches);
  return launches;
}

// main stream, once the side stream is done: the verdict
int groth16_agg_final(cudaStream_t st, const Groth16AggArgs& a, const Fp12* product) {
  k_groth16_agg_final3<32><<<1, 32, trio::trio_smem_bytes(32), st>>>((const AggScratch*)a.scratch, product, a.verdict);
  return 1;
}

}  // namespace launch
}  // namespace bn254
