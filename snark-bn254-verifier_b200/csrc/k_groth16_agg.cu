// Kernels of the opt-in aggregate Groth16 check (groth16_agg.cuh).  sm_100a only.
//   k_groth16_agg_miller  one proof per thread: validate, [r_i] A_i, [r_i] C_i, the single-pair Miller loop of (r_i A_i, B_i)
//   k_groth16_agg_fold    product / sum trees over the Miller values and the [r_i] C_i, eight to one per pass
//   k_groth16_agg_points  the batch's own three G1 points from the host-computed scalar sums (one thread)
//   k_groth16_agg_final3  three lanes (trio.cuh): the batch's three-pair Miller loop, times the folded product, final
//                         exponentiation, == 1
// Same barrier discipline as k_groth16.cu: nothing returns before the last block-wide barrier.
#include "kernels.h"
#include "groth16_agg.cuh"
#include "trio.cuh"

namespace bn254 {
namespace {

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_groth16_agg_miller(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                         const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs,
                         const uint8_t* __restrict__ rnd16, size_t n, uint8_t* __restrict__ status,
                         Fp12* __restrict__ fbuf, G1Jac* __restrict__ gbuf) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;  // spare threads of the last block walk the last proof and write nothing
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  Fp12 f;
  G1Jac rc;
  const int st = groth16_agg_one(f, rc, *vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i, n_inputs,
                                 rnd16 + 16 * i, live);
  if (!live) return;
  status[i] = (uint8_t)st;
  fbuf[i] = f;
  gbuf[i] = rc;
}

__global__ void __launch_bounds__(64)
    k_groth16_agg_fold(const Fp12* __restrict__ fin, const G1Jac* __restrict__ gin, size_t n, Fp12* __restrict__ fout,
                       G1Jac* __restrict__ gout, int per) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t lo = t * per;
  if (lo >= n) return;
  Fp12 f = fin[lo];
  G1Jac g = gin[lo];
  for (size_t k = lo + 1; k < n && k < lo + per; k++) groth16_agg_fold(f, g, fin[k], gin[k]);
  fout[t] = f;
  gout[t] = g;
}

struct AggPoints {
  G1Aff nsa, sl, sc;
  int ok;
};

__global__ void k_groth16_agg_points(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ scal_be,
                                     const G1Jac* __restrict__ sum_rc, AggPoints* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  out->ok = groth16_agg_points(out->nsa, out->sl, out->sc, *vk, scal_be, *sum_rc) ? 1 : 0;
}

template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_groth16_agg_final3(const Groth16VkDev* __restrict__ vk, const AggPoints* __restrict__ pts,
                         const Fp12* __restrict__ prod, uint8_t* __restrict__ verdict) {
  const bool live = trio::trio_lane_valid() && trio::trio_slot() == 0;
  const bool ok = pts->ok != 0;
  G1Aff A = vk->alpha, pf[2] = {vk->ic[0], vk->ic[0]};  // idle trios (and a degenerate batch) walk substitutes
  const G2Aff B = vk->beta;
  if (ok) A = pts->nsa, pf[0] = pts->sl, pf[1] = pts->sc;
  trio::S12 f;
  bool in_g2;
  trio::miller_loop_pairtab1_s(f, A, B, pf, vk->gd_pairs, &in_g2);
  const trio::S12 p = trio::fp12s_load(*prod);
  trio::fp12s_mul(f, f, p);
  trio::fp12s_final_exponentiation(f, f);
  const bool one = trio::fp12s_eq(f, trio::fp12s_one());
  if (live && trio::lane_j() == 0) *verdict = (ok && one) ? 1 : 0;
}

}  // namespace

namespace launch {

size_t groth16_agg_scratch_bytes() { return sizeof(AggPoints); }

int groth16_agg_miller(cudaStream_t st, const Groth16AggArgs& a, int sm_count) {
  const size_t m = a.m;
  const int shape = pick_shape(m, sm_count);
#define LA(TPB, MINB)                                                                                                     \
  k_groth16_agg_miller<TPB, MINB><<<(unsigned)((m + TPB - 1) / TPB), TPB, 0, st>>>(a.vk, a.proofs, a.stride, a.lens, a.inputs, \
                                                                                   a.n_inputs, a.rnd16, m, a.status, a.fbuf, a.gbuf)
  switch (shape) {
    case SHAPE_448: LA(448, 1); break;
    case SHAPE_384: LA(384, 1); break;
    default: LA(128, 2); break;
  }
#undef LA
  return 1;
}

// a.fbuf / a.gbuf hold m + (m + 7) / 8 entries: the passes ping-pong between [0, m) and [m, ..)
int groth16_agg_finish(cudaStream_t st, const Groth16AggArgs& a) {
  constexpr int PER = 8;
  int launches = 0;
  size_t cur = a.m;
  Fp12 *fa = a.fbuf, *fb = a.fbuf + a.m;
  G1Jac *ga = a.gbuf, *gb = a.gbuf + a.m;
  while (cur > 1) {
    const size_t nxt = (cur + PER - 1) / PER;
    k_groth16_agg_fold<<<(unsigned)((nxt + 63) / 64), 64, 0, st>>>(fa, ga, cur, fb, gb, PER);
    Fp12* tf = fa; fa = fb; fb = tf;
    G1Jac* tg = ga; ga = gb; gb = tg;
    cur = nxt;
    launches++;
  }
  AggPoints* pts = (AggPoints*)a.scratch;
  k_groth16_agg_points<<<1, 1, 0, st>>>(a.vk, a.scal_be, ga, pts);
  k_groth16_agg_final3<32><<<1, 32, trio::trio_smem_bytes(32), st>>>(a.vk, pts, fa, a.verdict);
  return launches + 2;
}

}  // namespace launch
}  // namespace bn254
