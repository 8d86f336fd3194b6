// Groth16 batch kernels (one proof per thread) and their launcher.  sm_100a only.
// Replaces, per proof, load_groth16_proof_from_bytes + verify_groth16 (reference verifier/src/groth16/converter.rs:14-26,
// verifier/src/groth16/verify.rs:53-78); the arithmetic is in groth16.cuh / pairing_body.inc.
//
// Every thread of a block -- spare threads of the last block and proofs that failed validation included -- runs the
// same sequence of Miller-loop / exponentiation iterations, because those contain block-wide phase barriers
// (field.cuh, BN_PHASE_SYNC): nothing returns before the last barrier.
#include <stdlib.h>

#include "kernels.h"
#include "trio.cuh"

namespace bn254 {
namespace {

__global__ void k_groth16_vk_prepare(Groth16VkDev* vk) {
  if (blockIdx.x == 0 && threadIdx.x == 0) groth16_vk_prepare(*vk);
}

// one thread per (base, window): builds the fixed-base window tables of VK-constant G1 bases (once per VK)
__global__ void k_g1_fixed_tables(const G1Aff* bases, G1Aff* table) {
  int b = blockIdx.x, w = threadIdx.x;
  if (w >= BN_IC_WINDOWS) return;
  groth16_ic_table_slice(table + ((size_t)b * BN_IC_WINDOWS + w) * BN_IC_ENTRIES, bases[b], w);
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_groth16_verify(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                     const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs,
                     size_t n, uint8_t* __restrict__ status, uint8_t* dbg_l, uint8_t* dbg_m, uint8_t* dbg_gt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;  // spare threads of the last block walk the last proof and write nothing
  Groth16Debug dbg{live && dbg_l ? dbg_l + 64 * i : nullptr, live && dbg_m ? dbg_m + 384 * i : nullptr,
                   live && dbg_gt ? dbg_gt + 384 * i : nullptr};
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  int st = groth16_verify_one(*vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i, n_inputs, dbg, live);
  if (live) status[i] = (uint8_t)st;
}

// Groth16 as three launches for big batches: k_groth16_prepare (further down; one thread per proof, small blocks, no
// barriers: decode, validate, prepare_inputs -> L parked in fbuf[i], status UNSET or the failure), k_groth16_miller
// (Miller loop -> fbuf[i]), k_groth16_finish (final exponentiation, verdict).  Each has a fraction of the code and stack
// of the fused kernel; measured against one fused launch: 2 % faster with the Miller loop and the exponentiation apart,
// 0.7 % (2^16 proofs) to 2 % (2^18) more with the prepare step in its own high-occupancy kernel.
template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_groth16_miller(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride, size_t n,
                         uint8_t* __restrict__ status, Fp12* __restrict__ fbuf, uint8_t* dbg_m) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;
  const bool ok = live && status[i] == BN254V_STATUS_UNSET;
  G1Aff A = vk->alpha, pf[2] = {vk->ic[0], vk->ic[0]};  // substitutes: block-wide barriers inside the loop
  G2Aff B = vk->beta;
  if (ok) {
    const uint8_t* pr = proofs + stride * i;
    load_g1_unchecked(A, pr);
    load_g2_unchecked(B, pr + 64);
    load_g1_unchecked(pf[1], pr + 192);
    pf[0] = *(const G1Aff*)&fbuf[i];
  }
  Fp12 f;
  bool in_g2;
  miller_loop_pairtab<1>(f, &A, &B, pf, vk->gd_pairs, &in_g2);
  if (!ok) return;
  if (!in_g2) {
    status[i] = BN254V_PANIC_NOT_IN_SUBGROUP;
    return;
  }
  if (dbg_m) fp12_to_bytes(dbg_m + 384 * i, f);
  fbuf[i] = f;
}
template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_groth16_finish(const Groth16VkDev* __restrict__ vk, size_t n, uint8_t* __restrict__ status,
                     const Fp12* __restrict__ fbuf, uint8_t* dbg_gt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool decide = i < n && status[i] == BN254V_STATUS_UNSET;
  Groth16Debug dbg{nullptr, nullptr, decide && dbg_gt ? dbg_gt + 384 * i : nullptr};
  Fp12 f = decide ? fbuf[i] : vk->target;  // decided / spare threads run the exponentiation for its barriers only
  int st = groth16_finish_one(f, *vk, dbg, decide);
  if (decide) status[i] = (uint8_t)st;
}

// ---- three lanes per proof (trio.cuh) for batches that cannot fill the GPU with one proof per thread ------------
// prepare (one thread per proof): decode, validate, prepare_inputs -> L, parked in the first 64 bytes of fbuf[i];
// miller3 / finish3 (three lanes per proof): the pairing.  Same statuses, same canonical values.
// (eight blocks per SM = 128 registers: the window-table loads want warps more than registers; 1.22 -> 1.13 ms at 2^16)
__global__ void __launch_bounds__(64, 8)
    k_groth16_prepare(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                      const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs, size_t n,
                      uint8_t* __restrict__ status, Fp12* __restrict__ fbuf, uint8_t* dbg_l) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;  // (no barriers in this kernel)
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  G1Aff A, C, L;
  G2Aff B;
  int st = groth16_parse_one(A, B, C, L, *vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i, n_inputs);
  if (st == BN254V_OK_TRUE) {
    *(G1Aff*)&fbuf[i] = L;
    status[i] = BN254V_STATUS_UNSET;
  } else {
    status[i] = (uint8_t)st;
  }
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_groth16_miller3(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride, size_t n,
                      uint8_t* __restrict__ status, Fp12* __restrict__ fbuf, uint8_t* dbg_l, uint8_t* dbg_m) {
  size_t i = trio::trio_slot();
  bool live = trio::trio_lane_valid() && i < n;
  if (!live) i = 0;
  if (status[i] != BN254V_STATUS_UNSET) live = false;  // failed in k_groth16_prepare: keeps its status
  G1Aff A = vk->alpha, pf[2] = {vk->ic[0], vk->ic[0]};  // substitutes of idle trios (block-wide barriers inside)
  G2Aff B = vk->beta;
  if (live) {
    const uint8_t* pr = proofs + stride * i;
    load_g1_unchecked(A, pr);  // validated by k_groth16_prepare
    load_g2_unchecked(B, pr + 64);
    load_g1_unchecked(pf[1], pr + 192);
    pf[0] = *(const G1Aff*)&fbuf[i];
  }
  trio::S12 f;
  bool in_g2;
  trio::miller_loop_pairtab1_s(f, A, B, pf, vk->gd_pairs, &in_g2);
  if (!live) return;
  if (!in_g2) {
    if (trio::lane_j() == 0) status[i] = BN254V_PANIC_NOT_IN_SUBGROUP;
    return;
  }
  if (dbg_l && trio::lane_j() == 0) store_g1(dbg_l + 64 * i, pf[0]);
  if (dbg_m) trio::fp12s_to_bytes(dbg_m + 384 * i, f);
  trio::fp12s_store(fbuf[i], f);
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_groth16_finish3(const Groth16VkDev* __restrict__ vk, size_t n, uint8_t* __restrict__ status,
                      const Fp12* __restrict__ fbuf, uint8_t* dbg_gt) {
  size_t i = trio::trio_slot();
  bool live = trio::trio_lane_valid() && i < n;
  if (!live) i = 0;
  if (status[i] != BN254V_STATUS_UNSET) live = false;
  const trio::S12 target = trio::fp12s_load(vk->target);
  trio::S12 f = live ? trio::fp12s_load(fbuf[i]) : target;
  trio::fp12s_final_exponentiation(f, f);
  if (!live) return;
  if (dbg_gt) trio::fp12s_to_bytes(dbg_gt + 384 * i, f);
  const bool ok = trio::fp12s_eq(f, target);
  if (trio::lane_j() == 0) status[i] = ok ? BN254V_OK_TRUE : BN254V_OK_FALSE;
}

}  // namespace

namespace launch {

// Launch shapes.  One proof per thread; the block is the unit that the phase barriers keep in step, and the grid should
// cover the SMs evenly: big batches use 448-thread blocks, one per SM (14 warps, 128 registers per thread; 2^16 proofs =
// 147 blocks on 148 SMs); 384 threads x 168 registers is 9 % faster per proof (a 16 K-register SMSP holds 3 warps at 168
// or 4 at 128) but 2^16 proofs do not fit one wave of it, so it is used once there are several waves; small batches use
// smaller blocks so that every SM gets work.
size_t trio_max_items(int sm_count) {
  static long forced = -2;
  if (forced == -2) {
    const char* e = getenv("BN254V_TRIO_MAX");
    forced = e ? atol(e) : -1;
  }
  if (forced >= 0) return (size_t)forced;
  return (size_t)sm_count * 128;  // below ~one warp per SM sub-partition with one item per thread
}

int pick_shape(size_t m, int sm_count) {
  if (m >= (size_t)sm_count * 384 * 4) return SHAPE_384;
  if (m >= (size_t)sm_count * 448 * 3 / 4) return SHAPE_448;
  if (m >= (size_t)sm_count * 128) return SHAPE_128;
  return SHAPE_32;
}

int groth16_vk_prepare(cudaStream_t st, Groth16VkDev* dv, int n_bases, G1Aff* table) {
  k_groth16_vk_prepare<<<1, 32, 0, st>>>(dv);
  if (n_bases > 0) k_g1_fixed_tables<<<n_bases, BN_IC_WINDOWS, 0, st>>>(&dv->ic[1], table);
  // IC_0 and alpha: the batch-wide points of the aggregate check (groth16_agg.cuh)
  G1Aff* agg = table + (size_t)n_bases * BN_IC_WINDOWS * BN_IC_ENTRIES;
  k_g1_fixed_tables<<<1, BN_IC_WINDOWS, 0, st>>>(&dv->ic[0], agg);
  k_g1_fixed_tables<<<1, BN_IC_WINDOWS, 0, st>>>(&dv->alpha, agg + (size_t)BN_IC_WINDOWS * BN_IC_ENTRIES);
  return n_bases > 0 ? 4 : 3;
}

// shared with the PlonK VK preparation (k_plonk.cu has its own copy of the kernel)
int groth16_verify(cudaStream_t st, const Groth16Args& a, int sm_count, bool* two_launch) {
  const int shape = pick_shape(a.m, sm_count);
  const size_t m = a.m;
  if (a.fbuf && m <= trio_max_items(sm_count)) {  // small batch: three lanes per proof
    constexpr int TPB = 128;
    const unsigned per_block = (TPB / 32) * BN_TRIOS_PER_WARP;
    const unsigned grid = (unsigned)((m + per_block - 1) / per_block);
    if (two_launch) *two_launch = true;
    k_groth16_prepare<<<(unsigned)((m + 63) / 64), 64, 0, st>>>(a.vk, a.proofs, a.stride, a.lens, a.inputs, a.n_inputs, m,
                                                               a.status, a.fbuf, a.dbg_l);
    if (a.pre) cudaEventRecord(a.pre, st);
    k_groth16_miller3<TPB, 3><<<grid, TPB, trio::trio_smem_bytes(TPB), st>>>(a.vk, a.proofs, a.stride, m, a.status, a.fbuf,
                                                                            a.dbg_l, a.dbg_m);
    if (a.mid) cudaEventRecord(a.mid, st);
    k_groth16_finish3<TPB, 3><<<grid, TPB, trio::trio_smem_bytes(TPB), st>>>(a.vk, m, a.status, a.fbuf, a.dbg_gt);
    return 3;
  }
  // Big batches: three launches (prepare | Miller loop | final exponentiation), see k_groth16_miller.
  if (a.fbuf && (shape == SHAPE_448 || shape == SHAPE_384)) {
    if (two_launch) *two_launch = true;
    k_groth16_prepare<<<(unsigned)((m + 63) / 64), 64, 0, st>>>(a.vk, a.proofs, a.stride, a.lens, a.inputs, a.n_inputs, m,
                                                               a.status, a.fbuf, a.dbg_l);
    if (a.pre) cudaEventRecord(a.pre, st);
    if (shape == SHAPE_448) {
      k_groth16_miller<448><<<(unsigned)((m + 447) / 448), 448, 0, st>>>(a.vk, a.proofs, a.stride, m, a.status, a.fbuf, a.dbg_m);
      if (a.mid) cudaEventRecord(a.mid, st);
      k_groth16_finish<448><<<(unsigned)((m + 447) / 448), 448, 0, st>>>(a.vk, m, a.status, a.fbuf, a.dbg_gt);
    } else {
      k_groth16_miller<384><<<(unsigned)((m + 383) / 384), 384, 0, st>>>(a.vk, a.proofs, a.stride, m, a.status, a.fbuf, a.dbg_m);
      if (a.mid) cudaEventRecord(a.mid, st);
      k_groth16_finish<384><<<(unsigned)((m + 383) / 384), 384, 0, st>>>(a.vk, m, a.status, a.fbuf, a.dbg_gt);
    }
    return 3;
  }
  if (two_launch) *two_launch = false;
#define LV(TPB, MINB)                                                                                                \
  k_groth16_verify<TPB, MINB><<<(unsigned)((m + TPB - 1) / TPB), TPB, 0, st>>>(a.vk, a.proofs, a.stride, a.lens,    \
                                                                               a.inputs, a.n_inputs, m, a.status, \
                                                                               a.dbg_l, a.dbg_m, a.dbg_gt)
  switch (shape) {
    case SHAPE_448: LV(448, 1); break;
    case SHAPE_384: LV(384, 1); break;
    case SHAPE_32: LV(32, 1); break;
    default: LV(128, 2); break;
  }
#undef LV
  return 1;
}

}  // namespace launch
}  // namespace bn254
