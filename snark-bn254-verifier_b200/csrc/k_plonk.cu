// PlonK batch kernels, staged (plonk.cuh): A (per proof) -> terms 0 (per proof x term) -> C (per survivor) -> terms 1
// -> E (per survivor: sums, 2-pair Miller loop, final exponentiation).  sm_100a only.
// Replaces, per proof, load_plonk_proof_from_bytes + verify_plonk (reference verifier/src/plonk/converter.rs:121-178,
// verifier/src/plonk/verify.rs:46-317, verifier/src/plonk/kzg.rs:46-190).
// `list` holds the indices of the proofs that are still alive after stage A (early rejects cost nothing further);
// a slot is set to -(i + 1) when stage C ends the proof.
#include <stdlib.h>

#include "kernels.h"
#include "trio.cuh"

namespace bn254 {
namespace {

struct PlonkDbgPtrs {
  uint8_t *g1, *fr, *m, *gt;
};
__device__ __forceinline__ PlonkDebug plonk_dbg(const PlonkDbgPtrs& d, size_t i) {
  return PlonkDebug{d.g1 ? d.g1 + 256 * i : nullptr, d.fr ? d.fr + 256 * i : nullptr, d.m ? d.m + 384 * i : nullptr,
                    d.gt ? d.gt + 384 * i : nullptr};
}

__global__ void k_plonk_vk_prepare(PlonkVkDev* vk) {
  if (blockIdx.x == 0 && threadIdx.x == 0) plonk_vk_prepare(*vk);
}
__global__ void k_plonk_fixed_tables(const G1Aff* bases, G1Aff* table) {
  int b = blockIdx.x, w = threadIdx.x;
  if (w >= BN_IC_WINDOWS) return;
  groth16_ic_table_slice(table + ((size_t)b * BN_IC_WINDOWS + w) * BN_IC_ENTRIES, bases[b], w);
}

__global__ void __launch_bounds__(64)
    k_plonk_stage_a(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                    const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs, size_t n,
                    uint8_t* __restrict__ status, PlonkWork* work, int* list, int* count, PlonkDbgPtrs dp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;  // (no barriers in stages A, C and the term kernels)
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  int st = plonk_stage_a(work[i], *vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i, n_inputs,
                         plonk_dbg(dp, i));
  if (st == BN254V_OK_TRUE) {
    list[atomicAdd(count, 1)] = (int)i;
    status[i] = BN254V_STATUS_UNSET;
  } else {
    status[i] = (uint8_t)st;
  }
}

// MINB blocks of 64 threads per SM: 8 = at most 128 registers, four warps per sub-partition (the per-term form and the
// second MSM round); 1 = no register cap (255), which the joint form of the first round prefers (2^16 proofs: 5.93
// against 6.28 ms; the second round 8.32 against 7.88 ms the other way round).
template <int MINB>
__global__ void __launch_bounds__(64, MINB)
    k_plonk_terms(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride, PlonkWork* work,
                  const int* __restrict__ list, const int* __restrict__ count, int stage, int joint) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  int i = list[slot];
  if (i < 0) return;
  // Blocks are scheduled in blockIdx order: the long terms (variable-base, ~2 200 multiplications) go first and the short
  // ones (fixed-base tables, ~350) last, so that the short ones fill the tail instead of leaving the long ones alone on
  // the machine at the end of the launch.
  if (joint) {
    // joint form (plonk.cuh plonk_item_joint): items of stage 1 in the order 0 2 3 4 1 5 (long ones first)
    const int item = stage == 0 ? (int)blockIdx.y : ((0x514320 >> (4 * blockIdx.y)) & 15);
    plonk_item_joint(work[i], *vk, proofs + stride * (size_t)i, stage, item);
  } else {
    plonk_term(work[i], *vk, proofs + stride * (size_t)i, stage, plonk_term_order(vk->n_qcp, stage, blockIdx.y));
  }
}

__global__ void __launch_bounds__(64)
    k_plonk_stage_c(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                    const uint8_t* __restrict__ rnd, uint8_t* __restrict__ status, PlonkWork* work, int* list,
                    const int* __restrict__ count, PlonkDbgPtrs dp, int joint) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  int i = list[slot];
  int st = plonk_stage_c(work[i], *vk, proofs + stride * (size_t)i, rnd + (size_t)32 * i, plonk_dbg(dp, i), joint != 0);
  if (st != BN254V_OK_TRUE) {
    status[i] = (uint8_t)st;
    list[slot] = -(i + 1);  // dead: the term kernel skips it, stage E walks the pairing on substitute points
  }
}

// Stage E contains the pairing's block-wide phase barriers: every thread of every launched block runs it to the end
// (a block with no survivor at all leaves as a whole -- a block-uniform exit).
template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_plonk_stage_e(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                    uint8_t* __restrict__ status, PlonkWork* work, const int* __restrict__ list,
                    const int* __restrict__ count, PlonkDbgPtrs dp, int joint) {
  const int cnt = *count;
  if ((int)(blockIdx.x * blockDim.x) >= cnt) return;  // uniform over the block
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = slot < cnt;
  int i = list[live ? slot : cnt - 1];
  if (i < 0) live = false, i = -(i + 1);
  PlonkDebug dbg = live ? plonk_dbg(dp, i) : PlonkDebug{nullptr, nullptr, nullptr, nullptr};
  int st = plonk_stage_e(work[i], *vk, proofs + stride * (size_t)i, dbg, live, joint != 0);
  if (live) status[i] = (uint8_t)st;
}

// ---- stage E as two kernels for batches that cannot fill the GPU with one proof per thread:
// D (one thread per survivor): G1 sums, identity checks, affine conversion -> the two pairing inputs;
// E3 (three lanes per survivor, trio.cuh): Miller loop over the pair table, final exponentiation, verdict.
__global__ void __launch_bounds__(64)
    k_plonk_stage_d(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                    uint8_t* __restrict__ status, PlonkWork* work, int* list, const int* __restrict__ count,
                    PlonkDbgPtrs dp, int joint) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  int i = list[slot];
  if (i < 0) return;
  G1Aff pf[2];
  int st = plonk_stage_d(pf, work[i], *vk, proofs + stride * (size_t)i, plonk_dbg(dp, i), joint != 0);
  if (st != BN254V_OK_TRUE) {
    status[i] = (uint8_t)st;
    list[slot] = -(i + 1);
  } else {
    work[i].pair[0] = pf[0], work[i].pair[1] = pf[1];
  }
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_plonk_stage_e3(const PlonkVkDev* __restrict__ vk, uint8_t* __restrict__ status, const PlonkWork* work,
                     const int* __restrict__ list, const int* __restrict__ count, PlonkDbgPtrs dp) {
  const int cnt = *count;
  const int per_block = (TPB / 32) * BN_TRIOS_PER_WARP;
  if ((int)blockIdx.x * per_block >= cnt) return;  // uniform over the block
  const size_t slot = trio::trio_slot();
  bool live = trio::trio_lane_valid() && slot < (size_t)cnt;
  int i = list[live ? slot : 0];
  if (i < 0) live = false, i = -(i + 1);
  G1Aff pf[2] = {vk->g1, vk->g1};  // substitute inputs of idle trios: the pairing contains block-wide barriers
  if (live) pf[0] = work[i].pair[0], pf[1] = work[i].pair[1];
  trio::S12 f;
  trio::miller_loop_pairtab0_s(f, pf, vk->g2_pairs);
  if (live && dp.m) trio::fp12s_to_bytes(dp.m + 384 * (size_t)i, f);
  trio::fp12s_final_exponentiation(f, f);
  if (live && dp.gt) trio::fp12s_to_bytes(dp.gt + 384 * (size_t)i, f);
  const bool one = trio::fp12s_eq(f, trio::fp12s_one());
  if (live && trio::lane_j() == 0) status[i] = one ? BN254V_OK_TRUE : BN254V_ERR_PAIRING_CHECK_FAILED;
}

}  // namespace

namespace launch {

int plonk_vk_prepare(cudaStream_t st, PlonkVkDev* dv, int n_fixed, G1Aff* bases_then_tables) {
  k_plonk_vk_prepare<<<1, 32, 0, st>>>(dv);
  k_plonk_fixed_tables<<<n_fixed, BN_IC_WINDOWS, 0, st>>>(bases_then_tables, bases_then_tables + n_fixed);
  return 2;
}

// Joint evaluation of the MSM terms (shared doublings) once a chunk has enough proofs to fill the GPU with a third of the
// threads; below that the per-term form wins on latency (BN254V_PLONK_JOINT_MIN overrides the switch-over point).
static size_t plonk_joint_min() {
  static long forced = -2;
  if (forced == -2) {
    const char* e = getenv("BN254V_PLONK_JOINT_MIN");
    forced = e ? atol(e) : -1;
  }
  return forced >= 0 ? (size_t)forced : (size_t)1 << 15;
}

int plonk_verify(cudaStream_t st, const PlonkArgs& a, int sm_count) {
  const size_t cm = a.m;
  PlonkDbgPtrs dp{a.dbg_g1, a.dbg_fr, a.dbg_m, a.dbg_gt};
  const unsigned g64 = (unsigned)((cm + 63) / 64);
  const int joint = cm >= plonk_joint_min() ? 1 : 0;
  const int n0 = plonk_n_items(a.n_qcp, 0, joint != 0), n1 = plonk_n_items(a.n_qcp, 1, joint != 0);
  cudaMemsetAsync(a.count, 0, sizeof(int), st);
  k_plonk_stage_a<<<g64, 64, 0, st>>>(a.vk, a.proofs, a.stride, a.lens, a.inputs, a.n_inputs, cm, a.status, a.work,
                                      a.list, a.count, dp);
  if (a.stage_ev) cudaEventRecord(a.stage_ev[0], st);
  if (joint) k_plonk_terms<1><<<dim3(g64, n0), 64, 0, st>>>(a.vk, a.proofs, a.stride, a.work, a.list, a.count, 0, joint);
  else k_plonk_terms<8><<<dim3(g64, n0), 64, 0, st>>>(a.vk, a.proofs, a.stride, a.work, a.list, a.count, 0, joint);
  if (a.stage_ev) cudaEventRecord(a.stage_ev[1], st);
  k_plonk_stage_c<<<g64, 64, 0, st>>>(a.vk, a.proofs, a.stride, a.rnd, a.status, a.work, a.list, a.count, dp, joint);
  if (a.stage_ev) cudaEventRecord(a.stage_ev[2], st);
  k_plonk_terms<8><<<dim3(g64, n1), 64, 0, st>>>(a.vk, a.proofs, a.stride, a.work, a.list, a.count, 1, joint);
  if (a.stage_ev) cudaEventRecord(a.stage_ev[3], st);
  // Three lanes per proof while one proof per thread would leave the SM sub-partitions short of warps (the number
  // of survivors is only known on the device: the choice goes by the chunk size).
  if (cm <= (size_t)trio_max_items(sm_count)) {
    k_plonk_stage_d<<<g64, 64, 0, st>>>(a.vk, a.proofs, a.stride, a.status, a.work, a.list, a.count, dp, joint);
    const unsigned per_block = (128 / 32) * BN_TRIOS_PER_WARP;
    k_plonk_stage_e3<128, 3><<<(unsigned)((cm + per_block - 1) / per_block), 128, trio::trio_smem_bytes(128), st>>>(
        a.vk, a.status, a.work, a.list, a.count, dp);
    return 6;
  }
  if (pick_shape(cm, sm_count) == SHAPE_32)
    k_plonk_stage_e<32><<<(unsigned)((cm + 31) / 32), 32, 0, st>>>(a.vk, a.proofs, a.stride, a.status, a.work, a.list,
                                                                   a.count, dp, joint);
  else
    k_plonk_stage_e<128><<<(unsigned)((cm + 127) / 128), 128, 0, st>>>(a.vk, a.proofs, a.stride, a.status, a.work,
                                                                       a.list, a.count, dp, joint);
  return 5;
}

}  // namespace launch
}  // namespace bn254
