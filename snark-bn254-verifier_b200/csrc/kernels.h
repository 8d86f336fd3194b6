// Launch interface between the host side of the library (bn254v.cu: C ABI, sharding, buffers) and the kernel
// translation units (k_groth16.cu, k_groth16_agg.cu, k_plonk.cu, k_pairing.cu, k_aux.cu).  Each kernel TU is compiled on its own (the
// pairing code is large; four nvcc processes in parallel build the library in a third of the time) and exposes plain
// host functions that enqueue its kernels on a stream and return how many launches they issued.
#pragma once
#include <cuda_runtime.h>

#include "plonk.cuh"
#include "synth.cuh"

namespace bn254 {
namespace launch {

// Launch shape of the per-proof pairing kernels for a batch of m items (see k_groth16.cu).
enum Shape { SHAPE_32 = 6, SHAPE_128 = 1, SHAPE_448 = 3, SHAPE_384 = 10 };
int pick_shape(size_t m, int sm_count);
// Largest batch that runs the pairing with three lanes per item (trio.cuh) instead of one item per thread
// (BN254V_TRIO_MAX overrides; 0 disables the three-lane kernels).
size_t trio_max_items(int sm_count);

// ---- VK-constant precomputation (once per VK and device)
int groth16_vk_prepare(cudaStream_t st, Groth16VkDev* dv, int n_bases, G1Aff* table);
int plonk_vk_prepare(cudaStream_t st, PlonkVkDev* dv, int n_fixed, G1Aff* bases_then_tables);

// ---- Groth16 batch on device-resident buffers
struct Groth16Args {
  const Groth16VkDev* vk;
  const uint8_t* proofs;
  size_t stride;
  const uint32_t* lens;  // or null
  const uint8_t* inputs;
  int n_inputs;
  size_t m;
  uint8_t* status;
  uint8_t *dbg_l, *dbg_m, *dbg_gt;  // or null
  Fp12* fbuf;                       // m Miller values (multi-launch shapes), or null
  cudaEvent_t mid;                  // recorded between the Miller and the final-exponentiation launch when non-null
  cudaEvent_t pre = nullptr;        // recorded after the prepare launch when non-null
};
int groth16_verify(cudaStream_t st, const Groth16Args& a, int sm_count, bool* two_launch);

// ---- opt-in aggregate Groth16 check (groth16_agg.cuh, k_groth16_agg.cu): "is every proof valid?"
struct Groth16AggArgs {
  const Groth16VkDev* vk;
  const uint8_t* proofs;
  size_t stride;
  const uint32_t* lens;  // or null
  const uint8_t* inputs;
  int n_inputs;
  const uint8_t* rnd16;  // 16 scalar bytes per proof
  size_t m;
  uint8_t* status;         // m per-proof validation statuses
  Fp12* fbuf;              // groth16_agg_slots(m) entries
  G1Jac* gbuf;             // groth16_agg_slots(m) entries
  const uint8_t* scal_be;  // (1 + n_inputs) x 32 bytes: the host-computed scalar sums (read by the side stream only)
  void* scratch;           // groth16_agg_scratch_bytes()
  uint8_t* verdict;        // 1 byte: 1 = the aggregate equation holds
};
size_t groth16_agg_scratch_bytes();
size_t groth16_agg_slots(size_t m);
// main stream: _prepare, then _miller (-> where the product of the Miller values lands), then -- after the side stream --
// _final; side stream, after _prepare: _side (sum tree, the batch's points and its own Miller value)
int groth16_agg_prepare(cudaStream_t st, const Groth16AggArgs& a);
int groth16_agg_side(cudaStream_t st, const Groth16AggArgs& a);
int groth16_agg_miller(cudaStream_t st, const Groth16AggArgs& a, int sm_count, Fp12** product);
int groth16_agg_final(cudaStream_t st, const Groth16AggArgs& a, const Fp12* product);

// ---- PlonK chunk (<= 2^16 proofs) on device-resident buffers
struct PlonkArgs {
  const PlonkVkDev* vk;
  int n_qcp;
  const uint8_t* proofs;
  size_t stride;
  const uint32_t* lens;  // or null
  const uint8_t* inputs;
  int n_inputs;
  const uint8_t* rnd;
  size_t m;
  uint8_t* status;
  uint8_t *dbg_g1, *dbg_fr, *dbg_m, *dbg_gt;  // or null
  PlonkWork* work;                            // m records
  int* list;                                  // m slots
  int* count;                                 // 1 counter
  cudaEvent_t* stage_ev;                      // 4 events recorded after stages A, terms 0, C, terms 1, or null
};
int plonk_verify(cudaStream_t st, const PlonkArgs& a, int sm_count);

// ---- raw pairing products
// fbuf: m Fp12 of scratch (big batches run as two launches, Miller loops | final exponentiations) or null (one fused
// launch); mid: an event recorded between the two launches, or null.  Returns the number of launches.
int pairing_product(cudaStream_t st, int k, const uint8_t* g1, const uint8_t* g2, size_t m, uint8_t* is_one,
                    uint8_t* miller_out, uint8_t* gt_out, int sm_count, Fp12* fbuf, cudaEvent_t mid);

// ---- synthetic workloads and the roofline probe (k_aux.cu)
int groth16_synth(cudaStream_t st, const Groth16Trapdoor& td, uint64_t seed, size_t first, size_t n, int n_public,
                  int sign_mode, uint8_t* proofs, uint8_t* inputs, uint8_t* expected);
int pairing_synth(cudaStream_t st, uint64_t seed, size_t first, size_t n, int k, uint8_t* g1, uint8_t* g2,
                  uint8_t* expected);
int imad_peak(cudaStream_t st, bool wide, int blocks, int threads, int iters, uint64_t* sink);

}  // namespace launch
}  // namespace bn254
