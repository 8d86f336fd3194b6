// bn254v: kernels + C ABI (include/bn254v.h).  sm_100a only; there is no CPU fallback -- every
// compute entry point fails with BN254V_E_NO_DEVICE when no CUDA device is usable.
//
// Kernel shape (v1): one proof per thread.  Per-proof data is ~0.3 KB in / 1 B out, so HBM traffic is
// negligible; the bound is the SM's 32-bit integer multiply-add pipe.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "gnark_host.h"
#include "groth16.cuh"
#include "plonk.cuh"
#include "lanepair.cuh"
#include "synth.cuh"

using namespace bn254;

// ------------------------------------------------------------------------------------------------
// library state
// ------------------------------------------------------------------------------------------------
namespace {

struct Dev {
  int id;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1, evm;  // evm: between the two Groth16 launches
};

std::mutex g_mu;
std::vector<Dev> g_devs;
bool g_inited = false;
std::atomic<uint64_t> g_launches{0};
thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) return fail(BN254V_E_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

int ensure_init() {
  if (g_inited) return 0;
  return bn254v_init(nullptr, 0);
}

// Scratch device memory is kept in a small per-device free list between calls: cudaMalloc / cudaFree cost
// milliseconds (and cudaFree synchronises the device), which is visible next to a 50 ms batch.
struct Block {
  void* p;
  size_t cap;
  int dev;
};
std::vector<Block> g_pool;  // guarded by g_pool_mu
std::mutex g_pool_mu;

struct DevBuf {  // RAII scratch allocation on the current device (returned to the pool, not freed)
  void* p = nullptr;
  size_t cap = 0;
  int dev = -1;
  ~DevBuf() {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(Block{p, cap, dev});
  }
  cudaError_t alloc(size_t n) {
    if (!n) n = 1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
      std::lock_guard<std::mutex> lk(g_pool_mu);
      int best = -1;
      for (int i = 0; i < (int)g_pool.size(); i++)
        if (g_pool[i].dev == dev && g_pool[i].cap >= n && g_pool[i].cap <= 2 * n + 4096 &&
            (best < 0 || g_pool[i].cap < g_pool[best].cap))
          best = i;
      if (best >= 0) {
        p = g_pool[best].p;
        cap = g_pool[best].cap;
        g_pool.erase(g_pool.begin() + best);
        return cudaSuccess;
      }
    }
    cap = n;
    e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {  // pool may be holding the memory: release it and retry once
      pool_release(dev);
      e = cudaMalloc(&p, n);
    }
    if (e != cudaSuccess) p = nullptr;
    return e;
  }
  static void pool_release(int device) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool.size();) {
      if (device < 0 || g_pool[i].dev == device) {
        cudaFree(g_pool[i].p);
        g_pool.erase(g_pool.begin() + i);
      } else {
        i++;
      }
    }
  }
  template <class T>
  T* as() { return (T*)p; }
};

// contiguous shard [lo, hi) of n items for device slot d of nd
inline void shard(size_t n, int d, int nd, size_t& lo, size_t& hi) {
  lo = n * (size_t)d / nd;
  hi = n * (size_t)(d + 1) / nd;
}

}  // namespace

struct bn254v_vk {
  int kind;  // 0 groth16, 1 plonk
  int n_public;
  int n_qcp = 0;  // PlonK: number of BSB22 commitments
  int sign_mode;
  std::vector<void*> dev;  // per device slot: Groth16VkDev* / PlonkVkDev*
  std::vector<void*> aux;  // per device slot: fixed-base tables (Groth16) or null
};

struct bn254v_batch {
  size_t n;
  int n_inputs;
  struct Part {
    size_t lo, hi;
    uint8_t *proofs, *inputs, *status;
    Fp12* fbuf;
  };
  std::vector<Part> parts;
};

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
#define BN_TPB 128

__global__ void k_groth16_vk_prepare(Groth16VkDev* vk) {
  if (blockIdx.x == 0 && threadIdx.x == 0) groth16_vk_prepare(*vk);
}

template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_groth16_verify(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                     const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs,
                     size_t n, uint8_t* __restrict__ status, uint8_t* dbg_l, uint8_t* dbg_m, uint8_t* dbg_gt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Groth16Debug dbg{dbg_l ? dbg_l + 64 * i : nullptr, dbg_m ? dbg_m + 384 * i : nullptr,
                   dbg_gt ? dbg_gt + 384 * i : nullptr};
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  status[i] = (uint8_t)groth16_verify_one(*vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i,
                                          n_inputs, dbg);
}

// one thread per (base, window): builds the fixed-base window tables of VK-constant G1 bases (once per VK)
__global__ void k_g1_fixed_tables(const G1Aff* bases, G1Aff* table) {
  int b = blockIdx.x, w = threadIdx.x;
  if (w >= BN_IC_WINDOWS) return;
  groth16_ic_table_slice(table + ((size_t)b * BN_IC_WINDOWS + w) * BN_IC_ENTRIES, bases[b], w);
}

__global__ void k_plonk_vk_prepare(PlonkVkDev* vk) {
  if (blockIdx.x == 0 && threadIdx.x == 0) plonk_vk_prepare(*vk);
}

// ---- lane-pair kernels (lanepair.cuh): two adjacent lanes per proof; no early return (phase barriers inside)
template <int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
    k_groth16_verify_lp(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                        const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs,
                        size_t n, uint8_t* __restrict__ status, uint8_t* dbg_l, uint8_t* dbg_m, uint8_t* dbg_gt) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const bool live = i < n;
  if (!live) i = n - 1;  // spare pairs of the last block redo the last proof and write nothing
  Groth16Debug dbg{live && dbg_l ? dbg_l + 64 * i : nullptr, live && dbg_m ? dbg_m + 384 * i : nullptr,
                   live && dbg_gt ? dbg_gt + 384 * i : nullptr};
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  int st = lp::groth16_verify_pair(*vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i, n_inputs, dbg);
  if (live && !(threadIdx.x & 1)) status[i] = (uint8_t)st;
}

template <int KP, int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_pairing_product_lp(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, size_t n,
                         uint8_t* __restrict__ is_one, uint8_t* miller_out, uint8_t* gt_out) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const bool live = i < n;
  if (!live) i = n - 1;
  bool one = lp::pairing_product_pair<KP>(g1 + (size_t)64 * KP * i, g2 + (size_t)128 * KP * i,
                                          live && miller_out ? miller_out + 384 * i : nullptr,
                                          live && gt_out ? gt_out + 384 * i : nullptr);
  if (live && !(threadIdx.x & 1)) is_one[i] = one ? 1 : 0;
}

// ---- PlonK, staged (plonk.cuh): A (per proof) -> terms 0 (per proof x term) -> C (per survivor) -> terms 1 -> E.
// `list` holds the indices of the proofs that are still alive after stage A (early rejects cost nothing further);
// a slot is set to -1 when a later stage ends the proof.
struct PlonkDbgPtrs {
  uint8_t *g1, *fr, *m, *gt;
};
__device__ __forceinline__ PlonkDebug plonk_dbg(const PlonkDbgPtrs& d, size_t i) {
  return PlonkDebug{d.g1 ? d.g1 + 256 * i : nullptr, d.fr ? d.fr + 256 * i : nullptr, d.m ? d.m + 384 * i : nullptr,
                    d.gt ? d.gt + 384 * i : nullptr};
}

__global__ void __launch_bounds__(64)
    k_plonk_stage_a(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                    const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs, size_t n,
                    uint8_t* __restrict__ status, PlonkWork* work, int* list, int* count, PlonkDbgPtrs dp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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

__global__ void __launch_bounds__(64)
    k_plonk_terms(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride, PlonkWork* work,
                  const int* __restrict__ list, const int* __restrict__ count, int stage) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  int i = list[slot];
  if (i < 0) return;
  plonk_term(work[i], *vk, proofs + stride * (size_t)i, stage, blockIdx.y);
}

__global__ void __launch_bounds__(64)
    k_plonk_stage_c(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                    const uint8_t* __restrict__ rnd, uint8_t* __restrict__ status, PlonkWork* work, int* list,
                    const int* __restrict__ count, PlonkDbgPtrs dp) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  int i = list[slot];
  int st = plonk_stage_c(work[i], *vk, proofs + stride * (size_t)i, rnd + (size_t)32 * i, plonk_dbg(dp, i));
  if (st != BN254V_OK_TRUE) {
    status[i] = (uint8_t)st;
    list[slot] = -(i + 1);  // dead: later stages skip it (the lane-pair stage E recomputes on it and writes nothing)
  }
}

template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_plonk_stage_e(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                    uint8_t* __restrict__ status, PlonkWork* work, const int* __restrict__ list,
                    const int* __restrict__ count, PlonkDbgPtrs dp) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  int i = list[slot];
  if (i < 0) return;
  status[i] = (uint8_t)plonk_stage_e(work[i], *vk, proofs + stride * (size_t)i, plonk_dbg(dp, i));
}

template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_plonk_stage_e_lp(const PlonkVkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                       uint8_t* __restrict__ status, PlonkWork* work, const int* __restrict__ list,
                       const int* __restrict__ count, PlonkDbgPtrs dp) {
  const int cnt = *count;
  if (cnt == 0) return;  // uniform over the grid
  int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  bool live = slot < cnt;
  if (!live) slot = cnt - 1;
  int i = list[slot];
  if (i < 0) {  // ended in stage C: recompute on its (well-formed) data, write nothing
    live = false;
    i = -(i + 1);
  }
  PlonkDebug dbg = live ? plonk_dbg(dp, i) : PlonkDebug{nullptr, nullptr, nullptr, nullptr};
  int st = lp::plonk_stage_e_pair(work[i], *vk, proofs + stride * (size_t)i, dbg);
  if (live && !(threadIdx.x & 1)) status[i] = (uint8_t)st;
}

// Groth16 as two launches (groth16.cuh): Miller values travel through `fbuf` (384 B per proof); a proof that failed
// in the first half keeps its status, the others are marked BN254V_STATUS_UNSET until the second half decides.
template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_groth16_miller(const Groth16VkDev* __restrict__ vk, const uint8_t* __restrict__ proofs, size_t stride,
                     const uint32_t* __restrict__ proof_len, const uint8_t* __restrict__ inputs, int n_inputs, size_t n,
                     uint8_t* __restrict__ status, Fp12* __restrict__ fbuf, uint8_t* dbg_l, uint8_t* dbg_m) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Groth16Debug dbg{dbg_l ? dbg_l + 64 * i : nullptr, dbg_m ? dbg_m + 384 * i : nullptr, nullptr};
  uint32_t len = proof_len ? proof_len[i] : (uint32_t)stride;
  if (len > stride) len = (uint32_t)stride;
  Fp12 f;
  int st = groth16_miller_one(f, *vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i, n_inputs, dbg);
  if (st == BN254V_OK_TRUE) {
    fbuf[i] = f;
    status[i] = BN254V_STATUS_UNSET;
  } else {
    status[i] = (uint8_t)st;
  }
}
template <int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_groth16_finish(const Groth16VkDev* __restrict__ vk, size_t n, uint8_t* __restrict__ status,
                     const Fp12* __restrict__ fbuf, uint8_t* dbg_gt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (status[i] != BN254V_STATUS_UNSET) return;
  Groth16Debug dbg{nullptr, nullptr, dbg_gt ? dbg_gt + 384 * i : nullptr};
  Fp12 f = fbuf[i];
  status[i] = (uint8_t)groth16_finish_one(f, *vk, dbg);
}

template <int KP, int TPB>
__global__ void __launch_bounds__(TPB, 1)
    k_pairing_product(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, size_t n,
                      uint8_t* __restrict__ is_one, uint8_t* miller_out, uint8_t* gt_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  is_one[i] = pairing_product_one<KP>(g1 + (size_t)64 * KP * i, g2 + (size_t)128 * KP * i,
                                      miller_out ? miller_out + 384 * i : nullptr,
                                      gt_out ? gt_out + 384 * i : nullptr)
                  ? 1
                  : 0;
}

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

// Launch shapes.  One proof per thread; the block is the unit that the phase barriers keep in step, and the grid should
// cover the SMs evenly.  `pick_shape`: big batches use 448-thread blocks, one per SM (14 warps, 128 registers/thread;
// 2^16 proofs = 147 blocks on 148 SMs); small batches use smaller blocks so that every SM gets work.
// BN254V_VARIANT overrides the choice for experiments (1: 128x2, 2: 128x4, 3: 448x1, 6: 32x1, 10: 384x1; 20/21/24: the
// lane-pair kernels at 448x1 / 448x2 / 512x1 -- measured equal or slower than one proof per thread).
static int g_sm_count = 148;
static bool g_two_launch = false;      // shape of the last Groth16 launch (for bn254v_last_kernel_split)
static float g_last_split_ms[2] = {0.f, 0.f};
static int pick_shape(size_t m) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("BN254V_VARIANT");
    forced = e ? atoi(e) : -1;
  }
  if (forced >= 1) return forced;
  // 384 threads x 168 registers is 9 % faster per proof than 448 x 128 (a 16 K-register SMSP holds 3 warps at 168 or 4 at
  // 128), but 2^16 proofs do not fit one wave of it (171 blocks on 148 SMs): use it once there are several waves.
  if (m >= (size_t)g_sm_count * 384 * 4) return 10;     // 384 x 1
  if (m >= (size_t)g_sm_count * 448 * 3 / 4) return 3;  // 448 x 1
  if (m >= (size_t)g_sm_count * 128) return 1;          // 128 x 2
  return 6;                                             // 32-thread blocks: spread thin batches over all SMs
}

static void launch_groth16_verify(cudaStream_t st, const Groth16VkDev* vk, const uint8_t* proofs, size_t stride,
                                  const uint32_t* lens, const uint8_t* inputs, int n_inputs, size_t m, uint8_t* status,
                                  uint8_t* l, uint8_t* ml, uint8_t* gt, Fp12* fbuf = nullptr,
                                  cudaEvent_t mid = nullptr) {
  // Big batches: two launches (Miller loop | final exponentiation), each with about half the code and stack of the fused
  // kernel -- measured 2 % faster at 2^16 and 2^18.  BN254V_VARIANT=42 / 43 force the fused 448 / 384 kernels.
  const int shape = pick_shape(m);
  if (fbuf && (shape == 3 || shape == 10)) {
    if (shape == 3) {
      k_groth16_miller<448><<<(unsigned)((m + 447) / 448), 448, 0, st>>>(vk, proofs, stride, lens, inputs, n_inputs, m,
                                                                        status, fbuf, l, ml);
      if (mid) cudaEventRecord(mid, st);
      k_groth16_finish<448><<<(unsigned)((m + 447) / 448), 448, 0, st>>>(vk, m, status, fbuf, gt);
    } else {
      k_groth16_miller<384><<<(unsigned)((m + 383) / 384), 384, 0, st>>>(vk, proofs, stride, lens, inputs, n_inputs, m,
                                                                        status, fbuf, l, ml);
      if (mid) cudaEventRecord(mid, st);
      k_groth16_finish<384><<<(unsigned)((m + 383) / 384), 384, 0, st>>>(vk, m, status, fbuf, gt);
    }
    g_launches++;  // (the caller counts the other one)
    g_two_launch = true;
    return;
  }
  g_two_launch = false;
#define LV(TPB, MINB)                                                                                              \
  k_groth16_verify<TPB, MINB><<<(unsigned)((m + TPB - 1) / TPB), TPB, 0, st>>>(vk, proofs, stride, lens, inputs, \
                                                                               n_inputs, m, status, l, ml, gt)
#define LVP(TPB, MINB)                                                                                  \
  k_groth16_verify_lp<TPB, MINB><<<(unsigned)((2 * m + TPB - 1) / TPB), TPB, 0, st>>>(                    \
      vk, proofs, stride, lens, inputs, n_inputs, m, status, l, ml, gt)
  switch (pick_shape(m)) {
    case 20: LVP(448, 1); return;  // experimental lane-pair kernels (lanepair.cuh): two lanes per proof
    case 21: LVP(448, 2); return;
    case 24: LVP(512, 1); return;
    default: break;
  }
#undef LVP
  switch (pick_shape(m)) {
    case 2: LV(128, 4); break;
    case 3: case 42: LV(448, 1); break;
    case 43: LV(384, 1); break;
    case 6: LV(32, 1); break;
    case 10: LV(384, 1); break;
    default: LV(128, 2); break;
  }
#undef LV
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

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int bn254v_init(const int* devices, int n_devices) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_inited) return BN254V_SUCCESS;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(BN254V_E_NO_DEVICE, "no CUDA device: %s", e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
  std::vector<int> ids;
  if (devices && n_devices > 0) {
    for (int i = 0; i < n_devices; i++) {
      if (devices[i] < 0 || devices[i] >= count) return fail(BN254V_E_BAD_ARG, "device %d out of range", devices[i]);
      ids.push_back(devices[i]);
    }
  } else {
    for (int i = 0; i < count; i++) ids.push_back(i);
  }
  for (int id : ids) {
    Dev d;
    d.id = id;
    CU(cudaSetDevice(id));
    CU(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, id));
    CU(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&d.ev0));
    CU(cudaEventCreate(&d.ev1));
    CU(cudaEventCreate(&d.evm));
    g_devs.push_back(d);
  }
  g_inited = true;
  return BN254V_SUCCESS;
}

void bn254v_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& d : g_devs) {
    cudaSetDevice(d.id);
    DevBuf::pool_release(d.id);
    cudaStreamDestroy(d.stream);
    cudaEventDestroy(d.ev0);
    cudaEventDestroy(d.ev1);
    cudaEventDestroy(d.evm);
  }
  g_devs.clear();
  g_inited = false;
}

int bn254v_device_count(void) { return (int)g_devs.size(); }
const char* bn254v_last_error(void) { return g_err.c_str(); }
uint64_t bn254v_launch_count(void) { return g_launches.load(); }

const char* bn254v_status_name(int s) {
  switch (s) {
    case BN254V_OK_TRUE: return "OK_TRUE";
    case BN254V_OK_FALSE: return "OK_FALSE";
    case BN254V_ERR_PREPARE_INPUTS: return "ERR_PREPARE_INPUTS";
    case BN254V_ERR_BSB22_MISMATCH: return "ERR_BSB22_MISMATCH";
    case BN254V_ERR_INVALID_WITNESS: return "ERR_INVALID_WITNESS";
    case BN254V_ERR_INVERSE_NOT_FOUND: return "ERR_INVERSE_NOT_FOUND";
    case BN254V_ERR_OPENING_POLY_MISMATCH: return "ERR_OPENING_POLY_MISMATCH";
    case BN254V_ERR_INVALID_NUMBER_OF_DIGESTS: return "ERR_INVALID_NUMBER_OF_DIGESTS";
    case BN254V_ERR_PAIRING_CHECK_FAILED: return "ERR_PAIRING_CHECK_FAILED";
    case BN254V_PANIC_FIELD_NOT_MEMBER: return "PANIC_FIELD_NOT_MEMBER";
    case BN254V_PANIC_NOT_ON_CURVE: return "PANIC_NOT_ON_CURVE";
    case BN254V_PANIC_NOT_IN_SUBGROUP: return "PANIC_NOT_IN_SUBGROUP";
    case BN254V_PANIC_IDENTITY: return "PANIC_IDENTITY";
    case BN254V_PANIC_SHORT_BUFFER: return "PANIC_SHORT_BUFFER";
    case BN254V_PANIC_DIV_BY_ZERO: return "PANIC_DIV_BY_ZERO";
    case BN254V_PANIC_INDEX_OUT_OF_RANGE: return "PANIC_INDEX_OUT_OF_RANGE";
    case BN254V_STATUS_UNSET: return "UNSET";
  }
  return "?";
}

// ---- verifying keys ----------------------------------------------------------------------------
int bn254v_groth16_vk_load(const uint8_t* vk_bytes, size_t len, int sign_mode, bn254v_vk** out) {
  if (!vk_bytes || !out || (sign_mode != 0 && sign_mode != 1)) return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  gnark::Groth16VkHost h;
  if (gnark::parse_groth16_vk(h, vk_bytes, len)) return fail(BN254V_E_VK_PARSE, "malformed Groth16 VK");
  if (h.k.empty() || h.k.size() > BN_MAX_IC)
    return fail(BN254V_E_UNSUPPORTED, "|IC| = %zu outside [1, %d]", h.k.size(), BN_MAX_IC);
  // h.beta2 is -beta_file (as the reference stores it).  sign_mode 0: (beta', gamma', delta') =
  // (-beta_file, gamma, -delta); sign_mode 1: (beta_file, -gamma, -delta).
  Groth16VkDev* hv = new Groth16VkDev();
  memset(hv, 0, sizeof *hv);
  hv->n_ic = (int)h.k.size();
  hv->alpha = h.alpha;
  hv->beta = sign_mode == 0 ? h.beta2 : neg(h.beta2);
  hv->gamma = sign_mode == 0 ? h.gamma2 : neg(h.gamma2);
  hv->delta = neg(h.delta2);
  for (size_t i = 0; i < h.k.size(); i++) hv->ic[i] = h.k[i];
  bn254v_vk* vk = new bn254v_vk();
  vk->kind = 0;
  vk->n_public = hv->n_ic - 1;
  vk->sign_mode = sign_mode;
  for (auto& d : g_devs) {
    cudaError_t e = cudaSetDevice(d.id);
    Groth16VkDev* dv = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&dv, sizeof(Groth16VkDev));
    G1Aff* table = nullptr;
    const int n_bases = hv->n_ic - 1;
    if (e == cudaSuccess && n_bases > 0)
      e = cudaMalloc(&table, sizeof(G1Aff) * (size_t)n_bases * BN_IC_WINDOWS * BN_IC_ENTRIES);
    hv->ic_table = table;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv, hv, sizeof(Groth16VkDev), cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
      k_groth16_vk_prepare<<<1, 32, 0, d.stream>>>(dv);
      g_launches++;
      if (n_bases > 0) {
        k_g1_fixed_tables<<<n_bases, BN_IC_WINDOWS, 0, d.stream>>>(&dv->ic[1], table);
        g_launches++;
      }
      e = cudaGetLastError();
    }
    vk->aux.push_back(table);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    if (e != cudaSuccess) {
      delete hv;
      bn254v_vk_free(vk);
      return fail(BN254V_E_CUDA, "vk upload/prepare: %s", cudaGetErrorString(e));
    }
    vk->dev.push_back(dv);
  }
  delete hv;
  *out = vk;
  return BN254V_SUCCESS;
}

int bn254v_plonk_vk_load(const uint8_t* vk_bytes, size_t len, bn254v_vk** out) {
  if (!vk_bytes || !out) return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  gnark::PlonkVkHost h;
  if (gnark::parse_plonk_vk(h, vk_bytes, len)) return fail(BN254V_E_VK_PARSE, "malformed PlonK VK");
  if (h.qcp.size() > BN_MAX_QCP || h.cci.size() != h.qcp.size() || h.nb_public > BN_MAX_PLONK_PUBLIC)
    return fail(BN254V_E_UNSUPPORTED, "VK shape outside compiled limits (nQcp %zu, nIdx %zu, nPublic %llu)",
                h.qcp.size(), h.cci.size(), (unsigned long long)h.nb_public);
  PlonkVkDev* hv = new PlonkVkDev();
  memset(hv, 0, sizeof *hv);
  hv->size = h.size;
  hv->n_public = (int)h.nb_public;
  hv->n_qcp = (int)h.qcp.size();
  hv->size_inv = fe_to_mont(h.size_inv);
  hv->generator = fe_to_mont(h.generator);
  hv->coset_shift = fe_to_mont(h.coset_shift);
  for (int i = 0; i < hv->n_qcp; i++) hv->w_pow_cci[i] = fr_pow_u64(hv->generator, h.nb_public + h.cci[i]);
  for (int i = 0; i < 3; i++) hv->s[i] = h.s[i];
  hv->ql = h.ql, hv->qr = h.qr, hv->qm = h.qm, hv->qo = h.qo, hv->qk = h.qk, hv->g1 = h.g1;
  hv->g2[0] = h.g2[0], hv->g2[1] = h.g2[1];
  for (int i = 0; i < hv->n_qcp; i++) hv->qcp[i] = h.qcp[i];
  {  // "gamma" | S1 S2 S3 Ql Qr Qm Qo Qk | Qcp..  (bind_public_data, verifier/src/plonk/verify.rs:325-335)
    sha256_init(hv->gamma_prefix);
    sha_bytes(hv->gamma_prefix, "gamma", 5);
    const G1Aff* pts[8] = {&hv->s[0], &hv->s[1], &hv->s[2], &hv->ql, &hv->qr, &hv->qm, &hv->qo, &hv->qk};
    uint8_t b[64];
    for (int i = 0; i < 8; i++) {
      store_g1(b, *pts[i]);
      sha256_update(hv->gamma_prefix, b, 64);
    }
    for (int i = 0; i < hv->n_qcp; i++) {
      store_g1(b, hv->qcp[i]);
      sha256_update(hv->gamma_prefix, b, 64);
    }
    store_g1(hv->kzg_vk_bytes, hv->s[0]);
    store_g1(hv->kzg_vk_bytes + 64, hv->s[1]);
    for (int i = 0; i < hv->n_qcp; i++) store_g1(hv->kzg_vk_bytes + 128 + 64 * i, hv->qcp[i]);
  }
  bn254v_vk* vk = new bn254v_vk();
  vk->kind = 1;
  vk->n_public = hv->n_public;
  vk->n_qcp = hv->n_qcp;
  vk->sign_mode = 0;
  for (auto& d : g_devs) {
    cudaError_t e = cudaSetDevice(d.id);
    PlonkVkDev* dv = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&dv, sizeof(PlonkVkDev));
    // fixed-base window tables of the VK-constant MSM bases: [bases | tables] in one allocation
    const int n_fixed = BN_PLONK_N_FIXED(hv->n_qcp);
    G1Aff* tab_mem = nullptr;
    if (e == cudaSuccess)
      e = cudaMalloc(&tab_mem, sizeof(G1Aff) * ((size_t)n_fixed + (size_t)n_fixed * BN_IC_WINDOWS * BN_IC_ENTRIES));
    if (e == cudaSuccess) {
      std::vector<G1Aff> bases(n_fixed);
      for (int i = 0; i < n_fixed; i++) bases[i] = plonk_fixed_base(*hv, i);
      e = cudaMemcpy(tab_mem, bases.data(), sizeof(G1Aff) * n_fixed, cudaMemcpyHostToDevice);
    }
    hv->fixed_tables = tab_mem ? tab_mem + n_fixed : nullptr;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv, hv, sizeof(PlonkVkDev), cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
      k_plonk_vk_prepare<<<1, 32, 0, d.stream>>>(dv);
      k_g1_fixed_tables<<<n_fixed, BN_IC_WINDOWS, 0, d.stream>>>(tab_mem, tab_mem + n_fixed);
      g_launches += 2;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    vk->aux.push_back(tab_mem);
    if (e != cudaSuccess) {
      delete hv;
      bn254v_vk_free(vk);
      return fail(BN254V_E_CUDA, "vk upload/prepare: %s", cudaGetErrorString(e));
    }
    vk->dev.push_back(dv);
  }
  delete hv;
  *out = vk;
  return BN254V_SUCCESS;
}

void bn254v_vk_free(bn254v_vk* vk) {
  if (!vk) return;
  for (size_t i = 0; i < vk->dev.size() && i < g_devs.size(); i++) {
    cudaSetDevice(g_devs[i].id);
    cudaFree(vk->dev[i]);
    if (i < vk->aux.size() && vk->aux[i]) cudaFree(vk->aux[i]);
  }
  delete vk;
}

int bn254v_vk_n_public(const bn254v_vk* vk) { return vk ? vk->n_public : -1; }

// ---- Groth16 batch -----------------------------------------------------------------------------
int bn254v_groth16_verify_batch(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs, size_t n,
                                uint8_t* status, const bn254v_debug* dbg) {
  if (!vk || vk->kind != 0 || !status || (n && (!proofs || (n_inputs > 0 && !inputs_be))) || n_inputs < 0 ||
      n_inputs > 64)
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) return BN254V_SUCCESS;
  const int nd = (int)g_devs.size();
  struct Part {
    DevBuf proofs, lens, inputs, status, l, m, gt, fbuf;
  };
  std::vector<Part> parts(nd);
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (int d = 0; d < nd; d++) {
    size_t lo, hi;
    shard(n, d, nd, lo, hi);
    size_t m = hi - lo;
    if (!m) continue;
    Part& p = parts[d];
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(p.proofs.alloc(m * proof_stride));
    CU(p.inputs.alloc(m * in_bytes));
    CU(p.status.alloc(m));
    CU(cudaMemcpyAsync(p.proofs.p, proofs + lo * proof_stride, m * proof_stride, cudaMemcpyHostToDevice, dev.stream));
    if (in_bytes)
      CU(cudaMemcpyAsync(p.inputs.p, inputs_be + lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, dev.stream));
    if (proof_len) {
      CU(p.lens.alloc(m * 4));
      CU(cudaMemcpyAsync(p.lens.p, proof_len + lo, m * 4, cudaMemcpyHostToDevice, dev.stream));
    }
    if (dbg && dbg->g1_out) CU(p.l.alloc(m * 64));
    if (dbg && dbg->miller_out) CU(p.m.alloc(m * 384));
    if (dbg && dbg->gt_out) CU(p.gt.alloc(m * 384));
    CU(p.fbuf.alloc(m * sizeof(Fp12)));
    launch_groth16_verify(dev.stream, (const Groth16VkDev*)vk->dev[d], p.proofs.as<uint8_t>(), proof_stride,
                          proof_len ? p.lens.as<uint32_t>() : nullptr, p.inputs.as<uint8_t>(), n_inputs, m,
                          p.status.as<uint8_t>(), p.l.as<uint8_t>(), p.m.as<uint8_t>(), p.gt.as<uint8_t>(),
                          p.fbuf.as<Fp12>());
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(status + lo, p.status.p, m, cudaMemcpyDeviceToHost, dev.stream));
    if (p.l.p) CU(cudaMemcpyAsync(dbg->g1_out + lo * 64, p.l.p, m * 64, cudaMemcpyDeviceToHost, dev.stream));
    if (p.m.p) CU(cudaMemcpyAsync(dbg->miller_out + lo * 384, p.m.p, m * 384, cudaMemcpyDeviceToHost, dev.stream));
    if (p.gt.p) CU(cudaMemcpyAsync(dbg->gt_out + lo * 384, p.gt.p, m * 384, cudaMemcpyDeviceToHost, dev.stream));
  }
  for (int d = 0; d < nd; d++) {
    CU(cudaSetDevice(g_devs[d].id));
    CU(cudaStreamSynchronize(g_devs[d].stream));
  }
  return BN254V_SUCCESS;
}

int bn254v_plonk_verify_batch(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                              const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                              const uint8_t* rnd_be, size_t n, uint8_t* status, const bn254v_debug* dbg) {
  if (!vk || vk->kind != 1 || !status || (n && (!proofs || !rnd_be || (n_inputs > 0 && !inputs_be))) || n_inputs < 0 ||
      n_inputs > BN_MAX_PLONK_PUBLIC)
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) return BN254V_SUCCESS;
  const int nd = (int)g_devs.size();
  struct Part {
    DevBuf proofs, lens, inputs, rnd, status, g1, fr, m, gt, work, list, count;
  };
  std::vector<Part> parts(nd);
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (int d = 0; d < nd; d++) {
    size_t lo, hi;
    shard(n, d, nd, lo, hi);
    size_t m = hi - lo;
    if (!m) continue;
    Part& p = parts[d];
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(p.proofs.alloc(m * proof_stride));
    CU(p.inputs.alloc(m * in_bytes));
    CU(p.rnd.alloc(m * 32));
    CU(p.status.alloc(m));
    CU(cudaMemcpyAsync(p.proofs.p, proofs + lo * proof_stride, m * proof_stride, cudaMemcpyHostToDevice, dev.stream));
    if (in_bytes)
      CU(cudaMemcpyAsync(p.inputs.p, inputs_be + lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, dev.stream));
    CU(cudaMemcpyAsync(p.rnd.p, rnd_be + lo * 32, m * 32, cudaMemcpyHostToDevice, dev.stream));
    if (proof_len) {
      CU(p.lens.alloc(m * 4));
      CU(cudaMemcpyAsync(p.lens.p, proof_len + lo, m * 4, cudaMemcpyHostToDevice, dev.stream));
    }
    if (dbg && dbg->g1_out) CU(p.g1.alloc(m * 256));
    if (dbg && dbg->fr_out) CU(p.fr.alloc(m * 256));
    if (dbg && dbg->miller_out) CU(p.m.alloc(m * 384));
    if (dbg && dbg->gt_out) CU(p.gt.alloc(m * 384));
    if (p.g1.p) CU(cudaMemsetAsync(p.g1.p, 0, m * 256, dev.stream));
    if (p.fr.p) CU(cudaMemsetAsync(p.fr.p, 0, m * 256, dev.stream));
    if (p.m.p) CU(cudaMemsetAsync(p.m.p, 0, m * 384, dev.stream));
    if (p.gt.p) CU(cudaMemsetAsync(p.gt.p, 0, m * 384, dev.stream));
    {
      // chunks bound the per-proof workspace (PlonkWork, ~1.9 KB): 2^16 proofs -> 125 MB
      const size_t CH = 1u << 16;
      const size_t mc = m < CH ? m : CH;
      DevBuf &work = p.work, &list = p.list, &count = p.count;  // live until the final synchronisation
      CU(work.alloc(mc * sizeof(PlonkWork)));
      CU(list.alloc(mc * sizeof(int)));
      CU(count.alloc(sizeof(int)));
      const PlonkVkDev* dvk = (const PlonkVkDev*)vk->dev[d];
      for (size_t c0 = 0; c0 < m; c0 += CH) {
        const size_t cm = m - c0 < CH ? m - c0 : CH;
        const uint8_t* cp = p.proofs.as<uint8_t>() + c0 * proof_stride;
        PlonkDbgPtrs dp{p.g1.p ? p.g1.as<uint8_t>() + c0 * 256 : nullptr, p.fr.p ? p.fr.as<uint8_t>() + c0 * 256 : nullptr,
                        p.m.p ? p.m.as<uint8_t>() + c0 * 384 : nullptr, p.gt.p ? p.gt.as<uint8_t>() + c0 * 384 : nullptr};
        uint8_t* cst = p.status.as<uint8_t>() + c0;
        const unsigned g64 = (unsigned)((cm + 63) / 64);
        const int n_terms = vk->n_qcp + 10;
        CU(cudaMemsetAsync(count.p, 0, sizeof(int), dev.stream));
        k_plonk_stage_a<<<g64, 64, 0, dev.stream>>>(dvk, cp, proof_stride, proof_len ? p.lens.as<uint32_t>() + c0 : nullptr,
                                                    p.inputs.as<uint8_t>() + c0 * in_bytes, n_inputs, cm, cst,
                                                    work.as<PlonkWork>(), list.as<int>(), count.as<int>(), dp);
        k_plonk_terms<<<dim3(g64, n_terms), 64, 0, dev.stream>>>(dvk, cp, proof_stride, work.as<PlonkWork>(),
                                                                 list.as<int>(), count.as<int>(), 0);
        k_plonk_stage_c<<<g64, 64, 0, dev.stream>>>(dvk, cp, proof_stride, p.rnd.as<uint8_t>() + c0 * 32, cst,
                                                    work.as<PlonkWork>(), list.as<int>(), count.as<int>(), dp);
        k_plonk_terms<<<dim3(g64, n_terms), 64, 0, dev.stream>>>(dvk, cp, proof_stride, work.as<PlonkWork>(),
                                                                 list.as<int>(), count.as<int>(), 1);
        if (pick_shape(cm) >= 20)
          k_plonk_stage_e_lp<128><<<(unsigned)((2 * cm + 127) / 128), 128, 0, dev.stream>>>(
              dvk, cp, proof_stride, cst, work.as<PlonkWork>(), list.as<int>(), count.as<int>(), dp);
        else if (pick_shape(cm) == 6)
          k_plonk_stage_e<32><<<(unsigned)((cm + 31) / 32), 32, 0, dev.stream>>>(dvk, cp, proof_stride, cst,
                                                                                 work.as<PlonkWork>(), list.as<int>(),
                                                                                 count.as<int>(), dp);
        else
          k_plonk_stage_e<128><<<(unsigned)((cm + 127) / 128), 128, 0, dev.stream>>>(dvk, cp, proof_stride, cst,
                                                                                     work.as<PlonkWork>(), list.as<int>(),
                                                                                     count.as<int>(), dp);
        g_launches += 5;
        CU(cudaGetLastError());
      }
    }
    CU(cudaMemcpyAsync(status + lo, p.status.p, m, cudaMemcpyDeviceToHost, dev.stream));
    if (p.g1.p) CU(cudaMemcpyAsync(dbg->g1_out + lo * 256, p.g1.p, m * 256, cudaMemcpyDeviceToHost, dev.stream));
    if (p.fr.p) CU(cudaMemcpyAsync(dbg->fr_out + lo * 256, p.fr.p, m * 256, cudaMemcpyDeviceToHost, dev.stream));
    if (p.m.p) CU(cudaMemcpyAsync(dbg->miller_out + lo * 384, p.m.p, m * 384, cudaMemcpyDeviceToHost, dev.stream));
    if (p.gt.p) CU(cudaMemcpyAsync(dbg->gt_out + lo * 384, p.gt.p, m * 384, cudaMemcpyDeviceToHost, dev.stream));
  }
  for (int d = 0; d < nd; d++) {
    CU(cudaSetDevice(g_devs[d].id));
    CU(cudaStreamSynchronize(g_devs[d].stream));
  }
  return BN254V_SUCCESS;
}

// ---- raw pairing products ----------------------------------------------------------------------
int bn254v_pairing_product_batch(const uint8_t* g1, const uint8_t* g2, int k, size_t n, uint8_t* is_one,
                                 uint8_t* miller_out, uint8_t* gt_out) {
  if (k < 1 || k > 4 || !is_one || (n && (!g1 || !g2))) return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) return BN254V_SUCCESS;
  const int nd = (int)g_devs.size();
  struct Part {
    DevBuf g1, g2, one, m, gt;
  };
  std::vector<Part> parts(nd);
  for (int d = 0; d < nd; d++) {
    size_t lo, hi;
    shard(n, d, nd, lo, hi);
    size_t m = hi - lo;
    if (!m) continue;
    Part& p = parts[d];
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(p.g1.alloc(m * 64 * k));
    CU(p.g2.alloc(m * 128 * k));
    CU(p.one.alloc(m));
    if (miller_out) CU(p.m.alloc(m * 384));
    if (gt_out) CU(p.gt.alloc(m * 384));
    CU(cudaMemcpyAsync(p.g1.p, g1 + lo * 64 * k, m * 64 * k, cudaMemcpyHostToDevice, dev.stream));
    CU(cudaMemcpyAsync(p.g2.p, g2 + lo * 128 * k, m * 128 * k, cudaMemcpyHostToDevice, dev.stream));
#define LAUNCH_PP2(KP, TPB)                                                                             \
  k_pairing_product<KP, TPB><<<(unsigned)((m + TPB - 1) / TPB), TPB, 0, dev.stream>>>(                 \
      p.g1.as<uint8_t>(), p.g2.as<uint8_t>(), m, p.one.as<uint8_t>(), p.m.as<uint8_t>(), p.gt.as<uint8_t>())
#define LAUNCH_PP(KP)                          \
  switch (pick_shape(m)) {                     \
    case 20: case 21: case 24:                                                                                    \
      k_pairing_product_lp<KP, 448><<<(unsigned)((2 * m + 447) / 448), 448, 0, dev.stream>>>(                     \
          p.g1.as<uint8_t>(), p.g2.as<uint8_t>(), m, p.one.as<uint8_t>(), p.m.as<uint8_t>(), p.gt.as<uint8_t>()); \
      break;                                   \
    case 3: case 10: LAUNCH_PP2(KP, 448); break; \
    case 6: LAUNCH_PP2(KP, 32); break;         \
    default: LAUNCH_PP2(KP, 128); break;       \
  }
    switch (k) {
      case 1: LAUNCH_PP(1); break;
      case 2: LAUNCH_PP(2); break;
      case 3: LAUNCH_PP(3); break;
      default: LAUNCH_PP(4); break;
    }
#undef LAUNCH_PP
#undef LAUNCH_PP2
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(is_one + lo, p.one.p, m, cudaMemcpyDeviceToHost, dev.stream));
    if (miller_out) CU(cudaMemcpyAsync(miller_out + lo * 384, p.m.p, m * 384, cudaMemcpyDeviceToHost, dev.stream));
    if (gt_out) CU(cudaMemcpyAsync(gt_out + lo * 384, p.gt.p, m * 384, cudaMemcpyDeviceToHost, dev.stream));
  }
  for (int d = 0; d < nd; d++) {
    CU(cudaSetDevice(g_devs[d].id));
    CU(cudaStreamSynchronize(g_devs[d].stream));
  }
  return BN254V_SUCCESS;
}

// ---- device-resident batches -------------------------------------------------------------------
int bn254v_groth16_batch_upload(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                const uint8_t* inputs_be, int n_inputs, size_t n, bn254v_batch** out) {
  if (!vk || vk->kind != 0 || !proofs || !out || proof_stride < 256 || n_inputs < 0 || (n_inputs && !inputs_be))
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  const int nd = (int)g_devs.size();
  bn254v_batch* b = new bn254v_batch();
  b->n = n;
  b->n_inputs = n_inputs;
  b->parts.resize(nd);
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (int d = 0; d < nd; d++) {
    auto& p = b->parts[d];
    shard(n, d, nd, p.lo, p.hi);
    size_t m = p.hi - p.lo;
    p.proofs = p.inputs = p.status = nullptr;
    p.fbuf = nullptr;
    if (!m) continue;
    Dev& dev = g_devs[d];
    cudaError_t e = cudaSetDevice(dev.id);
    if (e == cudaSuccess) e = cudaMalloc(&p.proofs, m * 256);
    if (e == cudaSuccess) e = cudaMalloc(&p.inputs, m * in_bytes + 1);
    if (e == cudaSuccess) e = cudaMalloc(&p.status, m);
    if (e == cudaSuccess) e = cudaMalloc(&p.fbuf, m * sizeof(Fp12));
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(p.proofs, 256, proofs + p.lo * proof_stride, proof_stride, 256, m, cudaMemcpyHostToDevice,
                            dev.stream);
    if (e == cudaSuccess && in_bytes)
      e = cudaMemcpyAsync(p.inputs, inputs_be + p.lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, dev.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dev.stream);
    if (e != cudaSuccess) {
      bn254v_batch_free(b);
      return fail(BN254V_E_CUDA, "batch upload: %s", cudaGetErrorString(e));
    }
  }
  *out = b;
  return BN254V_SUCCESS;
}

int bn254v_groth16_batch_verify(const bn254v_vk* vk, bn254v_batch* b, uint8_t* status, float* kernel_ms) {
  if (!vk || vk->kind != 0 || !b) return fail(BN254V_E_BAD_ARG, "bad argument");
  const int nd = (int)g_devs.size();
  if ((int)b->parts.size() != nd) return fail(BN254V_E_BAD_ARG, "batch was staged for another device set");
  for (int d = 0; d < nd; d++) {
    auto& p = b->parts[d];
    size_t m = p.hi - p.lo;
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(cudaEventRecord(dev.ev0, dev.stream));
    if (m) {
      launch_groth16_verify(dev.stream, (const Groth16VkDev*)vk->dev[d], p.proofs, 256, nullptr, p.inputs,
                            b->n_inputs, m, p.status, nullptr, nullptr, nullptr, p.fbuf, dev.evm);
      g_launches++;
      CU(cudaGetLastError());
    }
    CU(cudaEventRecord(dev.ev1, dev.stream));
  }
  float worst = 0.f;
  for (int d = 0; d < nd; d++) {
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(cudaStreamSynchronize(dev.stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, dev.ev0, dev.ev1));
    if (ms > worst) worst = ms;
    if (d == 0) {
      g_last_split_ms[0] = ms;
      g_last_split_ms[1] = 0.f;
      if (g_two_launch && b->parts[0].hi > b->parts[0].lo) {
        CU(cudaEventElapsedTime(&g_last_split_ms[0], dev.ev0, dev.evm));
        CU(cudaEventElapsedTime(&g_last_split_ms[1], dev.evm, dev.ev1));
      }
    }
  }
  if (kernel_ms) *kernel_ms = worst;
  if (status) {
    for (int d = 0; d < nd; d++) {
      auto& p = b->parts[d];
      if (p.hi == p.lo) continue;
      CU(cudaSetDevice(g_devs[d].id));
      CU(cudaMemcpy(status + p.lo, p.status, p.hi - p.lo, cudaMemcpyDeviceToHost));
    }
  }
  return BN254V_SUCCESS;
}

void bn254v_batch_free(bn254v_batch* b) {
  if (!b) return;
  for (size_t d = 0; d < b->parts.size() && d < g_devs.size(); d++) {
    cudaSetDevice(g_devs[d].id);
    cudaFree(b->parts[d].proofs);
    cudaFree(b->parts[d].inputs);
    cudaFree(b->parts[d].status);
    cudaFree(b->parts[d].fbuf);
  }
  delete b;
}

// ---- synthetic workloads -----------------------------------------------------------------------
static void host_g1_mul_gen(G1Aff& out, const Fr& k_mont) {
  Fr k = fe_from_mont(k_mont);
  to_affine(out, scalar_mul(g1_generator(), k.v));
}
static void host_g2_mul_gen(G2Aff& out, const Fr& k_mont) {
  Fr k = fe_from_mont(k_mont);
  to_affine(out, scalar_mul(g2_generator_dev(), k.v));
}

int bn254v_groth16_synth(uint64_t seed, int n_public, int sign_mode, size_t first_index, size_t n, uint8_t* vk_bytes,
                         size_t* vk_len, uint8_t* proofs, uint8_t* inputs_be, uint8_t* expected) {
  if (n_public < 1 || n_public + 1 > BN_MAX_IC_SYNTH || (sign_mode != 0 && sign_mode != 1))
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  Groth16Trapdoor td;
  trapdoor_init(td, seed, n_public);
  if (vk_bytes) {
    // gnark layout: alpha1 | beta1 | beta2 | gamma2 | delta1 | delta2 | u32 |K| | K.. |
    // u32 0 | Pedersen g, gRootSigmaNeg (parsed, unused: the G2 generator twice)
    size_t need = 288 + 4 + 32 * (size_t)(n_public + 1) + 4 + 128;
    if (!vk_len || *vk_len < need) return fail(BN254V_E_BAD_ARG, "vk buffer too small (%zu needed)", need);
    G1Aff a1, b1, d1;
    G2Aff b2, g2, d2, gen = g2_generator_dev();
    host_g1_mul_gen(a1, td.alpha);
    host_g1_mul_gen(b1, td.beta);
    host_g2_mul_gen(b2, td.beta);
    host_g2_mul_gen(g2, td.gamma);
    host_g1_mul_gen(d1, td.delta);
    host_g2_mul_gen(d2, td.delta);
    uint8_t* o = vk_bytes;
    gnark::compress_g1(o, a1);
    gnark::compress_g1(o + 32, b1);
    gnark::compress_g2(o + 64, b2);
    gnark::compress_g2(o + 128, g2);
    gnark::compress_g1(o + 192, d1);
    gnark::compress_g2(o + 224, d2);
    uint32_t nk = (uint32_t)(n_public + 1);
    o[288] = (uint8_t)(nk >> 24), o[289] = (uint8_t)(nk >> 16), o[290] = (uint8_t)(nk >> 8), o[291] = (uint8_t)nk;
    o += 292;
    for (uint32_t i = 0; i < nk; i++, o += 32) {
      G1Aff k;
      host_g1_mul_gen(k, td.ic[i]);
      gnark::compress_g1(o, k);
    }
    memset(o, 0, 4);
    o += 4;
    gnark::compress_g2(o, gen);
    gnark::compress_g2(o + 64, gen);
    *vk_len = need;
  }
  if (n == 0) return BN254V_SUCCESS;
  if (!proofs || !inputs_be || !expected) return fail(BN254V_E_BAD_ARG, "null output buffer");
  Dev& dev = g_devs[0];
  CU(cudaSetDevice(dev.id));
  DevBuf dp, di, de;
  CU(dp.alloc(n * 256));
  CU(di.alloc(n * 32 * n_public));
  CU(de.alloc(n));
  unsigned grid = (unsigned)((n + BN_TPB - 1) / BN_TPB);
  k_groth16_synth<<<grid, BN_TPB, 0, dev.stream>>>(td, seed, first_index, n, n_public, sign_mode, dp.as<uint8_t>(),
                                                   di.as<uint8_t>(), de.as<uint8_t>());
  g_launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(proofs, dp.p, n * 256, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaMemcpyAsync(inputs_be, di.p, n * 32 * n_public, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaMemcpyAsync(expected, de.p, n, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaStreamSynchronize(dev.stream));
  return BN254V_SUCCESS;
}

int bn254v_pairing_synth(uint64_t seed, int k, size_t first_index, size_t n, uint8_t* g1, uint8_t* g2,
                         uint8_t* expected_is_one) {
  if (k < 1 || k > 4 || !g1 || !g2 || !expected_is_one) return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) return BN254V_SUCCESS;
  Dev& dev = g_devs[0];
  CU(cudaSetDevice(dev.id));
  DevBuf d1, d2, de;
  CU(d1.alloc(n * 64 * k));
  CU(d2.alloc(n * 128 * k));
  CU(de.alloc(n));
  unsigned grid = (unsigned)((n + BN_TPB - 1) / BN_TPB);
  k_pairing_synth<<<grid, BN_TPB, 0, dev.stream>>>(seed, first_index, n, k, d1.as<uint8_t>(), d2.as<uint8_t>(),
                                                   de.as<uint8_t>());
  g_launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(g1, d1.p, n * 64 * k, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaMemcpyAsync(g2, d2.p, n * 128 * k, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaMemcpyAsync(expected_is_one, de.p, n, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaStreamSynchronize(dev.stream));
  return BN254V_SUCCESS;
}

// ---- measurement helpers -----------------------------------------------------------------------
int bn254v_last_kernel_split(float* miller_ms, float* finish_ms) {
  if (miller_ms) *miller_ms = g_last_split_ms[0];
  if (finish_ms) *finish_ms = g_last_split_ms[1];
  return BN254V_SUCCESS;
}

int bn254v_imad_peak(int iters, double* wide_mac_per_s, double* lo_mac_per_s, float* sm_clock_mhz) {
  if (iters < 1) return fail(BN254V_E_BAD_ARG, "iters < 1");
  int rc = ensure_init();
  if (rc) return rc;
  Dev& dev = g_devs[0];
  CU(cudaSetDevice(dev.id));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev.id));
  DevBuf sink;
  CU(sink.alloc(8));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  const double macs = (double)blocks * threads * (double)iters * 32.0;
  float ms;
  for (int pass = 0; pass < 2; pass++) {  // pass 0 warms up
    CU(cudaEventRecord(dev.ev0, dev.stream));
    k_imad_peak<true><<<blocks, threads, 0, dev.stream>>>(iters, 12345u, 6789u, sink.as<uint64_t>());
    CU(cudaEventRecord(dev.ev1, dev.stream));
    CU(cudaStreamSynchronize(dev.stream));
    CU(cudaEventElapsedTime(&ms, dev.ev0, dev.ev1));
    g_launches++;
  }
  if (wide_mac_per_s) *wide_mac_per_s = macs / (ms * 1e-3);
  for (int pass = 0; pass < 2; pass++) {
    CU(cudaEventRecord(dev.ev0, dev.stream));
    k_imad_peak<false><<<blocks, threads, 0, dev.stream>>>(iters, 12345u, 6789u, sink.as<uint64_t>());
    CU(cudaEventRecord(dev.ev1, dev.stream));
    CU(cudaStreamSynchronize(dev.stream));
    CU(cudaEventElapsedTime(&ms, dev.ev0, dev.ev1));
    g_launches++;
  }
  if (lo_mac_per_s) *lo_mac_per_s = macs / (ms * 1e-3);
  if (sm_clock_mhz) *sm_clock_mhz = prop.clockRate / 1000.f;
  return BN254V_SUCCESS;
}

}  // extern "C"
