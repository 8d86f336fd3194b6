// bn254v: the C ABI (include/bn254v.h, include/bn254v_bench.h) -- VK handles and their cache, batch sharding over
// devices, host <-> device copies, mixed-batch grouping.  The kernels live in k_*.cu behind kernels.h.
// sm_100a only; there is no CPU fallback: every compute entry point fails with BN254V_E_NO_DEVICE when no CUDA device
// is usable.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/random.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bn254v_bench.h"
#include "gnark_host.h"
#include "kernels.h"

using namespace bn254;

// ------------------------------------------------------------------------------------------------
// library state
// ------------------------------------------------------------------------------------------------
namespace {

struct Dev {
  int id;
  cudaStream_t stream;
  cudaEvent_t ev[8];  // stage boundaries of the device-resident runs: ev[0] start ... ev[k] end of stage k
  cudaStream_t side;  // second stream and its two ordering events (aggregate Groth16 check)
  cudaEvent_t side_ev[2];
  cudaEvent_t lane_ev[2];  // bn254v_verify_many: "the kernels of the chunk on lane 0 / 1 are done" (see Lane)
};

std::mutex g_mu;
std::vector<Dev> g_devs;
bool g_inited = false;
int g_sm_count = 148;
std::atomic<uint64_t> g_launches{0};
thread_local std::string g_err;
float g_stage_ms[8];
int g_stage_n = 0;

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
// milliseconds (and cudaFree synchronises the device), which is visible next to a 30 ms batch.
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

// Pinned host staging (mixed batches gather their groups into it so that the H2D copies run at full PCIe speed and
// asynchronously); cudaHostAlloc costs tens of milliseconds, so the blocks are kept for the life of the library.
struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~PinnedBuf() {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pinned().push_back(Block{p, cap, -1});
  }
  static std::vector<Block>& g_pinned() {
    static std::vector<Block> v;
    return v;
  }
  cudaError_t reserve(size_t n) {
    if (!n) n = 1;
    if (cap >= n) return cudaSuccess;
    {
      std::lock_guard<std::mutex> lk(g_pool_mu);
      if (p) g_pinned().push_back(Block{p, cap, -1});
      p = nullptr, cap = 0;
      auto& v = g_pinned();
      int best = -1;
      for (int i = 0; i < (int)v.size(); i++)
        if (v[i].cap >= n && (best < 0 || v[i].cap < v[best].cap)) best = i;
      if (best >= 0) {
        p = v[best].p, cap = v[best].cap;
        v.erase(v.begin() + best);
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocPortable);
    if (e == cudaSuccess) cap = n;
    else p = nullptr;
    return e;
  }
  static void release_all() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& b : g_pinned()) cudaFreeHost(b.p);
    g_pinned().clear();
  }
  template <class T>
  T* as() { return (T*)p; }
};

// Lane: which of a device's two streams a batch call runs on.  The exported entry points use lane 0 alone.
// bn254v_verify_many keeps two chunks in flight, one per lane: the copies of a chunk to the device run while the kernels
// of the chunk before it are still going, and its kernels are ordered behind them by an event (kernels of two chunks
// side by side would only slow each other down), so the device goes from the kernels of one chunk straight to those of
// the next.  `chained`: wait for the other lane's kernels before launching, record this lane's event after them.
struct Lane {
  int id = 0;
  bool chained = false;
  cudaStream_t stream(const Dev& d) const { return id ? d.side : d.stream; }
};
// Declared AFTER the per-device buffers of a batch call, so that it is destroyed BEFORE them on every exit path --
// including the early returns of CU() -- and no asynchronous copy still uses caller memory, or a scratch block that is
// about to go back to the pool, when the function returns.
struct SyncGuard {
  int lane;  // -1: both streams
  explicit SyncGuard(int lane_ = -1) : lane(lane_) {}
  ~SyncGuard() {
    for (auto& d : g_devs) {
      cudaSetDevice(d.id);
      if (lane != 1) cudaStreamSynchronize(d.stream);
      if (lane != 0) cudaStreamSynchronize(d.side);
    }
  }
};

// contiguous shard [lo, hi) of n items for device slot d of nd
inline void shard(size_t n, int d, int nd, size_t& lo, size_t& hi) {
  lo = n * (size_t)d / nd;
  hi = n * (size_t)(d + 1) / nd;
}

// 32 big-endian bytes == k * r for k in 0..5 (every 256-bit value that is 0 mod r)?
bool is_zero_mod_r(const uint8_t* b) {
  // the top 64 bits of k r lie in [k rt, k rt + k), rt = r >> 192: settles all but ~2^-59 of the values without arithmetic
  uint64_t hi;
  memcpy(&hi, b, 8);
  hi = __builtin_bswap64(hi);
  const uint64_t rt = 0x30644e72e131a029ull;
  bool maybe = false;
  for (uint64_t k = 0; k <= 5; k++) maybe |= hi - k * rt < 6;
  if (!maybe) return false;
  Fr t;
  fe_from_be_bytes(t, b);
  fe_reduce_full(t);
  return fe_is_zero(t);
}
// ---- ChaCha20 (RFC 8439 block function): expands one 32-byte getrandom(2) seed into the per-proof scalars of the
// aggregate Groth16 check (getrandom itself delivers ~0.4 GB/s: 40 ms for 2^20 proofs, on the critical path).
inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
void chacha20_block(uint8_t out[64], const uint8_t key[32], uint32_t counter, const uint8_t nonce[12]) {
  uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  memcpy(st + 4, key, 32);  // little-endian host
  st[12] = counter;
  memcpy(st + 13, nonce, 12);
  uint32_t x[16];
  memcpy(x, st, 64);
#define BN_QR(a, b, c, d)                          \
  x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 16);    \
  x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 12);    \
  x[a] += x[b], x[d] = rotl32(x[d] ^ x[a], 8);     \
  x[c] += x[d], x[b] = rotl32(x[b] ^ x[c], 7);
  for (int r = 0; r < 10; r++) {
    BN_QR(0, 4, 8, 12) BN_QR(1, 5, 9, 13) BN_QR(2, 6, 10, 14) BN_QR(3, 7, 11, 15)
    BN_QR(0, 5, 10, 15) BN_QR(1, 6, 11, 12) BN_QR(2, 7, 8, 13) BN_QR(3, 4, 9, 14)
  }
#undef BN_QR
  for (int i = 0; i < 16; i++) x[i] += st[i];
  memcpy(out, x, 64);
}
// out[0, n) = the ChaCha20 key stream of `key` with nonce (0, 0, block >> 32) and counter = block (low 32 bits);
// `first_block`: where in the stream out[0] lies
void chacha20_stream(uint8_t* out, size_t n, const uint8_t key[32], size_t first_block) {
  uint8_t nonce[12] = {0};
  uint8_t blk[64];
  for (size_t off = 0, b = first_block; off < n; off += 64, b++) {
    const uint32_t hi = (uint32_t)(b >> 32);
    memcpy(nonce + 8, &hi, 4);
    if (n - off >= 64) {
      chacha20_block(out + off, key, (uint32_t)b, nonce);
    } else {
      chacha20_block(blk, key, (uint32_t)b, nonce);
      memcpy(out + off, blk, n - off);
    }
  }
}
// the same stream, long outputs cut into block ranges over a few host threads (0.3 GB/s per thread)
void chacha20_expand(uint8_t* out, size_t n, const uint8_t key[32]) {
  const size_t blocks = (n + 63) / 64;
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t want = std::min<size_t>(std::min<size_t>(8, hw), blocks / 4096);  // >= 256 KB per thread
  if (want <= 1) return chacha20_stream(out, n, key, 0);
  const size_t per = (blocks + want - 1) / want;
  std::vector<std::thread> th;
  for (size_t t = 0; t < want; t++) {
    const size_t b0 = t * per, b1 = std::min(blocks, b0 + per);
    if (b0 >= b1) break;
    const size_t off = b0 * 64, len = std::min(n, b1 * 64) - off;
    th.emplace_back([=] { chacha20_stream(out + off, len, key, b0); });
  }
  for (auto& t : th) t.join();
}

// n fresh 32-byte scalars, none of them 0 mod r (kzg.rs:149-154: Fr::random(OsRng)): one 32-byte seed from the operating
// system's CSPRNG per call, expanded with ChaCha20 (getrandom(2) itself delivers 0.2-0.4 GB/s, which is 20-40 ms per
// 2^18 PlonK proofs in front of a 120 ms chunk)
int draw_rnd(std::vector<uint8_t>& out, size_t n) {
  uint8_t seed[32];
  if (getrandom(seed, sizeof seed, 0) != (ssize_t)sizeof seed) return fail(BN254V_E_BAD_ARG, "getrandom failed");
  out.resize(n * 32);
  chacha20_expand(out.data(), out.size(), seed);
  for (size_t i = 0; i < n; i++)
    while (is_zero_mod_r(out.data() + 32 * i))
      if (getrandom(out.data() + 32 * i, 32, 0) != 32) return fail(BN254V_E_BAD_ARG, "getrandom failed");
  return 0;
}

// ---- aggregate Groth16 check: the batch-wide scalars  s = sum r_i,  t_j = sum r_i x_ij  (mod r)  of one shard, with
// r_i = a_i + b_i lambda from the proof's 16 scalar bytes (groth16_agg.cuh).  Exact integer sums of the a- and b-halves
// (64 x 256-bit products accumulated in 384 bits), one reduction mod r at the end.
typedef unsigned __int128 u128;
struct Acc384 {
  uint64_t w[6] = {0, 0, 0, 0, 0, 0};
  void add_mul(const uint64_t* x, int nx, uint64_t k) {  // += k * x
    uint64_t carry = 0;
    int i = 0;
    for (; i < nx; i++) {
      u128 t = (u128)x[i] * k + w[i] + carry;
      w[i] = (uint64_t)t;
      carry = (uint64_t)(t >> 64);
    }
    for (; i < 6 && carry; i++) {
      u128 t = (u128)w[i] + carry;
      w[i] = (uint64_t)t;
      carry = (uint64_t)(t >> 64);
    }
  }
};
Fr fr_from_acc(const Acc384& a) {  // Montgomery form of the 384-bit integer mod r
  Fr lo, hi = fe_zero<FrCfg>(), two256;
  for (int i = 0; i < 4; i++) lo.v[2 * i] = (uint32_t)a.w[i], lo.v[2 * i + 1] = (uint32_t)(a.w[i] >> 32);
  for (int i = 0; i < 2; i++) hi.v[2 * i] = (uint32_t)a.w[4 + i], hi.v[2 * i + 1] = (uint32_t)(a.w[4 + i] >> 32);
  fe_reduce_full(lo);
  for (int i = 0; i < 8; i++) two256.v[i] = FrCfg::r1(i);  // 2^256 mod r as a plain value
  return fe_add(fe_to_mont(lo), fe_mul(fe_to_mont(hi), fe_to_mont(two256)));
}
void agg_host_sums(uint8_t* scal_be, const uint8_t* rnd16, const uint8_t* inputs_be, int n_inputs, size_t m) {
  std::vector<Acc384> acc(2 * (size_t)(1 + n_inputs));  // [a-half, b-half] of s, t_1, ..
  const uint64_t one = 1;
  for (size_t i = 0; i < m; i++) {
    uint64_t a, b;
    memcpy(&a, rnd16 + 16 * i, 8), memcpy(&b, rnd16 + 16 * i + 8, 8);  // little-endian host
    a |= 1;
    acc[0].add_mul(&one, 1, a);
    acc[1].add_mul(&one, 1, b);
    for (int j = 0; j < n_inputs; j++) {
      const uint8_t* x = inputs_be + ((size_t)i * n_inputs + j) * 32;
      uint64_t xw[4];
      for (int k = 0; k < 4; k++) {
        uint64_t t;
        memcpy(&t, x + 24 - 8 * k, 8);
        xw[k] = __builtin_bswap64(t);
      }
      acc[2 * (j + 1)].add_mul(xw, 4, a);
      acc[2 * (j + 1) + 1].add_mul(xw, 4, b);
    }
  }
  Fr lambda;
  for (int i = 0; i < 8; i++) lambda.v[i] = K::glv_lambda(i);
  for (int j = 0; j <= n_inputs; j++) {
    const Fr v = fe_add(fr_from_acc(acc[2 * j]), fe_mul(lambda, fr_from_acc(acc[2 * j + 1])));
    fe_to_be_bytes(scal_be + 32 * j, fe_from_mont(v));
  }
}

}  // namespace

struct bn254v_vk {
  int kind;  // enum bn254v_kind
  int n_public;
  int n_qcp = 0;  // PlonK: number of BSB22 commitments
  int sign_mode;
  bool cached = false;       // owned by the VK cache
  std::vector<int> dev_ids;  // device of each slot (so that free does not depend on the library state)
  std::vector<void*> dev;    // per device slot: Groth16VkDev* / PlonkVkDev*
  std::vector<void*> aux;    // per device slot: fixed-base tables
};

struct bn254v_batch {
  int kind;  // 0 groth16, 1 plonk, 2 pairing products
  size_t n;
  int n_inputs, k;
  struct Part {
    size_t lo = 0, hi = 0;
    int dev_id = -1;
    uint8_t *proofs = nullptr, *inputs = nullptr, *rnd = nullptr, *status = nullptr;
    Fp12* fbuf = nullptr;
    PlonkWork* work = nullptr;
    int *list = nullptr, *count = nullptr;
  };
  size_t stride;
  std::vector<Part> parts;
};

namespace {

// VK cache: key = kind | sign_mode | sha256(vk bytes)
std::mutex g_cache_mu;
std::map<std::string, bn254v_vk*> g_cache;

std::string vk_key(int kind, int sign_mode, const uint8_t* vk, size_t len) {
  Sha256 s;
  sha256_init(s);
  sha256_update(s, vk, (uint32_t)len);
  uint8_t dg[32];
  sha256_final(s, dg);
  std::string k(2 + 32, '\0');
  k[0] = (char)kind, k[1] = (char)sign_mode;
  memcpy(&k[2], dg, 32);
  return k;
}

void vk_destroy(bn254v_vk* vk) {
  for (size_t i = 0; i < vk->dev_ids.size(); i++) {
    cudaSetDevice(vk->dev_ids[i]);
    if (i < vk->dev.size() && vk->dev[i]) cudaFree(vk->dev[i]);
    if (i < vk->aux.size() && vk->aux[i]) cudaFree(vk->aux[i]);
  }
  delete vk;
}

#ifndef BN_PLONK_CHUNK
#define BN_PLONK_CHUNK ((size_t)1 << 18)  // PlonK proofs per pass of the staged kernels over one device's share
#endif
// ---- one device's share of a batch, on buffers already in device memory --------------------------------------
struct StageEvents {  // optional CUDA events around the stages (device-resident timing runs)
  Dev* dev = nullptr;
};

int run_plonk_chunks(Dev& dev, cudaStream_t st, const bn254v_vk* vk, int slot, const uint8_t* proofs, size_t stride, const uint32_t* lens,
                     const uint8_t* inputs, int n_inputs, const uint8_t* rnd, size_t m, uint8_t* status, uint8_t* g1,
                     uint8_t* fr, uint8_t* ml, uint8_t* gt, PlonkWork* work, int* list, int* count, bool timed) {
  // chunks bound the per-proof workspace (PlonkWork, ~1.9 KB): 2^18 proofs -> 500 MB.  (2^16-proof chunks spent 5 % of
  // their time in the short transcript / sum stages that cannot fill the GPU; they are amortised over a larger chunk.)
  const size_t CH = BN_PLONK_CHUNK;
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (size_t c0 = 0; c0 < m; c0 += CH) {
    const size_t cm = m - c0 < CH ? m - c0 : CH;
    launch::PlonkArgs a;
    a.vk = (const PlonkVkDev*)vk->dev[slot];
    a.n_qcp = vk->n_qcp;
    a.proofs = proofs + c0 * stride, a.stride = stride;
    a.lens = lens ? lens + c0 : nullptr;
    a.inputs = inputs + c0 * in_bytes, a.n_inputs = n_inputs;
    a.rnd = rnd + c0 * 32;
    a.m = cm;
    a.status = status + c0;
    a.dbg_g1 = g1 ? g1 + c0 * 256 : nullptr, a.dbg_fr = fr ? fr + c0 * 256 : nullptr;
    a.dbg_m = ml ? ml + c0 * 384 : nullptr, a.dbg_gt = gt ? gt + c0 * 384 : nullptr;
    a.work = work, a.list = list, a.count = count;
    a.stage_ev = (timed && c0 == 0) ? &dev.ev[1] : nullptr;  // stage split of the first chunk
    g_launches += launch::plonk_verify(st, a, g_sm_count);
    CU(cudaGetLastError());
    if (timed && c0 == 0) CU(cudaEventRecord(dev.ev[5], st));
  }
  return 0;
}

}  // namespace

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
    for (auto& ev : d.ev) CU(cudaEventCreate(&ev));
    CU(cudaStreamCreateWithFlags(&d.side, cudaStreamNonBlocking));
    for (auto& ev : d.side_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : d.lane_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    g_devs.push_back(d);
  }
  g_inited = true;
  return BN254V_SUCCESS;
}

void bn254v_shutdown(void) {
  bn254v_vk_cache_clear();
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& d : g_devs) {
    cudaSetDevice(d.id);
    cudaStreamSynchronize(d.stream);
    DevBuf::pool_release(d.id);
    cudaStreamDestroy(d.stream);
    for (auto& ev : d.ev) cudaEventDestroy(ev);
    cudaStreamSynchronize(d.side);
    cudaStreamDestroy(d.side);
    for (auto& ev : d.side_ev) cudaEventDestroy(ev);
    for (auto& ev : d.lane_ev) cudaEventDestroy(ev);
  }
  PinnedBuf::release_all();
  g_devs.clear();
  g_inited = false;
}

int bn254v_device_count(void) { return (int)g_devs.size(); }
const char* bn254v_last_error(void) { return g_err.c_str(); }
uint64_t bn254v_launch_count(void) { return g_launches.load(); }
void bn254v_chacha20_block(const uint8_t* key32, uint32_t counter, const uint8_t* nonce12, uint8_t* out64) {
  chacha20_block(out64, key32, counter, nonce12);
}
void bn254v_chacha20_expand(const uint8_t* key32, size_t n, uint8_t* out) { chacha20_expand(out, n, key32); }
void bn254v_agg_host_sums(const uint8_t* rnd16, const uint8_t* inputs_be, int n_inputs, size_t m, uint8_t* scal_be) {
  agg_host_sums(scal_be, rnd16, inputs_be, n_inputs, m);
}

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
    case BN254V_PANIC_VK_PARSE: return "PANIC_VK_PARSE";
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
  std::vector<uint8_t> hv_mem(sizeof(Groth16VkDev), 0);
  Groth16VkDev* hv = (Groth16VkDev*)hv_mem.data();
  hv->n_ic = (int)h.k.size();
  hv->alpha = h.alpha;
  hv->beta = sign_mode == 0 ? h.beta2 : neg(h.beta2);
  hv->gamma = sign_mode == 0 ? h.gamma2 : neg(h.gamma2);
  hv->delta = neg(h.delta2);
  for (size_t i = 0; i < h.k.size(); i++) hv->ic[i] = h.k[i];
  bn254v_vk* vk = new bn254v_vk();
  vk->kind = BN254V_KIND_GROTH16;
  vk->n_public = hv->n_ic - 1;
  vk->sign_mode = sign_mode;
  for (auto& d : g_devs) {
    // the slot is registered before anything that can fail, so that vk_destroy frees whatever was allocated
    vk->dev_ids.push_back(d.id);
    vk->dev.push_back(nullptr);
    vk->aux.push_back(nullptr);
    cudaError_t e = cudaSetDevice(d.id);
    if (e == cudaSuccess) e = cudaMalloc(&vk->dev.back(), sizeof(Groth16VkDev));
    const int n_bases = hv->n_ic - 1;
    // window tables of IC_1.. (prepare_inputs), then of IC_0 and alpha (the aggregate check's batch-wide points)
    if (e == cudaSuccess)
      e = cudaMalloc(&vk->aux.back(), sizeof(G1Aff) * (size_t)(n_bases + 2) * BN_IC_WINDOWS * BN_IC_ENTRIES);
    Groth16VkDev* dv = (Groth16VkDev*)vk->dev.back();
    G1Aff* table = (G1Aff*)vk->aux.back();
    hv->ic_table = n_bases > 0 ? table : nullptr;
    hv->agg_table = table + (size_t)n_bases * BN_IC_WINDOWS * BN_IC_ENTRIES;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv, hv, sizeof(Groth16VkDev), cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
      g_launches += launch::groth16_vk_prepare(d.stream, dv, n_bases, table);
      e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(d.stream);  // hv is read by the asynchronous copy: always drain
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) {
      vk_destroy(vk);
      return fail(BN254V_E_CUDA, "vk upload/prepare: %s", cudaGetErrorString(e));
    }
  }
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
  std::vector<uint8_t> hv_mem(sizeof(PlonkVkDev), 0);
  PlonkVkDev* hv = (PlonkVkDev*)hv_mem.data();
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
  vk->kind = BN254V_KIND_PLONK;
  vk->n_public = hv->n_public;
  vk->n_qcp = hv->n_qcp;
  vk->sign_mode = 0;
  const int n_fixed = BN_PLONK_N_FIXED(hv->n_qcp);
  std::vector<G1Aff> bases(n_fixed);
  for (int i = 0; i < n_fixed; i++) bases[i] = plonk_fixed_base(*hv, i);
  for (auto& d : g_devs) {
    vk->dev_ids.push_back(d.id);
    vk->dev.push_back(nullptr);
    vk->aux.push_back(nullptr);
    cudaError_t e = cudaSetDevice(d.id);
    if (e == cudaSuccess) e = cudaMalloc(&vk->dev.back(), sizeof(PlonkVkDev));
    // fixed-base window tables of the VK-constant MSM bases: [bases | tables] in one allocation
    if (e == cudaSuccess)
      e = cudaMalloc(&vk->aux.back(), sizeof(G1Aff) * ((size_t)n_fixed + (size_t)n_fixed * BN_IC_WINDOWS * BN_IC_ENTRIES));
    PlonkVkDev* dv = (PlonkVkDev*)vk->dev.back();
    G1Aff* tab_mem = (G1Aff*)vk->aux.back();
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(tab_mem, bases.data(), sizeof(G1Aff) * n_fixed, cudaMemcpyHostToDevice, d.stream);
    hv->fixed_tables = tab_mem ? tab_mem + n_fixed : nullptr;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv, hv, sizeof(PlonkVkDev), cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
      g_launches += launch::plonk_vk_prepare(d.stream, dv, n_fixed, tab_mem);
      e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(d.stream);
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) {
      vk_destroy(vk);
      return fail(BN254V_E_CUDA, "vk upload/prepare: %s", cudaGetErrorString(e));
    }
  }
  *out = vk;
  return BN254V_SUCCESS;
}

void bn254v_vk_free(bn254v_vk* vk) {
  if (!vk || vk->cached) return;  // cached handles belong to the library
  vk_destroy(vk);
}

int bn254v_vk_n_public(const bn254v_vk* vk) { return vk ? vk->n_public : -1; }

int bn254v_vk_cache_get(int kind, const uint8_t* vk_bytes, size_t len, int sign_mode, const bn254v_vk** out) {
  if (!vk_bytes || !out || (kind != BN254V_KIND_GROTH16 && kind != BN254V_KIND_PLONK))
    return fail(BN254V_E_BAD_ARG, "bad argument");
  if (kind == BN254V_KIND_PLONK) sign_mode = 0;
  const std::string key = vk_key(kind, sign_mode, vk_bytes, len);
  std::lock_guard<std::mutex> lk(g_cache_mu);
  auto it = g_cache.find(key);
  if (it == g_cache.end()) {
    bn254v_vk* vk = nullptr;
    int rc = kind == BN254V_KIND_GROTH16 ? bn254v_groth16_vk_load(vk_bytes, len, sign_mode, &vk)
                                         : bn254v_plonk_vk_load(vk_bytes, len, &vk);
    if (rc) return rc;
    vk->cached = true;
    it = g_cache.emplace(key, vk).first;
  }
  *out = it->second;
  return BN254V_SUCCESS;
}
size_t bn254v_vk_cache_size(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  return g_cache.size();
}
void bn254v_vk_cache_clear(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  for (auto& kv : g_cache) vk_destroy(kv.second);
  g_cache.clear();
}

// ---- Groth16 batch -----------------------------------------------------------------------------
// The batch entry points share the per-device streams, events and stage timers: calls from several host threads are
// serialised by g_call_mu in the exported wrappers (end of this block); the *_impl bodies are what
// bn254v_verify_many's helper thread calls while that call holds the lock.
static int groth16_verify_batch_impl(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                     const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs, size_t n,
                                     uint8_t* status, const bn254v_debug* dbg, Lane lane = Lane()) {
  if (!vk || vk->kind != BN254V_KIND_GROTH16 || !status || (n && (!proofs || (n_inputs > 0 && !inputs_be))) ||
      n_inputs < 0 || n_inputs > 64)
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) return BN254V_SUCCESS;
  const int nd = (int)g_devs.size();
  if ((int)vk->dev.size() != nd) return fail(BN254V_E_BAD_ARG, "VK was loaded for another device set");
  struct Part {
    DevBuf proofs, lens, inputs, status, l, m, gt, fbuf;
  };
  std::vector<Part> parts(nd);
  SyncGuard guard(lane.id);
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (int d = 0; d < nd; d++) {
    size_t lo, hi;
    shard(n, d, nd, lo, hi);
    size_t m = hi - lo;
    if (!m) continue;
    Part& p = parts[d];
    Dev& dev = g_devs[d];
    const cudaStream_t st = lane.stream(dev);
    CU(cudaSetDevice(dev.id));
    CU(p.proofs.alloc(m * proof_stride));
    CU(p.inputs.alloc(m * in_bytes));
    CU(p.status.alloc(m));
    CU(cudaMemcpyAsync(p.proofs.p, proofs + lo * proof_stride, m * proof_stride, cudaMemcpyHostToDevice, st));
    if (in_bytes)
      CU(cudaMemcpyAsync(p.inputs.p, inputs_be + lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, st));
    if (proof_len) {
      CU(p.lens.alloc(m * 4));
      CU(cudaMemcpyAsync(p.lens.p, proof_len + lo, m * 4, cudaMemcpyHostToDevice, st));
    }
    if (dbg && dbg->g1_out) CU(p.l.alloc(m * 64));
    if (dbg && dbg->miller_out) CU(p.m.alloc(m * 384));
    if (dbg && dbg->gt_out) CU(p.gt.alloc(m * 384));
    CU(p.fbuf.alloc(m * sizeof(Fp12)));
    launch::Groth16Args a{(const Groth16VkDev*)vk->dev[d], p.proofs.as<uint8_t>(), proof_stride,
                          proof_len ? p.lens.as<uint32_t>() : nullptr, p.inputs.as<uint8_t>(), n_inputs, m,
                          p.status.as<uint8_t>(), p.l.as<uint8_t>(), p.m.as<uint8_t>(), p.gt.as<uint8_t>(),
                          p.fbuf.as<Fp12>(), nullptr};
    if (lane.chained) CU(cudaStreamWaitEvent(st, dev.lane_ev[lane.id ^ 1], 0));
    g_launches += launch::groth16_verify(st, a, g_sm_count, nullptr);
    CU(cudaGetLastError());
    if (lane.chained) CU(cudaEventRecord(dev.lane_ev[lane.id], st));
    CU(cudaMemcpyAsync(status + lo, p.status.p, m, cudaMemcpyDeviceToHost, st));
    if (p.l.p) CU(cudaMemcpyAsync(dbg->g1_out + lo * 64, p.l.p, m * 64, cudaMemcpyDeviceToHost, st));
    if (p.m.p) CU(cudaMemcpyAsync(dbg->miller_out + lo * 384, p.m.p, m * 384, cudaMemcpyDeviceToHost, st));
    if (p.gt.p) CU(cudaMemcpyAsync(dbg->gt_out + lo * 384, p.gt.p, m * 384, cudaMemcpyDeviceToHost, st));
  }
  for (int d = 0; d < nd; d++) {
    CU(cudaSetDevice(g_devs[d].id));
    CU(cudaStreamSynchronize(lane.stream(g_devs[d])));
  }
  return BN254V_SUCCESS;
}

// ---- opt-in aggregate Groth16 check -------------------------------------------------------------
static int groth16_batch_all_valid_impl(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                        const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                                        const uint8_t* rnd16, size_t n, uint8_t* all_valid, uint8_t* status) {
  if (!vk || vk->kind != BN254V_KIND_GROTH16 || !all_valid || (n && (!proofs || (n_inputs > 0 && !inputs_be))) ||
      n_inputs < 0 || n_inputs > 64)
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  *all_valid = 1;  // the empty batch
  if (n == 0) return BN254V_SUCCESS;
  *all_valid = 0;
  // production path: the library draws the scalars itself, after the proofs are fixed -- one 32-byte seed from the
  // operating system's CSPRNG, expanded with ChaCha20 on a helper thread while the proofs travel to the devices
  std::vector<uint8_t> drawn;
  std::thread drawer;
  struct Joiner {
    std::thread& t;
    ~Joiner() {
      if (t.joinable()) t.join();
    }
  } joiner{drawer};
  if (!rnd16) {
    uint8_t seed[32];
    if (getrandom(seed, sizeof seed, 0) != (ssize_t)sizeof seed) return fail(BN254V_E_BAD_ARG, "getrandom failed");
    drawn.resize(n * 16);
    uint8_t* dst = drawn.data();
    const size_t bytes = drawn.size();
    std::vector<uint8_t> key(seed, seed + 32);
    drawer = std::thread([dst, bytes, key] { chacha20_expand(dst, bytes, key.data()); });
    rnd16 = drawn.data();
  }
  std::vector<uint8_t> own_status;
  if (!status) {
    own_status.resize(n);
    status = own_status.data();
  }
  const int nd = (int)g_devs.size();
  if ((int)vk->dev.size() != nd) return fail(BN254V_E_BAD_ARG, "VK was loaded for another device set");
  struct Part {
    DevBuf proofs, lens, inputs, rnd, status, fbuf, gbuf, scal, scratch, verdict;
    std::vector<uint8_t> scal_host;
    uint8_t verdict_host = 0;
    launch::Groth16AggArgs a;
    Fp12* product = nullptr;
    size_t lo = 0, m = 0;
  };
  std::vector<Part> parts(nd);
  SyncGuard guard;
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (int d = 0; d < nd; d++) {  // buffers and the proofs' way to the devices
    size_t lo, hi;
    shard(n, d, nd, lo, hi);
    size_t m = hi - lo;
    if (!m) continue;
    Part& p = parts[d];
    p.lo = lo, p.m = m;
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    const size_t slots = launch::groth16_agg_slots(m);
    CU(p.proofs.alloc(m * proof_stride));
    CU(p.inputs.alloc(m * in_bytes));
    CU(p.rnd.alloc(m * 16));
    CU(p.status.alloc(m));
    CU(p.fbuf.alloc(slots * sizeof(Fp12)));
    CU(p.gbuf.alloc(slots * sizeof(G1Jac)));
    CU(p.scal.alloc((size_t)32 * (1 + n_inputs)));
    CU(p.scratch.alloc(launch::groth16_agg_scratch_bytes()));
    CU(p.verdict.alloc(1));
    CU(cudaMemcpyAsync(p.proofs.p, proofs + lo * proof_stride, m * proof_stride, cudaMemcpyHostToDevice, dev.stream));
    if (in_bytes)
      CU(cudaMemcpyAsync(p.inputs.p, inputs_be + lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, dev.stream));
    if (proof_len) {
      CU(p.lens.alloc(m * 4));
      CU(cudaMemcpyAsync(p.lens.p, proof_len + lo, m * 4, cudaMemcpyHostToDevice, dev.stream));
    }
  }
  if (drawer.joinable()) drawer.join();  // the scalars are complete
  for (int d = 0; d < nd; d++) {
    Part& p = parts[d];
    if (!p.m) continue;
    const size_t lo = p.lo, m = p.m;
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(cudaMemcpyAsync(p.rnd.p, rnd16 + lo * 16, m * 16, cudaMemcpyHostToDevice, dev.stream));
    p.a = launch::Groth16AggArgs{(const Groth16VkDev*)vk->dev[d], p.proofs.as<uint8_t>(), proof_stride,
                                 proof_len ? p.lens.as<uint32_t>() : nullptr, p.inputs.as<uint8_t>(), n_inputs,
                                 p.rnd.as<uint8_t>(), m, p.status.as<uint8_t>(), p.fbuf.as<Fp12>(), p.gbuf.as<G1Jac>(),
                                 p.scal.as<uint8_t>(), p.scratch.p, p.verdict.as<uint8_t>()};
    if (d == 0) CU(cudaEventRecord(dev.ev[0], dev.stream));
    g_launches += launch::groth16_agg_prepare(dev.stream, p.a);
    CU(cudaEventRecord(dev.side_ev[0], dev.stream));
    g_launches += launch::groth16_agg_miller(dev.stream, p.a, g_sm_count, &p.product);
    CU(cudaGetLastError());
    if (d == 0) CU(cudaEventRecord(dev.ev[1], dev.stream));
  }
  const bool shape_ok = n_inputs == vk->n_public;  // otherwise every record carries ERR_PREPARE_INPUTS or an earlier failure
  for (int d = 0; d < nd; d++) {  // the scalar sums (host, while the devices run), then the batch's own pairing
    Part& p = parts[d];
    if (!p.m) continue;
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    if (shape_ok) {
      p.scal_host.resize((size_t)32 * (1 + n_inputs));
      agg_host_sums(p.scal_host.data(), rnd16 + p.lo * 16, inputs_be ? inputs_be + p.lo * in_bytes : nullptr, n_inputs,
                    p.m);
      // side stream: needs the [r_i] C_i of this device (side_ev[0]) and the sums; runs underneath the Miller kernel
      CU(cudaStreamWaitEvent(dev.side, dev.side_ev[0], 0));
      CU(cudaMemcpyAsync(p.scal.p, p.scal_host.data(), p.scal_host.size(), cudaMemcpyHostToDevice, dev.side));
      g_launches += launch::groth16_agg_side(dev.side, p.a);
      CU(cudaEventRecord(dev.side_ev[1], dev.side));
      CU(cudaStreamWaitEvent(dev.stream, dev.side_ev[1], 0));
      g_launches += launch::groth16_agg_final(dev.stream, p.a, p.product);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(&p.verdict_host, p.verdict.p, 1, cudaMemcpyDeviceToHost, dev.stream));
    }
    if (d == 0) CU(cudaEventRecord(dev.ev[2], dev.stream));
    CU(cudaMemcpyAsync(status + p.lo, p.status.p, p.m, cudaMemcpyDeviceToHost, dev.stream));
  }
  for (int d = 0; d < nd; d++) {
    CU(cudaSetDevice(g_devs[d].id));
    CU(cudaStreamSynchronize(g_devs[d].stream));
    CU(cudaStreamSynchronize(g_devs[d].side));
  }
  if (parts[0].m) {
    g_stage_n = 2;
    CU(cudaEventElapsedTime(&g_stage_ms[0], g_devs[0].ev[0], g_devs[0].ev[1]));
    CU(cudaEventElapsedTime(&g_stage_ms[1], g_devs[0].ev[1], g_devs[0].ev[2]));
  }
  bool all = true;
  for (int d = 0; d < nd; d++)
    if (parts[d].m && parts[d].verdict_host != 1) all = false;
  for (size_t i = 0; i < n && all; i++)
    if (status[i] != BN254V_OK_TRUE) all = false;
  *all_valid = all ? 1 : 0;
  return BN254V_SUCCESS;
}

static int plonk_verify_batch_impl(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                   const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                                   const uint8_t* rnd_be, size_t n, uint8_t* status, const bn254v_debug* dbg,
                                   Lane lane = Lane()) {
  if (!vk || vk->kind != BN254V_KIND_PLONK || !status || (n && (!proofs || (n_inputs > 0 && !inputs_be))) ||
      n_inputs < 0 || n_inputs > BN_MAX_PLONK_PUBLIC)
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) return BN254V_SUCCESS;
  std::vector<uint8_t> drawn;  // production path: the library draws the batch-opening scalars itself
  if (!rnd_be) {
    rc = draw_rnd(drawn, n);
    if (rc) return rc;
    rnd_be = drawn.data();
  }
  const int nd = (int)g_devs.size();
  if ((int)vk->dev.size() != nd) return fail(BN254V_E_BAD_ARG, "VK was loaded for another device set");
  struct Part {
    DevBuf proofs, lens, inputs, rnd, status, g1, fr, m, gt, work, list, count;
  };
  std::vector<Part> parts(nd);
  SyncGuard guard(lane.id);
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (int d = 0; d < nd; d++) {
    size_t lo, hi;
    shard(n, d, nd, lo, hi);
    size_t m = hi - lo;
    if (!m) continue;
    Part& p = parts[d];
    Dev& dev = g_devs[d];
    const cudaStream_t st = lane.stream(dev);
    CU(cudaSetDevice(dev.id));
    CU(p.proofs.alloc(m * proof_stride));
    CU(p.inputs.alloc(m * in_bytes));
    CU(p.rnd.alloc(m * 32));
    CU(p.status.alloc(m));
    CU(cudaMemcpyAsync(p.proofs.p, proofs + lo * proof_stride, m * proof_stride, cudaMemcpyHostToDevice, st));
    if (in_bytes)
      CU(cudaMemcpyAsync(p.inputs.p, inputs_be + lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(p.rnd.p, rnd_be + lo * 32, m * 32, cudaMemcpyHostToDevice, st));
    if (proof_len) {
      CU(p.lens.alloc(m * 4));
      CU(cudaMemcpyAsync(p.lens.p, proof_len + lo, m * 4, cudaMemcpyHostToDevice, st));
    }
    if (dbg && dbg->g1_out) CU(p.g1.alloc(m * 256));
    if (dbg && dbg->fr_out) CU(p.fr.alloc(m * 256));
    if (dbg && dbg->miller_out) CU(p.m.alloc(m * 384));
    if (dbg && dbg->gt_out) CU(p.gt.alloc(m * 384));
    if (p.g1.p) CU(cudaMemsetAsync(p.g1.p, 0, m * 256, st));
    if (p.fr.p) CU(cudaMemsetAsync(p.fr.p, 0, m * 256, st));
    if (p.m.p) CU(cudaMemsetAsync(p.m.p, 0, m * 384, st));
    if (p.gt.p) CU(cudaMemsetAsync(p.gt.p, 0, m * 384, st));
    const size_t mc = m < BN_PLONK_CHUNK ? m : BN_PLONK_CHUNK;
    CU(p.work.alloc(mc * sizeof(PlonkWork)));
    CU(p.list.alloc(mc * sizeof(int)));
    CU(p.count.alloc(sizeof(int)));
    if (lane.chained) CU(cudaStreamWaitEvent(st, dev.lane_ev[lane.id ^ 1], 0));
    rc = run_plonk_chunks(dev, st, vk, d, p.proofs.as<uint8_t>(), proof_stride, proof_len ? p.lens.as<uint32_t>() : nullptr,
                          p.inputs.as<uint8_t>(), n_inputs, p.rnd.as<uint8_t>(), m, p.status.as<uint8_t>(),
                          p.g1.as<uint8_t>(), p.fr.as<uint8_t>(), p.m.as<uint8_t>(), p.gt.as<uint8_t>(),
                          p.work.as<PlonkWork>(), p.list.as<int>(), p.count.as<int>(), false);
    if (rc) return rc;
    if (lane.chained) CU(cudaEventRecord(dev.lane_ev[lane.id], st));
    CU(cudaMemcpyAsync(status + lo, p.status.p, m, cudaMemcpyDeviceToHost, st));
    if (p.g1.p) CU(cudaMemcpyAsync(dbg->g1_out + lo * 256, p.g1.p, m * 256, cudaMemcpyDeviceToHost, st));
    if (p.fr.p) CU(cudaMemcpyAsync(dbg->fr_out + lo * 256, p.fr.p, m * 256, cudaMemcpyDeviceToHost, st));
    if (p.m.p) CU(cudaMemcpyAsync(dbg->miller_out + lo * 384, p.m.p, m * 384, cudaMemcpyDeviceToHost, st));
    if (p.gt.p) CU(cudaMemcpyAsync(dbg->gt_out + lo * 384, p.gt.p, m * 384, cudaMemcpyDeviceToHost, st));
  }
  for (int d = 0; d < nd; d++) {
    CU(cudaSetDevice(g_devs[d].id));
    CU(cudaStreamSynchronize(lane.stream(g_devs[d])));
  }
  return BN254V_SUCCESS;
}

// ---- raw pairing products ----------------------------------------------------------------------
static int pairing_product_batch_impl(const uint8_t* g1, const uint8_t* g2, int k, size_t n, uint8_t* is_one,
                                      uint8_t* miller_out, uint8_t* gt_out) {
  if (k < 1 || k > 4 || !is_one || (n && (!g1 || !g2))) return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) return BN254V_SUCCESS;
  const int nd = (int)g_devs.size();
  struct Part {
    DevBuf g1, g2, one, m, gt, fbuf;
  };
  std::vector<Part> parts(nd);
  SyncGuard guard;
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
    CU(p.fbuf.alloc(m * sizeof(Fp12)));
    g_launches += launch::pairing_product(dev.stream, k, p.g1.as<uint8_t>(), p.g2.as<uint8_t>(), m, p.one.as<uint8_t>(),
                                          p.m.as<uint8_t>(), p.gt.as<uint8_t>(), g_sm_count, p.fbuf.as<Fp12>(), nullptr);
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

// ---- mixed batches -----------------------------------------------------------------------------
// Items are grouped by (kind, VK, n_inputs); each group is packed into one contiguous host batch (ragged proofs keep
// their own length) and goes through the batch entry points above; statuses are scattered back in item order.
static int verify_many_impl(const bn254v_item* items, size_t n, int sign_mode, const uint8_t* rnd_be, uint8_t* status) {
  if ((n && !items) || !status || (sign_mode != 0 && sign_mode != 1)) return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  struct Group {
    int kind, n_inputs;
    const bn254v_vk* vk;  // null: the VK does not parse
    std::vector<size_t> pos;
    size_t stride = 1;
    bool ragged = false;
  };
  std::vector<Group> groups;
  std::map<std::string, int> by_key;                                     // (kind, n_inputs, vk hash) -> group
  std::map<std::pair<const uint8_t*, size_t>, std::string> key_of_ptr;    // identical VK pointers are hashed once
  struct Recent {
    const uint8_t* vk;
    size_t vk_len;
    int kind, n_inputs, group;
  };
  Recent recent[8];
  int n_recent = 0;
  for (size_t i = 0; i < n; i++) {
    const bn254v_item& it = items[i];
    if ((it.kind != BN254V_KIND_GROTH16 && it.kind != BN254V_KIND_PLONK) || !it.vk || (it.proof_len && !it.proof) ||
        it.n_inputs < 0 || (it.n_inputs && !it.inputs_be))
      return fail(BN254V_E_BAD_ARG, "item %zu: bad argument", i);
    // fast path: the (VK pointer, kind, n_inputs) combinations seen most recently (a few per call in practice)
    int hit = -1;
    for (int q = 0; q < n_recent; q++)
      if (recent[q].vk == it.vk && recent[q].vk_len == it.vk_len && recent[q].kind == it.kind &&
          recent[q].n_inputs == it.n_inputs) {
        hit = recent[q].group;
        break;
      }
    if (hit >= 0) {
      Group& gr = groups[hit];
      if (it.proof_len != gr.stride) {
        gr.ragged = true;
        if (it.proof_len > gr.stride) gr.stride = it.proof_len;
      }
      gr.pos.push_back(i);
      continue;
    }
    auto pk = std::make_pair(it.vk, it.vk_len);
    auto f = key_of_ptr.find(pk);
    if (f == key_of_ptr.end()) f = key_of_ptr.emplace(pk, vk_key(it.kind, it.kind ? 0 : sign_mode, it.vk, it.vk_len)).first;
    std::string key = f->second;
    key[0] = (char)it.kind;  // the same bytes offered as both kinds are two keys
    key += std::string(1, (char)it.n_inputs);
    auto g = by_key.find(key);
    if (g == by_key.end()) {
      Group ng;
      ng.kind = it.kind, ng.n_inputs = it.n_inputs;
      const bn254v_vk* h = nullptr;
      rc = bn254v_vk_cache_get(it.kind, it.vk, it.vk_len, sign_mode, &h);
      if (rc == BN254V_E_VK_PARSE || rc == BN254V_E_UNSUPPORTED) h = nullptr;
      else if (rc) return rc;
      ng.vk = h;
      groups.push_back(ng);
      g = by_key.emplace(key, (int)groups.size() - 1).first;
    }
    Group& gr = groups[g->second];
    if (!gr.pos.empty() && it.proof_len != gr.stride) gr.ragged = true;
    if (it.proof_len > gr.stride) gr.stride = it.proof_len;
    gr.pos.push_back(i);
    recent[n_recent < 8 ? n_recent++ : (i & 7)] = Recent{it.vk, it.vk_len, it.kind, it.n_inputs, g->second};
  }
  // A group is cut into chunks; while the devices verify chunk k (a helper thread inside the synchronous batch entry
  // point), the host cores gather chunk k + 1 into the other pinned staging set: the gather of a large mixed batch (GBs)
  // is hidden behind the kernels.
  const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  // Chunk sizes: 2^20 Groth16 proofs (the multi-wave kernels gain efficiency up to there), 2^18 PlonK proofs (one
  // workspace chunk); the very first chunk of the call is small so that the devices start early -- its gather is the
  // only one that nothing hides.
  struct Task {
    Group* gr;
    size_t first, m;
  };
  std::vector<Task> tasks;
  for (Group& gr : groups) {
    if (!gr.vk) {
      for (size_t j : gr.pos) status[j] = BN254V_PANIC_VK_PARSE;
      continue;
    }
    const size_t CH = (size_t)1 << (gr.kind == BN254V_KIND_GROTH16 ? 20 : 18);
    for (size_t f = 0; f < gr.pos.size();) {
      const size_t m = std::min(tasks.empty() ? (size_t)1 << 17 : CH, gr.pos.size() - f);
      tasks.push_back(Task{&gr, f, m});
      f += m;
    }
  }
  struct Stage {
    PinnedBuf proofs, inputs, rnd, st, lens;
    Task t{nullptr, 0, 0};
    bool has_rnd = false;
    int rc = 0;
    std::string err;
  };
  Stage stages[2];
  auto pack = [&](Stage& sg, const Task& t) -> int {
    Group& gr = *t.gr;
    const size_t in_bytes = (size_t)32 * gr.n_inputs;
    sg.t = t;
    sg.has_rnd = gr.kind == BN254V_KIND_PLONK && rnd_be;
    CU(sg.proofs.reserve(t.m * gr.stride));
    CU(sg.inputs.reserve(t.m * in_bytes + 1));
    CU(sg.st.reserve(t.m));
    if (gr.ragged) CU(sg.lens.reserve(t.m * 4));
    if (sg.has_rnd) CU(sg.rnd.reserve(t.m * 32));
    uint8_t *pp = sg.proofs.as<uint8_t>(), *pi = sg.inputs.as<uint8_t>(), *pr = sg.rnd.as<uint8_t>();
    uint32_t* pl = sg.lens.as<uint32_t>();
    auto work = [&](size_t a, size_t b) {
      for (size_t j = a; j < b; j++) {
        const size_t pos = gr.pos[t.first + j];
        const bn254v_item& it = items[pos];
        if (it.proof_len) memcpy(pp + j * gr.stride, it.proof, it.proof_len);
        if (it.proof_len < gr.stride) memset(pp + j * gr.stride + it.proof_len, 0, gr.stride - it.proof_len);
        if (in_bytes) memcpy(pi + j * in_bytes, it.inputs_be, in_bytes);
        if (gr.ragged) pl[j] = (uint32_t)it.proof_len;
        if (sg.has_rnd) memcpy(pr + j * 32, rnd_be + 32 * pos, 32);
      }
    };
    if (t.m < 4096 || hw == 1) {
      work(0, t.m);
    } else {
      std::vector<std::thread> th;
      for (unsigned k = 0; k < hw; k++) th.emplace_back(work, t.m * k / hw, t.m * (k + 1) / hw);
      for (auto& x : th) x.join();
    }
    return 0;
  };
  auto run = [&](Stage* sg, int lane_id) {
    Group& gr = *sg->t.gr;
    const Lane lane{lane_id, true};
    if (gr.kind == BN254V_KIND_GROTH16)
      sg->rc = groth16_verify_batch_impl(gr.vk, sg->proofs.as<uint8_t>(), gr.stride,
                                           gr.ragged ? sg->lens.as<uint32_t>() : nullptr, sg->inputs.as<uint8_t>(),
                                           gr.n_inputs, sg->t.m, sg->st.as<uint8_t>(), nullptr, lane);
    else
      sg->rc = plonk_verify_batch_impl(gr.vk, sg->proofs.as<uint8_t>(), gr.stride,
                                         gr.ragged ? sg->lens.as<uint32_t>() : nullptr, sg->inputs.as<uint8_t>(),
                                         gr.n_inputs, sg->has_rnd ? sg->rnd.as<uint8_t>() : nullptr, sg->t.m,
                                         sg->st.as<uint8_t>(), nullptr, lane);
    if (sg->rc) sg->err = g_err;  // (thread-local in the helper thread)
  };
  auto collect = [&](Stage& sg) -> int {
    if (sg.rc) return fail(sg.rc, "%s", sg.err.c_str());
    const uint8_t* st = sg.st.as<uint8_t>();
    for (size_t j = 0; j < sg.t.m; j++) status[sg.t.gr->pos[sg.t.first + j]] = st[j];
    return 0;
  };
  // Two chunks in flight, one per staging set and lane (see Lane): chunk k is gathered while chunk k - 1 runs and
  // chunk k - 2 has been collected; its copies to the device go out as soon as it is gathered, its kernels queue up
  // behind those of chunk k - 1.
  std::thread workers[2];
  int result = 0;
  for (size_t k = 0; k < tasks.size() && !result; k++) {
    Stage& cur = stages[k & 1];
    if (workers[k & 1].joinable()) {  // chunk k - 2: its staging set is the one to fill now
      workers[k & 1].join();
      result = collect(cur);
      if (result) break;
    }
    result = pack(cur, tasks[k]);
    if (!result) workers[k & 1] = std::thread(run, &cur, (int)(k & 1));
  }
  // the one or two chunks still running, older first
  const size_t done = tasks.size();
  for (size_t j = done >= 2 ? done - 2 : 0; j < done; j++) {
    if (!workers[j & 1].joinable()) continue;
    workers[j & 1].join();
    const int r2 = collect(stages[j & 1]);
    if (!result) result = r2;
  }
  return result;
}

// ---- the exported batch entry points: one call at a time -------------------------------------------------------
static std::mutex g_call_mu;
int bn254v_groth16_verify_batch(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs, size_t n,
                                uint8_t* status, const bn254v_debug* dbg) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  return groth16_verify_batch_impl(vk, proofs, proof_stride, proof_len, inputs_be, n_inputs, n, status, dbg);
}
int bn254v_groth16_batch_all_valid(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                   const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                                   const uint8_t* rnd16, size_t n, uint8_t* all_valid, uint8_t* status) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  return groth16_batch_all_valid_impl(vk, proofs, proof_stride, proof_len, inputs_be, n_inputs, rnd16, n, all_valid, status);
}
int bn254v_plonk_verify_batch(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                              const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                              const uint8_t* rnd_be, size_t n, uint8_t* status, const bn254v_debug* dbg) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  return plonk_verify_batch_impl(vk, proofs, proof_stride, proof_len, inputs_be, n_inputs, rnd_be, n, status, dbg);
}
int bn254v_pairing_product_batch(const uint8_t* g1, const uint8_t* g2, int k, size_t n, uint8_t* is_one,
                                 uint8_t* miller_out, uint8_t* gt_out) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  return pairing_product_batch_impl(g1, g2, k, n, is_one, miller_out, gt_out);
}
int bn254v_verify_many(const bn254v_item* items, size_t n, int sign_mode, const uint8_t* rnd_be, uint8_t* status) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  return verify_many_impl(items, n, sign_mode, rnd_be, status);
}

// ---- device-resident batches (bn254v_bench.h) ----------------------------------------------------
static int batch_alloc_common(bn254v_batch* b, size_t n) {
  const int nd = (int)g_devs.size();
  b->n = n;
  b->parts.resize(nd);
  for (int d = 0; d < nd; d++) {
    shard(n, d, nd, b->parts[d].lo, b->parts[d].hi);
    b->parts[d].dev_id = g_devs[d].id;
  }
  return 0;
}
#define BCU(call)                                                                    \
  do {                                                                               \
    cudaError_t _e = (call);                                                         \
    if (_e != cudaSuccess) {                                                         \
      cudaDeviceSynchronize();                                                       \
      bn254v_batch_free(b);                                                          \
      return fail(BN254V_E_CUDA, "batch upload: %s", cudaGetErrorString(_e));        \
    }                                                                                \
  } while (0)

int bn254v_groth16_batch_upload(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                const uint8_t* inputs_be, int n_inputs, size_t n, bn254v_batch** out) {
  if (!vk || vk->kind != BN254V_KIND_GROTH16 || !proofs || !out || proof_stride < 256 || n_inputs < 0 ||
      (n_inputs && !inputs_be))
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  bn254v_batch* b = new bn254v_batch();
  b->kind = 0, b->n_inputs = n_inputs, b->k = 0, b->stride = 256;
  batch_alloc_common(b, n);
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (size_t d = 0; d < b->parts.size(); d++) {
    auto& p = b->parts[d];
    size_t m = p.hi - p.lo;
    if (!m) continue;
    Dev& dev = g_devs[d];
    BCU(cudaSetDevice(dev.id));
    BCU(cudaMalloc(&p.proofs, m * 256));
    BCU(cudaMalloc(&p.inputs, m * in_bytes + 1));
    BCU(cudaMalloc(&p.status, m));
    BCU(cudaMalloc(&p.fbuf, m * sizeof(Fp12)));
    BCU(cudaMemcpy2DAsync(p.proofs, 256, proofs + p.lo * proof_stride, proof_stride, 256, m, cudaMemcpyHostToDevice,
                          dev.stream));
    if (in_bytes)
      BCU(cudaMemcpyAsync(p.inputs, inputs_be + p.lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, dev.stream));
    BCU(cudaStreamSynchronize(dev.stream));
  }
  *out = b;
  return BN254V_SUCCESS;
}

int bn254v_plonk_batch_upload(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                              const uint8_t* inputs_be, int n_inputs, const uint8_t* rnd_be, size_t n,
                              bn254v_batch** out) {
  if (!vk || vk->kind != BN254V_KIND_PLONK || !proofs || !out || !rnd_be || n_inputs < 0 || (n_inputs && !inputs_be))
    return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  bn254v_batch* b = new bn254v_batch();
  b->kind = 1, b->n_inputs = n_inputs, b->k = 0, b->stride = proof_stride;
  batch_alloc_common(b, n);
  const size_t in_bytes = (size_t)32 * n_inputs;
  for (size_t d = 0; d < b->parts.size(); d++) {
    auto& p = b->parts[d];
    size_t m = p.hi - p.lo;
    if (!m) continue;
    Dev& dev = g_devs[d];
    const size_t mc = m < BN_PLONK_CHUNK ? m : BN_PLONK_CHUNK;
    BCU(cudaSetDevice(dev.id));
    BCU(cudaMalloc(&p.proofs, m * proof_stride));
    BCU(cudaMalloc(&p.inputs, m * in_bytes + 1));
    BCU(cudaMalloc(&p.rnd, m * 32));
    BCU(cudaMalloc(&p.status, m));
    BCU(cudaMalloc(&p.work, mc * sizeof(PlonkWork)));
    BCU(cudaMalloc(&p.list, mc * sizeof(int)));
    BCU(cudaMalloc(&p.count, sizeof(int)));
    BCU(cudaMemcpyAsync(p.proofs, proofs + p.lo * proof_stride, m * proof_stride, cudaMemcpyHostToDevice, dev.stream));
    if (in_bytes)
      BCU(cudaMemcpyAsync(p.inputs, inputs_be + p.lo * in_bytes, m * in_bytes, cudaMemcpyHostToDevice, dev.stream));
    BCU(cudaMemcpyAsync(p.rnd, rnd_be + p.lo * 32, m * 32, cudaMemcpyHostToDevice, dev.stream));
    BCU(cudaStreamSynchronize(dev.stream));
  }
  *out = b;
  return BN254V_SUCCESS;
}

int bn254v_pairing_batch_upload(const uint8_t* g1, const uint8_t* g2, int k, size_t n, bn254v_batch** out) {
  if (k < 1 || k > 4 || !g1 || !g2 || !out) return fail(BN254V_E_BAD_ARG, "bad argument");
  int rc = ensure_init();
  if (rc) return rc;
  bn254v_batch* b = new bn254v_batch();
  b->kind = 2, b->n_inputs = 0, b->k = k, b->stride = 0;
  batch_alloc_common(b, n);
  for (size_t d = 0; d < b->parts.size(); d++) {
    auto& p = b->parts[d];
    size_t m = p.hi - p.lo;
    if (!m) continue;
    Dev& dev = g_devs[d];
    BCU(cudaSetDevice(dev.id));
    BCU(cudaMalloc(&p.proofs, m * 64 * k));   // G1 points
    BCU(cudaMalloc(&p.inputs, m * 128 * k));  // G2 points
    BCU(cudaMalloc(&p.status, m));
    BCU(cudaMalloc(&p.fbuf, m * sizeof(Fp12)));
    BCU(cudaMemcpyAsync(p.proofs, g1 + p.lo * 64 * k, m * 64 * k, cudaMemcpyHostToDevice, dev.stream));
    BCU(cudaMemcpyAsync(p.inputs, g2 + p.lo * 128 * k, m * 128 * k, cudaMemcpyHostToDevice, dev.stream));
    BCU(cudaStreamSynchronize(dev.stream));
  }
  *out = b;
  return BN254V_SUCCESS;
}

// runs the staged batch on every device; stage events on device slot 0
static int batch_run(const bn254v_vk* vk, bn254v_batch* b, uint8_t* status, float* kernel_ms) {
  std::lock_guard<std::mutex> call_lock(g_call_mu);
  const int nd = (int)g_devs.size();
  if ((int)b->parts.size() != nd) return fail(BN254V_E_BAD_ARG, "batch was staged for another device set");
  if (vk && (int)vk->dev.size() != nd) return fail(BN254V_E_BAD_ARG, "VK was loaded for another device set");
  SyncGuard guard;
  int n_stage = 1;
  for (int d = 0; d < nd; d++) {
    auto& p = b->parts[d];
    size_t m = p.hi - p.lo;
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(cudaEventRecord(dev.ev[0], dev.stream));
    if (m && b->kind == 0) {
      bool two = false;
      launch::Groth16Args a{(const Groth16VkDev*)vk->dev[d], p.proofs, 256, nullptr, p.inputs, b->n_inputs, m, p.status,
                            nullptr, nullptr, nullptr, p.fbuf, dev.ev[1], dev.ev[3]};
      CU(cudaEventRecord(dev.ev[3], dev.stream));  // (stays at the start when there is no separate prepare launch)
      g_launches += launch::groth16_verify(dev.stream, a, g_sm_count, &two);
      if (!two) CU(cudaEventRecord(dev.ev[1], dev.stream));
      CU(cudaEventRecord(dev.ev[2], dev.stream));
      if (d == 0) n_stage = 2;
    } else if (m && b->kind == 1) {
      int rc = run_plonk_chunks(dev, dev.stream, vk, d, p.proofs, b->stride, nullptr, p.inputs, b->n_inputs, p.rnd, m, p.status,
                                nullptr, nullptr, nullptr, nullptr, p.work, p.list, p.count, true);
      if (rc) return rc;
      if (d == 0) n_stage = 5;
    } else if (m) {
      CU(cudaEventRecord(dev.ev[1], dev.stream));  // (stays at the start of a fused launch)
      const int nl = launch::pairing_product(dev.stream, b->k, p.proofs, p.inputs, m, p.status, nullptr, nullptr, g_sm_count,
                                             p.fbuf, dev.ev[1]);
      g_launches += nl;
      CU(cudaEventRecord(dev.ev[2], dev.stream));
      if (d == 0) n_stage = 2;  // [0] Miller loops (or nothing), [1] final exponentiations (or the fused launch)
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(dev.ev[7], dev.stream));
  }
  float worst = 0.f;
  for (int d = 0; d < nd; d++) {
    Dev& dev = g_devs[d];
    CU(cudaSetDevice(dev.id));
    CU(cudaStreamSynchronize(dev.stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, dev.ev[0], dev.ev[7]));
    if (ms > worst) worst = ms;
    if (d == 0 && b->parts[0].hi > b->parts[0].lo) {
      g_stage_n = n_stage;
      for (int s = 0; s < n_stage; s++) CU(cudaEventElapsedTime(&g_stage_ms[s], dev.ev[s], dev.ev[s + 1]));
      if (b->kind == 0) {  // Groth16: [0] prepare + Miller, [1] final exponentiation, [2] the prepare launch alone
        CU(cudaEventElapsedTime(&g_stage_ms[2], dev.ev[0], dev.ev[3]));
        g_stage_n = 3;
      }
      // PlonK batches above 2^16 proofs run in chunks: the split is that of the first chunk, scaled to the whole time
      if (b->kind == 1 && b->parts[0].hi - b->parts[0].lo > BN_PLONK_CHUNK) {
        float sum = 0.f;
        for (int s = 0; s < n_stage; s++) sum += g_stage_ms[s];
        if (sum > 0.f)
          for (int s = 0; s < n_stage; s++) g_stage_ms[s] *= ms / sum;
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

int bn254v_groth16_batch_verify(const bn254v_vk* vk, bn254v_batch* b, uint8_t* status, float* kernel_ms) {
  if (!vk || vk->kind != BN254V_KIND_GROTH16 || !b || b->kind != 0) return fail(BN254V_E_BAD_ARG, "bad argument");
  return batch_run(vk, b, status, kernel_ms);
}
int bn254v_plonk_batch_verify(const bn254v_vk* vk, bn254v_batch* b, uint8_t* status, float* kernel_ms) {
  if (!vk || vk->kind != BN254V_KIND_PLONK || !b || b->kind != 1) return fail(BN254V_E_BAD_ARG, "bad argument");
  return batch_run(vk, b, status, kernel_ms);
}
int bn254v_pairing_batch_verify(bn254v_batch* b, uint8_t* is_one, float* kernel_ms) {
  if (!b || b->kind != 2) return fail(BN254V_E_BAD_ARG, "bad argument");
  return batch_run(nullptr, b, is_one, kernel_ms);
}

void bn254v_batch_free(bn254v_batch* b) {
  if (!b) return;
  for (auto& p : b->parts) {
    if (p.dev_id < 0) continue;
    cudaSetDevice(p.dev_id);
    cudaFree(p.proofs), cudaFree(p.inputs), cudaFree(p.rnd), cudaFree(p.status), cudaFree(p.fbuf);
    cudaFree(p.work), cudaFree(p.list), cudaFree(p.count);
  }
  delete b;
}

int bn254v_last_stage_ms(float* out, int cap) {
  int n = g_stage_n < cap ? g_stage_n : cap;
  for (int i = 0; i < n; i++) out[i] = g_stage_ms[i];
  return n;
}
int bn254v_last_kernel_split(float* miller_ms, float* finish_ms) {
  if (miller_ms) *miller_ms = g_stage_n >= 1 ? g_stage_ms[0] : 0.f;
  if (finish_ms) *finish_ms = g_stage_n >= 2 ? g_stage_ms[1] : 0.f;
  return BN254V_SUCCESS;
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
  SyncGuard guard;
  CU(dp.alloc(n * 256));
  CU(di.alloc(n * 32 * n_public));
  CU(de.alloc(n));
  g_launches += launch::groth16_synth(dev.stream, td, seed, first_index, n, n_public, sign_mode, dp.as<uint8_t>(),
                                      di.as<uint8_t>(), de.as<uint8_t>());
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
  SyncGuard guard;
  CU(d1.alloc(n * 64 * k));
  CU(d2.alloc(n * 128 * k));
  CU(de.alloc(n));
  g_launches += launch::pairing_synth(dev.stream, seed, first_index, n, k, d1.as<uint8_t>(), d2.as<uint8_t>(),
                                      de.as<uint8_t>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(g1, d1.p, n * 64 * k, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaMemcpyAsync(g2, d2.p, n * 128 * k, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaMemcpyAsync(expected_is_one, de.p, n, cudaMemcpyDeviceToHost, dev.stream));
  CU(cudaStreamSynchronize(dev.stream));
  return BN254V_SUCCESS;
}

// ---- measurement helpers -----------------------------------------------------------------------
int bn254v_imad_peak(int iters, double* wide_mac_per_s, double* lo_mac_per_s, float* sm_clock_mhz) {
  if (iters < 1) return fail(BN254V_E_BAD_ARG, "iters < 1");
  int rc = ensure_init();
  if (rc) return rc;
  Dev& dev = g_devs[0];
  CU(cudaSetDevice(dev.id));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev.id));
  DevBuf sink;
  SyncGuard guard;
  CU(sink.alloc(8));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  const double macs = (double)blocks * threads * (double)iters * 32.0;
  float ms = 0.f;
  for (int wide = 1; wide >= 0; wide--) {
    for (int pass = 0; pass < 2; pass++) {  // pass 0 warms up
      CU(cudaEventRecord(dev.ev[0], dev.stream));
      g_launches += launch::imad_peak(dev.stream, wide != 0, blocks, threads, iters, sink.as<uint64_t>());
      CU(cudaEventRecord(dev.ev[1], dev.stream));
      CU(cudaStreamSynchronize(dev.stream));
      CU(cudaEventElapsedTime(&ms, dev.ev[0], dev.ev[1]));
    }
    if (wide && wide_mac_per_s) *wide_mac_per_s = macs / (ms * 1e-3);
    if (!wide && lo_mac_per_s) *lo_mac_per_s = macs / (ms * 1e-3);
  }
  if (sm_clock_mhz) *sm_clock_mhz = prop.clockRate / 1000.f;
  return BN254V_SUCCESS;
}

}  // extern "C"
