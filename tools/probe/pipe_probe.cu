// Which instructions overlap with the multiplier's IMAD.WIDE stream on a B200 SMSP?  Loop bodies are built from the
// production primitives of csrc/field.cuh (so the SASS is the production pattern: carry-chained IMAD.WIDE.U32.X rows,
// IADD3.X chains) and run at 2..16 warps per SMSP; the host prints cycles per loop iteration per SMSP-resident warp.
// If two instruction kinds run on separate pipes, time(A + B) ~ max(time(A), time(B)); if they share an issue
// resource it is the sum.
// Build (no GPU needed):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I snark-bn254-verifier_b200/csrc \
//        -o tools/probe/pipe_probe tools/probe/pipe_probe.cu
#include <stdint.h>
#include <stdio.h>
#include "field.cuh"
using namespace bn254;

enum { M_MUL = 0, M_ADD, M_MUL_ADD, M_SPLIT, M_MUL_LDS, M_MUL_STS, M_LDS, M_MUL_LDL, M_REDC, M_MUL2_ADD, M_N };
static const char* NAMES[M_N] = {
    "A: wide product (64 MAC)", "B: 4 x 16-word add chain (64 IADD3.X)", "A + B in the same warp",
    "warp split: even warps A, odd warps B (time per pair)", "A + 64 LDS", "A + 64 STS", "64 LDS + 64 xor",
    "A + 32 LDG/STG pairs (local-like)", "wide reduction (72 MAC)", "2 A + B in the same warp"};

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, Fp* io, uint32_t* scratch) {
  __shared__ uint32_t sm[16 * 256];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fp a = io[tid], b = io[tid ^ 1];
  uint32_t X[16], Y[16], T[16], U[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    X[i] = a.v[i & 7] + i, Y[i] = b.v[i & 7] * 3 + i, T[i] = i, U[i] = 0;
    sm[i * 256 + threadIdx.x] = a.v[i & 7] ^ i;
  }
  uint32_t* lp = scratch + (size_t)tid * 16;
  const bool odd = ((threadIdx.x >> 5) >> 2) & 1;  // warps w and w + 4 of a block share an SMSP
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (MODE == M_MUL || MODE == M_MUL_ADD || MODE == M_MUL_LDS || MODE == M_MUL_STS || MODE == M_MUL_LDL ||
        MODE == M_MUL2_ADD || (MODE == M_SPLIT && !odd)) {
      fe_mul_wide(T, a, b);
#pragma unroll
      for (int i = 0; i < 8; i++) a.v[i] ^= T[i + 8];  // dependency through the loop (8 LOP3)
    }
    if (MODE == M_MUL2_ADD) {
      fe_mul_wide(U, b, a);
#pragma unroll
      for (int i = 0; i < 8; i++) b.v[i] ^= U[i + 4];
    }
    if (MODE == M_REDC) {
      Fp r = fe_redc_wide<FpCfg>(X);
#pragma unroll
      for (int i = 0; i < 8; i++) X[i] ^= r.v[i];
    }
    if (MODE == M_ADD || MODE == M_MUL_ADD || MODE == M_MUL2_ADD || (MODE == M_SPLIT && odd)) {
      wide_add(X, Y);
      wide_add(Y, X);
      wide_add(X, Y);
      wide_add(Y, X);
    }
    if (MODE == M_MUL_LDS || MODE == M_LDS) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) X[i] ^= ((volatile uint32_t*)sm)[((i + r) & 15) * 256 + threadIdx.x];
    }
    if (MODE == M_MUL_STS) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) ((volatile uint32_t*)sm)[((i + r) & 15) * 256 + threadIdx.x] = T[i];
    }
    if (MODE == M_MUL_LDL) {
#pragma unroll
      for (int i = 0; i < 16; i++) ((volatile uint32_t*)lp)[i] = T[i];
#pragma unroll
      for (int i = 0; i < 16; i++) X[i] ^= ((volatile uint32_t*)lp)[i];
#pragma unroll
      for (int i = 0; i < 16; i++) ((volatile uint32_t*)lp)[i] = X[i];
#pragma unroll
      for (int i = 0; i < 16; i++) Y[i] ^= ((volatile uint32_t*)lp)[i];
    }
  }
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = a.v[i] ^ b.v[i] ^ X[i] ^ X[i + 8] ^ Y[i] ^ Y[i + 8] ^ T[i] ^ U[i];
  io[tid] = r;
}

template <int MODE>
static float run(int blocks, int iters, Fp* io, uint32_t* scratch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(iters, io, scratch);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  if (cudaGetLastError() != cudaSuccess) printf("CUDA error in mode %d\n", MODE);
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const double ghz = prop.clockRate / 1e6;
  const size_t maxthreads = (size_t)sms * 8 * 256;
  Fp* io;
  uint32_t* scratch;
  cudaMalloc(&io, maxthreads * sizeof(Fp));
  cudaMemset(io, 0x5a, maxthreads * sizeof(Fp));
  cudaMalloc(&scratch, maxthreads * 16 * 4);
  const int iters = 2048;
  printf("device %s, %d SMs, %.3f GHz (nominal max)\n", prop.name, sms, ghz);
  for (int mode = 0; mode < M_N; mode++) {
    for (int bps = 1; bps <= 8; bps *= 2) {  // blocks of 8 warps per SM: 2 warps per SMSP each
      const int blocks = sms * bps;
      float ms = 0;
      switch (mode) {
#define C(M) case M: ms = run<M>(blocks, iters, io, scratch); break;
        C(M_MUL) C(M_ADD) C(M_MUL_ADD) C(M_SPLIT) C(M_MUL_LDS) C(M_MUL_STS) C(M_LDS) C(M_MUL_LDL) C(M_REDC) C(M_MUL2_ADD)
#undef C
      }
      const double cyc_iter_warp = ghz * 1e9 * ms * 1e-3 / iters;  // latency of one iteration of one warp
      const double cyc_iter_smsp = cyc_iter_warp / (2.0 * bps);    // SMSP time per warp-iteration (throughput view)
      printf("%-58s warps/SMSP %2d: %8.3f ms  %8.1f cyc/iter (one warp)  %7.1f cyc per warp-iteration on the SMSP\n",
             NAMES[mode], 2 * bps, ms, cyc_iter_warp, cyc_iter_smsp);
    }
  }
  return 0;
}
