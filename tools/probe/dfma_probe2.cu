// Finer probe of the FP64 pipe patterns used by a DFMA-based multiplier.  Each kernel runs `iters` iterations of 12 independent operations per warp.
#include <stdio.h>
#include <stdint.h>
template <int MODE>
__global__ void __launch_bounds__(128) k(double* x, uint64_t* z, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double f[12], g[12];
  uint64_t u[4];
#pragma unroll
  for (int k = 0; k < 12; k++) { f[k] = x[i] + k; g[k] = x[i] - k; }
#pragma unroll
  for (int k = 0; k < 4; k++) u[k] = z[i] + k;
  double m = x[i + 1];
  const uint32_t w0 = (uint32_t)z[i + 1], w1 = (uint32_t)(z[i + 1] >> 32);
  for (int j = 0; j < iters; j++) {
#pragma unroll
    for (int k = 0; k < 12; k++) {
      if (MODE == 0) f[k] = __fma_rz(f[k], m, 0x1p-30);            // DFMA reg, reg, imm
      if (MODE == 1) f[k] = __fma_rz(g[k], m, f[k]);               // DFMA reg, reg, reg
      if (MODE == 2) f[k] = __dadd_rn(0x1p60, -f[k]);              // DADD
      if (MODE == 3) { const double hi = __fma_rz(f[k], m, 0x1p104); f[k] = __fma_rz(f[k], m, (0x1p104 + 0x1p52) - hi) - 0x1p52; }  // product pattern, 4 FP64 ops
      if (MODE == 4 || MODE == 5) {                                                                                  // product pattern + integer accumulation
        const double hi = __fma_rz(f[k], m, 0x1p104);
        const double lo = __fma_rz(f[k], m, (0x1p104 + 0x1p52) - hi);
        if (MODE == 4) { u[k & 3] += (uint64_t)__double_as_longlong(hi); u[(k + 1) & 3] += (uint64_t)__double_as_longlong(lo); }
        else u[k & 3] += (uint64_t)__double_as_longlong(hi) + (uint64_t)__double_as_longlong(lo);
        f[k] = lo - 0x1p52;
      }
      if (MODE == 7) {  // product pattern; integer adds of unrelated registers
        const double hi = __fma_rz(f[k], m, 0x1p104);
        f[k] = __fma_rz(f[k], m, (0x1p104 + 0x1p52) - hi) - 0x1p52;
        uint32_t lo0 = (uint32_t)u[k & 3], hi0 = (uint32_t)(u[k & 3] >> 32);
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3; add.cc.u32 %0, %0, %3; addc.u32 %1, %1, %2;" : "+r"(lo0), "+r"(hi0) : "r"(w0), "r"(w1));
        u[k & 3] = ((uint64_t)hi0 << 32) | lo0;
      }
      if (MODE == 8) {  // product pattern; integer adds consume the previous iteration's results
        u[k & 3] += (uint64_t)__double_as_longlong(g[k]); u[(k + 1) & 3] += (uint64_t)__double_as_longlong(f[k]);
        const double hi = __fma_rz(f[k], m, 0x1p104);
        g[k] = hi;
        f[k] = __fma_rz(f[k], m, (0x1p104 + 0x1p52) - hi) - 0x1p52;
      }
      if (MODE == 6) { f[k] = __fma_rz(g[k], m, f[k]); g[k] = __fma_rn(f[k], m, g[k]); }  // 2 dependent DFMA reg reg reg
    }
  }
  double s = 0; uint64_t t = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) s += f[k] + g[k];
#pragma unroll
  for (int k = 0; k < 4; k++) t += u[k];
  x[i] = s; z[i] = t;
}
int main() {
  const int threads = 128, iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[9] = {"12 DFMA r,r,imm", "12 DFMA r,r,r", "12 DADD", "12 x (DFMA,DADD,DFMA,DADD)", "12 x (DFMA,DADD,DFMA,DADD) + 2 separate 64-bit adds", "12 x (DFMA,DADD,DFMA,DADD) + one 3-input 64-bit add", "12 x 2 DFMA r,r,r", "12 x (DFMA,DADD,DFMA,DADD) + 2 unrelated 64-bit adds", "12 x (DFMA,DADD,DFMA,DADD) + 2 64-bit adds of the previous iteration's results"};
  for (int wps = 4; wps <= 8; wps *= 2) {
    int blocks = 148 * wps; size_t n = (size_t)blocks * threads + 1;
    double* x; uint64_t* z; cudaMalloc(&x, n * 8); cudaMalloc(&z, n * 8); cudaMemset(x, 0, n * 8); cudaMemset(z, 1, n * 8);
    for (int mode = 0; mode < 9; mode++) {
      float best = 1e9, ms;
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        switch (mode) { case 0: k<0><<<blocks, threads>>>(x, z, iters); break; case 1: k<1><<<blocks, threads>>>(x, z, iters); break; case 2: k<2><<<blocks, threads>>>(x, z, iters); break;
          case 3: k<3><<<blocks, threads>>>(x, z, iters); break; case 4: k<4><<<blocks, threads>>>(x, z, iters); break; case 5: k<5><<<blocks, threads>>>(x, z, iters); break; case 6: k<6><<<blocks, threads>>>(x, z, iters); break; case 7: k<7><<<blocks, threads>>>(x, z, iters); break; default: k<8><<<blocks, threads>>>(x, z, iters); }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("warps/SMSP %d  %-84s %7.1f cycles per iteration per warp\n", wps, names[mode], 1.965e9 * best * 1e-3 / ((double)iters * wps));
    }
    cudaFree(x); cudaFree(z);
  }
  return 0;
}
