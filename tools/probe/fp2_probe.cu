// Throughput probe for the field primitives (register-resident loops) at several occupancies.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I snark-bn254-verifier_b200/csrc -o fp2_probe tools/probe/fp2_probe.cu
#include <stdio.h>
#include "tower.cuh"
using namespace bn254;

template <int MODE>
__global__ void __launch_bounds__(128) k(Fp2* x, const Fp2* y, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp2 a = x[i], b = y[i];
  for (int j = 0; j < iters; j++) {
    if (MODE == 0) a = mul(a, b);                       // lazy Fp2 mul: 336 MACs
    else if (MODE == 1) a = sqr(a);                     // 272 MACs
    else if (MODE == 2) { a.c0 = fe_mul(a.c0, b.c0); a.c1 = fe_mul(a.c1, b.c1); }  // 2 x 136
    else if (MODE == 3) { a = add(a, b); b = sub(b, a); }  // 4 Fp add/sub
    else { a = mul(add(a, b), sub(a, b)); a = add(mul_xi(a), b); }  // mixed: 1 mul + 2 add + mul_xi(10 adds)+1 add
  }
  x[i] = a;
}

int main() {
  const int threads = 128;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[5] = {"fp2 mul (lazy, 336 MAC)", "fp2 sqr (272 MAC)", "2 x fe_mul (272 MAC)", "fp2 add+sub (0 MAC)", "mixed (336 MAC + 14 fp2 adds)"};
  const double macs[5] = {336, 272, 272, 0, 336};
  for (int mode = 0; mode < 5; mode++) {
    for (int wps = 1; wps <= 8; wps *= 2) {  // warps per SMSP
      int blocks = 148 * wps;                // 128 threads = 4 warps = 1 warp per SMSP per block
      size_t n = (size_t)blocks * threads;
      Fp2 *x, *y; cudaMalloc(&x, n * sizeof(Fp2)); cudaMalloc(&y, n * sizeof(Fp2));
      cudaMemset(x, 1, n * sizeof(Fp2)); cudaMemset(y, 2, n * sizeof(Fp2));
      int iters = 2000;
      float best = 1e9;
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        switch (mode) { case 0: k<0><<<blocks, threads>>>(x, y, iters); break; case 1: k<1><<<blocks, threads>>>(x, y, iters); break;
          case 2: k<2><<<blocks, threads>>>(x, y, iters); break; case 3: k<3><<<blocks, threads>>>(x, y, iters); break; default: k<4><<<blocks, threads>>>(x, y, iters); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double ops = (double)n * iters;
      double cyc = 1.965e9 * best * 1e-3 / ((double)iters * wps);  // cycles per iteration per warp-slot on an SMSP
      printf("%-32s warps/SMSP %d: %7.3f ms  %6.2f T MAC/s  %7.1f cycles per op per SMSP-warp\n", names[mode], wps, best, ops * macs[mode] / best / 1e9, cyc);
      cudaFree(x); cudaFree(y);
    }
  }
  return 0;
}
