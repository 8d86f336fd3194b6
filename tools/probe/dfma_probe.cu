// Probe: Montgomery multiplication mod p on the FP64 pipe (52-bit limbs, DFMA hi/lo splitting) against the production
// 32-bit-limb IMAD.WIDE multiplication, plus the raw DFMA issue rate with and without integer adds beside it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I snark-bn254-verifier_b200/csrc -o dfma_probe tools/probe/dfma_probe.cu
#include <stdio.h>
#include <stdlib.h>
#include "tower.cuh"
using namespace bn254;

#define M52 0xfffffffffffffull
#define NP52 0x20782e4866389ull
#define E1 (0x467ull << 52)  // bits of 2^104
#define E2 (0x433ull << 52)  // bits of 2^52

struct F52 {
  double l[5];
};

__device__ __forceinline__ void prod(uint64_t* c, int k, double a, double b) {
  const double hi = __fma_rz(a, b, 0x1p104);
  const double lo = __fma_rz(a, b, (0x1p104 + 0x1p52) - hi);
  c[k] += (uint64_t)__double_as_longlong(lo);
  c[k + 1] += (uint64_t)__double_as_longlong(hi);
}

__device__ __forceinline__ constexpr int cnt_lo(int k) { return k > 8 ? 0 : (k < 5 ? k + 1 : 9 - k); }

__device__ __forceinline__ F52 mul52(const F52& a, const F52& b) {
  constexpr double N52[5] = {(double)0x8c16d87cfd47ull, (double)0x916871ca8d3c2ull, (double)0x181585d97816aull, (double)0xa029b85045b68ull,
                             (double)0x30644e72e131ull};
  uint64_t c[11];
#pragma unroll
  for (int k = 0; k < 11; k++) {
    const int cl = cnt_lo(k), ch = k > 0 ? cnt_lo(k - 1) : 0;
    c[k] = 0ull - 2ull * ((uint64_t)cl * E2 + (uint64_t)ch * E1);
  }
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) prod(c, i + j, a.l[i], b.l[j]);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const uint64_t q = (c[i] * NP52) & M52;
    const double qd = __longlong_as_double((long long)(q | E2)) - 0x1p52;
#pragma unroll
    for (int j = 0; j < 5; j++) prod(c, i + j, qd, N52[j]);
    c[i + 1] += c[i] >> 52;
  }
  F52 r;
#pragma unroll
  for (int k = 5; k < 10; k++) {
    if (k < 9) c[k + 1] += c[k] >> 52;
    r.l[k - 5] = __longlong_as_double((long long)((c[k] & M52) | E2)) - 0x1p52;
  }
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(128) k52(F52* x, const F52* y, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  F52 a = x[2 * i], a2 = x[2 * i + 1], b = y[i];
  for (int j = 0; j < iters; j++) {
    a = mul52(a, b);
    if (MODE == 1) a2 = mul52(a2, b);
  }
  x[2 * i] = a;
  x[2 * i + 1] = a2;
}

__global__ void __launch_bounds__(128) k32(Fp* x, const Fp* y, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp a = x[2 * i], a2 = x[2 * i + 1], b = y[i];
  for (int j = 0; j < iters; j++) {
    a = fe_mul_inl(a, b);
    a2 = fe_mul_inl(a2, b);
  }
  x[2 * i] = a;
  x[2 * i + 1] = a2;
}

// raw pipes: NF independent DFMA chains and NI 64-bit integer adds per iteration
template <int NF, int NI>
__global__ void __launch_bounds__(128) kraw(double* x, uint64_t* z, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double f[NF > 0 ? NF : 1];
  uint64_t u[NI > 0 ? NI : 1];
#pragma unroll
  for (int k = 0; k < NF; k++) f[k] = x[i] + k;
#pragma unroll
  for (int k = 0; k < NI; k++) u[k] = z[i] + k;
  const double m = x[i + 1];
  const uint64_t w = z[i + 1];
  for (int j = 0; j < iters; j++) {
#pragma unroll
    for (int k = 0; k < NF; k++) f[k] = __fma_rz(f[k], m, 0x1p-30);
#pragma unroll
    for (int k = 0; k < NI; k++) {
      uint32_t lo = (uint32_t)u[k], hi = (uint32_t)(u[k] >> 32);
      asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"((uint32_t)w), "r"((uint32_t)(w >> 32)));
      u[k] = ((uint64_t)hi << 32) | lo;
    }
  }
  double s = 0;
  uint64_t t = 0;
#pragma unroll
  for (int k = 0; k < NF; k++) s += f[k];
#pragma unroll
  for (int k = 0; k < NI; k++) t += u[k];
  x[i] = s;
  z[i] = t;
}

typedef unsigned __int128 u128;
struct Big { uint64_t w[9]; };
static const uint64_t P64[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};

// host check: (a * b * 2^-260) mod p via plain long arithmetic on 64-bit words
static void to_words(const F52& a, uint64_t* w) {  // 5 x 52 -> 5 x 64 (value < 2^260)
  u128 acc = 0; int bits = 0, o = 0;
  for (int k = 0; k < 5; k++) {
    acc |= (u128)(uint64_t)a.l[k] << bits; bits += 52;
    while (bits >= 64) { w[o++] = (uint64_t)acc; acc >>= 64; bits -= 64; }
  }
  w[o++] = (uint64_t)acc;  // o == 5
}
static void mod_p(uint64_t* x, int n) {  // x (n words) mod p by shift-subtract; result in x[0..3]
  for (int bit = n * 64 - 254; bit >= 0; bit--) {
    // compare x with p << bit
    uint64_t s[12] = {0};
    for (int k = 0; k < 4; k++) {
      int wi = bit / 64, sh = bit % 64;
      s[k + wi] |= P64[k] << sh;
      if (sh) s[k + wi + 1] |= P64[k] >> (64 - sh);
    }
    int ge = 1;
    for (int k = n; k >= 0; k--) { uint64_t xv = k < n ? x[k] : 0; if (xv != s[k]) { ge = xv > s[k]; break; } }
    if (ge) { u128 br = 0; for (int k = 0; k < n; k++) { u128 d = (u128)x[k] - s[k] - (uint64_t)br; x[k] = (uint64_t)d; br = (d >> 64) & 1; } }
  }
}
static void mulmod(const uint64_t* a, const uint64_t* b, int na, int nb, uint64_t* out) {  // out = a*b mod p (4 words)
  uint64_t t[12] = {0};
  for (int i = 0; i < na; i++) { u128 cy = 0; for (int j = 0; j < nb; j++) { u128 v = (u128)a[i] * b[j] + t[i + j] + (uint64_t)cy; t[i + j] = (uint64_t)v; cy = v >> 64; } t[i + nb] += (uint64_t)cy; }
  mod_p(t, na + nb);
  for (int k = 0; k < 4; k++) out[k] = t[k];
}

int main() {
  const int threads = 128;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // ---- correctness of mul52 (one iteration): r * 2^260 == a * b (mod p)
  {
    const int n = 148 * 128;
    F52* h = (F52*)malloc(3 * n * sizeof(F52));
    srand(7);
    for (int i = 0; i < 3 * n; i++) for (int k = 0; k < 5; k++) {
      uint64_t v = ((uint64_t)rand() << 40) ^ ((uint64_t)rand() << 20) ^ rand();
      h[i].l[k] = (double)(k < 4 ? (v & M52) : (v & 0x3fffffffffffull));  // values < 2^254 .. some above p
      if (i % 97 == 0) h[i].l[k] = (double)(k < 4 ? M52 : 0x3fffffffffffull);
    }
    F52 *x, *y; cudaMalloc(&x, 2 * n * sizeof(F52)); cudaMalloc(&y, n * sizeof(F52));
    cudaMemcpy(x, h, 2 * n * sizeof(F52), cudaMemcpyHostToDevice); cudaMemcpy(y, h + 2 * n, n * sizeof(F52), cudaMemcpyHostToDevice);
    k52<1><<<148, 128>>>(x, y, 1);
    F52* r = (F52*)malloc(2 * n * sizeof(F52));
    cudaMemcpy(r, x, 2 * n * sizeof(F52), cudaMemcpyDeviceToHost);
    uint64_t R260[5] = {0, 0, 0, 0, 16};
    int bad = 0; double maxtop = 0;
    for (int i = 0; i < 2 * n && bad < 5; i++) {
      uint64_t a[5], b[5], rr[5], lhs[4], rhs[4];
      to_words(h[i], a); to_words(h[2 * n + i / 2], b); to_words(r[i], rr);
      for (int k = 0; k < 5; k++) if (r[i].l[k] < 0 || r[i].l[k] >= 0x1p52) bad++;
      if (r[i].l[4] > maxtop) maxtop = r[i].l[4];
      mulmod(a, b, 5, 5, lhs); mulmod(rr, R260, 5, 5, rhs);
      if (memcmp(lhs, rhs, 32)) { bad++; printf("mismatch at %d\n", i); }
    }
    printf("mul52 check: %s (max top limb %.0f = %.3f p)\n", bad ? "FAILED" : "ok", maxtop, maxtop / (double)0x30644e72e131ull);
    cudaFree(x); cudaFree(y); free(h); free(r);
  }
  const int iters = 2000;
  for (int wps = 1; wps <= 8; wps *= 2) {
    int blocks = 148 * wps; size_t n = (size_t)blocks * threads;
    F52 *x, *y; cudaMalloc(&x, 2 * n * sizeof(F52)); cudaMalloc(&y, n * sizeof(F52));
    cudaMemset(x, 0, 2 * n * sizeof(F52)); cudaMemset(y, 0, n * sizeof(F52));
    Fp *x3, *y3; cudaMalloc(&x3, 2 * n * sizeof(Fp)); cudaMalloc(&y3, n * sizeof(Fp));
    cudaMemset(x3, 1, 2 * n * sizeof(Fp)); cudaMemset(y3, 2, n * sizeof(Fp));
    float b0 = 1e9, b1 = 1e9, b2 = 1e9, ms;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0); k52<0><<<blocks, threads>>>(x, y, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); if (ms < b0) b0 = ms;
      cudaEventRecord(e0); k52<1><<<blocks, threads>>>(x, y, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); if (ms < b1) b1 = ms;
      cudaEventRecord(e0); k32<<<blocks, threads>>>(x3, y3, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); if (ms < b2) b2 = ms;
    }
    const double cyc = 1.965e9 * 1e-3 / ((double)iters * wps);
    printf("warps/SMSP %d: mul52 x1 %7.1f cycles/mul   mul52 x2 %7.1f cycles/mul   fe_mul (imad) x2 %7.1f cycles/mul   [%s]\n", wps, b0 * cyc, b1 * cyc / 2, b2 * cyc / 2,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(x); cudaFree(y); cudaFree(x3); cudaFree(y3);
  }
  {
    const int wps = 4, blocks = 148 * wps; size_t n = (size_t)blocks * threads + 1;
    double* x; uint64_t* z; cudaMalloc(&x, n * 8); cudaMalloc(&z, n * 8); cudaMemset(x, 0, n * 8); cudaMemset(z, 1, n * 8);
    float ms; const double cyc = 1.965e9 * 1e-3 / ((double)iters * wps);
#define RAW(NF, NI)                                                                                               \
  { float best = 1e9; for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); kraw<NF, NI><<<blocks, threads>>>(x, z, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); \
      cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }                                                  \
    printf("raw: %2d DFMA + %2d 64-bit adds (2 instr each) per iteration: %6.1f cycles per iteration per warp (4 warps/SMSP)\n", NF, NI, best * cyc); }
    RAW(12, 0) RAW(0, 12) RAW(12, 4) RAW(12, 8) RAW(12, 12) RAW(12, 16)
  }
  return 0;
}
