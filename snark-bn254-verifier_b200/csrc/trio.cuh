// Three lanes per proof ("trio"): the pairing of ONE proof runs on three adjacent lanes of a warp, sliced at
// Fq2-multiplication granularity -- the thread-group kernel shape of BASELINE.json's north_star.  Lane j of a trio
// holds coefficient j of every Fq6 value (an Fq2, 16 registers), so an Fq12 = (c0, c1) is two Fq2 per lane and the
// Miller accumulator lives in registers instead of a 4 KB per-thread stack.  The Karatsuba Fq6 multiplication splits
// evenly: each lane forms v_j = a_j b_j and one cross product, exchanging operands and partial products with
// warp shuffles; no lane executes a multiplication the sequential algorithm does not have (zero extra multiply-adds
// on dense operands).  One proof finishes three times sooner than with one proof per thread, which is what batches
// below ~2^15 proofs need: there a thread-per-proof launch leaves most of the 592 SM sub-partitions with less than
// one warp (PlonK at 2^14: stage E was 52 % of the step at 0.22 of the roofline).
//
// Every value is the same field element as in the sequential code (tower_body.inc / pairing_body.inc), hence the same
// canonical bytes: the formulas only re-associate which lane forms which product.
//
// The code is written once over the "lane vector" types V1 / V2 below.  Under nvcc a lane vector is that lane's
// value and the data movement primitives are __shfl_sync; under a plain C++ compiler (tests/hostsim, no GPU) it is the triple of the
// three lanes' values and the same primitives permute the triple, so the CPU tests execute the identical schedule in
// lock-step and compare it with the oracle.  Control flow never depends on the lane index: lanes differ only through
// tri_sel / tri_get, so a warp never diverges.
#pragma once
#include "pairing.cuh"

namespace bn254 {
namespace trio {

#if defined(__CUDACC__)  // (nvcc, host and device passes alike: the kernels must see the same types in both)
// ---------------------------------------------------------------------------------------------- device: one lane
typedef Fp V1;
typedef Fp2 V2;
#define TRIO_DEV 1
__device__ __forceinline__ int lane_j() { return (int)((threadIdx.x & 31u) % 3u); }
__device__ __forceinline__ int lane_base() { return (int)(threadIdx.x & 31u) - lane_j(); }
// Data exchange inside a trio goes through a per-warp area of shared memory (2 KB per warp: 4 x 16-byte quads per
// lane in a [quad][lane] layout, so a warp's 128-bit stores and loads are conflict-free): tri_put publishes this lane's
// value, tri_fetch reads the value published by lane s_j of the trio; one put serves any number of fetches.  Measured
// against __shfl_sync: a shuffle moves one word and, inside these out-of-line functions, ptxas brackets every pair with
// WARPSYNC / ENDCOLLECTIVE (368 instructions per Fq6 multiplication against 56 this way).
// Kernels that run trio code reserve trio_smem_bytes(threads) of dynamic shared memory.
extern __shared__ uint4 bn_trio_xchg[];
__host__ __device__ constexpr size_t trio_smem_bytes(int threads) { return (size_t)(threads / 32) * 128 * sizeof(uint4); }
__device__ __forceinline__ uint4* trio_area() { return bn_trio_xchg + (threadIdx.x >> 5) * 128; }
__device__ __forceinline__ int trio_src(int s0, int s1, int s2) {
  const int j = lane_j();
  return (lane_base() + (j == 0 ? s0 : (j == 1 ? s1 : s2))) & 31;
}
__device__ __forceinline__ void tri_put(const V1& x) {
  uint4* p = trio_area() + (threadIdx.x & 31u);
  __syncwarp();
  p[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  p[32] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  __syncwarp();
}
__device__ __forceinline__ void tri_put(const V2& x) {
  uint4* p = trio_area() + (threadIdx.x & 31u);
  __syncwarp();
  p[0] = make_uint4(x.c0.v[0], x.c0.v[1], x.c0.v[2], x.c0.v[3]);
  p[32] = make_uint4(x.c0.v[4], x.c0.v[5], x.c0.v[6], x.c0.v[7]);
  p[64] = make_uint4(x.c1.v[0], x.c1.v[1], x.c1.v[2], x.c1.v[3]);
  p[96] = make_uint4(x.c1.v[4], x.c1.v[5], x.c1.v[6], x.c1.v[7]);
  __syncwarp();
}
__device__ __forceinline__ V1 tri_fetch1(int s0, int s1, int s2) {
  const uint4* p = trio_area() + trio_src(s0, s1, s2);
  const uint4 a = p[0], b = p[32];
  V1 r;
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ V2 tri_fetch(int s0, int s1, int s2) {
  const uint4* p = trio_area() + trio_src(s0, s1, s2);
  const uint4 a = p[0], b = p[32], c = p[64], d = p[96];
  V2 r;
  r.c0.v[0] = a.x, r.c0.v[1] = a.y, r.c0.v[2] = a.z, r.c0.v[3] = a.w;
  r.c0.v[4] = b.x, r.c0.v[5] = b.y, r.c0.v[6] = b.z, r.c0.v[7] = b.w;
  r.c1.v[0] = c.x, r.c1.v[1] = c.y, r.c1.v[2] = c.z, r.c1.v[3] = c.w;
  r.c1.v[4] = d.x, r.c1.v[5] = d.y, r.c1.v[6] = d.z, r.c1.v[7] = d.w;
  return r;
}
// lane j receives x from lane s_j of its trio
__device__ __forceinline__ V2 tri_get(const V2& x, int s0, int s1, int s2) {
  tri_put(x);
  return tri_fetch(s0, s1, s2);
}
// lane j takes x_j.  Written with opaque bit masks: from a ternary on the lane index ptxas builds a small divergent
// branch per word (BSSY / BRA / BSYNC, measured: a third of the issue slots of the first version of these kernels).
__device__ __forceinline__ V1 tri_sel(const V1& x0, const V1& x1, const V1& x2) {
  const int j = lane_j();
  uint32_t m0 = j == 0 ? 0xffffffffu : 0u, m01 = j <= 1 ? 0xffffffffu : 0u;
  asm volatile("" : "+r"(m0), "+r"(m01));
  V1 r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t t = x1.v[i] ^ ((x0.v[i] ^ x1.v[i]) & m0);
    r.v[i] = x2.v[i] ^ ((t ^ x2.v[i]) & m01);
  }
  return r;
}
__device__ __forceinline__ V2 tri_sel(const V2& x0, const V2& x1, const V2& x2) {
  return V2{tri_sel(x0.c0, x1.c0, x2.c0), tri_sel(x0.c1, x1.c1, x2.c1)};
}
// two-way forms (one logic instruction per word): lane 0 takes x0, lanes 1 and 2 take x12 / lanes 0 and 1 take x01, lane 2 x2
__device__ __forceinline__ V2 tri_sel_0(const V2& x0, const V2& x12) {
  uint32_t m = lane_j() == 0 ? 0xffffffffu : 0u;
  asm volatile("" : "+r"(m));
  V2 r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.c0.v[i] = x12.c0.v[i] ^ ((x0.c0.v[i] ^ x12.c0.v[i]) & m);
    r.c1.v[i] = x12.c1.v[i] ^ ((x0.c1.v[i] ^ x12.c1.v[i]) & m);
  }
  return r;
}
__device__ __forceinline__ V2 tri_sel_2(const V2& x01, const V2& x2) {
  uint32_t m = lane_j() == 2 ? 0xffffffffu : 0u;
  asm volatile("" : "+r"(m));
  V2 r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.c0.v[i] = x01.c0.v[i] ^ ((x2.c0.v[i] ^ x01.c0.v[i]) & m);
    r.c1.v[i] = x01.c1.v[i] ^ ((x2.c1.v[i] ^ x01.c1.v[i]) & m);
  }
  return r;
}
// true on every lane of the trio iff `ok` holds on all three
__device__ __forceinline__ bool tri_all(bool ok) {
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  return ((m >> lane_base()) & 7u) == 7u;
}
#define TRIO_FN __device__ __forceinline__
#define TRIO_FN_NOINLINE static __device__ __noinline__
TRIO_FN V1 v_add(const V1& a, const V1& b) { return fe_add(a, b); }
TRIO_FN V1 v_mul(const V1& a, const V1& b) { return fe_mul(a, b); }
TRIO_FN V2 v_add(const V2& a, const V2& b) { return add(a, b); }
TRIO_FN V2 v_add_nr(const V2& a, const V2& b) { return V2{fe_add_nr(a.c0, b.c0), fe_add_nr(a.c1, b.c1)}; }
TRIO_FN V2 v_sub(const V2& a, const V2& b) { return sub(a, b); }
TRIO_FN V2 v_neg(const V2& a) { return neg(a); }
TRIO_FN V2 v_dbl(const V2& a) { return dbl(a); }
TRIO_FN V2 v_conj(const V2& a) { return conj(a); }
TRIO_FN V2 v_mul(const V2& a, const V2& b) { return mul(a, b); }
TRIO_FN V2 v_mul_inl(const V2& a, const V2& b) { return mul_inl(a, b); }
TRIO_FN V2 v_sqr(const V2& a) { return sqr(a); }
TRIO_FN V2 v_xi(const V2& a) { return mul_xi(a); }
TRIO_FN V2 v_scale(const V2& a, const V1& k) { return scale(a, k); }
TRIO_FN V2 v_scale2_add(const V2& a, const V1& j, const V2& b, const V1& k) { return scale2_add(a, j, b, k); }
TRIO_FN V2 v_inv(const V2& a) { return inv(a); }
TRIO_FN V2 v_halve(const V2& a) { return fp2_halve(a); }
TRIO_FN V2 v_embed(const V1& k) { return V2{k, fe_zero<FpCfg>()}; }
TRIO_FN bool v_is_zero_lane(const V2& a) { return is_zero(a); }
TRIO_FN bool v_eq(const V2& a, const V2& b) { return tri_all(eq(a, b)); }
// lane 0: a == b, lane 1: a == b, lane 2: a != 0 -- all three must hold (end-point test of the G2 chain)
TRIO_FN bool v_eq_eq_nz(const V2& a, const V2& b) { return tri_all(lane_j() == 2 ? !is_zero(a) : eq(a, b)); }
TRIO_FN V2 v_zero() { return fp2_zero(); }
TRIO_FN V2 v_const(const Fp2& x0, const Fp2& x1, const Fp2& x2) { return tri_sel(x0, x1, x2); }
TRIO_FN V1 v_const(const Fp& x0, const Fp& x1, const Fp& x2) { return tri_sel(x0, x1, x2); }
TRIO_FN V1 v_bcast(const Fp& x) { return x; }
TRIO_FN V2 v_bcast(const Fp2& x) { return x; }
// this lane's slice of a full value stored in memory: element j of an array of three
TRIO_FN V2 v_load3(const Fp2* p) { return p[lane_j()]; }
// lane j loads element i_j of an array
TRIO_FN V2 v_pick(const Fp2* p, int i0, int i1, int i2) {
  const int j = lane_j();
  return p[j == 0 ? i0 : (j == 1 ? i1 : i2)];
}
TRIO_FN void v_store3(Fp2* p, const V2& x) { p[lane_j()] = x; }
#else
// ---------------------------------------------------------------------------------------------- host: the three lanes
struct V1 {
  Fp l[3];
};
struct V2 {
  Fp2 l[3];
};
#define TRIO_DEV 0
inline V2& trio_area() {
  static thread_local V2 area;
  return area;
}
inline void tri_put(const V2& x) { trio_area() = x; }
inline V2 tri_fetch(int s0, int s1, int s2) {
  const V2& x = trio_area();
  return V2{{x.l[s0], x.l[s1], x.l[s2]}};
}
inline V2 tri_get(const V2& x, int s0, int s1, int s2) { return V2{{x.l[s0], x.l[s1], x.l[s2]}}; }
inline V1 tri_sel(const V1& x0, const V1& x1, const V1& x2) { return V1{{x0.l[0], x1.l[1], x2.l[2]}}; }
inline V2 tri_sel(const V2& x0, const V2& x1, const V2& x2) { return V2{{x0.l[0], x1.l[1], x2.l[2]}}; }
inline V2 tri_sel_0(const V2& x0, const V2& x12) { return V2{{x0.l[0], x12.l[1], x12.l[2]}}; }
inline V2 tri_sel_2(const V2& x01, const V2& x2) { return V2{{x01.l[0], x01.l[1], x2.l[2]}}; }
#define TRIO_FN inline
#define TRIO_FN_NOINLINE static
#define TRIO_EACH(expr) \
  {                     \
    for (int _j = 0; _j < 3; _j++) { expr; } \
  }
inline V1 v_add(const V1& a, const V1& b) { V1 r; TRIO_EACH(r.l[_j] = fe_add(a.l[_j], b.l[_j])) return r; }
inline V1 v_mul(const V1& a, const V1& b) { V1 r; TRIO_EACH(r.l[_j] = fe_mul(a.l[_j], b.l[_j])) return r; }
inline V2 v_add(const V2& a, const V2& b) { V2 r; TRIO_EACH(r.l[_j] = add(a.l[_j], b.l[_j])) return r; }
inline V2 v_add_nr(const V2& a, const V2& b) {
  V2 r;
  TRIO_EACH(r.l[_j] = (Fp2{fe_add_nr(a.l[_j].c0, b.l[_j].c0), fe_add_nr(a.l[_j].c1, b.l[_j].c1)}))
  return r;
}
inline V2 v_sub(const V2& a, const V2& b) { V2 r; TRIO_EACH(r.l[_j] = sub(a.l[_j], b.l[_j])) return r; }
inline V2 v_neg(const V2& a) { V2 r; TRIO_EACH(r.l[_j] = neg(a.l[_j])) return r; }
inline V2 v_dbl(const V2& a) { V2 r; TRIO_EACH(r.l[_j] = dbl(a.l[_j])) return r; }
inline V2 v_conj(const V2& a) { V2 r; TRIO_EACH(r.l[_j] = conj(a.l[_j])) return r; }
inline V2 v_mul(const V2& a, const V2& b) { V2 r; TRIO_EACH(r.l[_j] = mul(a.l[_j], b.l[_j])) return r; }
inline V2 v_mul_inl(const V2& a, const V2& b) { return v_mul(a, b); }
inline V2 v_sqr(const V2& a) { V2 r; TRIO_EACH(r.l[_j] = sqr(a.l[_j])) return r; }
inline V2 v_xi(const V2& a) { V2 r; TRIO_EACH(r.l[_j] = mul_xi(a.l[_j])) return r; }
inline V2 v_scale(const V2& a, const V1& k) { V2 r; TRIO_EACH(r.l[_j] = scale(a.l[_j], k.l[_j])) return r; }
inline V2 v_scale2_add(const V2& a, const V1& j, const V2& b, const V1& k) {
  V2 r;
  TRIO_EACH(r.l[_j] = scale2_add(a.l[_j], j.l[_j], b.l[_j], k.l[_j]))
  return r;
}
inline V2 v_inv(const V2& a) { V2 r; TRIO_EACH(r.l[_j] = inv(a.l[_j])) return r; }
inline V2 v_halve(const V2& a) { V2 r; TRIO_EACH(r.l[_j] = fp2_halve(a.l[_j])) return r; }
inline V2 v_embed(const V1& k) { V2 r; TRIO_EACH(r.l[_j] = (Fp2{k.l[_j], fe_zero<FpCfg>()})) return r; }
inline bool v_eq(const V2& a, const V2& b) { return eq(a.l[0], b.l[0]) && eq(a.l[1], b.l[1]) && eq(a.l[2], b.l[2]); }
inline bool v_eq_eq_nz(const V2& a, const V2& b) { return eq(a.l[0], b.l[0]) && eq(a.l[1], b.l[1]) && !is_zero(a.l[2]); }
inline V2 v_zero() { return V2{{fp2_zero(), fp2_zero(), fp2_zero()}}; }
inline V2 v_const(const Fp2& x0, const Fp2& x1, const Fp2& x2) { return V2{{x0, x1, x2}}; }
inline V1 v_const(const Fp& x0, const Fp& x1, const Fp& x2) { return V1{{x0, x1, x2}}; }
inline V1 v_bcast(const Fp& x) { return V1{{x, x, x}}; }
inline V2 v_bcast(const Fp2& x) { return V2{{x, x, x}}; }
inline V2 v_load3(const Fp2* p) { return V2{{p[0], p[1], p[2]}}; }
inline V2 v_pick(const Fp2* p, int i0, int i1, int i2) { return V2{{p[i0], p[i1], p[i2]}}; }
inline void v_store3(Fp2* p, const V2& x) { p[0] = x.l[0], p[1] = x.l[1], p[2] = x.l[2]; }
#endif

// ------------------------------------------------------------------------------------------ sliced Fq6 / Fq12
// An Fq6 value is one V2 (lane j: coefficient c_j); an Fq12 value is the pair below (lane j: c0.c_j and c1.c_j).
struct S12 {
  V2 c0, c1;
};

// c = a b in Fq6 (Karatsuba, the formulas of tower_body.inc mul(Fp6)), and optionally vc = v c.
//   v_j = a_j b_j                                     (local)
//   t_j: lane 0 (a1+a2)(b1+b2), lane 1 (a0+a1)(b0+b1), lane 2 (a0+a2)(b0+b2)
//   c0 = v0 + xi (t0 - v1 - v2),  c1 = t1 - v0 - v1 + xi v2,  c2 = t2 - v0 - v2 + v1
// Lane 2 has no use for the multiplication by xi that the other two need, so it forms xi c2 there -- which is
// coefficient 0 of v c = (xi c2, c0, c1): the product by v comes with one rotation and no further arithmetic.
TRIO_FN_NOINLINE void fp6s_mul(V2& c, V2* vc, const V2& a, const V2& b) {
  // operand sums: lane 0 a1 + a2, lane 1 a0 + a1, lane 2 a0 + a2 (reading from oneself replaces a select)
  tri_put(a);
  const V2 sa = v_add_nr(tri_fetch(1, 0, 0), tri_fetch(2, 1, 2));
  tri_put(b);
  const V2 sb = v_add(tri_fetch(1, 0, 0), tri_fetch(2, 1, 2));  // one side reduced: the lazy product needs a b' < 4 p^2
  const V2 v = v_mul_inl(a, b);
  const V2 t = v_mul_inl(sa, sb);
  // u: lane 0 t - v1 - v2, lane 1 t - v0 - v1, lane 2 t - v0 - v2
  tri_put(v);
  const V2 u = v_sub(v_sub(t, tri_fetch(1, 0, 0)), tri_fetch(2, 1, 2));
  const V2 vo = tri_fetch(0, 2, 1);  // lane 1: v2, lane 2: v1 (lane 0: unused)
  const V2 c2 = v_add(u, vo);        // meaningful on lane 2
  const V2 xw = v_xi(tri_sel(u, vo, c2));  // lane 0: xi u, lane 1: xi v2, lane 2: xi c2
  c = tri_sel_2(v_add(tri_sel_0(v, u), xw), c2);  // lane 0: v0 + xi u, lane 1: u + xi v2, lane 2: c2
  if (vc) *vc = tri_get(tri_sel_2(c, xw), 2, 0, 1);  // (xi c2, c0, c1)
}
TRIO_FN V2 fp6s_mul(const V2& a, const V2& b) {
  V2 c;
  fp6s_mul(c, nullptr, a, b);
  return c;
}
// v a = (xi a2, a0, a1) without a product
TRIO_FN_NOINLINE V2 fp6s_mul_v(const V2& a) {
  const V2 r = tri_get(a, 2, 0, 1);
  return tri_sel_0(v_xi(r), r);
}

TRIO_FN S12 fp12s_one() {
  return S12{v_const(fp2_one(), fp2_zero(), fp2_zero()), v_zero()};
}
TRIO_FN bool fp12s_eq(const S12& a, const S12& b) {
  const bool e0 = v_eq(a.c0, b.c0), e1 = v_eq(a.c1, b.c1);
  return e0 && e1;
}
TRIO_FN S12 fp12s_conj(const S12& a) { return S12{a.c0, v_neg(a.c1)}; }

// Karatsuba over Fq6 (tower_body.inc mul(Fp12)): r0 = a0 b0 + v a1 b1, r1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1
TRIO_FN_NOINLINE void fp12s_mul(S12& r, const S12& a, const S12& b) {
  V2 A, B, vB;
  fp6s_mul(A, nullptr, a.c0, b.c0);
  fp6s_mul(B, &vB, a.c1, b.c1);
  const V2 C = fp6s_mul(v_add(a.c0, a.c1), v_add(b.c0, b.c1));
  r.c0 = v_add(A, vB);
  r.c1 = v_sub(v_sub(C, A), B);
}
// complex squaring (tower_body.inc sqr(Fp12)): m = a0 a1, r0 = (a0 + a1)(a0 + v a1) - m - v m, r1 = 2 m
TRIO_FN_NOINLINE void fp12s_sqr(S12& r, const S12& a) {
  V2 m, vm;
  const V2 s = fp6s_mul(v_add(a.c0, a.c1), v_add(a.c0, fp6s_mul_v(a.c1)));
  fp6s_mul(m, &vm, a.c0, a.c1);
  r.c0 = v_sub(v_sub(s, m), vm);
  r.c1 = v_dbl(m);
}

// Granger-Scott squaring in the cyclotomic subgroup (tower_body.inc cyclotomic_sqr).  With z0 = c0.c0, z1 = c1.c1,
// z2 = c1.c0, z3 = c0.c2, z4 = c0.c1, z5 = c1.c2 the three Fq4 squarings (z0, z1), (z4, z5), (z2, z3) fall one per lane:
// lane j holds one member in its c0 slice and fetches the other from the c1 slice of lane j + 1.
TRIO_FN_NOINLINE void fp12s_cyclotomic_sqr(S12& r, const S12& a) {
  const V2 y = tri_get(a.c1, 1, 2, 0);         // lane 0: z1, lane 1: z5, lane 2: z2
  const V2 za = tri_sel_2(a.c0, y);            // (z0, z4, z2)
  const V2 zb = tri_sel_2(y, a.c0);            // (z1, z5, z3)
  // fp4_sqr: tmp = za zb, t0 = (za + zb)(za + xi zb) - tmp - xi tmp, t1 = 2 tmp
  const V2 tmp = v_mul(za, zb);
  const V2 t0 = v_sub(v_sub(v_mul(v_add(za, zb), v_add(za, v_xi(zb))), tmp), v_xi(tmp));
  const V2 t1 = v_dbl(tmp);
  // c0 slices: lane 0 <- t0 of (z0, z1) [lane 0], lane 1 <- t0 of (z2, z3) [lane 2], lane 2 <- t0 of (z4, z5) [lane 1]
  const V2 T0 = tri_get(t0, 0, 2, 1);
  // c1 slices: lane 0 <- xi t1 of (z4, z5) [lane 1], lane 1 <- t1 of (z0, z1) [lane 0], lane 2 <- t1 of (z2, z3) [lane 2]
  const V2 xt1 = v_xi(t1);
  const V2 T1 = tri_get(tri_sel(t1, xt1, t1), 1, 0, 2);
  r.c0 = v_add(v_dbl(v_sub(T0, a.c0)), T0);  // 3 t - 2 z
  r.c1 = v_add(v_dbl(v_add(T1, a.c1)), T1);  // 3 t + 2 z
}

// Frobenius^K: coefficient of w^i -> conj^K(.) * xi^(i (p^K - 1) / 6); c0.c_j is the coefficient of w^(2j), c1.c_j of w^(2j+1)
template <int KK>
TRIO_FN_NOINLINE void fp12s_frobenius(S12& r, const S12& a) {
  if (KK & 1) {
    const V2 g0 = v_const(fp2_one(), frob_coeff<KK>(2), frob_coeff<KK>(4));
    const V2 g1 = v_const(frob_coeff<KK>(1), frob_coeff<KK>(3), frob_coeff<KK>(5));
    r.c0 = v_mul(v_conj(a.c0), g0);
    r.c1 = v_mul(v_conj(a.c1), g1);
  } else {
    const V1 g0 = v_const(fe_one<FpCfg>(), frob_coeff_fp<KK>(2), frob_coeff_fp<KK>(4));
    const V1 g1 = v_const(frob_coeff_fp<KK>(1), frob_coeff_fp<KK>(3), frob_coeff_fp<KK>(5));
    r.c0 = v_scale(a.c0, g0);
    r.c1 = v_scale(a.c1, g1);
  }
}

// Fq6 inverse (tower_body.inc inv(Fp6)):  c0 = a0^2 - xi a1 a2, c1 = xi a2^2 - a0 a1, c2 = a1^2 - a0 a2,
// n = a0 c0 + xi (a2 c1 + a1 c2), r = c / n.  The Fq2 inversion (one Fermat chain in Fq) runs on all lanes at once.
TRIO_FN_NOINLINE V2 fp6s_inv(const V2& a) {
  const V2 a0 = tri_get(a, 0, 0, 0), a1 = tri_get(a, 1, 1, 1), a2 = tri_get(a, 2, 2, 2);
  const V2 sq = v_sqr(tri_sel(a0, a2, a1));
  const V2 pr = v_mul(tri_sel(a1, a0, a0), tri_sel(a2, a1, a2));
  const V2 x = v_xi(tri_sel(pr, sq, pr));  // lane 0: xi a1 a2, lane 1: xi a2^2
  const V2 c = tri_sel(v_sub(sq, x), v_sub(x, pr), v_sub(sq, pr));
  const V2 p = v_mul(tri_sel(a0, a2, a1), c);  // a0 c0, a2 c1, a1 c2
  const V2 p0 = tri_get(p, 0, 0, 0), p1 = tri_get(p, 1, 1, 1), p2 = tri_get(p, 2, 2, 2);
  const V2 n = v_add(p0, v_xi(v_add(p1, p2)));
  return v_mul(c, v_inv(n));
}
// Fq12 inverse (tower_body.inc inv(Fp12)): t = (a0^2 - v a1^2)^-1, r = (a0 t, -a1 t)
TRIO_FN_NOINLINE void fp12s_inv(S12& r, const S12& a) {
  V2 s1, vs1;
  const V2 s0 = fp6s_mul(a.c0, a.c0);
  fp6s_mul(s1, &vs1, a.c1, a.c1);
  const V2 t = fp6s_inv(v_sub(s0, vs1));
  r.c0 = fp6s_mul(a.c0, t);
  r.c1 = v_neg(fp6s_mul(a.c1, t));
}

// r = conj(a^x) over the width-4 non-adjacent form of x (pairing_body.inc exp_by_neg_z: same digits, same element)
TRIO_FN_NOINLINE void fp12s_exp_by_neg_z(S12& r, const S12& a) {
  const signed char DIG[63] = {1, 0, 0, 0, -1, 0, 0, 0, 0, 5, 0, 0, 0, 0, 0, 0, -7, 0, 0, 0, 7, 0, 0, 0, 0, 5, 0, 0, 0, 0, 1, 0, 0, 0, -3, 0, 0, 0, -5, 0, 0, 0, 5, 0, 0, 0, 0, 3, 0, 0, 0, -3, 0, 0, 0, 0, 5, 0, 0, 0, 0, 0, 1};
  S12 tab[4], t;
  tab[0] = a;
  fp12s_cyclotomic_sqr(t, a);
  for (int k = 1; k < 4; k++) fp12s_mul(tab[k], tab[k - 1], t);
  r = a;
  for (int i = 61; i >= 0; i--) {
    if ((i & (BN_SYNC_PERIOD_EXP - 1)) == 0) BN_PHASE_SYNC();
    fp12s_cyclotomic_sqr(r, r);
    const int d = DIG[i];
    if (d > 0) {
      fp12s_mul(r, r, tab[d >> 1]);
    } else if (d < 0) {
      fp12s_mul(r, r, fp12s_conj(tab[(-d) >> 1]));
    }
  }
  BN_PHASE_SYNC();
  r = fp12s_conj(r);
}

// Fq12::final_exponentiation: the chain of pairing_body.inc final_exponentiation, statement by statement
TRIO_FN_NOINLINE void fp12s_final_exponentiation(S12& r, const S12& f) {
  S12 T, A, B, D, E, Kk, L, X;
  fp12s_inv(A, f);
  fp12s_mul(T, fp12s_conj(f), A);  // f^(p^6 - 1)
  fp12s_frobenius<2>(A, T);
  fp12s_mul(T, A, T);              // t = f^((p^6-1)(p^2+1))
  fp12s_exp_by_neg_z(A, T);        // a
  fp12s_cyclotomic_sqr(B, A);      // b
  fp12s_cyclotomic_sqr(X, B);      // c
  fp12s_mul(D, X, B);              // d
  fp12s_exp_by_neg_z(E, D);        // e
  fp12s_cyclotomic_sqr(X, E);      // f
  fp12s_exp_by_neg_z(A, X);        // g
  A = fp12s_conj(A);               // i = conj(g)
  fp12s_mul(Kk, A, E);             // j = i e
  fp12s_mul(Kk, Kk, fp12s_conj(D));  // k = j h
  fp12s_mul(L, Kk, B);             // l = k b
  fp12s_mul(X, Kk, E);             // m = k e
  fp12s_mul(X, T, X);              // n = t m
  fp12s_frobenius<1>(A, L);        // o
  fp12s_mul(X, A, X);              // p = o n
  fp12s_frobenius<2>(A, Kk);       // q
  fp12s_mul(X, A, X);              // r = q p
  fp12s_mul(A, fp12s_conj(T), L);  // t' = s l
  fp12s_frobenius<3>(B, A);        // u
  fp12s_mul(r, B, X);
}

// ------------------------------------------------------------------------------------------ Miller loop, two
// VK-constant G2 points through their pair table (pairing_body.inc eval_line_pair / miller_loop_pairtab<0>):
//   M = (m0 + m1 v + m2 v^2) + (n0 + n1 v) w,   m0 = K0 + K1 (Y1 Y2), m1 = K2 (X1 X2), m2 = K3 X2 + K4 X1,
//   n0 = K5 (X1 Y2) + K6 (Y1 X2), n1 = K7 Y2 + K8 Y1.
// Lane j evaluates (m_j, n_j), each as one sum of two Fq2-by-Fq scalings (K0 = K0 * 1; missing terms scale by zero:
// exact, multiplying a reduced element by the Montgomery one or by zero returns it or zero).
struct PairScalarsS {
  V1 ma, mb, na, nb;  // the scalars of this lane's (m, n): m = KMA ma + KMB mb, n = KNA na + KNB nb
};
TRIO_FN PairScalarsS pair_scalars_s(const G1Aff& p1, const G1Aff& p2) {
  const PairScalars s = pair_scalars(p1, p2);
  const Fp one = fe_one<FpCfg>(), zero = fe_zero<FpCfg>();
  PairScalarsS r;
  r.ma = v_const(one, s.x1x2, s.x2);      // K0 * 1,      K2 (X1 X2),  K3 X2
  r.mb = v_const(s.y1y2, zero, s.x1);     // K1 (Y1 Y2),  -,           K4 X1
  r.na = v_const(s.x1y2, s.y2, zero);     // K5 (X1 Y2),  K7 Y2,       -
  r.nb = v_const(s.y1x2, s.y1, zero);     // K6 (Y1 X2),  K8 Y1,       -
  return r;
}
TRIO_FN_NOINLINE void eval_line_pair_s(S12& M, const LinePairKF& k, const PairScalarsS& s) {
  const V2 kma = v_pick(k.k, 0, 2, 3), kmb = v_pick(k.k, 1, 0, 4);  // (index 0 with a zero scalar stands for "no term")
  const V2 kna = v_pick(k.k, 5, 7, 0), knb = v_pick(k.k, 6, 8, 0);
  M.c0 = v_scale2_add(kma, s.ma, kmb, s.mb);
  M.c1 = v_scale2_add(kna, s.na, knb, s.nb);
}
// f = Miller value of e(P1, Q1) e(P2, Q2) for the two VK-constant G2 points behind `ptab`
TRIO_FN_NOINLINE void miller_loop_pairtab0_s(S12& f, const G1Aff* pf, const LinePairKF* ptab) {
  const PairScalarsS ps = pair_scalars_s(pf[0], pf[1]);
  S12 M;
  int idx = 0;
  for (int k = 0; k < 64; k++) {
    if ((k & (BN_SYNC_PERIOD - 1)) == 0) BN_PHASE_SYNC();
    eval_line_pair_s(M, ptab[idx], ps);
    if (k > 0) {
      fp12s_sqr(f, f);
      fp12s_mul(f, f, M);
    } else {
      f = M;  // 1 * M
    }
    idx++;
    if (K::ate_digit(k) != 0) {
      eval_line_pair_s(M, ptab[idx], ps);
      fp12s_mul(f, f, M);
      idx++;
    }
  }
  BN_PHASE_SYNC();
  eval_line_pair_s(M, ptab[idx], ps);
  fp12s_mul(f, f, M);
  eval_line_pair_s(M, ptab[idx + 1], ps);
  fp12s_mul(f, f, M);
}


// ------------------------------------------------------------------------------------------ G2 steps, sliced
// The running point R = (X, Y, Z) of a variable G2 point lives one coordinate per lane (lane 0 X, lane 1 Y, lane 2 Z).
// A line (x0 + x2 v^2 + x4 v w, pairing_body.inc Line3) is the sparse S12 {c0: (x0, 0, x2), c1: (0, x4, 0)}, i.e. lane 0
// holds (x0, 0), lane 1 (0, x4), lane 2 (x2, 0).  The step formulas are those of pairing_body.inc (doubling_step_at,
// addition_step_at); every round below is one multiplication per lane on per-lane operands.
TRIO_FN S12 line_pack(const V2& at0, const V2& at1, const V2& at2) {  // x0 held by lane 0, x4 by lane 1, x2 by lane 2
  const V2 z = v_zero();
  return S12{tri_sel(at0, z, at2), tri_sel(z, at1, z)};
}
TRIO_FN S12 line_one_s() { return fp12s_one(); }

// R <- 2R; line = tangent at R evaluated at P = (px, py)
TRIO_FN_NOINLINE void doubling_step_s(S12& line, V2& r, const V1& px, const V1& py) {
  const V2 s1 = v_sqr(r);                                   // (j = X^2, b = Y^2, c = Z^2)
  tri_put(r);
  const V2 o = tri_fetch(1, 2, 2);                          // lane 0: Y, lane 1: Z
  const V2 yz = v_add(r, o);                                // lane 1: Y + Z
  const V2 s3 = v_add(v_dbl(s1), s1);                       // lane 0: vv = 3 j, lane 2: 3 c
  const V2 m2 = v_mul(tri_sel(r, yz, v_bcast(fp2_b2())), tri_sel(o, yz, s3));  // (X Y, (Y + Z)^2, e = b' 3 c)
  tri_put(s1);
  const V2 sc = tri_fetch(1, 2, 1);                         // lane 0: b, lane 1: c, lane 2: b
  const V2 h = v_sub(v_sub(m2, s1), sc);                    // lane 1: h = (Y + Z)^2 - b - c
  tri_put(m2);
  const V2 e = tri_fetch(2, 2, 2);
  const V2 f = v_add(v_dbl(e), e);                          // 3 e
  const V2 bb = tri_sel(sc, s1, sc);                        // b on every lane
  const V2 g = v_halve(v_add(bb, f));
  const V2 m3 = v_mul(tri_sel(v_halve(m2), g, e), tri_sel(v_sub(bb, f), g, e));  // (X' = a (b - f), g^2, e^2)
  tri_put(m3);
  const V2 e2 = tri_fetch(2, 2, 2);
  const V2 y3 = v_sub(m3, v_add(v_dbl(e2), e2));            // lane 1: Y' = g^2 - 3 e^2
  tri_put(h);
  const V2 hh = tri_fetch(1, 1, 1);
  // (x2 = vv px on lane 0, x4 = -h py on lane 1, Z' = b h on lane 2)
  const V2 m4 = v_mul(tri_sel(s3, v_neg(h), bb), tri_sel(v_embed(px), v_embed(py), hh));
  const V2 x0 = v_xi(v_sub(e, bb));                          // every lane has e and b
  r = tri_sel(m3, y3, m4);
  tri_put(m4);
  line = line_pack(x0, m4, tri_fetch(0, 0, 0));
}

// R <- R + Q for the affine point Q = (qx, qy) (the same values on the three lanes); line = chord evaluated at P
TRIO_FN_NOINLINE void addition_step_s(S12& line, V2& r, const V2& qx, const V2& qy, const V1& px, const V1& py) {
  tri_put(r);
  const V2 z = tri_fetch(2, 2, 2), x = tri_fetch(0, 0, 0);
  const V2 de = v_sub(r, v_mul(z, tri_sel(qx, qy, qx)));    // lane 0: d = X - Z qx, lane 1: e = Y - Z qy
  tri_put(de);
  const V2 d = tri_fetch(0, 0, 0), e = tri_fetch(1, 1, 1);
  const V2 m2a = v_mul(tri_sel(d, e, e), tri_sel(d, e, qx));                       // (f = d^2, e^2, e qx)
  const V2 m2b = v_mul(tri_sel(d, v_neg(e), d), tri_sel(v_embed(py), v_embed(px), qy));  // (x4 = d py, x2 = -e px, d qy)
  const V2 x0 = v_xi(v_sub(m2a, m2b));                       // lane 2: xi (e qx - d qy)
  tri_put(tri_sel_2(m2b, x0));
  line = line_pack(tri_fetch(2, 2, 2), tri_fetch(0, 0, 0), tri_fetch(1, 1, 1));
  tri_put(m2a);
  const V2 ff = tri_fetch(0, 0, 0);
  const V2 m3 = v_mul(tri_sel(d, z, x), tri_sel(ff, m2a, ff));  // (h = d f, Z e^2, i = X f)
  tri_put(m3);
  const V2 hh = tri_fetch(0, 0, 0), ze2 = tri_fetch(1, 1, 1), ii = tri_fetch(2, 2, 2);
  const V2 jj = v_sub(v_add(ze2, hh), v_dbl(ii));            // j = Z e^2 + h - 2 i
  const V2 m4a = v_mul(tri_sel(d, e, z), tri_sel(jj, v_sub(ii, jj), hh));  // (X' = d j, e (i - j), Z' = Z h)
  const V2 m4b = v_mul(hh, r);                               // lane 1: h Y
  r = tri_sel(m4a, v_sub(m4a, m4b), m4a);
}

// M = l1 l2 for two sparse lines (pairing_body.inc mul_lines): 6 Fq2 multiplications in two rounds
TRIO_FN_NOINLINE void mul_lines_s(S12& M, const S12& l1, const S12& l2) {
  const V2 a = tri_sel(l1.c0, l1.c1, l1.c0), b = tri_sel(l2.c0, l2.c1, l2.c0);  // (x0, x4, x2), (y0, y4, y2)
  const V2 p = v_mul(a, b);                                  // (p00, p44, p22)
  tri_put(a);
  const V2 sa = v_add(a, tri_fetch(2, 0, 1));                // (x0 + x2, x4 + x0, x2 + x4)
  tri_put(b);
  const V2 sb = v_add(b, tri_fetch(2, 0, 1));
  const V2 cr = v_mul(sa, sb);
  tri_put(p);
  const V2 pa = tri_fetch(1, 2, 0), pb = tri_fetch(2, 0, 1);  // (p44, p22, p00), (p22, p00, p44)
  const V2 u = v_sub(v_sub(cr, p), pb);                      // lane 0: m2, lane 1: n1, lane 2: n0 / xi
  const V2 xw = v_xi(tri_sel_2(pa, u));                      // lane 0: xi p44, lane 1: m1 = xi p22, lane 2: n0
  tri_put(tri_sel_2(u, xw));
  const V2 g = tri_fetch(2, 2, 0);                           // lane 0: n0, lane 2: m2
  M.c0 = tri_sel(v_add(p, xw), xw, g);
  M.c1 = tri_sel(g, u, v_zero());
}

// psi(Q) for an affine G2 point given on all lanes (curve_body.inc g2_psi)
TRIO_FN void g2_psi_s(V2& ox, V2& oy, const V2& qx, const V2& qy) {
  ox = v_mul(v_conj(qx), v_bcast(frob_coeff<1>(2)));
  oy = v_mul(v_conj(qy), v_bcast(frob_coeff<1>(3)));
}
// R == -psi^3(Q) with Z != 0 (pairing_body.inc ate_endpoint_in_g2); (q2x, q2y) = -psi^2(Q)
TRIO_FN_NOINLINE bool ate_endpoint_in_g2_s(const V2& r, const V2& q2x, const V2& q2y) {
  V2 tx, ty;
  g2_psi_s(tx, ty, q2x, q2y);
  tri_put(r);
  const V2 z = tri_fetch(2, 2, 2);
  const V2 m = v_mul(tri_sel(tx, ty, v_bcast(fp2_one())), z);  // (tx Z, ty Z, Z)
  return v_eq_eq_nz(tri_sel(r, r, m), m);
}

// Shared-accumulator Miller loop: one variable pair (A, B) and the two VK-constant pairs behind the pair table
// (pairing_body.inc miller_loop_pairtab<1>: same order of operations, same value).  in_g2: B (on the curve) lies in G2.
TRIO_FN_NOINLINE void miller_loop_pairtab1_s(S12& f, const G1Aff& pa, const G2Aff& qb, const G1Aff* pf,
                                             const LinePairKF* ptab, bool* in_g2) {
  const PairScalarsS ps = pair_scalars_s(pf[0], pf[1]);
  const V1 px = v_bcast(pa.x), py = v_bcast(pa.y);
  const V2 qx = v_bcast(qb.x), qy = v_bcast(qb.y), nqy = v_neg(qy);
  V2 r = v_const(qb.x, qb.y, fp2_one());
  S12 M, l1, l2;
  int idx = 0;
  for (int k = 0; k < 64; k++) {
    if ((k & (BN_SYNC_PERIOD - 1)) == 0) BN_PHASE_SYNC();
    eval_line_pair_s(M, ptab[idx], ps);
    if (k > 0) {
      fp12s_sqr(f, f);
      fp12s_mul(f, f, M);
    } else {
      f = M;
    }
    idx++;
    doubling_step_s(l1, r, px, py);
    const int d = K::ate_digit(k);
    if (d != 0) {
      eval_line_pair_s(M, ptab[idx], ps);
      fp12s_mul(f, f, M);
      idx++;
      addition_step_s(l2, r, qx, d == 1 ? qy : nqy, px, py);
      mul_lines_s(M, l1, l2);
      fp12s_mul(f, f, M);
    } else {
      fp12s_mul(f, f, l1);
    }
  }
  BN_PHASE_SYNC();
  eval_line_pair_s(M, ptab[idx], ps);
  fp12s_mul(f, f, M);
  eval_line_pair_s(M, ptab[idx + 1], ps);
  fp12s_mul(f, f, M);
  V2 q1x, q1y, q2x, q2y;
  g2_psi_s(q1x, q1y, qx, qy);
  g2_psi_s(q2x, q2y, q1x, q1y);
  q2y = v_neg(q2y);
  addition_step_s(l1, r, q1x, q1y, px, py);
  addition_step_s(l2, r, q2x, q2y, px, py);
  if (in_g2) *in_g2 = ate_endpoint_in_g2_s(r, q2x, q2y);
  mul_lines_s(M, l1, l2);
  fp12s_mul(f, f, M);
}

// Shared-accumulator Miller loop over NV variable pairs (pairing_body.inc miller_loop<NV, 0>); skip: pairs with an
// identity member contribute the line 1.
template <int NV>
TRIO_FN_NOINLINE void miller_loop_var_s(S12& f, const G1Aff* pv, const G2Aff* qv, uint32_t skip) {
  V2 r[NV];
  S12 ls[2 * NV], M;
  for (int v = 0; v < NV; v++) r[v] = v_const(qv[v].x, qv[v].y, fp2_one());
  f = fp12s_one();
  for (int k = 0; k < 64; k++) {
    if ((k & (BN_SYNC_PERIOD - 1)) == 0) BN_PHASE_SYNC();
    if (k > 0) fp12s_sqr(f, f);
    for (int v = 0; v < NV; v++) {
      doubling_step_s(ls[v], r[v], v_bcast(pv[v].x), v_bcast(pv[v].y));
      if ((skip >> v) & 1) ls[v] = line_one_s();
    }
    const int d = K::ate_digit(k);
    int nl = NV;
    if (d != 0) {
      for (int v = 0; v < NV; v++) {
        const V2 qy = v_bcast(qv[v].y);
        addition_step_s(ls[NV + v], r[v], v_bcast(qv[v].x), d == 1 ? qy : v_neg(qy), v_bcast(pv[v].x), v_bcast(pv[v].y));
        if ((skip >> v) & 1) ls[NV + v] = line_one_s();
      }
      nl = 2 * NV;
    }
    for (int j = 0; j + 1 < nl; j += 2) {
      mul_lines_s(M, ls[j], ls[j + 1]);
      fp12s_mul(f, f, M);
    }
    if (nl & 1) fp12s_mul(f, f, ls[nl - 1]);
  }
  BN_PHASE_SYNC();
  for (int v = 0; v < NV; v++) {
    const V1 px = v_bcast(pv[v].x), py = v_bcast(pv[v].y);
    V2 q1x, q1y, q2x, q2y;
    g2_psi_s(q1x, q1y, v_bcast(qv[v].x), v_bcast(qv[v].y));
    g2_psi_s(q2x, q2y, q1x, q1y);
    addition_step_s(ls[v], r[v], q1x, q1y, px, py);
    addition_step_s(ls[NV + v], r[v], q2x, v_neg(q2y), px, py);
    if ((skip >> v) & 1) ls[v] = line_one_s(), ls[NV + v] = line_one_s();
  }
  for (int j = 0; j + 1 < 2 * NV; j += 2) {
    mul_lines_s(M, ls[j], ls[j + 1]);
    fp12s_mul(f, f, M);
  }
}

// ------------------------------------------------------------------------------------------ full values <-> slices
// An Fp12 in memory is c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2: the c0 slices are elements 0..2, the c1 slices 3..5.
TRIO_FN S12 fp12s_load(const Fp12& x) {
  const Fp2* p = (const Fp2*)&x;
  return S12{v_load3(p), v_load3(p + 3)};
}
TRIO_FN void fp12s_store(Fp12& x, const S12& s) {
  Fp2* p = (Fp2*)&x;
  v_store3(p, s.c0);
  v_store3(p + 3, s.c1);
}

#if TRIO_DEV
// ------------------------------------------------------------------------------------------ device plumbing
// Lanes 0..29 of a warp form ten trios; lanes 30 and 31 walk along on substitute data (warp-wide shuffles need them)
// and never write.  trio_slot(): index of this lane's trio among all trios of the grid.
#define BN_TRIOS_PER_WARP 10
__device__ __forceinline__ bool trio_lane_valid() { return (threadIdx.x & 31u) < 30u; }
__device__ __forceinline__ size_t trio_slot() {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  return warp * BN_TRIOS_PER_WARP + (threadIdx.x & 31u) / 3u;
}
// canonical bytes of this lane's two coefficients into a 384-byte Fq12 record (fp12_to_bytes order)
__device__ __forceinline__ void fp12s_to_bytes(uint8_t* out, const S12& a) {
  const int j = lane_j();
  fe_to_be_bytes(out + 64 * j, fe_from_mont(a.c0.c0));
  fe_to_be_bytes(out + 64 * j + 32, fe_from_mont(a.c0.c1));
  fe_to_be_bytes(out + 192 + 64 * j, fe_from_mont(a.c1.c0));
  fe_to_be_bytes(out + 192 + 64 * j + 32, fe_from_mont(a.c1.c1));
}
#endif

}  // namespace trio
}  // namespace bn254
