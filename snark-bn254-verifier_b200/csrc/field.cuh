// 254-bit prime-field arithmetic on 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Replaces substrate-bn's Fq / Fr (reference call sites: verifier/src/groth16/verify.rs:61,
// verifier/src/plonk/verify.rs:98-250).  Values are kept fully reduced in [0, m).
//
// Device path: PTX carry chains.  Each `mad.lo.cc / madc.hi.cc` pair is fused by ptxas into one
// IMAD.WIDE.U32(.X) with a predicate carry, so a multiplication costs 8*(8+8) wide MACs + 8 low
// IMADs = 136 integer-FMA-pipe issues plus a handful of IADD3.  The product is accumulated in two
// interleaved limb arrays (pairs starting at even / odd word positions) so every 64-bit product lands
// on an aligned register pair and the only inter-array traffic is one word per row.
//
// Host path (`!__CUDA_ARCH__`): the same algorithm with the carry flag emulated in a thread-local --
// used only by the host-side VK preparation and by tests/hostsim to debug kernel logic without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN static __host__ __device__ __noinline__
#else
#define HD inline
#define HDN static __attribute__((noinline))
#endif

// Phase barrier.  Every thread of a block runs the same data-independent sequence of big operations, but warps drift
// apart and then each one pulls a different part of the (hundreds of KB of) pairing code through the SM's instruction
// cache.  A block-wide barrier at the boundaries of the big operations keeps all warps of the block inside the same
// function at the same time, so the instruction working set is one function, not the whole kernel.  Threads that
// finished early (rejected / malformed proofs) have exited and do not take part in the barrier.
// RULE: only code that every live thread of the block executes the same number of times may contain the barrier
// (miller_loop, exp_by_neg_z, final_exponentiation, the subgroup check); generic helpers that callers may invoke
// under a data-dependent branch (fe_pow_words, scalar_mul<false>, jac_*) must not.
#if defined(__CUDA_ARCH__)
#define BN_PHASE_SYNC() __syncthreads()
#else
#define BN_PHASE_SYNC()
#endif
#ifndef BN_SYNC_PERIOD
#define BN_SYNC_PERIOD 1  // a barrier every BN_SYNC_PERIOD iterations of the Miller / exponentiation loops (power of 2)
#endif
#ifndef BN_SYNC_PERIOD_EXP
#define BN_SYNC_PERIOD_EXP 4  // the same for the much shorter iterations of the exponentiation by x (measured: 11.18 / 11.06 / 11.00 ms per launch at 1 / 2 / 4)
#endif
#ifdef BN_SYNC_FINE  // measured on B200: the per-iteration barrier alone is 3 % faster than barriers between all big operations
#define BN_PHASE_SYNC_FINE() BN_PHASE_SYNC()  // between the big operations inside one loop iteration
#else
#define BN_PHASE_SYNC_FINE()
#endif

namespace bn254 {

// ---------------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------------
namespace cc {
#if defined(__CUDA_ARCH__)
#define BN_ASM3(name, ins)                                                                   \
  __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b) {                         \
    uint32_t r;                                                                              \
    asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                              \
    return r;                                                                                \
  }
#define BN_ASM4(name, ins)                                                                   \
  __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b, uint32_t c) {             \
    uint32_t r;                                                                              \
    asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));                  \
    return r;                                                                                \
  }
BN_ASM3(add_cc, "add.cc.u32")
BN_ASM3(addc_cc, "addc.cc.u32")
BN_ASM3(addc, "addc.u32")
BN_ASM3(sub_cc, "sub.cc.u32")
BN_ASM3(subc_cc, "subc.cc.u32")
BN_ASM3(subc, "subc.u32")
BN_ASM4(mad_lo_cc, "mad.lo.cc.u32")
BN_ASM4(mad_hi_cc, "mad.hi.cc.u32")
BN_ASM4(madc_lo_cc, "madc.lo.cc.u32")
BN_ASM4(madc_hi_cc, "madc.hi.cc.u32")
BN_ASM4(madc_hi, "madc.hi.u32")
#undef BN_ASM3
#undef BN_ASM4
#else
// Host emulation of the PTX condition-code register (CC.CF).
inline uint32_t& CF() {
  static thread_local uint32_t cf = 0;
  return cf;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; CF() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + CF(); CF() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; CF() = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - CF(); CF() = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF(); }
inline uint32_t lo32(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t hi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(lo32(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(hi32(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(lo32(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(hi32(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(hi32(a, b), c); }
#endif
}  // namespace cc

// ---------------------------------------------------------------------------------------------
// Field parameter packs (little-endian 32-bit limbs).  Generated by tools/gen_constants.py.
// ---------------------------------------------------------------------------------------------
#define BN_LIMBS(name, ...)                                         \
  static HD constexpr uint32_t name(int i) {                        \
    constexpr uint32_t t[8] = {__VA_ARGS__};                        \
    return t[i];                                                    \
  }

struct FpCfg {  // base field, p
  BN_LIMBS(mod, 0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
  BN_LIMBS(r1, 0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
  BN_LIMBS(r2, 0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u)
  static constexpr uint32_t inv = 0xe4866389u;  // -p^-1 mod 2^32
};

struct FrCfg {  // scalar field, r
  BN_LIMBS(mod, 0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
  BN_LIMBS(r1, 0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
  BN_LIMBS(r2, 0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u)
  static constexpr uint32_t inv = 0xefffffffu;  // -r^-1 mod 2^32
};

template <class C>
struct alignas(16) Fe {
  uint32_t v[8];
};

template <class C>
HD Fe<C> fe_zero() {
  Fe<C> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}
template <class C>
HD Fe<C> fe_one() {  // Montgomery one
  Fe<C> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = C::r1(i);
  return r;
}
template <class C>
HD bool fe_is_zero(const Fe<C>& a) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t |= a.v[i];
  return t == 0;
}
template <class C>
HD bool fe_eq(const Fe<C>& a, const Fe<C>& b) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t |= a.v[i] ^ b.v[i];
  return t == 0;
}

// r = a - m if a >= m else a   (a < 2m)
template <class C>
HD void fe_reduce_once(uint32_t* a) {
  uint32_t t[8];
  t[0] = cc::sub_cc(a[0], C::mod(0));
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = cc::subc_cc(a[i], C::mod(i));
  uint32_t borrow = cc::subc(0, 0);  // 0 or 0xffffffff
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
}

template <class C>
HD Fe<C> fe_add(const Fe<C>& a, const Fe<C>& b) {
  Fe<C> r;
  r.v[0] = cc::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = cc::addc_cc(a.v[i], b.v[i]);
  // m < 2^254 so a+b < 2^255: no carry out
  fe_reduce_once<C>(r.v);
  return r;
}

// a + b without the conditional subtraction: result < 2m < 2^255.  Only valid as a multiplier operand:
// fe_mul(x, y) is fully reduced whenever x * y < 2^256 * m, which holds for x, y < 2m because 4m < 2^256.
template <class C>
HD Fe<C> fe_add_nr(const Fe<C>& a, const Fe<C>& b) {
  Fe<C> r;
  r.v[0] = cc::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = cc::addc_cc(a.v[i], b.v[i]);
  r.v[7] = cc::addc(a.v[7], b.v[7]);
  return r;
}

template <class C>
HD Fe<C> fe_sub(const Fe<C>& a, const Fe<C>& b) {
  Fe<C> r;
  r.v[0] = cc::sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = cc::subc_cc(a.v[i], b.v[i]);
  const uint32_t borrow = cc::subc(0, 0);
  if (borrow) {  // ptxas predicates these eight adds (16 instructions in all, against 25 with a masked add-back)
    r.v[0] = cc::add_cc(r.v[0], C::mod(0));
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = cc::addc_cc(r.v[i], C::mod(i));
    r.v[7] = cc::addc(r.v[7], C::mod(7));
  }
  return r;
}

// a / 2: (a + m) >> 1 when a is odd, a >> 1 otherwise (the representation -- Montgomery or plain -- does not matter)
template <class C>
HD Fe<C> fe_halve(const Fe<C>& a) {
  const uint32_t mask = 0u - (a.v[0] & 1u);
  uint32_t t[9];
  t[0] = cc::add_cc(a.v[0], C::mod(0) & mask);
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = cc::addc_cc(a.v[i], C::mod(i) & mask);
  t[8] = cc::addc(0u, 0u);
  Fe<C> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
  return r;
}

// m - b for b in [0, m): a value in (0, m] (m itself stands for 0; fine as an operand of fe_mul9_add / fe_mul)
template <class C>
HD Fe<C> fe_mod_minus(const Fe<C>& b) {
  Fe<C> r;
  r.v[0] = cc::sub_cc(C::mod(0), b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = cc::subc_cc(C::mod(i), b.v[i]);
  r.v[7] = cc::subc(C::mod(7), b.v[7]);
  return r;
}

// (9 x + y) mod m for x in [0, m), y in [0, m] -- the multiplication by xi = 9 + u needs exactly this twice.
// w = 9 x + y <= 10 m is formed as a 9-word integer with 16 small multiply-adds, the quotient is estimated from the top
// bits (q' = ((w >> 250) * 5416) >> 16 is floor(w / m) or one less for every w < 11 m; checked exhaustively over the
// 134 possible top values), and w - q' m < 2 m takes one conditional subtraction.  ~65 instructions instead of the
// ~120 of three doublings and two additions with their reductions.
template <class C>
HD Fe<C> fe_mul9_add(const Fe<C>& x, const Fe<C>& y) {
  uint32_t t[9], u[9];
  t[0] = cc::mad_lo_cc(x.v[0], 9u, y.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = cc::madc_lo_cc(x.v[i], 9u, y.v[i]);
  t[8] = cc::addc(0u, 0u);
  t[1] = cc::mad_hi_cc(x.v[0], 9u, t[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) t[i + 1] = cc::madc_hi_cc(x.v[i], 9u, t[i + 1]);
  t[8] = cc::madc_hi(x.v[7], 9u, t[8]);
  const uint32_t h = (t[8] << 6) | (t[7] >> 26);
  const uint32_t q = (h * 5416u) >> 16;
  u[0] = cc::mad_lo_cc(C::mod(0), q, 0u);
#pragma unroll
  for (int i = 1; i < 8; i++) u[i] = cc::madc_lo_cc(C::mod(i), q, 0u);
  u[8] = cc::addc(0u, 0u);
  u[1] = cc::mad_hi_cc(C::mod(0), q, u[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) u[i + 1] = cc::madc_hi_cc(C::mod(i), q, u[i + 1]);
  u[8] = cc::madc_hi(C::mod(7), q, u[8]);
  Fe<C> r;
  r.v[0] = cc::sub_cc(t[0], u[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = cc::subc_cc(t[i], u[i]);
  fe_reduce_once<C>(r.v);  // w - q' m < 2 m < 2^256: the ninth word is zero
  return r;
}

template <class C>
HD Fe<C> fe_neg(const Fe<C>& a) {
  return fe_sub(fe_zero<C>(), a);
}

template <class C>
HD Fe<C> fe_dbl(const Fe<C>& a) {
  return fe_add(a, a);
}

// Montgomery product a*b/2^256 mod m, fully reduced.
// W[s][q] holds the limb at absolute word position q of the accumulator whose 64-bit pairs start at
// positions of parity s.  Row i adds a_j*b_i at position i+j and q_i*m_j likewise; after the row the
// word at position i is zero.  All indices are compile-time after unrolling: no data movement.
template <class C>
HD Fe<C> fe_mul_inl(const Fe<C>& a, const Fe<C>& b) {
  uint32_t W[2][18];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t* Cw = W[i & 1];        // pairs aligned with position i
    uint32_t* Dw = W[(i & 1) ^ 1];  // pairs aligned with position i+1
    const uint32_t bi = b.v[i];
    if (i == 0) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint64_t t = (uint64_t)a.v[j] * bi;
        Cw[j] = (uint32_t)t;
        Cw[j + 1] = (uint32_t)(t >> 32);
        uint64_t u = (uint64_t)a.v[j + 1] * bi;
        Dw[j + 1] = (uint32_t)u;
        Dw[j + 2] = (uint32_t)(u >> 32);
      }
      Cw[8] = 0;
    } else {
      // fold the orphan word (high half of the pair whose low half was just cleared)
      Cw[i] = cc::add_cc(Cw[i], Dw[i]);
      // odd-j products, carry-in from the fold
#pragma unroll
      for (int j = 1; j < 8; j += 2) {
        Dw[i + j] = cc::madc_lo_cc(a.v[j], bi, Dw[i + j]);
        Dw[i + j + 1] = (j == 7) ? cc::madc_hi_cc(a.v[j], bi, 0u) : cc::madc_hi_cc(a.v[j], bi, Dw[i + j + 1]);
      }
      // even-j products
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        Cw[i + j] = (j == 0) ? cc::mad_lo_cc(a.v[j], bi, Cw[i + j]) : cc::madc_lo_cc(a.v[j], bi, Cw[i + j]);
        Cw[i + j + 1] = cc::madc_hi_cc(a.v[j], bi, Cw[i + j + 1]);
      }
      Cw[i + 8] = cc::addc(0u, 0u);
    }
    const uint32_t q = Cw[i] * C::inv;
#pragma unroll
    for (int j = 1; j < 8; j += 2) {
      Dw[i + j] = (j == 1) ? cc::mad_lo_cc(C::mod(j), q, Dw[i + j]) : cc::madc_lo_cc(C::mod(j), q, Dw[i + j]);
      Dw[i + j + 1] = cc::madc_hi_cc(C::mod(j), q, Dw[i + j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      Cw[i + j] = (j == 0) ? cc::mad_lo_cc(C::mod(j), q, Cw[i + j]) : cc::madc_lo_cc(C::mod(j), q, Cw[i + j]);
      Cw[i + j + 1] = cc::madc_hi_cc(C::mod(j), q, Cw[i + j + 1]);
    }
    Cw[i + 8] = cc::addc(Cw[i + 8], 0u);
  }
  // merge: positions 8..15 of both arrays (W[1] = C of the last row, W[0] = D)
  Fe<C> r;
  r.v[0] = cc::add_cc(W[0][8], W[1][8]);
#pragma unroll
  for (int k = 1; k < 8; k++) r.v[k] = cc::addc_cc(W[0][8 + k], W[1][8 + k]);
  fe_reduce_once<C>(r.v);
  return r;
}

// The one out-of-line copy of the multiplier per field.  Operands travel BY VALUE so that ptxas keeps
// them in registers across the call (a by-reference operand would be forced into local memory); the
// call costs ~20 register moves against ~190 instructions of multiplier, and keeps the pairing
// kernels' code small enough for the instruction cache.
#if !defined(__CUDA_ARCH__) && defined(BN254_COUNT_MULS)
// host-simulation builds only: counts 32x32+64 limb multiply-adds (the algorithmic work unit of DESIGN.md):
// 136 per Montgomery multiplication, 64 per wide product, 72 per wide reduction
inline unsigned long long& fe_mac_counter() {
  static thread_local unsigned long long c = 0;
  return c;
}
#define BN_COUNT_MACS(n) (fe_mac_counter() += (n))
#else
#define BN_COUNT_MACS(n)
#endif

// ---------------------------------------------------------------------------------------------
// Lazy reduction building blocks (used by the Fp2 multiplication): a full 512-bit product without reduction,
// and a Montgomery reduction of a 512-bit value.  Fp2 Karatsuba then needs 3 products + 2 reductions
// (3*64 + 2*72 = 336 multiply-adds) instead of 3 full Montgomery multiplications (408).
// ---------------------------------------------------------------------------------------------
// T[0..15] = a * b (plain integer product).  Products a_j * b_i whose position i+j is even accumulate in E, the others
// in O, so every 64-bit product lands on an aligned word pair of its accumulator (one IMAD.WIDE.U32 each); the
// carry out of a row's chain falls on a word that is either still zero or holds earlier carries (<= 2): no ripple.
template <class C>
HD void fe_mul_wide(uint32_t* T, const Fe<C>& a, const Fe<C>& b) {
  BN_COUNT_MACS(64);
  uint32_t E[17], O[17];
#pragma unroll
  for (int k = 0; k < 17; k++) E[k] = O[k] = 0;
  {
    const uint32_t b0 = b.v[0];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      uint64_t t = (uint64_t)a.v[j] * b0;
      E[j] = (uint32_t)t;
      E[j + 1] = (uint32_t)(t >> 32);
      uint64_t u = (uint64_t)a.v[j + 1] * b0;
      O[j + 1] = (uint32_t)u;
      O[j + 2] = (uint32_t)(u >> 32);
    }
  }
#pragma unroll
  for (int i = 1; i < 8; i++) {
    const uint32_t bi = b.v[i];
    const int je = i & 1;        // first j with i + j even
    const int jo = 1 - (i & 1);  // first j with i + j odd
#pragma unroll
    for (int j = je; j < 8; j += 2) {
      E[i + j] = (j == je) ? cc::mad_lo_cc(a.v[j], bi, E[i + j]) : cc::madc_lo_cc(a.v[j], bi, E[i + j]);
      E[i + j + 1] = cc::madc_hi_cc(a.v[j], bi, E[i + j + 1]);
    }
    E[i + je + 8] = cc::addc(E[i + je + 8], 0u);
#pragma unroll
    for (int j = jo; j < 8; j += 2) {
      O[i + j] = (j == jo) ? cc::mad_lo_cc(a.v[j], bi, O[i + j]) : cc::madc_lo_cc(a.v[j], bi, O[i + j]);
      O[i + j + 1] = cc::madc_hi_cc(a.v[j], bi, O[i + j + 1]);
    }
    O[i + jo + 8] = cc::addc(O[i + jo + 8], 0u);
  }
  T[0] = cc::add_cc(E[0], O[0]);
#pragma unroll
  for (int k = 1; k < 15; k++) T[k] = cc::addc_cc(E[k], O[k]);
  T[15] = cc::addc(E[15], O[15]);
}

// T (16 words, T < m * 2^256) -> T / 2^256 mod m, fully reduced.
//   REDC(T) = T_hi + (T_lo + Q m) / 2^256,  Q = -T_lo / m mod 2^256
// The second term is the word-serial reduction of the low half alone (the reduction rows of fe_mul_inl with the
// accumulator initialised to T_lo); it is <= m, so the sum is < 2m and one conditional subtraction finishes.
template <class C>
HD Fe<C> fe_redc_wide(const uint32_t* T) {
  BN_COUNT_MACS(72);
  uint32_t W[2][18];
#pragma unroll
  for (int k = 0; k < 18; k++) W[0][k] = W[1][k] = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) W[0][k] = T[k];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t* Cw = W[i & 1];
    uint32_t* Dw = W[(i & 1) ^ 1];
    if (i > 0) Cw[i] = cc::add_cc(Cw[i], Dw[i]);  // fold; the carry enters the odd chain below
    const uint32_t q = Cw[i] * C::inv;
#pragma unroll
    for (int j = 1; j < 8; j += 2) {
      Dw[i + j] = (j == 1 && i == 0) ? cc::mad_lo_cc(C::mod(j), q, Dw[i + j]) : cc::madc_lo_cc(C::mod(j), q, Dw[i + j]);
      Dw[i + j + 1] = cc::madc_hi_cc(C::mod(j), q, Dw[i + j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      Cw[i + j] = (j == 0) ? cc::mad_lo_cc(C::mod(j), q, Cw[i + j]) : cc::madc_lo_cc(C::mod(j), q, Cw[i + j]);
      Cw[i + j + 1] = cc::madc_hi_cc(C::mod(j), q, Cw[i + j + 1]);
    }
    Cw[i + 8] = cc::addc(Cw[i + 8], 0u);
  }
  // U = W[0][8..15] + W[1][8..15] (<= m), r = U + T_hi (< 2m)
  Fe<C> r;
  r.v[0] = cc::add_cc(W[0][8], W[1][8]);
#pragma unroll
  for (int k = 1; k < 8; k++) r.v[k] = cc::addc_cc(W[0][8 + k], W[1][8 + k]);
  r.v[0] = cc::add_cc(r.v[0], T[8]);
#pragma unroll
  for (int k = 1; k < 8; k++) r.v[k] = cc::addc_cc(r.v[k], T[8 + k]);
  fe_reduce_once<C>(r.v);
  return r;
}

// X -= Y over 16 words; returns the borrow mask (0 or 0xffffffff)
HD uint32_t wide_sub(uint32_t* X, const uint32_t* Y) {
  X[0] = cc::sub_cc(X[0], Y[0]);
#pragma unroll
  for (int k = 1; k < 16; k++) X[k] = cc::subc_cc(X[k], Y[k]);
  return cc::subc(0u, 0u);
}
// X += m << 256   (wraps modulo 2^512: cancels the borrow of a preceding wide_sub; call under `if (borrow)`)
template <class C>
HD void wide_add_mod_hi(uint32_t* X) {
  X[8] = cc::add_cc(X[8], C::mod(0));
#pragma unroll
  for (int k = 1; k < 7; k++) X[8 + k] = cc::addc_cc(X[8 + k], C::mod(k));
  X[15] = cc::addc(X[15], C::mod(7));
}

// X += Y over 16 words (the caller's bound keeps the sum below 2^512)
HD void wide_add(uint32_t* X, const uint32_t* Y) {
  X[0] = cc::add_cc(X[0], Y[0]);
#pragma unroll
  for (int k = 1; k < 15; k++) X[k] = cc::addc_cc(X[k], Y[k]);
  X[15] = cc::addc(X[15], Y[15]);
}
// a b + c d (mod m) with one reduction for both products: a b + c d < 2 m^2 < m 2^256 (200 multiply-adds and no
// modular addition, against 272 and one).  Same residue as fe_add(fe_mul(a, b), fe_mul(c, d)), fully reduced.
template <class C>
HDN Fe<C> fe_mul2_add(Fe<C> a, Fe<C> b, Fe<C> c, Fe<C> d) {
  uint32_t T0[16], T1[16];
  fe_mul_wide(T0, a, b);
  fe_mul_wide(T1, c, d);
  wide_add(T0, T1);
  return fe_redc_wide<C>(T0);
}

// ---------------------------------------------------------------------------------------------
// Signed 512-bit values (two's complement, 16 words) for lazy reduction one level up (Fq6): sums and differences of
// unreduced products are formed with wide_add / wide_sub and reduced once per output coefficient.
// ---------------------------------------------------------------------------------------------
// Montgomery reduction of a signed wide value V, -m 2^256 < V < K m 2^256 (K = 1, or 2 with TWICE): V / 2^256 mod m,
// fully reduced.  A negative V gets m 2^256 added first (same residue); fe_redc_wide's U + T_hi is then < (K + 1) m.
template <class C, bool TWICE>
HD Fe<C> fe_redc_wide_signed(uint32_t* T) {
  if (T[15] >> 31) wide_add_mod_hi<C>(T);
  Fe<C> r = fe_redc_wide<C>(T);
  if (TWICE) fe_reduce_once<C>(r.v);
  return r;
}

// Z = 9 X + Y (SUB = false) or 9 X - Y (SUB = true) modulo m 2^256, as a signed value with |Z| <= (1/2 + 1e-4) m 2^256.
// X, Y signed wide values with -4.2 m 2^256 < 9 X +- Y < 7.2 m 2^256 (the two uses in the Fq6 multiplication).
// The 17-word sum A gets 5 m 2^256 added (A' > 0), q = round(A' / (m 2^256)) is estimated from the top 26 bits of A'
// (a' = A' >> 488, q = (a' * 43337 + 2^36) >> 37: m / 2^232 = 3171406.45 and 2^37 / 3171406.45 = 43336.9, so the
// estimate is off by < 1e-4 before rounding), and q m is subtracted from the high half.  BN254 base field only.
template <class C, bool SUB>
HD void wide_mul9_addsub(uint32_t* Z, const uint32_t* X, const uint32_t* Y) {
  constexpr uint32_t FIVE_M[8] = {0x3a70f263u, 0x2ca2bc72u, 0x0a38f4c2u, 0xf58714d7u,
                                  0x8786b9d3u, 0x99915c90u, 0x65f820d0u, 0xf1f5883eu};
  uint32_t t[17];
  const uint32_t nx = X[15] >> 31, ny = Y[15] >> 31;
  if (SUB) {
    t[0] = X[0] * 9u;
    t[1] = cc::mad_hi_cc(X[0], 9u, 0u);
#pragma unroll
    for (int i = 1; i < 15; i++) t[i + 1] = cc::madc_hi_cc(X[i], 9u, 0u);
    t[16] = cc::madc_hi(X[15], 9u, 0u);
    t[1] = cc::mad_lo_cc(X[1], 9u, t[1]);
#pragma unroll
    for (int i = 2; i < 16; i++) t[i] = cc::madc_lo_cc(X[i], 9u, t[i]);
    t[16] = cc::addc(t[16], 0u);
    t[0] = cc::sub_cc(t[0], Y[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) t[i] = cc::subc_cc(t[i], Y[i]);
    t[16] = cc::subc(t[16], 0u);
    t[16] = t[16] - 9u * nx + ny;  // sign extensions of X and Y (unsigned words stood for value + 2^512)
  } else {
    t[0] = cc::mad_lo_cc(X[0], 9u, Y[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) t[i] = cc::madc_lo_cc(X[i], 9u, Y[i]);
    t[16] = cc::addc(0u, 0u);
    t[1] = cc::mad_hi_cc(X[0], 9u, t[1]);
#pragma unroll
    for (int i = 1; i < 15; i++) t[i + 1] = cc::madc_hi_cc(X[i], 9u, t[i + 1]);
    t[16] = cc::madc_hi(X[15], 9u, t[16]);
    t[16] = t[16] - 9u * nx - ny;
  }
  // A' = A + 5 m 2^256
  t[8] = cc::add_cc(t[8], FIVE_M[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) t[8 + i] = cc::addc_cc(t[8 + i], FIVE_M[i]);
  t[16] = cc::addc(t[16], 0u);
  const uint32_t ap = (t[16] << 24) | (t[15] >> 8);
  const uint32_t q = (uint32_t)(((uint64_t)ap * 43337u + (1ull << 36)) >> 37);
  uint32_t u[9];
  u[0] = cc::mad_lo_cc(C::mod(0), q, 0u);
#pragma unroll
  for (int i = 1; i < 8; i++) u[i] = cc::madc_lo_cc(C::mod(i), q, 0u);
  u[8] = cc::addc(0u, 0u);
  u[1] = cc::mad_hi_cc(C::mod(0), q, u[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) u[i + 1] = cc::madc_hi_cc(C::mod(i), q, u[i + 1]);
  u[8] = cc::madc_hi(C::mod(7), q, u[8]);
#pragma unroll
  for (int i = 0; i < 8; i++) Z[i] = t[i];
  Z[8] = cc::sub_cc(t[8], u[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) Z[8 + i] = cc::subc_cc(t[8 + i], u[i]);
  Z[15] = cc::subc(t[15], u[7]);  // word 16 (t[16] - u[8] - borrow) is the sign extension of word 15
}

template <class C>
HDN Fe<C> fe_mul(Fe<C> a, Fe<C> b) {
  BN_COUNT_MACS(136);
  return fe_mul_inl(a, b);
}

template <class C>
HD Fe<C> fe_sqr(const Fe<C>& a) {
  return fe_mul(a, a);
}

// T[0..15] = a * a.  The 28 products a_i a_j, i < j, accumulate as in fe_mul_wide (position i + j even: E, odd: O; every
// carry word is written before any product reaches it), their sum is doubled with funnel shifts and the 8 squares
// a_i^2 go onto the aligned pairs (2i, 2i + 1) in one carry chain: 36 multiply-adds instead of 64.
template <class C>
HD void fe_sqr_wide(uint32_t* T, const Fe<C>& a) {
  BN_COUNT_MACS(36);
  uint32_t E[16], O[16];
#pragma unroll
  for (int k = 0; k < 16; k++) E[k] = O[k] = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const uint32_t ai = a.v[i];
    // j = i + 2, i + 4, ..: position i + j even
    if (i + 2 < 8) {
#pragma unroll
      for (int j = i + 2; j < 8; j += 2) {
        E[i + j] = (j == i + 2) ? cc::mad_lo_cc(a.v[j], ai, E[i + j]) : cc::madc_lo_cc(a.v[j], ai, E[i + j]);
        E[i + j + 1] = cc::madc_hi_cc(a.v[j], ai, E[i + j + 1]);
      }
      const int top = 2 * i + ((7 - i) & ~1);  // position of the last pair of the chain
      if (top + 2 < 16) E[top + 2] = cc::addc(E[top + 2], 0u);
    }
    // j = i + 1, i + 3, ..: position i + j odd
#pragma unroll
    for (int j = i + 1; j < 8; j += 2) {
      O[i + j] = (j == i + 1) ? cc::mad_lo_cc(a.v[j], ai, O[i + j]) : cc::madc_lo_cc(a.v[j], ai, O[i + j]);
      O[i + j + 1] = cc::madc_hi_cc(a.v[j], ai, O[i + j + 1]);
    }
    {
      const int top = 2 * i + 1 + ((6 - i) & ~1);
      if (top + 2 < 16) O[top + 2] = cc::addc(O[top + 2], 0u);
    }
  }
  uint32_t S[16];
  S[0] = 0;
  S[1] = cc::add_cc(E[1], O[1]);
#pragma unroll
  for (int k = 2; k < 15; k++) S[k] = cc::addc_cc(E[k], O[k]);
  S[15] = cc::addc(E[15], O[15]);
  // T = 2 S (S < 2^511), then + sum a_i^2 2^(64 i)
  T[0] = 0;
#pragma unroll
  for (int k = 1; k < 16; k++) T[k] = (S[k] << 1) | (S[k - 1] >> 31);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    T[2 * i] = (i == 0) ? cc::mad_lo_cc(a.v[i], a.v[i], T[2 * i]) : cc::madc_lo_cc(a.v[i], a.v[i], T[2 * i]);
    T[2 * i + 1] = (i == 7) ? cc::madc_hi(a.v[i], a.v[i], T[2 * i + 1]) : cc::madc_hi_cc(a.v[i], a.v[i], T[2 * i + 1]);
  }
}
// a^2 / 2^256 mod m, fully reduced, for a < 2m (a^2 < m 2^256): 108 multiply-adds against the 136 of fe_mul(a, a).
// Its own out-of-line copy: used by the G1 arithmetic (5 of the 7 multiplications of a doubling are squarings).
template <class C>
HDN Fe<C> fe_sqr_short(Fe<C> a) {
  uint32_t T[16];
  fe_sqr_wide(T, a);
  return fe_redc_wide<C>(T);
}

template <class C>
HD Fe<C> fe_to_mont(const Fe<C>& a) {
  Fe<C> r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.v[i] = C::r2(i);
  return fe_mul(a, r2);
}
template <class C>
HD Fe<C> fe_from_mont(const Fe<C>& a) {
  Fe<C> one = fe_zero<C>();
  one.v[0] = 1;
  return fe_mul(a, one);
}

// a^e for a 256-bit exponent given as 8 LE words (plain integer, not Montgomery), MSB first.
template <class C>
HDN Fe<C> fe_pow_words(const Fe<C>& a, const uint32_t* e) {
  Fe<C> r = fe_one<C>();
  for (int i = 255; i >= 0; i--) {
    r = fe_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) r = fe_mul(r, a);
  }
  return r;
}

// Fermat inverse a^(m-2); returns zero for zero.  (Kept as the cross-check of fe_inv: tests/hostsim.)
template <class C>
HD Fe<C> fe_inv_fermat(const Fe<C>& a) {
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = C::mod(i);
  e[0] -= 2;  // low limb of both moduli is >= 2
  return fe_pow_words(a, e);
}

// ---- modular inverse by divsteps (Bernstein, Yang: "Fast constant-time gcd computation and modular inversion", 2019),
// in the batched form with signed 30-bit limbs: 20 rounds of 30 divsteps on the low limbs of (f, g) give a 2 x 2
// transition matrix t / 2^30 that is then applied to the full (f, g) -- exactly -- and to the Bezout pair (d, e) modulo m.
// 600 >= 590 divsteps suffice for 256-bit inputs.  Data-independent control flow (no divergence inside a warp), ~2 000
// multiply-adds and ~15 000 other integer instructions instead of the 380 multiplications (52 000 multiply-adds,
// 150 000 instructions) of the Fermat chain; the inverse is unique, so every value downstream is unchanged.
struct Signed30 {
  int32_t v[9];  // value = sum v[i] 2^(30 i); limbs 0..7 in [0, 2^30), limb 8 signed
};
struct Trans30 {
  int32_t u, v, q, r;
};
// 30 divsteps on the low 30 bits of f and g.  zeta = -(delta + 1/2).  On return [f; g] <- (1 / 2^30) [[u v]; [q r]] [f; g].
HD int32_t divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, Trans30& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 1
  for (int i = 0; i < 30; i++) {
    uint32_t m1 = (uint32_t)(zeta >> 31);  // zeta < 0
    const uint32_t m2 = 0u - (g & 1u);     // g odd
    const uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;  // (f, u, v) negated when zeta < 0
    g += x & m2, q += y & m2, r += z & m2;
    m1 &= m2;
    zeta = (int32_t)((uint32_t)zeta ^ m1) - 1;  // -zeta - 2 or zeta - 1
    f += g & m1, u += q & m1, v += r & m1;
    g >>= 1, u <<= 1, v <<= 1;
  }
  t.u = (int32_t)u, t.v = (int32_t)v, t.q = (int32_t)q, t.r = (int32_t)r;
  return zeta;
}
// (f, g) <- t / 2^30 (f, g): the bottom 30 bits of both combinations are zero by construction
HD void update_fg_30(Signed30& f, Signed30& g, const Trans30& t) {
  const int32_t M30 = 0x3fffffff;
  int64_t cf = (int64_t)t.u * f.v[0] + (int64_t)t.v * g.v[0];
  int64_t cg = (int64_t)t.q * f.v[0] + (int64_t)t.r * g.v[0];
  cf >>= 30, cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    const int32_t fi = f.v[i], gi = g.v[i];
    cf += (int64_t)t.u * fi + (int64_t)t.v * gi;
    cg += (int64_t)t.q * fi + (int64_t)t.r * gi;
    f.v[i - 1] = (int32_t)cf & M30, cf >>= 30;
    g.v[i - 1] = (int32_t)cg & M30, cg >>= 30;
  }
  f.v[8] = (int32_t)cf, g.v[8] = (int32_t)cg;
}
// (d, e) <- t / 2^30 (d, e) mod m, d and e in (-2m, m): a multiple of m is added that clears the bottom 30 bits
template <class C>
HD void update_de_30(Signed30& d, Signed30& e, const Trans30& t, const Signed30& m, uint32_t m_inv30) {
  const int32_t M30 = 0x3fffffff;
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
  int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
  int64_t cd = (int64_t)t.u * d.v[0] + (int64_t)t.v * e.v[0];
  int64_t ce = (int64_t)t.q * d.v[0] + (int64_t)t.r * e.v[0];
  md -= (int32_t)((m_inv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((m_inv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)m.v[0] * md, ce += (int64_t)m.v[0] * me;
  cd >>= 30, ce >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    const int32_t di = d.v[i], ei = e.v[i];
    cd += (int64_t)t.u * di + (int64_t)t.v * ei + (int64_t)m.v[i] * md;
    ce += (int64_t)t.q * di + (int64_t)t.r * ei + (int64_t)m.v[i] * me;
    d.v[i - 1] = (int32_t)cd & M30, cd >>= 30;
    e.v[i - 1] = (int32_t)ce & M30, ce >>= 30;
  }
  d.v[8] = (int32_t)cd, e.v[8] = (int32_t)ce;
}
HD Signed30 signed30_from_words(const uint32_t* w) {  // 8 x 32 -> 9 x 30 (value < 2^256)
  Signed30 r;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
    uint32_t x = w[k] >> sh;
    if (sh > 2 && k + 1 < 8) x |= w[k + 1] << (32 - sh);
    r.v[i] = (int32_t)(i < 8 ? (x & 0x3fffffffu) : x);
  }
  return r;
}
// r in (-2m, m) -> sign-adjusted (negated when `sign` < 0) and brought into [0, m); then back to 8 x 32
template <class C>
HD void signed30_normalize_to_words(uint32_t* w, Signed30 r, int32_t sign, const Signed30& m) {
  const int32_t M30 = 0x3fffffff;
  // r < 0: add m;  then negate if sign < 0;  carry;  r < 0 again (only after the negation): add m
  int32_t cond_add = r.v[8] >> 31;
  const int32_t cond_negate = sign >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] = ((r.v[i] + (m.v[i] & cond_add)) ^ cond_negate) - cond_negate;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i + 1] += r.v[i] >> 30, r.v[i] &= M30;
  cond_add = r.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] += m.v[i] & cond_add;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i + 1] += r.v[i] >> 30, r.v[i] &= M30;
#pragma unroll
  for (int k = 0; k < 8; k++) {  // word k = bits [32k, 32k + 32)
    const int i = (32 * k) / 30, sh = 32 * k - 30 * i;
    uint32_t x = (uint32_t)r.v[i] >> sh;
    if (i + 1 < 9) x |= (uint32_t)r.v[i + 1] << (30 - sh);
    if (i + 2 < 9 && 60 - sh < 32) x |= (uint32_t)r.v[i + 2] << (60 - sh);
    w[k] = x;
  }
}
// 1 / a in Montgomery form (a in Montgomery form, fully reduced); zero for zero.
template <class C>
HDN Fe<C> fe_inv(const Fe<C>& a) {
  BN_COUNT_MACS(20 * 92);
  uint32_t mw[8];
#pragma unroll
  for (int i = 0; i < 8; i++) mw[i] = C::mod(i);
  const Signed30 m = signed30_from_words(mw);
  const uint32_t m_inv30 = (0u - C::inv) & 0x3fffffffu;  // m^-1 mod 2^30 (C::inv = -m^-1 mod 2^32)
  Signed30 d, e, f = m, g = signed30_from_words(a.v);
#pragma unroll
  for (int i = 0; i < 9; i++) d.v[i] = 0, e.v[i] = 0;
  e.v[0] = 1;
  int32_t zeta = -1;
#pragma unroll 1
  for (int it = 0; it < 20; it++) {
    Trans30 t;
    zeta = divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    update_de_30<C>(d, e, t, m, m_inv30);
    update_fg_30(f, g, t);
  }
  // g = 0 and f = +-gcd: d = +-(a R)^-1 (plain integer).  Back to Montgomery form: times R^3, as a Montgomery product.
  Fe<C> x;
  signed30_normalize_to_words<C>(x.v, d, f.v[8], m);
  Fe<C> r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.v[i] = C::r2(i);
  return fe_mul(x, fe_mul(r2, r2));  // (R^2 R^2 / R) = R^3;  x R^3 / R = a^-1 R
}

// 32 big-endian bytes -> 8 little-endian words.  On the device a record that is 16-byte aligned (every record of a
// 256-byte-stride Groth16 batch, the points of a pairing-product batch) is read with two 128-bit loads and one byte
// permutation per word, a 4-byte aligned one (ragged / 904-byte PlonK records) with eight 32-bit loads, instead of
// 32 byte loads and 24 shift-or pairs.
HD void be32_to_words(uint32_t* w, const uint8_t* b) {
#if defined(__CUDA_ARCH__)
  const uintptr_t a = (uintptr_t)b;
  if ((a & 15) == 0) {
    const uint4 hi = *(const uint4*)b, lo = *(const uint4*)(b + 16);  // hi.x holds bytes 0..3: the top word
    w[7] = __byte_perm(hi.x, 0, 0x0123), w[6] = __byte_perm(hi.y, 0, 0x0123);
    w[5] = __byte_perm(hi.z, 0, 0x0123), w[4] = __byte_perm(hi.w, 0, 0x0123);
    w[3] = __byte_perm(lo.x, 0, 0x0123), w[2] = __byte_perm(lo.y, 0, 0x0123);
    w[1] = __byte_perm(lo.z, 0, 0x0123), w[0] = __byte_perm(lo.w, 0, 0x0123);
    return;
  }
  if ((a & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = __byte_perm(*(const uint32_t*)(b + 28 - 4 * i), 0, 0x0123);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* p = b + 28 - 4 * i;
    w[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
  }
}

// 32-byte big-endian -> plain limbs; returns false when the value is >= m (Fq::from_slice semantics).
template <class C>
HD bool fe_from_be_bytes(Fe<C>& out, const uint8_t* b) {
  be32_to_words(out.v, b);
  // out < m ?
  cc::sub_cc(out.v[0], C::mod(0));
  uint32_t d = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) d = cc::subc_cc(out.v[i], C::mod(i));
  (void)d;
  uint32_t borrow = cc::subc(0, 0);
  return borrow != 0;
}

template <class C>
HD void fe_to_be_bytes(uint8_t* b, const Fe<C>& a) {  // a: plain limbs
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* p = b + 28 - 4 * i;
    p[0] = (uint8_t)(a.v[i] >> 24);
    p[1] = (uint8_t)(a.v[i] >> 16);
    p[2] = (uint8_t)(a.v[i] >> 8);
    p[3] = (uint8_t)a.v[i];
  }
}

// Reduce an arbitrary 256-bit plain value mod m (value < 2^256 < 6m): at most 5 subtractions.
template <class C>
HD void fe_reduce_full(Fe<C>& a) {
  for (int k = 0; k < 5; k++) {
    uint32_t t[8];
    t[0] = cc::sub_cc(a.v[0], C::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = cc::subc_cc(a.v[i], C::mod(i));
    uint32_t borrow = cc::subc(0, 0);
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = borrow ? a.v[i] : t[i];
  }
}

typedef Fe<FpCfg> Fp;
typedef Fe<FrCfg> Fr;

}  // namespace bn254
