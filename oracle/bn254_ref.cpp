// CPU ORACLE / CPU BASELINE (test infrastructure only -- never linked into or called by the product path).
//
// C++ restatement of the reference algorithm (succinctlabs/snark-bn254-verifier) IN THE REFERENCE'S SHAPE,
// used (a) as the fast second oracle for parity at sizes the Python oracle cannot reach and (b) as the
// `cpu_baseline` / `--impl reference` arm of bench.py.  The Rust crate itself cannot be built here: there is
// no Rust toolchain and its arithmetic dependency substrate-bn 0.7.0 (git sp1-patches/bn @ 3c53d256...,
// reference Cargo.lock:405-407) is not on disk; its published algorithms (libff alt_bn128 / paritytech bn,
// SURVEY.md Appendix B) are restated below.  PARITY PINNING: this file is pinned against the Python oracle
// (oracle/bn254_oracle.py, itself pinned on the 4 bundled PlonK fixtures and Appendix C values) by
// tests/test_ref_cpu.py on the committed golden vectors; Groth16 verdicts / raw Miller values have no
// reference-held vector ("parity unpinned" by the reference itself).
//
// What follows the reference, per call (no hoisting, no caching -- that is what the crate does):
//   Groth16Verifier::verify            verifier/src/lib.rs:44-49
//     load_groth16_proof_from_bytes    verifier/src/groth16/converter.rs:14-26  (on-curve, G2 subgroup check)
//     load_groth16_verifying_key_...   verifier/src/groth16/converter.rs:28-89  (point decompression: sqrt)
//     verify_groth16                   verifier/src/groth16/verify.rs:65-78     pairing(alpha,-beta) + 3-pair batch
//     prepare_inputs                   verifier/src/groth16/verify.rs:53-63     naive double-and-add
// Independent code: 4 x 64-bit limbs with unsigned __int128 (the CUDA path uses 8 x 32-bit limbs and PTX
// carry chains); nothing is shared with snark-bn254-verifier_b200/csrc.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <utility>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

#ifdef REF_COUNT
static thread_local u64 g_fp_mul_count = 0;
#define COUNT_MUL() (g_fp_mul_count++)
#else
#define COUNT_MUL()
#endif

// ------------------------------------------------------------------------------------------ fields
struct Mod {
  u64 m[4], r1[4], r2[4], inv;
};
static bool geq(const u64* a, const u64* b) {
  for (int i = 3; i >= 0; i--)
    if (a[i] != b[i]) return a[i] > b[i];
  return true;
}
static u64 sub_n(u64* r, const u64* a, const u64* b) {
  u64 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - b[i] - br;
    r[i] = (u64)t;
    br = (u64)(t >> 64) & 1;
  }
  return br;
}
static u64 add_n(u64* r, const u64* a, const u64* b) {
  u64 c = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  return c;
}
static Mod make_mod(const u64 m[4]) {
  Mod M;
  memcpy(M.m, m, 32);
  u64 x = 1;  // -m^-1 mod 2^64 by Newton iteration
  for (int i = 0; i < 6; i++) x *= 2 - m[0] * x;
  M.inv = (u64)0 - x;
  u64 t[4] = {1, 0, 0, 0};
  for (int i = 0; i < 512; i++) {  // t = 2^i mod m
    u64 c = add_n(t, t, t);
    if (c || geq(t, m)) sub_n(t, t, m);
    if (i == 255) memcpy(M.r1, t, 32);
  }
  memcpy(M.r2, t, 32);
  return M;
}
static const u64 P_LIMBS[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 R_LIMBS[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const Mod MP = make_mod(P_LIMBS), MR = make_mod(R_LIMBS);

// CIOS Montgomery product with the modulus as compile-time constants (fully unrolled by the compiler).
template <u64 M0, u64 M1, u64 M2, u64 M3, u64 INV>
static inline __attribute__((always_inline)) void mont_mul_c(u64* r, const u64* a, const u64* b) {
  const u64 m[4] = {M0, M1, M2, M3};
  u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
#pragma GCC unroll 4
  for (int i = 0; i < 4; i++) {
    const u64 bi = b[i];
    u128 x = (u128)a[0] * bi + t0;
    t0 = (u64)x;
    x = (u128)a[1] * bi + t1 + (u64)(x >> 64);
    t1 = (u64)x;
    x = (u128)a[2] * bi + t2 + (u64)(x >> 64);
    t2 = (u64)x;
    x = (u128)a[3] * bi + t3 + (u64)(x >> 64);
    t3 = (u64)x;
    u128 y = (u128)t4 + (u64)(x >> 64);
    t4 = (u64)y;
    u64 t5 = (u64)(y >> 64);
    const u64 q = t0 * INV;
    x = (u128)q * m[0] + t0;
    x = (u128)q * m[1] + t1 + (u64)(x >> 64);
    t0 = (u64)x;
    x = (u128)q * m[2] + t2 + (u64)(x >> 64);
    t1 = (u64)x;
    x = (u128)q * m[3] + t3 + (u64)(x >> 64);
    t2 = (u64)x;
    y = (u128)t4 + (u64)(x >> 64);
    t3 = (u64)y;
    t4 = t5 + (u64)(y >> 64);
  }
  u64 t[4] = {t0, t1, t2, t3};
  if (t4 || geq(t, m)) sub_n(t, t, m);
  r[0] = t[0], r[1] = t[1], r[2] = t[2], r[3] = t[3];
}
static inline void mont_mul(u64* r, const u64* a, const u64* b, const Mod& M) {
  if (&M == &MP)
    mont_mul_c<0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull,
               0x87d20782e4866389ull>(r, a, b);
  else
    mont_mul_c<0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull,
               0xc2e1f593efffffffull>(r, a, b);
}

struct Fp {
  u64 v[4];
};
static inline Fp fp_zero() { return Fp{{0, 0, 0, 0}}; }
static inline Fp fp_one() {
  Fp r;
  memcpy(r.v, MP.r1, 32);
  return r;
}
static inline Fp operator+(const Fp& a, const Fp& b) {
  Fp r;
  u64 c = add_n(r.v, a.v, b.v);
  if (c || geq(r.v, MP.m)) sub_n(r.v, r.v, MP.m);
  return r;
}
static inline Fp operator-(const Fp& a, const Fp& b) {
  Fp r;
  if (sub_n(r.v, a.v, b.v)) add_n(r.v, r.v, MP.m);
  return r;
}
static inline Fp operator-(const Fp& a) { return fp_zero() - a; }
static inline Fp operator*(const Fp& a, const Fp& b) {
  COUNT_MUL();
  Fp r;
  mont_mul(r.v, a.v, b.v, MP);
  return r;
}
static inline bool operator==(const Fp& a, const Fp& b) { return memcmp(a.v, b.v, 32) == 0; }
static inline bool is_zero(const Fp& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static inline Fp dbl(const Fp& a) { return a + a; }
static Fp fp_pow(const Fp& a, const u64 e[4]) {
  Fp r = fp_one();
  for (int i = 255; i >= 0; i--) {
    r = r * r;
    if ((e[i >> 6] >> (i & 63)) & 1) r = r * a;
  }
  return r;
}
static Fp fp_inv(const Fp& a) {
  u64 e[4];
  memcpy(e, MP.m, 32);
  e[0] -= 2;
  return fp_pow(a, e);
}
static Fp fp_from_u64(u64 x) {
  Fp t{{x, 0, 0, 0}}, r2, r;
  memcpy(r2.v, MP.r2, 32);
  mont_mul(r.v, t.v, r2.v, MP);
  return r;
}
// 32-byte big-endian -> canonical limbs; false when >= m
static bool limbs_from_be(u64* out, const uint8_t* b, const Mod& M) {
  for (int i = 0; i < 4; i++) {
    u64 w = 0;
    for (int k = 0; k < 8; k++) w = (w << 8) | b[(3 - i) * 8 + k];
    out[i] = w;
  }
  return !geq(out, M.m);
}
static void limbs_to_be(uint8_t* b, const u64* v) {
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 8; k++) b[(3 - i) * 8 + k] = (uint8_t)(v[i] >> (56 - 8 * k));
}
static bool fp_from_be(Fp& out, const uint8_t* b) {  // Fq::from_slice
  u64 t[4];
  bool ok = limbs_from_be(t, b, MP);
  mont_mul(out.v, t, MP.r2, MP);
  return ok;
}
static void fp_to_be(uint8_t* b, const Fp& a) {
  u64 one[4] = {1, 0, 0, 0}, t[4];
  mont_mul(t, a.v, one, MP);
  limbs_to_be(b, t);
}
static bool fp_sqrt(Fp& out, const Fp& a) {  // p = 3 mod 4: a^((p+1)/4)
  u64 e[4], one[4] = {1, 0, 0, 0};
  add_n(e, MP.m, one);
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
    e[3] >>= 1;
  }
  Fp y = fp_pow(a, e);
  if (!(y * y == a)) return false;
  out = y;
  return true;
}
static int fp_cmp(const Fp& a, const Fp& b) {  // canonical integer order
  u64 one[4] = {1, 0, 0, 0}, x[4], y[4];
  mont_mul(x, a.v, one, MP);
  mont_mul(y, b.v, one, MP);
  for (int i = 3; i >= 0; i--)
    if (x[i] != y[i]) return x[i] > y[i] ? 1 : -1;
  return 0;
}

// Fr: only what the Groth16 path and the workload generator need (plain canonical limbs for scalars)
struct Fr {
  u64 v[4];  // Montgomery
};
static Fr fr_mul(const Fr& a, const Fr& b) {
  Fr r;
  mont_mul(r.v, a.v, b.v, MR);
  return r;
}
static Fr fr_add(const Fr& a, const Fr& b) {
  Fr r;
  u64 c = add_n(r.v, a.v, b.v);
  if (c || geq(r.v, MR.m)) sub_n(r.v, r.v, MR.m);
  return r;
}
static Fr fr_sub(const Fr& a, const Fr& b) {
  Fr r;
  if (sub_n(r.v, a.v, b.v)) add_n(r.v, r.v, MR.m);
  return r;
}
static Fr fr_from_plain(const u64* p) {
  Fr r;
  mont_mul(r.v, p, MR.r2, MR);
  return r;
}
static void fr_to_plain(u64* p, const Fr& a) {
  u64 one[4] = {1, 0, 0, 0};
  mont_mul(p, a.v, one, MR);
}
static Fr fr_inv(const Fr& a) {
  u64 e[4];
  memcpy(e, MR.m, 32);
  e[0] -= 2;
  Fr r;
  memcpy(r.v, MR.r1, 32);
  for (int i = 255; i >= 0; i--) {
    r = fr_mul(r, r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = fr_mul(r, a);
  }
  return r;
}

// ------------------------------------------------------------------------------------------ tower
struct Fp2 {
  Fp c0, c1;
};
static inline Fp2 operator+(const Fp2& a, const Fp2& b) { return Fp2{a.c0 + b.c0, a.c1 + b.c1}; }
static inline Fp2 operator-(const Fp2& a, const Fp2& b) { return Fp2{a.c0 - b.c0, a.c1 - b.c1}; }
static inline Fp2 operator-(const Fp2& a) { return Fp2{-a.c0, -a.c1}; }
static inline Fp2 operator*(const Fp2& a, const Fp2& b) {
  Fp aa = a.c0 * b.c0, bb = a.c1 * b.c1;
  return Fp2{aa - bb, (a.c0 + a.c1) * (b.c0 + b.c1) - aa - bb};
}
static inline Fp2 sqr(const Fp2& a) {
  Fp ab = a.c0 * a.c1;
  return Fp2{(a.c0 + a.c1) * (a.c0 - a.c1), ab + ab};
}
static inline Fp2 scale(const Fp2& a, const Fp& k) { return Fp2{a.c0 * k, a.c1 * k}; }
static inline Fp2 dbl(const Fp2& a) { return a + a; }
static inline Fp2 conj(const Fp2& a) { return Fp2{a.c0, -a.c1}; }
static inline bool operator==(const Fp2& a, const Fp2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
static inline bool is_zero(const Fp2& a) { return is_zero(a.c0) && is_zero(a.c1); }
static inline Fp2 fp2_zero() { return Fp2{fp_zero(), fp_zero()}; }
static inline Fp2 fp2_one() { return Fp2{fp_one(), fp_zero()}; }
static inline Fp2 mul_xi(const Fp2& a) {  // (9 + u) a
  Fp e0 = dbl(dbl(dbl(a.c0))), e1 = dbl(dbl(dbl(a.c1)));
  return Fp2{e0 + a.c0 - a.c1, e1 + a.c1 + a.c0};
}
static Fp2 inv(const Fp2& a) {
  Fp n = fp_inv(a.c0 * a.c0 + a.c1 * a.c1);
  return Fp2{a.c0 * n, -(a.c1 * n)};
}
static Fp2 fp2_pow(const Fp2& a, const std::vector<u64>& e) {
  Fp2 r = fp2_one();
  for (int i = (int)e.size() * 64 - 1; i >= 0; i--) {
    r = sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = r * a;
  }
  return r;
}

struct Fp6 {
  Fp2 c0, c1, c2;
};
static inline Fp6 operator+(const Fp6& a, const Fp6& b) { return Fp6{a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
static inline Fp6 operator-(const Fp6& a, const Fp6& b) { return Fp6{a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
static inline Fp6 operator-(const Fp6& a) { return Fp6{-a.c0, -a.c1, -a.c2}; }
static inline Fp6 mul_v(const Fp6& a) { return Fp6{mul_xi(a.c2), a.c0, a.c1}; }
static Fp6 operator*(const Fp6& a, const Fp6& b) {
  Fp2 aa = a.c0 * b.c0, bb = a.c1 * b.c1, cc = a.c2 * b.c2;
  return Fp6{mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - bb - cc) + aa, (a.c0 + a.c1) * (b.c0 + b.c1) - aa - bb + mul_xi(cc),
             (a.c0 + a.c2) * (b.c0 + b.c2) - aa + bb - cc};
}
static Fp6 sqr(const Fp6& a) {
  Fp2 s0 = sqr(a.c0), ab = a.c0 * a.c1, s1 = ab + ab, s2 = sqr(a.c0 - a.c1 + a.c2), bc = a.c1 * a.c2, s3 = bc + bc,
      s4 = sqr(a.c2);
  return Fp6{s0 + mul_xi(s3), s1 + mul_xi(s4), s1 + s2 + s3 - s0 - s4};
}
static Fp6 inv(const Fp6& a) {
  Fp2 c0 = sqr(a.c0) - mul_xi(a.c1 * a.c2), c1 = mul_xi(sqr(a.c2)) - a.c0 * a.c1, c2 = sqr(a.c1) - a.c0 * a.c2;
  Fp2 t = inv(mul_xi(a.c2 * c1 + a.c1 * c2) + a.c0 * c0);
  return Fp6{t * c0, t * c1, t * c2};
}
static inline bool operator==(const Fp6& a, const Fp6& b) { return a.c0 == b.c0 && a.c1 == b.c1 && a.c2 == b.c2; }

struct Fp12 {
  Fp6 c0, c1;
};
static inline Fp12 fp12_one() {
  return Fp12{Fp6{fp2_one(), fp2_zero(), fp2_zero()}, Fp6{fp2_zero(), fp2_zero(), fp2_zero()}};
}
static Fp12 operator*(const Fp12& a, const Fp12& b) {
  Fp6 aa = a.c0 * b.c0, bb = a.c1 * b.c1;
  return Fp12{mul_v(bb) + aa, (a.c0 + a.c1) * (b.c0 + b.c1) - aa - bb};
}
static Fp12 sqr(const Fp12& a) {
  Fp6 ab = a.c0 * a.c1;
  return Fp12{(mul_v(a.c1) + a.c0) * (a.c0 + a.c1) - ab - mul_v(ab), ab + ab};
}
static inline Fp12 conj(const Fp12& a) { return Fp12{a.c0, -a.c1}; }
static Fp12 inv(const Fp12& a) {
  Fp6 t = inv(sqr(a.c0) - mul_v(sqr(a.c1)));
  return Fp12{a.c0 * t, -(a.c1 * t)};
}
static inline bool operator==(const Fp12& a, const Fp12& b) { return a.c0 == b.c0 && a.c1 == b.c1; }

// ---- constants derived at start-up from p and xi (no tables copied from anywhere)
struct Consts {
  Fp two_inv, three;
  Fp2 b2;                 // 3 / xi
  Fp2 frob[4][6];         // frob[k][i] = xi^(i (p^k - 1)/6)
  Fp2 g2x, g2y;           // G2 generator
  Consts() {
    two_inv = fp_inv(fp_from_u64(2));
    three = fp_from_u64(3);
    Fp2 xi{fp_from_u64(9), fp_one()};
    b2 = scale(inv(xi), three);
    // (p^k - 1)/6 as multi-word integers
    std::vector<u64> pk(1, 1);
    for (int k = 1; k <= 3; k++) {
      std::vector<u64> nx(pk.size() + 4, 0);  // pk *= p
      for (size_t i = 0; i < pk.size(); i++) {
        u64 c = 0;
        for (int j = 0; j < 4; j++) {
          u128 t = (u128)pk[i] * P_LIMBS[j] + nx[i + j] + c;
          nx[i + j] = (u64)t;
          c = (u64)(t >> 64);
        }
        nx[i + 4] += c;
      }
      pk = nx;
      std::vector<u64> e = pk;
      e[0] -= 1;  // p^k is odd
      u64 rem = 0;  // e /= 6
      for (int i = (int)e.size() - 1; i >= 0; i--) {
        u128 cur = ((u128)rem << 64) | e[i];
        e[i] = (u64)(cur / 6);
        rem = (u64)(cur % 6);
      }
      Fp2 g = fp2_pow(xi, e);
      frob[k][0] = fp2_one();
      for (int i = 1; i < 6; i++) frob[k][i] = frob[k][i - 1] * g;
    }
    static const uint8_t G2X1[32] = {0x19, 0x8e, 0x93, 0x93, 0x92, 0x0d, 0x48, 0x3a, 0x72, 0x60, 0xbf, 0xb7, 0x31, 0xfb, 0x5d, 0x25,
                                     0xf1, 0xaa, 0x49, 0x33, 0x35, 0xa9, 0xe7, 0x12, 0x97, 0xe4, 0x85, 0xb7, 0xae, 0xf3, 0x12, 0xc2};
    static const uint8_t G2X0[32] = {0x18, 0x00, 0xde, 0xef, 0x12, 0x1f, 0x1e, 0x76, 0x42, 0x6a, 0x00, 0x66, 0x5e, 0x5c, 0x44, 0x79,
                                     0x67, 0x43, 0x22, 0xd4, 0xf7, 0x5e, 0xda, 0xdd, 0x46, 0xde, 0xbd, 0x5c, 0xd9, 0x92, 0xf6, 0xed};
    static const uint8_t G2Y1[32] = {0x09, 0x06, 0x89, 0xd0, 0x58, 0x5f, 0xf0, 0x75, 0xec, 0x9e, 0x99, 0xad, 0x69, 0x0c, 0x33, 0x95,
                                     0xbc, 0x4b, 0x31, 0x33, 0x70, 0xb3, 0x8e, 0xf3, 0x55, 0xac, 0xda, 0xdc, 0xd1, 0x22, 0x97, 0x5b};
    static const uint8_t G2Y0[32] = {0x12, 0xc8, 0x5e, 0xa5, 0xdb, 0x8c, 0x6d, 0xeb, 0x4a, 0xab, 0x71, 0x80, 0x8d, 0xcb, 0x40, 0x8f,
                                     0xe3, 0xd1, 0xe7, 0x69, 0x0c, 0x43, 0xd3, 0x7b, 0x4c, 0xe6, 0xcc, 0x01, 0x66, 0xfa, 0x7d, 0xaa};
    fp_from_be(g2x.c1, G2X1);
    fp_from_be(g2x.c0, G2X0);
    fp_from_be(g2y.c1, G2Y1);
    fp_from_be(g2y.c0, G2Y0);
  }
};
static const Consts KC;

static Fp12 frobenius(const Fp12& a, int k) {
  // w-power order: a0=c0.c0, a1=c1.c0, a2=c0.c1, a3=c1.c1, a4=c0.c2, a5=c1.c2; a_i -> conj^k(a_i) * frob[k][i]
  auto f = [&](const Fp2& x, int i) { return ((k & 1) ? conj(x) : x) * KC.frob[k][i]; };
  return Fp12{Fp6{f(a.c0.c0, 0), f(a.c0.c1, 2), f(a.c0.c2, 4)}, Fp6{f(a.c1.c0, 1), f(a.c1.c1, 3), f(a.c1.c2, 5)}};
}
// f * (ell_0 + ell_vv v^2 + ell_vw v w)   [substrate-bn Fq12::mul_by_024(ell_0, ell_vw, ell_vv)]
static Fp12 mul_by_024(const Fp12& f, const Fp2& ell_0, const Fp2& ell_vw, const Fp2& ell_vv) {
  Fp12 s{Fp6{ell_0, fp2_zero(), ell_vv}, Fp6{fp2_zero(), ell_vw, fp2_zero()}};
  return f * s;  // the fork spends 17 Fq2 multiplications here; a dense product is 18
}
static Fp12 cyclotomic_sqr(const Fp12& a) {
  Fp2 z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2;
  auto fp4 = [](Fp2& t0, Fp2& t1, const Fp2& x, const Fp2& y) {
    Fp2 tmp = x * y;
    t0 = (x + y) * (x + mul_xi(y)) - tmp - mul_xi(tmp);
    t1 = tmp + tmp;
  };
  Fp2 t0, t1, t2, t3, t4, t5;
  fp4(t0, t1, z0, z1);
  fp4(t2, t3, z2, z3);
  fp4(t4, t5, z4, z5);
  z0 = dbl(t0 - z0) + t0;
  z1 = dbl(t1 + z1) + t1;
  Fp2 tmp = mul_xi(t5);
  z2 = dbl(tmp + z2) + tmp;
  z3 = dbl(t4 - z3) + t4;
  z4 = dbl(t2 - z4) + t2;
  z5 = dbl(t3 + z5) + t3;
  return Fp12{Fp6{z0, z4, z3}, Fp6{z2, z1, z5}};
}
static void fp12_to_be(uint8_t* out, const Fp12& a) {
  const Fp2* cs[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
  for (int i = 0; i < 6; i++) {
    fp_to_be(out + 64 * i, cs[i]->c0);
    fp_to_be(out + 64 * i + 32, cs[i]->c1);
  }
}

// ------------------------------------------------------------------------------------------ groups
template <class F> struct Aff { F x, y; };
template <class F> struct Jac { F x, y, z; };
typedef Aff<Fp> G1A;
typedef Aff<Fp2> G2A;
static inline Fp one_of(const Fp*) { return fp_one(); }
static inline Fp2 one_of(const Fp2*) { return fp2_one(); }
static inline Fp zero_of(const Fp*) { return fp_zero(); }
static inline Fp2 zero_of(const Fp2*) { return fp2_zero(); }
static inline Fp sqr(const Fp& a) { return a * a; }
static inline Fp coeff_b(const Fp*) { return KC.three; }
static inline Fp2 coeff_b(const Fp2*) { return KC.b2; }

template <class F> static Jac<F> jac_identity() { return Jac<F>{one_of((F*)0), one_of((F*)0), zero_of((F*)0)}; }
template <class F> static bool on_curve(const Aff<F>& p) { return sqr(p.y) == sqr(p.x) * p.x + coeff_b((F*)0); }
template <class F> static Jac<F> jac_double(const Jac<F>& p) {
  if (is_zero(p.z)) return p;
  F a = sqr(p.x), b = sqr(p.y), c = sqr(b);
  F d = sqr(p.x + b) - a - c;
  d = d + d;
  F e = a + a + a, f = sqr(e);
  F x3 = f - (d + d);
  F c8 = dbl(dbl(dbl(c)));
  F y3 = e * (d - x3) - c8;
  F yz = p.y * p.z;
  return Jac<F>{x3, y3, yz + yz};
}
template <class F> static Jac<F> jac_add(const Jac<F>& p, const Jac<F>& q) {
  if (is_zero(p.z)) return q;
  if (is_zero(q.z)) return p;
  F z1z1 = sqr(p.z), z2z2 = sqr(q.z);
  F u1 = p.x * z2z2, u2 = q.x * z1z1;
  F s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
  if (u1 == u2) {
    if (s1 == s2) return jac_double(p);
    return jac_identity<F>();
  }
  F h = u2 - u1, i = sqr(h + h), j = h * i, rr = s2 - s1;
  rr = rr + rr;
  F v = u1 * i;
  F x3 = sqr(rr) - j - (v + v);
  F s1j = s1 * j;
  F y3 = rr * (v - x3) - (s1j + s1j);
  F z3 = (sqr(p.z + q.z) - z1z1 - z2z2) * h;
  return Jac<F>{x3, y3, z3};
}
template <class F> static Jac<F> to_jac(const Aff<F>& p) { return Jac<F>{p.x, p.y, one_of((F*)0)}; }
static Fp inv(const Fp& a) { return fp_inv(a); }
template <class F> static bool to_affine(Aff<F>& out, const Jac<F>& p) {
  if (is_zero(p.z)) return false;
  F zi = inv(p.z), zi2 = sqr(zi);
  out.x = p.x * zi2;
  out.y = p.y * zi2 * zi;
  return true;
}
// MSB-first double-and-add over all 256 bits of a canonical scalar (substrate-bn G::mul)
template <class F> static Jac<F> scalar_mul(const Aff<F>& p, const u64 k[4]) {
  Jac<F> acc = jac_identity<F>(), base = to_jac(p);
  for (int i = 255; i >= 0; i--) {
    acc = jac_double(acc);
    if ((k[i >> 6] >> (i & 63)) & 1) acc = jac_add(acc, base);
  }
  return acc;
}
template <class F> static Aff<F> neg(const Aff<F>& p) { return Aff<F>{p.x, -p.y}; }

// ------------------------------------------------------------------------------------------ pairing
static const uint8_t ATE_NAF[64] = {1, 0, 1, 0, 0, 0, 3, 0, 3, 0, 0, 0, 3, 0, 1, 0, 3, 0, 0, 3, 0, 0, 0, 0, 0, 1, 0, 0, 3, 0, 1, 0,
                                    0, 3, 0, 0, 0, 0, 3, 0, 1, 0, 0, 0, 3, 0, 3, 0, 0, 1, 0, 0, 0, 3, 0, 0, 3, 0, 1, 0, 1, 0, 0, 0};
static const u64 BN_X = 0x44e992b44a6909f1ull;

struct Ell {
  Fp2 ell_0, ell_vw, ell_vv;
};
static Ell doubling_step(Jac<Fp2>& r) {
  Fp2 a = scale(r.x * r.y, KC.two_inv), b = sqr(r.y), c = sqr(r.z), d = c + c + c, e = KC.b2 * d, f = e + e + e;
  Fp2 g = scale(b + f, KC.two_inv), h = sqr(r.y + r.z) - (b + c), i = e - b, j = sqr(r.x), e2 = sqr(e);
  r.x = a * (b - f);
  r.y = sqr(g) - (e2 + e2 + e2);
  r.z = b * h;
  return Ell{mul_xi(i), -h, j + j + j};
}
static Ell addition_step(Jac<Fp2>& r, const G2A& q) {
  Fp2 d = r.x - r.z * q.x, e = r.y - r.z * q.y, f = sqr(d), g = sqr(e), h = d * f, i = r.x * f;
  Fp2 j = r.z * g + h - (i + i);
  Ell l{mul_xi(e * q.x - d * q.y), d, -e};
  r.x = d * j;
  r.y = e * (i - j) - h * r.y;
  r.z = r.z * h;
  return l;
}
static G2A mul_by_q(const G2A& q) { return G2A{conj(q.x) * KC.frob[1][2], conj(q.y) * KC.frob[1][3]}; }
static void g2_precompute(std::vector<Ell>& out, const G2A& q) {
  Jac<Fp2> r = to_jac(q);
  G2A nq = neg(q);
  out.clear();
  for (int k = 0; k < 64; k++) {
    out.push_back(doubling_step(r));
    if (ATE_NAF[k] == 1) out.push_back(addition_step(r, q));
    else if (ATE_NAF[k] == 3) out.push_back(addition_step(r, nq));
  }
  G2A q1 = mul_by_q(q), q2 = neg(mul_by_q(q1));
  out.push_back(addition_step(r, q1));
  out.push_back(addition_step(r, q2));
}
static Fp12 miller_loop_batch(const std::vector<std::vector<Ell>>& pre, const std::vector<G1A>& ps) {
  Fp12 f = fp12_one();
  size_t idx = 0;
  auto apply = [&]() {
    for (size_t t = 0; t < ps.size(); t++) {
      const Ell& c = pre[t][idx];
      f = mul_by_024(f, c.ell_0, scale(c.ell_vw, ps[t].y), scale(c.ell_vv, ps[t].x));
    }
    idx++;
  };
  for (int k = 0; k < 64; k++) {
    f = sqr(f);
    apply();
    if (ATE_NAF[k]) apply();
  }
  apply();
  apply();
  return f;
}
static Fp12 exp_by_neg_z(const Fp12& a) {
  Fp12 r = a;
  for (int i = 61; i >= 0; i--) {
    r = cyclotomic_sqr(r);
    if ((BN_X >> i) & 1) r = r * a;
  }
  return conj(r);
}
static Fp12 final_exponentiation(const Fp12& f) {
  Fp12 t = conj(f) * inv(f);
  t = frobenius(t, 2) * t;
  Fp12 a = exp_by_neg_z(t), b = cyclotomic_sqr(a), c = cyclotomic_sqr(b), d = c * b, e = exp_by_neg_z(d);
  Fp12 ff = cyclotomic_sqr(e), g = exp_by_neg_z(ff), h = conj(d), i = conj(g), j = i * e, k = j * h, l = k * b, m = k * e;
  Fp12 n = t * m, o = frobenius(l, 1), p = o * n, q = frobenius(k, 2), rr = q * p, s = conj(t), t2 = s * l;
  Fp12 u = frobenius(t2, 3);
  return u * rr;
}
// bn::pairing_batch without the final exponentiation
static Fp12 miller_product(const G1A* ps, const G2A* qs, int k) {
  std::vector<std::vector<Ell>> pre(k);
  std::vector<G1A> pv(ps, ps + k);
  for (int j = 0; j < k; j++) g2_precompute(pre[j], qs[j]);
  return miller_loop_batch(pre, pv);
}

// ------------------------------------------------------------------------------------------ gnark wire format
enum {
  ST_OK_TRUE = 0, ST_OK_FALSE = 1, ST_ERR_PREPARE_INPUTS = 2, ST_PANIC_FIELD = 16, ST_PANIC_CURVE = 17,
  ST_PANIC_SUBGROUP = 18, ST_PANIC_IDENTITY = 19, ST_PANIC_SHORT = 20, ST_PANIC_VK = 22
};
static int load_g1(G1A& p, const uint8_t* b) {  // verifier/src/converter.rs:78-88
  bool ok = fp_from_be(p.x, b);
  ok = fp_from_be(p.y, b + 32) && ok;
  if (!ok) return ST_PANIC_FIELD;
  if (!on_curve(p)) return ST_PANIC_CURVE;
  return 0;
}
static int load_g2(G2A& q, const uint8_t* b) {  // verifier/src/converter.rs:135-153; AffineG2::new checks the order
  bool ok = fp_from_be(q.x.c1, b);
  ok = fp_from_be(q.x.c0, b + 32) && ok;
  ok = fp_from_be(q.y.c1, b + 64) && ok;
  ok = fp_from_be(q.y.c0, b + 96) && ok;
  if (!ok) return ST_PANIC_FIELD;
  if (!on_curve(q)) return ST_PANIC_CURVE;
  Jac<Fp2> t = scalar_mul(q, R_LIMBS);
  if (!is_zero(t.z)) return ST_PANIC_SUBGROUP;
  return 0;
}
static void store_g1(uint8_t* b, const G1A& p) {
  fp_to_be(b, p.x);
  fp_to_be(b + 32, p.y);
}
static void store_g2(uint8_t* b, const G2A& q) {
  fp_to_be(b, q.x.c1);
  fp_to_be(b + 32, q.x.c0);
  fp_to_be(b + 64, q.y.c1);
  fp_to_be(b + 96, q.y.c0);
}
static Fp fp_from_be_reduced(const uint8_t* b) {
  u64 t[4];
  limbs_from_be(t, b, MP);
  while (geq(t, MP.m)) sub_n(t, t, MP.m);
  Fp r;
  mont_mul(r.v, t, MP.r2, MP);
  return r;
}
static int decompress_g1(G1A& out, const uint8_t* buf) {  // verifier/src/converter.rs:23-43,62-76
  uint8_t flag = buf[0] & 0xC0;
  if (flag != 0x80 && flag != 0xC0) return -1;
  uint8_t tmp[32];
  memcpy(tmp, buf, 32);
  tmp[0] &= 0x3F;
  Fp x = fp_from_be_reduced(tmp), y;
  if (!fp_sqrt(y, x * x * x + KC.three)) return -1;
  Fp ny = -y;
  bool y_larger = fp_cmp(y, ny) > 0;
  out.x = x;
  out.y = (flag == 0xC0) == y_larger ? y : ny;  // 0xC0 = larger root
  return 0;
}
static bool fp2_sqrt(Fp2& out, const Fp2& a) {
  if (is_zero(a.c1)) {
    Fp s;
    if (fp_sqrt(s, a.c0)) { out = Fp2{s, fp_zero()}; return true; }
    if (fp_sqrt(s, -a.c0)) { out = Fp2{fp_zero(), s}; return true; }
    return false;
  }
  Fp alpha;
  if (!fp_sqrt(alpha, a.c0 * a.c0 + a.c1 * a.c1)) return false;
  Fp x0;
  if (!fp_sqrt(x0, (a.c0 + alpha) * KC.two_inv) && !fp_sqrt(x0, (a.c0 - alpha) * KC.two_inv)) return false;
  Fp x1 = a.c1 * fp_inv(x0 + x0);
  Fp2 c{x0, x1};
  if (!(sqr(c) == a)) return false;
  out = c;
  return true;
}
static bool fp2_lex_gt(const Fp2& a, const Fp2& b) {
  int c = fp_cmp(a.c1, b.c1);
  if (c) return c > 0;
  return fp_cmp(a.c0, b.c0) > 0;
}
static int decompress_g2(G2A& out, const uint8_t* buf) {  // verifier/src/converter.rs:113-133
  uint8_t flag = buf[0] & 0xC0;
  if (flag == 0x40) { out = G2A{KC.g2x, KC.g2y}; return 0; }  // reference quirk: AffineG2::one()
  if (flag != 0x80 && flag != 0xC0) return -1;
  uint8_t tmp[32];
  memcpy(tmp, buf, 32);
  tmp[0] &= 0x3F;
  Fp2 x{fp_from_be_reduced(buf + 32), fp_from_be_reduced(tmp)}, y;
  if (!fp2_sqrt(y, sqr(x) * x + KC.b2)) return -1;
  Fp2 ny = -y;
  bool y_larger = fp2_lex_gt(y, ny);
  out.x = x;
  out.y = (flag == 0xC0) == y_larger ? y : ny;
  return 0;
}
static void compress_g1(uint8_t* o, const G1A& p) {
  fp_to_be(o, p.x);
  o[0] |= fp_cmp(p.y, -p.y) > 0 ? 0xC0 : 0x80;
}
static void compress_g2(uint8_t* o, const G2A& q) {
  fp_to_be(o, q.x.c1);
  fp_to_be(o + 32, q.x.c0);
  o[0] |= fp2_lex_gt(q.y, -q.y) ? 0xC0 : 0x80;
}
static uint32_t be32(const uint8_t* b) { return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }

struct Groth16Vk {  // verifier/src/groth16/verify.rs:6-36 (fields verify_groth16 reads)
  G1A alpha;
  G2A beta_neg, gamma, delta;
  std::vector<G1A> k;
};
static int parse_groth16_vk(Groth16Vk& vk, const uint8_t* buf, size_t len) {  // verifier/src/groth16/converter.rs:28-89
  if (len < 292) return -1;
  G1A b1, d1;
  G2A b2;
  if (decompress_g1(vk.alpha, buf) || decompress_g1(b1, buf + 32) || decompress_g2(b2, buf + 64) ||
      decompress_g2(vk.gamma, buf + 128) || decompress_g1(d1, buf + 192) || decompress_g2(vk.delta, buf + 224))
    return -1;
  vk.beta_neg = neg(b2);
  uint32_t nk = be32(buf + 288);
  size_t off = 292;
  if (nk > 4096 || len < off + 32ull * nk + 4) return -1;
  vk.k.resize(nk);
  for (uint32_t i = 0; i < nk; i++, off += 32)
    if (decompress_g1(vk.k[i], buf + off)) return -1;
  uint32_t narr = be32(buf + off);
  off += 4;
  for (uint32_t a = 0; a < narr; a++) {
    if (len < off + 4) return -1;
    off += 4 + 4ull * be32(buf + off);
  }
  if (len < off + 128) return -1;
  G2A ck;
  if (decompress_g2(ck, buf + off) || decompress_g2(ck, buf + off + 64)) return -1;
  return 0;
}

// Groth16Verifier::verify for one proof, exactly the work the crate does per call.
static int groth16_verify_one(const uint8_t* vk_bytes, size_t vk_len, const uint8_t* proof, size_t proof_len,
                              const uint8_t* inputs_be, int n_inputs, uint8_t* dbg_l, uint8_t* dbg_ml, uint8_t* dbg_gt) {
  if (proof_len < 256) return ST_PANIC_SHORT;
  G1A A, C;
  G2A B;
  int st;
  if ((st = load_g1(A, proof))) return st;
  if ((st = load_g2(B, proof + 64))) return st;
  if ((st = load_g1(C, proof + 192))) return st;
  Groth16Vk vk;
  if (parse_groth16_vk(vk, vk_bytes, vk_len)) return ST_PANIC_VK;
  // verify_groth16 (verifier/src/groth16/verify.rs:65-78)
  Fp12 alpha_beta = final_exponentiation(miller_product(&vk.alpha, &vk.beta_neg, 1));
  if ((size_t)n_inputs + 1 != vk.k.size()) return ST_ERR_PREPARE_INPUTS;
  Jac<Fp> acc = to_jac(vk.k[0]);
  for (int i = 0; i < n_inputs; i++) {  // prepare_inputs (:53-63): affine accumulation, identity panics
    u64 x[4];
    if (!limbs_from_be(x, inputs_be + 32 * i, MR)) return ST_PANIC_FIELD;
    Jac<Fp> term = scalar_mul(vk.k[i + 1], x);
    G1A ta;
    if (!to_affine(ta, term)) return ST_PANIC_IDENTITY;
    acc = jac_add(acc, to_jac(ta));
    G1A aa;
    if (!to_affine(aa, acc)) return ST_PANIC_IDENTITY;
    acc = to_jac(aa);
  }
  G1A L;
  to_affine(L, acc);
  if (dbg_l) store_g1(dbg_l, L);
  G1A ps[3] = {A, L, C};
  G2A qs[3] = {B, vk.gamma, neg(vk.delta)};
  Fp12 ml = miller_product(ps, qs, 3);
  if (dbg_ml) fp12_to_be(dbg_ml, ml);
  Fp12 gt = final_exponentiation(ml);
  if (dbg_gt) fp12_to_be(dbg_gt, gt);
  return gt == alpha_beta ? ST_OK_TRUE : ST_OK_FALSE;
}

template <class Fn> static void parallel_for(size_t n, int threads, Fn fn) {
  if (threads < 1) threads = 1;
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) return;
      fn(i);
    }
  };
  std::vector<std::thread> ts;
  for (int t = 1; t < threads; t++) ts.emplace_back(worker);
  worker();
  for (auto& t : ts) t.join();
}

// ------------------------------------------------------------------------------------------ workload generator
// Same definition as oracle/bn254_oracle.py (splitmix64, synth_scalar, Groth16Trapdoor).
static u64 splitmix64(u64& s) {
  s += 0x9E3779B97F4A7C15ull;
  u64 z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void synth_scalar(u64 out[4], u64 seed, u64 index, u64 slot) {  // canonical, in [1, r)
  u64 st = seed * 0xD1342543DE82EF95ull + index * 0x2545F4914F6CDD1Dull + slot * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  for (int w = 0; w < 4; w++) out[w] = splitmix64(st);
  while (geq(out, MR.m)) sub_n(out, out, MR.m);
  if (!(out[0] | out[1] | out[2] | out[3])) out[0] = 1;
}
struct Trapdoor {
  Fr alpha, beta, gamma, delta, delta_inv, ic[16];
  int n_public, sign_mode;
  u64 seed;
};
static void trapdoor_init(Trapdoor& td, u64 seed, int n_public, int sign_mode) {
  u64 t[4];
  const u64 big = 1ull << 40;
  auto get = [&](u64 slot) { synth_scalar(t, seed, big, slot); return fr_from_plain(t); };
  td.alpha = get(0), td.beta = get(1), td.gamma = get(2), td.delta = get(3);
  td.delta_inv = fr_inv(td.delta);
  for (int i = 0; i <= n_public; i++) td.ic[i] = get(4 + i);
  td.n_public = n_public, td.sign_mode = sign_mode, td.seed = seed;
}
static void base_scalars(const Trapdoor& td, u64 index, u64 xs[][4], Fr& a, Fr& b, Fr& c) {
  u64 t[4];
  for (int i = 0; i < td.n_public; i++) synth_scalar(xs[i], td.seed, index, 8 + i);
  xs[0][3] &= 0x00ffffffffffffffull;
  if (!(xs[0][0] | xs[0][1] | xs[0][2] | xs[0][3])) xs[0][0] = 1;
  synth_scalar(t, td.seed, index, 0);
  a = fr_from_plain(t);
  synth_scalar(t, td.seed, index, 1);
  b = fr_from_plain(t);
  Fr ell = td.ic[0];
  for (int i = 0; i < td.n_public; i++) ell = fr_add(ell, fr_mul(fr_from_plain(xs[i]), td.ic[i + 1]));
  Fr ab = fr_mul(a, b), lg = fr_mul(ell, td.gamma), al = fr_mul(td.alpha, td.beta);
  Fr num = td.sign_mode == 0 ? fr_add(fr_add(ab, lg), al) : fr_sub(fr_sub(ab, lg), al);
  c = fr_mul(num, td.delta_inv);
}
static G1A g1_gen() { return G1A{fp_one(), fp_from_u64(2)}; }
static void g1_mul_gen(G1A& out, const Fr& k) {
  u64 p[4];
  fr_to_plain(p, k);
  to_affine(out, scalar_mul(g1_gen(), p));
}
static void g2_mul_gen(G2A& out, const Fr& k) {
  u64 p[4];
  fr_to_plain(p, k);
  to_affine(out, scalar_mul(G2A{KC.g2x, KC.g2y}, p));
}

// ------------------------------------------------------------------------------------------ PlonK (reference shape)
// verify_plonk and its helpers restated per call, as the crate does it (verifier/src/plonk/verify.rs:46-396,
// verifier/src/plonk/kzg.rs:46-190, verifier/src/transcript.rs:68-107, verifier/src/hash_to_field.rs:30-97,
// verifier/src/plonk/converter.rs:18-178): VK decompressed on every call, naive AffineG1::msm (one 256-bit
// double-and-add per term), generic Fr::pow, three separate Fr inversions, a 2-pair pairing_batch with both G2
// precomputations.  Statuses as include/bn254v.h.
enum {
  ST_ERR_BSB22 = 3, ST_ERR_WITNESS = 4, ST_ERR_INVERSE = 5, ST_ERR_OPENING = 6, ST_ERR_NDIGESTS = 7, ST_ERR_PAIRING = 8,
  ST_PANIC_DIV0 = 21, ST_PANIC_INDEX = 22
};

struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  u64 len;
};
static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_compress(uint32_t* h, const uint8_t* b) {
  static const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
      0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
      0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
      0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
      0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
      0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
      0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = ((uint32_t)b[4 * i] << 24) | (b[4 * i + 1] << 16) | (b[4 * i + 2] << 8) | b[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = hh + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
    uint32_t t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
    hh = g, g = f, f = e, e = d + t1, d = c, c = bb, bb = a, a = t1 + t2;
  }
  h[0] += a, h[1] += bb, h[2] += c, h[3] += d, h[4] += e, h[5] += f, h[6] += g, h[7] += hh;
}
static void sha_init(Sha256& s) {
  static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(s.h, iv, 32);
  s.len = 0;
}
static void sha_update(Sha256& s, const void* data, size_t n) {
  const uint8_t* p = (const uint8_t*)data;
  for (size_t i = 0; i < n; i++) {
    s.buf[s.len++ & 63] = p[i];
    if ((s.len & 63) == 0) sha_compress(s.h, s.buf);
  }
}
static void sha_final(Sha256& s, uint8_t* out) {
  u64 bits = s.len * 8;
  uint8_t pad = 0x80;
  sha_update(s, &pad, 1);
  pad = 0;
  while ((s.len & 63) != 56) sha_update(s, &pad, 1);
  uint8_t lb[8];
  for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
  sha_update(s, lb, 8);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = s.h[i] >> 24, out[4 * i + 1] = s.h[i] >> 16, out[4 * i + 2] = s.h[i] >> 8, out[4 * i + 3] = s.h[i];
  }
}

static Fr fr_zero() { return Fr{{0, 0, 0, 0}}; }
static Fr fr_one() {
  Fr r;
  memcpy(r.v, MR.r1, 32);
  return r;
}
static bool fr_is_zero(const Fr& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static bool fr_eq(const Fr& a, const Fr& b) { return memcmp(a.v, b.v, 32) == 0; }
static Fr fr_neg(const Fr& a) { return fr_sub(fr_zero(), a); }
static bool fr_from_be(Fr& out, const uint8_t* b) {  // Fr::from_slice: must be < r
  u64 t[4];
  bool ok = limbs_from_be(t, b, MR);
  out = fr_from_plain(t);
  return ok;
}
static Fr fr_from_be_mod_order(const uint8_t* b) {  // Fr::from_bytes_be_mod_order
  u64 t[4];
  limbs_from_be(t, b, MR);
  while (geq(t, MR.m)) sub_n(t, t, MR.m);
  return fr_from_plain(t);
}
static void fr_to_be(uint8_t* b, const Fr& a) {
  u64 t[4];
  fr_to_plain(t, a);
  limbs_to_be(b, t);
}
static Fr fr_pow_fr(const Fr& a, const Fr& e_mont) {  // Fr::pow(Fr): generic 256-bit square-and-multiply
  u64 e[4];
  fr_to_plain(e, e_mont);
  Fr r = fr_one();
  for (int i = 255; i >= 0; i--) {
    r = fr_mul(r, r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = fr_mul(r, a);
  }
  return r;
}
static Fr fr_from_u64(u64 x) {
  u64 t[4] = {x, 0, 0, 0};
  return fr_from_plain(t);
}

struct PlonkVk {
  u64 size, nb_public;
  Fr size_inv, generator, coset_shift;
  G1A s[3], ql, qr, qm, qo, qk, g1;
  std::vector<G1A> qcp;
  G2A g2[2];
  std::vector<u64> cci;
};
static u64 be64(const uint8_t* b) { return ((u64)be32(b) << 32) | be32(b + 4); }
static int parse_plonk_vk(PlonkVk& vk, const uint8_t* buf, size_t len) {  // plonk/converter.rs:18-119
  if (len < 372) return -1;
  vk.size = be64(buf);
  if (!fr_from_be(vk.size_inv, buf + 8) || !fr_from_be(vk.generator, buf + 40)) return -1;
  vk.nb_public = be64(buf + 72);
  if (!fr_from_be(vk.coset_shift, buf + 80)) return -1;
  G1A* pts[8] = {&vk.s[0], &vk.s[1], &vk.s[2], &vk.ql, &vk.qr, &vk.qm, &vk.qo, &vk.qk};
  for (int i = 0; i < 8; i++)
    if (decompress_g1(*pts[i], buf + 112 + 32 * i)) return -1;
  uint32_t nq = be32(buf + 368);
  size_t off = 372;
  if (nq > 64 || len < off + 32ull * nq + 160 + 33788 + 8) return -1;
  vk.qcp.resize(nq);
  for (uint32_t i = 0; i < nq; i++, off += 32)
    if (decompress_g1(vk.qcp[i], buf + off)) return -1;
  if (decompress_g1(vk.g1, buf + off) || decompress_g2(vk.g2[0], buf + off + 32) || decompress_g2(vk.g2[1], buf + off + 96)) return -1;
  off += 160 + 33788;
  u64 nidx = be64(buf + off);
  off += 8;
  if (nidx > 64 || len < off + 8 * nidx) return -1;
  vk.cci.resize(nidx);
  for (u64 i = 0; i < nidx; i++) vk.cci[i] = be64(buf + off + 8 * i);
  return 0;
}

static G1A g1_neg(const G1A& p) { return G1A{p.x, -p.y}; }
// AffineG1::msm: naive sum of AffineG1 * Fr; the result converts to affine (identity panics)
static bool g1_msm(G1A& out, const std::vector<G1A>& pts, const std::vector<Fr>& ks) {
  Jac<Fp> acc = jac_identity<Fp>();
  size_t n = pts.size() < ks.size() ? pts.size() : ks.size();
  for (size_t i = 0; i < n; i++) {
    u64 k[4];
    fr_to_plain(k, ks[i]);
    acc = jac_add(acc, scalar_mul(pts[i], k));
  }
  return to_affine(out, acc);
}
static Fr transcript_challenge(const char* id, const uint8_t* prev, const std::vector<std::pair<const uint8_t*, size_t>>& binds,
                               uint8_t* digest) {
  Sha256 s;
  sha_init(s);
  sha_update(s, id, strlen(id));
  if (prev) sha_update(s, prev, 32);
  for (auto& b : binds) sha_update(s, b.first, b.second);
  sha_final(s, digest);
  return fr_from_be_mod_order(digest);
}
static Fr hash_to_field_bsb22(const uint8_t* pt64) {  // hash_to_field.rs: expand_message_xmd, 48 bytes, mod r
  static const uint8_t dst[12] = {'B', 'S', 'B', '2', '2', '-', 'P', 'l', 'o', 'n', 'k', 11};
  uint8_t b0[32], b1[32], b2[32], z[64] = {0}, l[3] = {0, 48, 0}, one = 1, two = 2, x[32];
  Sha256 s;
  sha_init(s), sha_update(s, z, 64), sha_update(s, pt64, 64), sha_update(s, l, 3), sha_update(s, dst, 12), sha_final(s, b0);
  sha_init(s), sha_update(s, b0, 32), sha_update(s, &one, 1), sha_update(s, dst, 12), sha_final(s, b1);
  for (int i = 0; i < 32; i++) x[i] = b0[i] ^ b1[i];
  sha_init(s), sha_update(s, x, 32), sha_update(s, &two, 1), sha_update(s, dst, 12), sha_final(s, b2);
  // BE(b1 | b2[0..16]) mod r = (b1 mod r) * 2^128 + b2_hi
  Fr hi = fr_from_be_mod_order(b1);
  u64 t128[4] = {0, 0, 1, 0};
  uint8_t lo_be[32] = {0};
  memcpy(lo_be + 16, b2, 16);
  return fr_add(fr_mul(hi, fr_from_plain(t128)), fr_from_be_mod_order(lo_be));
}

static int plonk_verify_one(const uint8_t* vk_bytes, size_t vk_len, const uint8_t* pr, size_t len, const uint8_t* inputs_be,
                            int n_inputs, const uint8_t* rnd_be, uint8_t* dbg_gt) {
  for (int i = 0; i < n_inputs; i++) {
    u64 t[4];
    if (!limbs_from_be(t, inputs_be + 32 * i, MR)) return ST_PANIC_FIELD;
  }
  // load_plonk_proof_from_bytes
  if (len < 516) return ST_PANIC_SHORT;
  G1A P8[8];
  int st;
  for (int i = 0; i < 8; i++)
    if ((st = load_g1(P8[i], pr + 64 * i))) return st;
  uint32_t ncl = be32(pr + 512);
  size_t off = 516;
  std::vector<Fr> cl;
  for (uint32_t i = 0; i < ncl; i++) {
    if (off + 32 > len) return ST_PANIC_SHORT;
    Fr t;
    if (!fr_from_be(t, pr + off)) return ST_PANIC_FIELD;
    cl.push_back(t);
    off += 32;
  }
  if (off + 100 > len) return ST_PANIC_SHORT;
  const size_t off_zsh = off;
  G1A zs_h;
  if ((st = load_g1(zs_h, pr + off))) return st;
  Fr zu;
  if (!fr_from_be(zu, pr + off + 64)) return ST_PANIC_FIELD;
  uint32_t nbsb = be32(pr + off + 96);
  off += 100;
  const size_t off_bsb = off;
  std::vector<G1A> bsb;
  for (uint32_t i = 0; i < nbsb; i++) {
    if (off + 64 > len) return ST_PANIC_SHORT;
    G1A t;
    if ((st = load_g1(t, pr + off))) return st;
    bsb.push_back(t);
    off += 64;
  }
  PlonkVk vk;
  if (parse_plonk_vk(vk, vk_bytes, vk_len)) return ST_PANIC_VK;
  // verify_plonk
  if (bsb.size() != vk.qcp.size()) return ST_ERR_BSB22;
  if ((u64)n_inputs != vk.nb_public) return ST_ERR_WITNESS;
  uint8_t vkpts[64 * 72], dg[32], dgb[32], dga[32], dgz[32];
  size_t nvk = 0;
  {
    const G1A* pts[8] = {&vk.s[0], &vk.s[1], &vk.s[2], &vk.ql, &vk.qr, &vk.qm, &vk.qo, &vk.qk};
    for (int i = 0; i < 8; i++) store_g1(vkpts + 64 * nvk++, *pts[i]);
    for (auto& q : vk.qcp) store_g1(vkpts + 64 * nvk++, q);
  }
  Fr gamma = transcript_challenge("gamma", nullptr, {{vkpts, 64 * nvk}, {inputs_be, (size_t)32 * n_inputs}, {pr, 192}}, dg);
  Fr beta = transcript_challenge("beta", dg, {}, dgb);
  Fr alpha = transcript_challenge("alpha", dgb, {{pr + off_bsb, 64 * (size_t)nbsb}, {pr + 192, 64}}, dga);
  Fr zeta = transcript_challenge("zeta", dga, {{pr + 256, 192}}, dgz);
  const Fr one = fr_one();
  Fr zeta_n = fr_pow_fr(zeta, fr_from_u64(vk.size));
  Fr zh_zeta = fr_sub(zeta_n, one);
  Fr zm1 = fr_sub(zeta, one);
  if (fr_is_zero(zm1)) return ST_ERR_INVERSE;
  Fr lagrange_one = fr_mul(fr_mul(fr_inv(zm1), zh_zeta), vk.size_inv);
  Fr pi = fr_zero();
  {
    std::vector<Fr> dens;
    Fr accw = one;
    for (int i = 0; i < n_inputs; i++) {
      dens.push_back(fr_sub(zeta, accw));
      accw = fr_mul(accw, vk.generator);
    }
    // batch_invert (zeros skipped)
    std::vector<Fr> prod;
    Fr tmp = one;
    for (auto& d : dens)
      if (!fr_is_zero(d)) {
        tmp = fr_mul(tmp, d);
        prod.push_back(tmp);
      }
    std::vector<Fr> inv(dens.size(), fr_zero());
    if (!prod.empty()) {
      tmp = fr_inv(tmp);
      int k = (int)prod.size() - 1;
      for (int i = (int)dens.size() - 1; i >= 0; i--) {
        if (fr_is_zero(dens[i])) continue;
        Fr sfx = k > 0 ? prod[k - 1] : one;
        inv[i] = fr_mul(tmp, sfx);
        tmp = fr_mul(tmp, dens[i]);
        k--;
      }
    }
    accw = one;
    for (int i = 0; i < n_inputs; i++) {
      Fr w;
      fr_from_be(w, inputs_be + 32 * i);
      pi = fr_add(pi, fr_mul(fr_mul(fr_mul(fr_mul(zh_zeta, inv[i]), vk.size_inv), accw), w));
      accw = fr_mul(accw, vk.generator);
    }
  }
  for (size_t i = 0; i < vk.cci.size(); i++) {
    if (i >= bsb.size()) return ST_PANIC_INDEX;
    Fr hc = hash_to_field_bsb22(pr + off_bsb + 64 * i);
    Fr wpi = fr_pow_fr(vk.generator, fr_from_u64(vk.nb_public + vk.cci[i]));
    Fr den = fr_sub(zeta, wpi);
    if (fr_is_zero(den)) return ST_PANIC_DIV0;
    Fr lag = fr_mul(fr_mul(fr_mul(zh_zeta, wpi), fr_inv(den)), vk.size_inv);
    pi = fr_add(pi, fr_mul(lag, hc));
  }
  if (cl.size() < 6) return ST_PANIC_INDEX;
  const Fr &l = cl[1], &r = cl[2], &o = cl[3], &s1 = cl[4], &s2 = cl[5];
  Fr a2l1 = fr_mul(fr_mul(lagrange_one, alpha), alpha);
  Fr t1 = fr_add(fr_add(fr_mul(beta, s1), gamma), l), t2 = fr_add(fr_add(fr_mul(beta, s2), gamma), r);
  Fr const_lin = fr_mul(fr_mul(fr_mul(fr_mul(t1, t2), fr_add(o, gamma)), alpha), zu);
  const_lin = fr_neg(fr_add(fr_sub(const_lin, a2l1), pi));
  if (!fr_eq(const_lin, cl[0])) return ST_ERR_OPENING;
  Fr s1c = fr_mul(fr_mul(fr_mul(fr_mul(t1, t2), beta), alpha), zu);
  const Fr& u = vk.coset_shift;
  Fr bz = fr_mul(beta, zeta), buz = fr_mul(bz, u);
  Fr s2c = fr_mul(fr_mul(fr_add(fr_add(bz, gamma), l), fr_add(fr_add(buz, gamma), r)), fr_add(fr_add(fr_mul(buz, u), gamma), o));
  s2c = fr_neg(fr_mul(s2c, alpha));
  Fr zn2 = fr_pow_fr(zeta, fr_from_u64(vk.size + 2));
  std::vector<G1A> pts(bsb);
  for (const G1A* q : {&vk.ql, &vk.qr, &vk.qm, &vk.qo, &vk.qk, &vk.s[2], &P8[3], &P8[4], &P8[5], &P8[6]}) pts.push_back(*q);
  std::vector<Fr> sc(cl.begin() + 6, cl.end());
  for (const Fr& k : {l, r, fr_mul(l, r), o, one, s1c, fr_add(a2l1, s2c), fr_neg(zh_zeta), fr_neg(fr_mul(zn2, zh_zeta)),
                      fr_neg(fr_mul(fr_mul(zn2, zn2), zh_zeta))})
    sc.push_back(k);
  G1A lin;
  if (!g1_msm(lin, pts, sc)) return ST_PANIC_IDENTITY;
  // kzg::fold_proof
  std::vector<G1A> digests = {lin, P8[0], P8[1], P8[2], vk.s[0], vk.s[1]};
  for (auto& q : vk.qcp) digests.push_back(q);
  if (digests.size() != cl.size()) return ST_ERR_NDIGESTS;
  uint8_t zb[32], dbytes[64 * 72], dk[32];
  fr_to_be(zb, zeta);
  for (size_t i = 0; i < digests.size(); i++) store_g1(dbytes + 64 * i, digests[i]);
  Fr kg = transcript_challenge("gamma", nullptr, {{zb, 32}, {dbytes, 64 * digests.size()}, {pr + 516, 32 * (size_t)ncl}, {pr + off_zsh + 64, 32}}, dk);
  std::vector<Fr> gi = {one};
  for (size_t i = 1; i < digests.size(); i++) gi.push_back(fr_mul(gi.back(), kg));
  Fr folded_eval = fr_zero();
  for (size_t i = 0; i < cl.size(); i++) folded_eval = fr_add(folded_eval, fr_mul(cl[i], gi[i]));
  G1A folded_digest;
  if (!g1_msm(folded_digest, digests, gi)) return ST_PANIC_IDENTITY;
  // kzg::batch_verify_multi_points
  Fr rnd = fr_from_be_mod_order(rnd_be);
  std::vector<Fr> rn = {one, rnd};
  std::vector<G1A> quot = {P8[7], zs_h};
  G1A fq, fdg, t;
  if (!g1_msm(fq, quot, rn)) return ST_PANIC_IDENTITY;
  if (!g1_msm(fdg, {folded_digest, P8[3]}, rn)) return ST_PANIC_IDENTITY;
  Fr fev = fr_add(folded_eval, fr_mul(zu, rnd));
  {
    u64 k[4];
    fr_to_plain(k, fev);
    if (!to_affine(t, scalar_mul(vk.g1, k))) return ST_PANIC_IDENTITY;
    if (!to_affine(fdg, jac_add(to_jac(fdg), to_jac(g1_neg(t))))) return ST_PANIC_IDENTITY;
  }
  Fr shifted = fr_mul(zeta, vk.generator);
  std::vector<Fr> rn2 = {zeta, fr_mul(rnd, shifted)};
  if (!g1_msm(t, quot, rn2)) return ST_PANIC_IDENTITY;
  if (!to_affine(fdg, jac_add(to_jac(fdg), to_jac(t)))) return ST_PANIC_IDENTITY;
  fq = g1_neg(fq);
  G1A ps[2] = {fdg, fq};
  Fp12 gt = final_exponentiation(miller_product(ps, vk.g2, 2));
  if (dbg_gt) fp12_to_be(dbg_gt, gt);
  return gt == fp12_one() ? ST_OK_TRUE : ST_ERR_PAIRING;
}

extern "C" {

// status[i] as include/bn254v.h (22 = the VK itself failed to parse: the reference would panic on every call)
int ref_groth16_verify_batch(const uint8_t* vk, size_t vk_len, const uint8_t* proofs, size_t stride, const uint32_t* lens,
                             const uint8_t* inputs_be, int n_inputs, size_t n, uint8_t* status, uint8_t* dbg_l,
                             uint8_t* dbg_ml, uint8_t* dbg_gt, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    status[i] = (uint8_t)groth16_verify_one(vk, vk_len, proofs + stride * i, lens ? lens[i] : stride,
                                            inputs_be + (size_t)32 * n_inputs * i, n_inputs, dbg_l ? dbg_l + 64 * i : 0,
                                            dbg_ml ? dbg_ml + 384 * i : 0, dbg_gt ? dbg_gt + 384 * i : 0);
  });
  return 0;
}

// bn::pairing_batch on trusted points: miller / gt canonical Fq12, is_one
int ref_pairing_product_batch(const uint8_t* g1, const uint8_t* g2, int k, size_t n, uint8_t* is_one, uint8_t* ml_out,
                              uint8_t* gt_out, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    G1A ps[8];
    G2A qs[8];
    for (int j = 0; j < k; j++) {
      fp_from_be(ps[j].x, g1 + (i * k + j) * 64);
      fp_from_be(ps[j].y, g1 + (i * k + j) * 64 + 32);
      const uint8_t* b = g2 + (i * k + j) * 128;
      fp_from_be(qs[j].x.c1, b);
      fp_from_be(qs[j].x.c0, b + 32);
      fp_from_be(qs[j].y.c1, b + 64);
      fp_from_be(qs[j].y.c0, b + 96);
    }
    Fp12 ml = miller_product(ps, qs, k);
    Fp12 gt = final_exponentiation(ml);
    if (ml_out) fp12_to_be(ml_out + 384 * i, ml);
    if (gt_out) fp12_to_be(gt_out + 384 * i, gt);
    is_one[i] = gt == fp12_one();
  });
  return 0;
}

int ref_groth16_synth(u64 seed, int n_public, int sign_mode, size_t first, size_t n, uint8_t* vk_bytes, size_t* vk_len,
                      uint8_t* proofs, uint8_t* inputs_be, uint8_t* expected, int threads) {
  if (n_public < 1 || n_public > 15) return -1;
  Trapdoor td;
  trapdoor_init(td, seed, n_public, sign_mode);
  if (vk_bytes) {
    G1A a1, b1, d1, kk;
    G2A b2, g2, d2, gen{KC.g2x, KC.g2y};
    g1_mul_gen(a1, td.alpha), g1_mul_gen(b1, td.beta), g2_mul_gen(b2, td.beta), g2_mul_gen(g2, td.gamma);
    g1_mul_gen(d1, td.delta), g2_mul_gen(d2, td.delta);
    uint8_t* o = vk_bytes;
    memset(o, 0, 292 + 32 * (n_public + 1) + 4 + 128);
    compress_g1(o, a1), compress_g1(o + 32, b1), compress_g2(o + 64, b2), compress_g2(o + 128, g2);
    compress_g1(o + 192, d1), compress_g2(o + 224, d2);
    o[291] = (uint8_t)(n_public + 1);
    o += 292;
    for (int i = 0; i <= n_public; i++, o += 32) {
      g1_mul_gen(kk, td.ic[i]);
      compress_g1(o, kk);
    }
    o += 4;
    compress_g2(o, gen), compress_g2(o + 64, gen);
    *vk_len = 292 + 32 * (n_public + 1) + 4 + 128;
  }
  parallel_for(n, threads, [&](size_t ii) {
    u64 index = first + ii, xs[16][4], xs2[16][4];
    Fr a, b, c, a2, b2;
    base_scalars(td, index, xs, a, b, c);
    u64 j = index >> 1, st = seed ^ (j * 0xA24BAED4963EE407ull) ^ 0x9FB21C651E98DF25ull;
    u64 z = splitmix64(st);
    bool bad = (index & 1) == (z & 1);
    int klass = (int)(j % 5);
    if (bad) {
      if (klass == 0) {
        u64 one[4] = {1, 0, 0, 0};
        add_n(xs[0], xs[0], one);
        if (geq(xs[0], MR.m)) sub_n(xs[0], xs[0], MR.m);
      } else if (klass == 1) a = fr_add(a, a);
      else if (klass == 2) { Fr zero{{0, 0, 0, 0}}; c = fr_sub(zero, c); }
      else if (klass == 3) b = fr_add(b, b);
      else base_scalars(td, index ^ 1, xs2, a2, b2, c);
    }
    G1A pa, pc;
    G2A pb;
    g1_mul_gen(pa, a), g2_mul_gen(pb, b), g1_mul_gen(pc, c);
    store_g1(proofs + 256 * ii, pa), store_g2(proofs + 256 * ii + 64, pb), store_g1(proofs + 256 * ii + 192, pc);
    for (int i = 0; i < n_public; i++) limbs_to_be(inputs_be + (ii * n_public + i) * 32, xs[i]);
    expected[ii] = bad ? ST_OK_FALSE : ST_OK_TRUE;
  });
  return 0;
}

int ref_plonk_verify_batch(const uint8_t* vk, size_t vk_len, const uint8_t* proofs, size_t stride, const uint32_t* lens,
                            const uint8_t* inputs_be, int n_inputs, const uint8_t* rnd_be, size_t n, uint8_t* status,
                            uint8_t* dbg_gt, int threads) {
  parallel_for(n, threads, [&](size_t i) {
    status[i] = (uint8_t)plonk_verify_one(vk, vk_len, proofs + stride * i, lens ? lens[i] : stride,
                                          inputs_be + (size_t)32 * n_inputs * i, n_inputs, rnd_be + 32 * i,
                                          dbg_gt ? dbg_gt + 384 * i : 0);
  });
  return 0;
}

u64 ref_fp_mul_count(void) {
#ifdef REF_COUNT
  return g_fp_mul_count;
#else
  return 0;
#endif
}
void ref_fp_mul_count_reset(void) {
#ifdef REF_COUNT
  g_fp_mul_count = 0;
#endif
}

}  // extern "C"
