// Host-side gnark wire-format framing and point decompression ("parsing stays on the host").
// Restates, in C++, the VK half of the reference's parsers:
//   deserialize_with_flags                 verifier/src/converter.rs:23-43
//   unchecked_compressed_x_to_g1_point     verifier/src/converter.rs:62-76
//   unchecked_compressed_x_to_g2_point     verifier/src/converter.rs:113-133
//   load_groth16_verifying_key_from_bytes  verifier/src/groth16/converter.rs:28-89
//   load_plonk_verifying_key_from_bytes    verifier/src/plonk/converter.rs:18-119
// Runs once per VK.  Field arithmetic is the host build of field.cuh (VK preparation only; no
// per-proof work is ever done on the host).
#pragma once
#include <stdint.h>

#include <vector>

#include "curve.cuh"

namespace bn254 {
namespace gnark {

static const uint8_t MASK = 0xC0, FLAG_POS = 0x80, FLAG_NEG = 0xC0, FLAG_INF = 0x40;

inline Fp fp_from_be_mod_order(const uint8_t* b) {  // BE bytes -> Montgomery, reduced mod p
  Fp t;
  fe_from_be_bytes(t, b);
  fe_reduce_full(t);
  return fe_to_mont(t);
}

// lexicographic compare of canonical integers (Montgomery inputs)
inline int fp_cmp(const Fp& a, const Fp& b) {
  Fp x = fe_from_mont(a), y = fe_from_mont(b);
  for (int i = 7; i >= 0; i--) {
    if (x.v[i] != y.v[i]) return x.v[i] > y.v[i] ? 1 : -1;
  }
  return 0;
}

inline bool fp_sqrt(Fp& out, const Fp& a) {
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = K::sqrt_exp(i);
  Fp y = fe_pow_words(a, e);
  if (!fe_eq(fe_sqr(y), a)) return false;
  out = y;
  return true;
}

inline bool fp2_sqrt(Fp2& out, const Fp2& a) {
  if (fe_is_zero(a.c1)) {
    Fp s;
    if (fp_sqrt(s, a.c0)) {
      out = Fp2{s, fe_zero<FpCfg>()};
      return true;
    }
    if (fp_sqrt(s, fe_neg(a.c0))) {
      out = Fp2{fe_zero<FpCfg>(), s};
      return true;
    }
    return false;
  }
  Fp alpha;
  if (!fp_sqrt(alpha, fe_add(fe_sqr(a.c0), fe_sqr(a.c1)))) return false;
  Fp delta = fp_halve(fe_add(a.c0, alpha));
  Fp x0;
  if (!fp_sqrt(x0, delta)) {
    delta = fp_halve(fe_sub(a.c0, alpha));
    if (!fp_sqrt(x0, delta)) return false;
  }
  Fp x1 = fe_mul(a.c1, fe_inv(fe_dbl(x0)));
  Fp2 cand{x0, x1};
  if (!eq(sqr(cand), a)) return false;
  out = cand;
  return true;
}

// returns 0 ok, <0 error (the reference would panic / return Err inside the VK parser)
inline int deserialize_with_flags(Fp& x, uint8_t& flag, const uint8_t* buf) {
  flag = buf[0] & MASK;
  if (flag != FLAG_POS && flag != FLAG_NEG && flag != FLAG_INF) return -1;  // constants.rs:24 panic
  if (flag == FLAG_INF) {
    if (buf[0] & ~MASK) return -1;
    for (int i = 1; i < 32; i++)
      if (buf[i]) return -1;
    x = fe_zero<FpCfg>();
    return 0;
  }
  uint8_t tmp[32];
  memcpy(tmp, buf, 32);
  tmp[0] &= (uint8_t)~MASK;
  x = fp_from_be_mod_order(tmp);
  return 0;
}

inline int decompress_g1(G1Aff& out, const uint8_t* buf) {
  Fp x;
  uint8_t flag;
  if (deserialize_with_flags(x, flag, buf)) return -1;
  Fp y;
  if (!fp_sqrt(y, fe_add(fe_mul(fe_sqr(x), x), fp_three()))) return -1;
  Fp ny = fe_neg(y);
  bool y_greater = fp_cmp(y, ny) > 0;
  // Positive flag -> smaller root, Negative -> larger root
  Fp final_y = y;
  if (y_greater) {
    if (flag == FLAG_POS) final_y = ny;
  } else if (flag == FLAG_NEG) {
    final_y = ny;
  }
  out = G1Aff{x, final_y};
  return 0;
}

inline G2Aff g2_generator() {
  G2Aff g;
  BN_LOAD_FP(g.x.c0, K::g2_gen, 0);
  BN_LOAD_FP(g.x.c1, K::g2_gen, 1);
  BN_LOAD_FP(g.y.c0, K::g2_gen, 2);
  BN_LOAD_FP(g.y.c1, K::g2_gen, 3);
  return g;
}

inline bool fp2_lex_gt(const Fp2& a, const Fp2& b) {  // gnark E2 ordering: A1 first, then A0
  int c = fp_cmp(a.c1, b.c1);
  if (c) return c > 0;
  return fp_cmp(a.c0, b.c0) > 0;
}

inline int decompress_g2(G2Aff& out, const uint8_t* buf) {
  Fp x1;
  uint8_t flag;
  if (deserialize_with_flags(x1, flag, buf)) return -1;
  Fp x0 = fp_from_be_mod_order(buf + 32);
  if (flag == FLAG_INF) {  // reference quirk: AffineG2::one()
    out = g2_generator();
    return 0;
  }
  Fp2 x{x0, x1};
  Fp2 y;
  if (!fp2_sqrt(y, add(mul(sqr(x), x), fp2_b2()))) return -1;
  Fp2 ny = neg(y);
  Fp2 lo = y, hi = ny;
  if (fp2_lex_gt(y, ny)) {
    lo = ny;
    hi = y;
  }
  out = G2Aff{x, flag == FLAG_NEG ? hi : lo};
  return 0;
}

inline void compress_g1(uint8_t* out32, const G1Aff& p) {
  fe_to_be_bytes(out32, fe_from_mont(p.x));
  out32[0] |= (fp_cmp(p.y, fe_neg(p.y)) > 0) ? FLAG_NEG : FLAG_POS;
}
inline void compress_g2(uint8_t* out64, const G2Aff& q) {
  fe_to_be_bytes(out64, fe_from_mont(q.x.c1));
  fe_to_be_bytes(out64 + 32, fe_from_mont(q.x.c0));
  out64[0] |= fp2_lex_gt(q.y, neg(q.y)) ? FLAG_NEG : FLAG_POS;
}

inline uint32_t be32(const uint8_t* b) { return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }
inline uint64_t be64(const uint8_t* b) { return ((uint64_t)be32(b) << 32) | be32(b + 4); }

struct Groth16VkHost {
  G1Aff alpha, beta1, delta1;      // beta1 stored negated as the reference does (unused by verify)
  G2Aff beta2, gamma2, delta2;     // beta2 stored negated (groth16/converter.rs:79)
  std::vector<G1Aff> k;
};

inline int parse_groth16_vk(Groth16VkHost& vk, const uint8_t* buf, size_t len) {
  if (len < 292) return -1;
  G1Aff b1;
  G2Aff b2;
  if (decompress_g1(vk.alpha, buf) || decompress_g1(b1, buf + 32) || decompress_g2(b2, buf + 64) ||
      decompress_g2(vk.gamma2, buf + 128) || decompress_g1(vk.delta1, buf + 192) || decompress_g2(vk.delta2, buf + 224))
    return -1;
  vk.beta1 = neg(b1);
  vk.beta2 = neg(b2);
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
    uint32_t n = be32(buf + off);
    off += 4 + 4ull * n;
  }
  if (len < off + 128) return -1;
  G2Aff ck;
  if (decompress_g2(ck, buf + off) || decompress_g2(ck, buf + off + 64)) return -1;  // parsed, unused
  return 0;
}

struct PlonkVkHost {
  uint64_t size, nb_public;
  Fr size_inv, generator, coset_shift;  // plain (non-Montgomery) limbs
  G1Aff s[3], ql, qr, qm, qo, qk, g1;
  std::vector<G1Aff> qcp;
  G2Aff g2[2];
  std::vector<uint64_t> cci;
};

inline int parse_plonk_vk(PlonkVkHost& vk, const uint8_t* buf, size_t len) {
  if (len < 372) return -1;
  vk.size = be64(buf);
  if (!fe_from_be_bytes(vk.size_inv, buf + 8)) return -1;
  if (!fe_from_be_bytes(vk.generator, buf + 40)) return -1;
  vk.nb_public = be64(buf + 72);
  if (!fe_from_be_bytes(vk.coset_shift, buf + 80)) return -1;
  G1Aff* pts[8] = {&vk.s[0], &vk.s[1], &vk.s[2], &vk.ql, &vk.qr, &vk.qm, &vk.qo, &vk.qk};
  for (int i = 0; i < 8; i++)
    if (decompress_g1(*pts[i], buf + 112 + 32 * i)) return -1;
  uint32_t nqcp = be32(buf + 368);
  size_t off = 372;
  if (nqcp > 64 || len < off + 32ull * nqcp + 160 + 33788 + 8) return -1;
  vk.qcp.resize(nqcp);
  for (uint32_t i = 0; i < nqcp; i++, off += 32)
    if (decompress_g1(vk.qcp[i], buf + off)) return -1;
  if (decompress_g1(vk.g1, buf + off) || decompress_g2(vk.g2[0], buf + off + 32) || decompress_g2(vk.g2[1], buf + off + 96))
    return -1;
  off += 160 + 33788;
  uint64_t nidx = be64(buf + off);
  off += 8;
  if (nidx > 64 || len < off + 8 * nidx) return -1;
  vk.cci.resize(nidx);
  for (uint64_t i = 0; i < nidx; i++) vk.cci[i] = be64(buf + off + 8 * i);
  return 0;
}

}  // namespace gnark
}  // namespace bn254
