// SHA-256 (FIPS 180-4) for the on-device Fiat-Shamir transcript and hash-to-field.
// Replaces the `sha2` crate as used by Transcript::compute_challenge (reference
// verifier/src/transcript.rs:68-107) and WrappedHashToField (verifier/src/hash_to_field.rs:30-97).
// __host__ __device__: the host build is used once per VK to cache the midstate of the VK-constant
// prefix of the gamma transcript.
#pragma once
#include "field.cuh"

namespace bn254 {

struct Sha256 {
  uint32_t h[8];
  uint32_t w[16];   // current block, big-endian words
  uint32_t fill;    // bytes buffered in w (0..63)
  uint32_t blocks;  // full blocks compressed so far
};

HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

HDN void sha256_compress(uint32_t* h, uint32_t* w) {
  const uint32_t K256[64] = {
      0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u,
      0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u,
      0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u,
      0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u,
      0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u,
      0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au,
      0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u,
      0xc67178f2u};
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
  for (int t = 0; t < 64; t++) {
    uint32_t wt;
    if (t < 16) {
      wt = w[t];
    } else {
      uint32_t w15 = w[(t - 15) & 15], w2 = w[(t - 2) & 15];
      uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      wt = w[t & 15] + s0 + w[(t - 7) & 15] + s1;
      w[t & 15] = wt;
    }
    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + K256[t] + wt;
    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g, g = f, f = e, e = d + t1, d = c, c = b, b = a, a = t1 + t2;
  }
  h[0] += a, h[1] += b, h[2] += c, h[3] += d, h[4] += e, h[5] += f, h[6] += g, h[7] += hh;
}

HD void sha256_init(Sha256& s) {
  s.h[0] = 0x6a09e667u, s.h[1] = 0xbb67ae85u, s.h[2] = 0x3c6ef372u, s.h[3] = 0xa54ff53au;
  s.h[4] = 0x510e527fu, s.h[5] = 0x9b05688cu, s.h[6] = 0x1f83d9abu, s.h[7] = 0x5be0cd19u;
  for (int i = 0; i < 16; i++) s.w[i] = 0;
  s.fill = 0;
  s.blocks = 0;
}

HD void sha256_put(Sha256& s, uint8_t byte) {
  uint32_t i = s.fill >> 2, sh = 24 - 8 * (s.fill & 3);
  s.w[i] = (s.w[i] & ~(0xffu << sh)) | ((uint32_t)byte << sh);
  if (++s.fill == 64) {
    sha256_compress(s.h, s.w);
    s.fill = 0;
    s.blocks++;
  }
}

HDN void sha256_update(Sha256& s, const uint8_t* data, uint32_t len) {
  for (uint32_t i = 0; i < len; i++) sha256_put(s, data[i]);
}

// digest as 8 big-endian words / 32 bytes; `s` is consumed
HDN void sha256_final(Sha256& s, uint8_t* out) {
  uint64_t bits = ((uint64_t)s.blocks * 64 + s.fill) * 8;
  sha256_put(s, 0x80);
  while (s.fill != 56) sha256_put(s, 0);
  for (int i = 7; i >= 0; i--) sha256_put(s, (uint8_t)(bits >> (8 * i)));
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(s.h[i] >> 24);
    out[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
    out[4 * i + 2] = (uint8_t)(s.h[i] >> 8);
    out[4 * i + 3] = (uint8_t)s.h[i];
  }
}

}  // namespace bn254
