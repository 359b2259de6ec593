// sha256.cuh -- SHA-256 (FIPS 180-4) and the reference's hash-to-scalar rule.
//
// Replaces for the batched path cybozu::Sha256 (reference: third-parties/mcl/include/cybozu/sha2.hpp)
// and Fr::setHashOf (mcl/include/mcl/fp.hpp:430-435 -> fp.cpp:552-556 sha256 -> fp.cpp:612-662
// copyAndMask(SmallMask)): digest read as a LITTLE-endian integer, masked to 255 bits, and to 254
// bits if still >= r.  Not a modular reduction.
#pragma once
#include "fp.cuh"

PSB_CONST_ARRAY(SHA_K, 64, 0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u)

namespace psb {

struct Sha256 {
  uint32_t h[8];
  uint32_t w[16];   // current block, big-endian words
  uint32_t fill;    // bytes in the current block
  uint64_t total;   // bytes absorbed
};

PSB_HD PSB_INL uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

PSB_HD PSB_INL uint32_t sha_k(int i) { return PSB_K(SHA_K)[i]; }

PSB_HD PSB_INL void sha256_init(Sha256& s) {
  s.h[0] = 0x6a09e667u; s.h[1] = 0xbb67ae85u; s.h[2] = 0x3c6ef372u; s.h[3] = 0xa54ff53au;
  s.h[4] = 0x510e527fu; s.h[5] = 0x9b05688cu; s.h[6] = 0x1f83d9abu; s.h[7] = 0x5be0cd19u;
  for (int i = 0; i < 16; i++) s.w[i] = 0;
  s.fill = 0;
  s.total = 0;
}

PSB_HD PSB_NOINL void sha256_compress(Sha256& s) {
  uint32_t w[16];
  for (int i = 0; i < 16; i++) w[i] = s.w[i];
  uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t wi;
    if (i < 16) {
      wi = w[i];
    } else {
      const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
      w[i & 15] = wi;
    }
    const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    const uint32_t ch = (e & f) ^ (~e & g);
    const uint32_t t1 = h + S1 + ch + sha_k(i) + wi;
    const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    const uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
  for (int i = 0; i < 16; i++) s.w[i] = 0;
  s.fill = 0;
}

PSB_HD PSB_INL void sha256_put(Sha256& s, uint8_t byte) {
  s.w[s.fill >> 2] |= (uint32_t)byte << (24 - 8 * (s.fill & 3));
  s.fill++;
  s.total++;
  if (s.fill == 64) sha256_compress(s);
}
PSB_HD PSB_INL void sha256_update(Sha256& s, const uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; i++) sha256_put(s, p[i]);
}
// state words (big-endian digest words) after padding
PSB_HD PSB_INL void sha256_final(Sha256& s, uint32_t out[8]) {
  const uint64_t bits = s.total * 8;
  sha256_put(s, 0x80);
  while (s.fill != 56) sha256_put(s, 0);
  s.w[14] = (uint32_t)(bits >> 32);
  s.w[15] = (uint32_t)bits;
  sha256_compress(s);
  for (int i = 0; i < 8; i++) out[i] = s.h[i];
}
PSB_HD PSB_INL uint32_t bswap32(uint32_t x) {
  return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
}

// SmallMask on the digest (mcl fp.cpp:612-662): little-endian integer of the 32 digest bytes, masked to bitSize(r)
// bits (255 for BLS12-381, 254 for BN254) and to one bit less if still >= r.  Output: NORMAL form scalar (8 LE limbs).
constexpr uint32_t kFrMask1 = (1u << (PSB_FR_BITS - 224)) - 1u;
constexpr uint32_t kFrMask2 = (1u << (PSB_FR_BITS - 225)) - 1u;
PSB_HD PSB_INL void digest_to_fr_normal(uint32_t k[8], const uint32_t digest_words[8]) {
  for (int i = 0; i < 8; i++) k[i] = bswap32(digest_words[i]);
  k[7] &= kFrMask1;
  if (fr_geq_modulus(k)) k[7] &= kFrMask2;
}
// 32 raw little-endian bytes (e.g. a CSPRNG draw, mcl Fr::setByCSPRNG fp.hpp:408-414) -> scalar
PSB_HD PSB_INL void bytes_to_fr_normal(uint32_t k[8], const uint8_t* b) {
  for (int i = 0; i < 8; i++)
    k[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  k[7] &= kFrMask1;
  if (fr_geq_modulus(k)) k[7] &= kFrMask2;
}

// Fr::setHashOf(msg) -> normal-form scalar
PSB_HD PSB_INL void fr_set_hash_of(uint32_t k[8], const uint8_t* msg, size_t len) {
  Sha256 s;
  sha256_init(s);
  sha256_update(s, msg, len);
  uint32_t d[8];
  sha256_final(s, d);
  digest_to_fr_normal(k, d);
}

}  // namespace psb
