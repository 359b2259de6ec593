// hash_to_curve.cuh -- mcl's hashAndMapToG1 for BLS12-381, per lane.
//
// Reference semantics (third-parties/mcl):
//   include/mcl/bn.hpp:2088-2097  hashAndMapToG1 = Fp::setHashOf + mapToG1
//   include/mcl/fp.hpp:430-435    Fp::setHashOf: op.hash = SHA-512 for fields wider than 256 bits (src/fp.cpp:552-556),
//                                 the first 48 digest bytes as a little-endian integer, masked to 381 bits and, if
//                                 still >= p, to 380 bits (copyAndMask SmallMask, src/fp.cpp:612-662)
//   include/mcl/bn.hpp:337-366    MapTo::calcBN: Shallue-van de Woestijne / Fouque-Tibouchi map with the sign of y
//                                 taken from the Legendre symbol of t; constants c1 = sqrt(-3), c2 = (c1-1)/2 (:470-489)
//   include/mcl/bn.hpp:422-425    mulByCofactorBLS12: multiplication by (z-1)^2/3 (mulGeneric: the point is not yet in
//                                 the r-torsion, so no GLV)
// The EL PASSO verifier needs H(service_name) once per batch (computed by the host); this batched version is for
// relying parties that mix many service names in one batch.  __host__ __device__ like the other lane functions.
#pragma once
#include "protocol.cuh"

namespace psb {

// ---- SHA-512 (FIPS 180-4), one-shot over a byte string --------------------------------------------------------
PSB_HD PSB_INL uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
PSB_HD PSB_INL uint64_t sha512_k(int i) {
  // first 64 bits of the fractional parts of the cube roots of the first 80 primes
  const uint64_t K[80] = {
      0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull, 0x3956c25bf348b538ull,
      0x59f111f1b605d019ull, 0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull, 0xd807aa98a3030242ull, 0x12835b0145706fbeull,
      0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull, 0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull,
      0xc19bf174cf692694ull, 0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull, 0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull,
      0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull, 0x983e5152ee66dfabull,
      0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull, 0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull,
      0x06ca6351e003826full, 0x142929670a0e6e70ull, 0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull,
      0x53380d139d95b3dfull, 0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull,
      0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull, 0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull, 0xd192e819d6ef5218ull,
      0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull, 0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull,
      0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull, 0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull, 0x5b9cca4f7763e373ull,
      0x682e6ff3d6b2b8a3ull, 0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull,
      0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull, 0xca273eceea26619cull,
      0xd186b8c721c0c207ull, 0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull, 0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull,
      0x113f9804bef90daeull, 0x1b710b35131c471bull, 0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull,
      0x431d67c49c100d4cull, 0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull, 0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull};
  return K[i];
}
PSB_HD PSB_NOINL void sha512_compress(uint64_t h[8], const uint64_t blk[16]) {
  uint64_t w[16];
  for (int i = 0; i < 16; i++) w[i] = blk[i];
  uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 80; i++) {
    if (i >= 16) {
      const uint64_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const uint64_t s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
      const uint64_t s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
      w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
    }
    const uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
    const uint64_t ch = (e & f) ^ (~e & g);
    const uint64_t t1 = hh + S1 + ch + sha512_k(i) + w[i & 15];
    const uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
    const uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
    const uint64_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
// digest words (big-endian 64-bit) of msg
PSB_HD PSB_NOINL void sha512(uint64_t h[8], const uint8_t* msg, size_t len) {
  h[0] = 0x6a09e667f3bcc908ull; h[1] = 0xbb67ae8584caa73bull; h[2] = 0x3c6ef372fe94f82bull; h[3] = 0xa54ff53a5f1d36f1ull;
  h[4] = 0x510e527fade682d1ull; h[5] = 0x9b05688c2b3e6c1full; h[6] = 0x1f83d9abfb41bd6bull; h[7] = 0x5be0cd19137e2179ull;
  uint64_t blk[16];
  for (int i = 0; i < 16; i++) blk[i] = 0;
  int fill = 0;
  // message bytes, then 0x80, zero padding, 128-bit big-endian bit length
  const size_t total = len + 1;
  for (size_t i = 0; i < total; i++) {
    const uint8_t byte = i < len ? msg[i] : 0x80;
    blk[fill >> 3] |= (uint64_t)byte << (56 - 8 * (fill & 7));
    if (++fill == 128) {
      sha512_compress(h, blk);
      for (int k = 0; k < 16; k++) blk[k] = 0;
      fill = 0;
    }
  }
  if (fill > 112) {
    sha512_compress(h, blk);
    for (int k = 0; k < 16; k++) blk[k] = 0;
  }
  blk[15] = (uint64_t)len * 8;   // messages are far below 2^61 bytes: the high length word stays 0
  sha512_compress(h, blk);
}

// Fp::setHashOf(msg): canonical Montgomery element
PSB_HD PSB_NOINL void fp_set_hash_of(Fp& t, const uint8_t* msg, size_t len) {
  Fp n;
#if PSB_FP_BITS > 256
  uint64_t h[8];
  sha512(h, msg, len);
  // the first 48 digest BYTES as a little-endian integer: digest byte j = byte (7 - j%8) of word j/8
  for (int i = 0; i < PSB_NL; i++) {
    uint32_t v = 0;
    for (int b = 0; b < 4; b++) {
      const int j = 4 * i + b;
      const uint8_t byte = (uint8_t)(h[j >> 3] >> (56 - 8 * (j & 7)));
      v |= (uint32_t)byte << (8 * b);
    }
    n.v[i] = v;
  }
#else
  // fields of at most 256 bits hash with SHA-256 (mcl/src/fp.cpp:552-556): 32 digest bytes, little-endian
  Sha256 s;
  sha256_init(s);
  sha256_update(s, msg, len);
  uint32_t d[8];
  sha256_final(s, d);
  for (int i = 0; i < PSB_NL; i++) n.v[i] = bswap32(d[i]);
#endif
  constexpr int top = PSB_FP_BITS - 32 * (PSB_NL - 1);              // bits of p in the top limb
  n.v[PSB_NL - 1] &= (1u << top) - 1u;                              // bitSize(p) bits
  uint32_t tmp[PSB_NL];
  if (sub_mod_n<FpT>(tmp, n.v) == 0) n.v[PSB_NL - 1] &= (1u << (top - 1)) - 1u;   // >= p: one bit less
  fp_to_mont(t, n);
}

// Legendre symbol by Euler's criterion: 1, -1 or 0  (mcl uses gmp::legendre; same value)
PSB_HD PSB_NOINL int fp_legendre(const Fp& a) {
  if (fp_is_zero(a)) return 0;
  Fp r, one;
  fp_pow_nib(r, a, PSB_K(FP_PM1D2_NIB));
  fp_set_one(one);
  return fp_eq(r, one) ? 1 : -1;
}

// MapTo::calcBN for G1.  false for the exceptional inputs (t = 0, t^2 + b + 1 = 0), like mcl.
PSB_HD PSB_NOINL bool map_to_g1(G1J& P, const Fp& t) {
  pt_set_zero(P);
  const int leg = fp_legendre(t);
  if (leg == 0) return false;
  const bool negative = leg < 0;
  Fp one, four, c1, c2, w, x, y, u;
  fp_set_one(one);
  fp_set_curve_b(four);                   // the curve constant b (4 for BLS12-381, 2 for BN254)
  for (int i = 0; i < PSB_NL; i++) { c1.v[i] = PSB_K(MAPTO_C1)[i]; c2.v[i] = PSB_K(MAPTO_C2)[i]; }
  fp_sqr(w, t);
  fp_add(w, w, four);
  fp_add(w, w, one);                      // t^2 + b + 1
  if (fp_is_zero(w)) return false;
  fp_inv(w, w);
  fp_mul(w, w, c1);
  fp_mul(w, w, t);                        // w = sqrt(-3) t / (1 + b + t^2)
  for (int i = 0; i < 3; i++) {
    if (i == 0) { fp_mul(x, t, w); fp_neg(x, x); fp_add(x, x, c2); }
    else if (i == 1) { fp_neg(x, x); fp_sub(x, x, one); }
    else { fp_sqr(x, w); fp_inv(x, x); fp_add(x, x, one); }
    fp_sqr(u, x); fp_mul(u, u, x); fp_add(u, u, four);     // x^3 + b
    if (fp_sqrt(y, u)) {
      if (negative) fp_neg(y, y);
      P.x = x; P.y = y; fp_set_one(P.z);
      return true;
    }
  }
  return false;
}

// P <- [(z-1)^2 / 3] P, plain windowed multiplication (valid for any curve point), normalised
PSB_HD PSB_NOINL void g1_clear_cofactor(G1J& R, const G1J& P) {
#if PSB_IS_BN
  R = P;                                  // BN curves: E(Fp) has prime order, the map output is already in G1 with z = 1
#else
  G1J T;
  pt_mul_window(T, P, PSB_K(G1_COFACTOR));
  pt_normalize(R, T);
#endif
}

// hashAndMapToG1(msg): normalised point; returns false only for the exceptional hash values mcl asserts against
PSB_HD PSB_NOINL bool hash_and_map_to_g1(G1J& P, const uint8_t* msg, size_t len) {
  Fp t;
  fp_set_hash_of(t, msg, len);
  G1J Q;
  if (!map_to_g1(Q, t)) { pt_set_zero(P); return false; }
  g1_clear_cofactor(P, Q);
  return true;
}

}  // namespace psb
