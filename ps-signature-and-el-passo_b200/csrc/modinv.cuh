// modinv.cuh -- Fp inversion by Bernstein-Yang "safegcd" division steps (constant schedule, branch-free).
//
// Replaces, for the batched path, mcl's Fp::inv (reference: third-parties/mcl/src/fp.cpp:215-246 fp_invMontOpC:
// a plain modular inverse by extended GCD on the host big integer, then a multiplication by R^3 to return to
// Montgomery form; include/mcl/fp.hpp:488-494).  The VALUE is the unique inverse in [0, p), so every byte downstream
// (normalised points, GT) is the same as with any other inversion algorithm; inv(0) = 0 like mcl's.
//
// Why not Fermat: a^(p-2) is ~476 Montgomery products = 143 k MAC32 on the integer-multiply pipe, all serial.  The
// half-delta divstep iteration (Bernstein & Yang, "Fast constant-time gcd computation and modular inversion", CHES
// 2019; bound floor((45907 b + 26313)/19929) steps for b-bit moduli) needs 30 batches (BLS12-381) of
//   30 divsteps on the low 30 bits of (f, g)      -- shifts / adds / masks on 32-bit words, no multiplies
//   one 2x2 matrix applied to (f, g) and (d, e)   -- 6 x 13 signed 32x32->64 MACs
// i.e. ~5 k wide MACs + ~25 k ALU instructions, with no data-dependent branch: every lane of a warp runs the same
// schedule.  Values live in 30-bit signed limbs (PSB_N30 of them, constants from tools/gen_constants.py).
#pragma once
#include "fp.cuh"

namespace psb {
namespace modinv {

constexpr int L = PSB_N30;
constexpr int32_t M30 = 0x3fffffff;

struct Mat { int32_t u, v, q, r; };

// 30 half-delta divsteps on the low bits of f (odd) and g; returns the new zeta = -(delta + 1/2) and the transition
// matrix t with  2^30 [f'; g'] = t [f; g]
PSB_HD PSB_INL int32_t divsteps30(int32_t zeta, uint32_t f, uint32_t g, Mat& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
#ifdef __CUDA_ARCH__
#pragma unroll 6
#endif
  for (int i = 0; i < 30; i++) {
    uint32_t m1 = (uint32_t)(zeta >> 31);        // delta > 0
    const uint32_t m2 = 0u - (g & 1u);           // g odd
    const uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;   // (f, u, v) negated when delta > 0
    g += x & m2; q += y & m2; r += z & m2;       // g <- g -+ f when g is odd
    m1 &= m2;                                    // swap: delta > 0 and g odd
    zeta = (int32_t)((uint32_t)zeta ^ m1) - 1;
    f += g & m1; u += q & m1; v += r & m1;       // f <- old g
    g >>= 1; u <<= 1; v <<= 1;
  }
  t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
  return zeta;
}

// (f, g) <- t (f, g) / 2^30   (exact)
PSB_HD PSB_INL void update_fg(int32_t* f, int32_t* g, const Mat& t) {
  int64_t cf = (int64_t)t.u * f[0] + (int64_t)t.v * g[0];
  int64_t cg = (int64_t)t.q * f[0] + (int64_t)t.r * g[0];
  cf >>= 30; cg >>= 30;
  PSB_UNROLL
  for (int i = 1; i < L; i++) {
    cf += (int64_t)t.u * f[i] + (int64_t)t.v * g[i];
    cg += (int64_t)t.q * f[i] + (int64_t)t.r * g[i];
    f[i - 1] = (int32_t)cf & M30; cf >>= 30;
    g[i - 1] = (int32_t)cg & M30; cg >>= 30;
  }
  f[L - 1] = (int32_t)cf;
  g[L - 1] = (int32_t)cg;
}

// (d, e) <- t (d, e) / 2^30 mod p, both kept in (-2p, p): a multiple of p is added so that the division is exact
PSB_HD PSB_INL void update_de(int32_t* d, int32_t* e, const Mat& t) {
  const int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
  int32_t md = (t.u & sd) + (t.v & se);
  int32_t me = (t.q & sd) + (t.r & se);
  int64_t cd = (int64_t)t.u * d[0] + (int64_t)t.v * e[0];
  int64_t ce = (int64_t)t.q * d[0] + (int64_t)t.r * e[0];
  md -= (int32_t)((PSB_P_INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((PSB_P_INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)(int32_t)PSB_K(FP_P30)[0] * md;
  ce += (int64_t)(int32_t)PSB_K(FP_P30)[0] * me;
  cd >>= 30; ce >>= 30;
  PSB_UNROLL
  for (int i = 1; i < L; i++) {
    const int32_t pi = (int32_t)PSB_K(FP_P30)[i];
    cd += (int64_t)t.u * d[i] + (int64_t)t.v * e[i];
    ce += (int64_t)t.q * d[i] + (int64_t)t.r * e[i];
    cd += (int64_t)pi * md;
    ce += (int64_t)pi * me;
    d[i - 1] = (int32_t)cd & M30; cd >>= 30;
    e[i - 1] = (int32_t)ce & M30; ce >>= 30;
  }
  d[L - 1] = (int32_t)cd;
  e[L - 1] = (int32_t)ce;
}

// r in (-2p, p) -> sign * r in [0, p), limbs in [0, 2^30)
PSB_HD PSB_INL void normalize(int32_t* r, int32_t sign_mask /* -1: negate */) {
  int32_t add = r[L - 1] >> 31;
  PSB_UNROLL
  for (int i = 0; i < L; i++) {
    int32_t x = r[i] + ((int32_t)PSB_K(FP_P30)[i] & add);      // (-p, p)
    r[i] = (x ^ sign_mask) - sign_mask;
  }
  PSB_UNROLL
  for (int i = 0; i < L - 1; i++) { r[i + 1] += r[i] >> 30; r[i] &= M30; }
  add = r[L - 1] >> 31;
  PSB_UNROLL
  for (int i = 0; i < L; i++) r[i] += (int32_t)PSB_K(FP_P30)[i] & add;   // [0, p)
  PSB_UNROLL
  for (int i = 0; i < L - 1; i++) { r[i + 1] += r[i] >> 30; r[i] &= M30; }
}

// 32-bit limbs (value < 2^(32 PSB_NL)) -> 30-bit limbs
PSB_HD PSB_INL void to30(int32_t* o, const uint32_t* x) {
  PSB_UNROLL
  for (int i = 0; i < L; i++) {
    const int bit = 30 * i, w = bit >> 5, s = bit & 31;
    uint32_t v = 0;
    if (w < PSB_NL) v = x[w] >> s;
    if (s > 2 && w + 1 < PSB_NL) v |= x[w + 1] << (32 - s);
    o[i] = (int32_t)(v & (uint32_t)M30);
  }
}
// 30-bit limbs in [0, 2^30), value < 2^(32 PSB_NL) -> 32-bit limbs
PSB_HD PSB_INL void from30(uint32_t* x, const int32_t* v) {
  PSB_UNROLL
  for (int j = 0; j < PSB_NL; j++) {
    const int bit = 32 * j, i = bit / 30, s = bit % 30;
    uint32_t w = (uint32_t)v[i] >> s;
    if (i + 1 < L) w |= (uint32_t)v[i + 1] << (30 - s);
    if (s > 28 && i + 2 < L) w |= (uint32_t)v[i + 2] << (60 - s);
    x[j] = w;
  }
}

// x^-1 mod p for a plain integer x in [0, p) (32-bit limbs); 0 -> 0
PSB_HD PSB_NOINL void inv_plain(uint32_t* out, const uint32_t* x) {
  int32_t d[L], e[L], f[L], g[L];
  PSB_UNROLL
  for (int i = 0; i < L; i++) { d[i] = 0; e[i] = 0; f[i] = (int32_t)PSB_K(FP_P30)[i]; }
  e[0] = 1;
  to30(g, x);
  int32_t zeta = -1;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int it = 0; it < PSB_BY_BATCHES; it++) {
    Mat t;
    zeta = divsteps30(zeta, (uint32_t)f[0], (uint32_t)g[0], t);
    update_de(d, e, t);
    update_fg(f, g, t);
  }
  // g = 0 now, f = +-gcd = +-1 (or +-p for x = 0, where d = 0), d = +-x^-1
  normalize(d, f[L - 1] >> 31);
  from30(out, d);
}

}  // namespace modinv

// a^-1 in Montgomery form: (aR)^-1 = a^-1 R^-1 as a plain integer, times R^3 through one Montgomery product
PSB_HD PSB_INL void fp_inv_by(Fp& r, const Fp& a) {
  Fp t, r3;
  PSB_OPCOUNT(INV);
  modinv::inv_plain(t.v, a.v);
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) r3.v[i] = PSB_K(FP_R3)[i];
  fp_mul(r, t, r3);
}

PSB_HD PSB_INL void fp_inv(Fp& r, const Fp& a) {
#ifdef PSB_INV_FERMAT
  fp_inv_fermat(r, a);
#else
  fp_inv_by(r, a);
#endif
}

}  // namespace psb
