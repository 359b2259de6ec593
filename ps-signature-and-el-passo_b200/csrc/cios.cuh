// cios.cuh -- device-only Montgomery multiplication kernels for the 381-bit field (12 x 32-bit limbs).
//
// Replaces mcl's x86-64 back-ends for Fp::mul / Fp2::mul / Fp2::sqr (reference:
// third-parties/mcl/src/fp_generator.hpp:827-868 gen_montMul6 (Xbyak JIT), src/asm/x86-64.bmi2.s
// mcl_fp_mont6Lbmi2, src/low_func.hpp:554-652; Fp2: include/mcl/fp_tower.hpp:528-550,713-741).
//
// Algorithm: operand-scanning CIOS Montgomery with the running sum split into an EVEN and an ODD
// accumulator array, T = E + 2^32 O.  The 64-bit product of an even multiplicand limb lands on an
// aligned limb pair of E, of an odd limb on an aligned pair of O, so one row is two carry chains
// of six wide MACs (mad.lo.cc + madc.hi.cc, which ptxas fuses into IMAD.WIDE.U32.X with predicate
// carries) and needs no separate carry-propagation instructions.  The per-row shift by one limb is
// folded into the MAC chain of the odd array (destination limb j, addend limb j+2) and the two
// arrays swap roles every row.  Several products can be accumulated in one row before the single
// reduction step of that row ("dot" products: a0 b0 - a1 b1 of an Fp2 product costs one reduction).
//
// Fp2 products / squares run two accumulator pairs (real, imaginary) side by side in one fully
// unrolled function: four independent carry chains give the IMAD pipe enough ILP at the low
// occupancy this register-heavy code runs at, and the multiplicand limbs are loaded once.
//
// Bounds: inputs < 2p (multiplicands of squarings are unreduced sums), sum of products < 4 p^2
// => T < 2p at every row and the final value < 2p: one conditional subtraction canonicalises.
#pragma once
#include "fp_cios.cuh"

namespace psb {
namespace cios {

typedef uint32_t L12[12];
#define PSB_EV(v) v[0], v[2], v[4], v[6], v[8], v[10]
#define PSB_OD(v) v[1], v[3], v[5], v[7], v[9], v[11]
#define PSB_ALL(v) v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]
#define PSB_X(m, ...) m(__VA_ARGS__)  // expand the limb lists before the row macro counts its arguments

// first row of an accumulator pair: E = a_even * b, O = a_odd * b
__device__ PSB_INL void first(L12& E, L12& O, const L12& a, uint32_t b) {
  PSB_X(PSB_ROW_FIRST_EVEN, PSB_ALL(E), PSB_EV(a), b);
  PSB_X(PSB_ROW_FIRST_ODD, PSB_ALL(O), PSB_OD(a), b);
}
// first product of a later row: merge the stray limb, shift the odd array, accumulate a * b
__device__ PSB_INL void mac_shift(L12& E, L12& O, const L12& a, uint32_t b) {
  PSB_X(PSB_ROW_ODD_RSHIFT, E[0], PSB_ALL(O), PSB_OD(a), b);
  PSB_X(PSB_ROW_EVEN, PSB_ALL(E), O[11], PSB_EV(a), b);
}
// further product of the same row, in place
__device__ PSB_INL void mac(L12& E, L12& O, const L12& a, uint32_t b) {
  PSB_X(PSB_ROW_ODD, PSB_ALL(O), PSB_OD(a), b);
  PSB_X(PSB_ROW_EVEN, PSB_ALL(E), O[11], PSB_EV(a), b);
}
// Montgomery step of the row: m = T mod 2^32 * (-p^-1), T += m p  (low limb of E becomes 0)
__device__ PSB_INL void reduce(L12& E, L12& O) {
  const uint32_t m = E[0] * PSB_FP_N0;
  PSB_X(PSB_RED_ODD, PSB_ALL(O), m);
  PSB_X(PSB_RED_EVEN, PSB_ALL(E), O[11], m);
}
// after the last row (even = E with E[0] == 0, odd = O): T = (E >> 32) + O, canonicalise, store
__device__ PSB_INL void finish(uint32_t* r, const L12& E, const L12& O) {
  uint32_t t[12];
  t[0] = ptx::add_cc(E[1], O[0]);
  PSB_UNROLL
  for (int i = 1; i < 11; i++) t[i] = ptx::addc_cc(E[i + 1], O[i]);
  t[11] = ptx::addc(0, O[11]);
  cond_sub_mod<FpT>(t);
  uint4* q = reinterpret_cast<uint4*>(r);   // every Fp is 16-byte aligned: three 128-bit stores
  q[0] = make_uint4(t[0], t[1], t[2], t[3]);
  q[1] = make_uint4(t[4], t[5], t[6], t[7]);
  q[2] = make_uint4(t[8], t[9], t[10], t[11]);
}
__device__ PSB_INL void load12(L12& d, const uint32_t* s) {
  const uint4* q = reinterpret_cast<const uint4*>(s);   // three 128-bit loads
  const uint4 v0 = q[0], v1 = q[1], v2 = q[2];
  d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w;
  d[4] = v1.x; d[5] = v1.y; d[6] = v1.z; d[7] = v1.w;
  d[8] = v2.x; d[9] = v2.y; d[10] = v2.z; d[11] = v2.w;
}

// r = a * b / R mod p
__device__ PSB_NOINL void mul(uint32_t* r, const uint32_t* a_, const uint32_t* b_) {
  L12 a, b, X, Y;
  load12(a, a_);
  load12(b, b_);
  first(X, Y, a, b[0]);
  reduce(X, Y);
  PSB_UNROLL
  for (int i = 1; i < 12; i += 2) {
    mac_shift(Y, X, a, b[i]);
    reduce(Y, X);
    if (i + 1 < 12) {
      mac_shift(X, Y, a, b[i + 1]);
      reduce(X, Y);
    }
  }
  finish(r, Y, X);
}

// r = (a b + c d) / R mod p, one reduction per row
__device__ PSB_NOINL void dot2(uint32_t* r, const uint32_t* a_, const uint32_t* b_, const uint32_t* c_, const uint32_t* d_) {
  L12 a, b, c, d, X, Y;
  load12(a, a_);
  load12(c, c_);
  load12(b, b_);
  load12(d, d_);
  first(X, Y, a, b[0]);
  mac(X, Y, c, d[0]);
  reduce(X, Y);
  PSB_UNROLL
  for (int i = 1; i < 12; i += 2) {
    mac_shift(Y, X, a, b[i]);
    mac(Y, X, c, d[i]);
    reduce(Y, X);
    if (i + 1 < 12) {
      mac_shift(X, Y, a, b[i + 1]);
      mac(X, Y, c, d[i + 1]);
      reduce(X, Y);
    }
  }
  finish(r, Y, X);
}

// Fp2 product: r = x * y,  x = a + c i,  y = y0 + y1 i
//   re = a y0 + (p - c) y1,   im = a y1 + c y0       (two accumulator pairs, 4 + 2 chains per row)
__device__ PSB_NOINL void fp2_mul(uint32_t* r, const uint32_t* x, const uint32_t* y) {
  L12 a, c, nc, X1, Y1, X2, Y2;
  load12(a, x);
  load12(c, x + 12);
  nc[0] = ptx::sub_cc(FpT::p(0), c[0]);
  PSB_UNROLL
  for (int i = 1; i < 11; i++) nc[i] = ptx::subc_cc(FpT::p(i), c[i]);
  nc[11] = ptx::subc(FpT::p(11), c[11]);   // p - c in (0, p]: fine as a multiplicand
  {
    const uint32_t y0 = y[0], y1 = y[12];
    first(X1, Y1, a, y0); mac(X1, Y1, nc, y1); reduce(X1, Y1);
    first(X2, Y2, a, y1); mac(X2, Y2, c, y0); reduce(X2, Y2);
  }
  PSB_UNROLL
  for (int i = 1; i < 12; i += 2) {
    {
      const uint32_t y0 = y[i], y1 = y[12 + i];
      mac_shift(Y1, X1, a, y0); mac(Y1, X1, nc, y1); reduce(Y1, X1);
      mac_shift(Y2, X2, a, y1); mac(Y2, X2, c, y0); reduce(Y2, X2);
    }
    if (i + 1 < 12) {
      const uint32_t y0 = y[i + 1], y1 = y[13 + i];
      mac_shift(X1, Y1, a, y0); mac(X1, Y1, nc, y1); reduce(X1, Y1);
      mac_shift(X2, Y2, a, y1); mac(X2, Y2, c, y0); reduce(X2, Y2);
    }
  }
  finish(r, Y1, X1);
  finish(r + 12, Y2, X2);
}

// Fp2 square: (a + b i)^2 = (a + b)(a - b) + (2a) b i    (multiplicands a+b, 2a are unreduced, < 2p)
__device__ PSB_NOINL void fp2_sqr(uint32_t* r, const uint32_t* x) {
  L12 a, b, s, t, d, X1, Y1, X2, Y2;
  load12(a, x);
  load12(b, x + 12);
  add_n<12>(s, a, b);
  add_n<12>(t, a, a);
  mod_sub<FpT>(d, a, b);
  first(X1, Y1, s, d[0]); reduce(X1, Y1);
  first(X2, Y2, t, b[0]); reduce(X2, Y2);
  PSB_UNROLL
  for (int i = 1; i < 12; i += 2) {
    mac_shift(Y1, X1, s, d[i]); reduce(Y1, X1);
    mac_shift(Y2, X2, t, b[i]); reduce(Y2, X2);
    if (i + 1 < 12) {
      mac_shift(X1, Y1, s, d[i + 1]); reduce(X1, Y1);
      mac_shift(X2, Y2, t, b[i + 1]); reduce(X2, Y2);
    }
  }
  finish(r, Y1, X1);
  finish(r + 12, Y2, X2);
}

// Fp2 * Fp: (a + c i) k
__device__ PSB_NOINL void fp2_mul_fp(uint32_t* r, const uint32_t* x, const uint32_t* k) {
  L12 a, c, X1, Y1, X2, Y2;
  load12(a, x);
  load12(c, x + 12);
  first(X1, Y1, a, k[0]); reduce(X1, Y1);
  first(X2, Y2, c, k[0]); reduce(X2, Y2);
  PSB_UNROLL
  for (int i = 1; i < 12; i += 2) {
    mac_shift(Y1, X1, a, k[i]); reduce(Y1, X1);
    mac_shift(Y2, X2, c, k[i]); reduce(Y2, X2);
    if (i + 1 < 12) {
      mac_shift(X1, Y1, a, k[i + 1]); reduce(X1, Y1);
      mac_shift(X2, Y2, c, k[i + 1]); reduce(X2, Y2);
    }
  }
  finish(r, Y1, X1);
  finish(r + 12, Y2, X2);
}

}  // namespace cios
}  // namespace psb
