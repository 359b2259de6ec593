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
#ifdef PSB_BUILD_BN254
#include "fp_cios_bn254.cuh"
#else
#include "fp_cios.cuh"
#endif

namespace psb {
namespace cios {

typedef uint32_t L12[PSB_NL];   // one field element in registers (12 limbs for BLS12-381, 8 for BN254)
#if PSB_NL == 12
#define PSB_EV(v) v[0], v[2], v[4], v[6], v[8], v[10]
#define PSB_OD(v) v[1], v[3], v[5], v[7], v[9], v[11]
#define PSB_ALL(v) v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]
#else
#define PSB_EV(v) v[0], v[2], v[4], v[6]
#define PSB_OD(v) v[1], v[3], v[5], v[7]
#define PSB_ALL(v) v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]
#endif
#define PSB_TOP (PSB_NL - 1)
#define PSB_X(m, ...) m(__VA_ARGS__)  // expand the limb lists before the row macro counts its arguments

// first row of an accumulator pair: E = a_even * b, O = a_odd * b
__device__ PSB_INL void first(L12& E, L12& O, const L12& a, uint32_t b) {
  PSB_X(PSB_ROW_FIRST_EVEN, PSB_ALL(E), PSB_EV(a), b);
  PSB_X(PSB_ROW_FIRST_ODD, PSB_ALL(O), PSB_OD(a), b);
}
// first product of a later row: merge the stray limb, shift the odd array, accumulate a * b
__device__ PSB_INL void mac_shift(L12& E, L12& O, const L12& a, uint32_t b) {
  PSB_X(PSB_ROW_ODD_RSHIFT, E[0], PSB_ALL(O), PSB_OD(a), b);
  PSB_X(PSB_ROW_EVEN, PSB_ALL(E), O[PSB_TOP], PSB_EV(a), b);
}
// further product of the same row, in place
__device__ PSB_INL void mac(L12& E, L12& O, const L12& a, uint32_t b) {
  PSB_X(PSB_ROW_ODD, PSB_ALL(O), PSB_OD(a), b);
  PSB_X(PSB_ROW_EVEN, PSB_ALL(E), O[PSB_TOP], PSB_EV(a), b);
}
// Montgomery step of the row: m = T mod 2^32 * (-p^-1), T += m p  (low limb of E becomes 0)
__device__ PSB_INL void reduce(L12& E, L12& O) {
  const uint32_t m = E[0] * PSB_FP_N0;
  PSB_X(PSB_RED_ODD, PSB_ALL(O), m);
  PSB_X(PSB_RED_EVEN, PSB_ALL(E), O[PSB_TOP], m);
}
// after the last row (even = E with E[0] == 0, odd = O): T = (E >> 32) + O, canonicalised into registers
__device__ PSB_INL void finish_rr(uint32_t* t, const L12& E, const L12& O) {
  t[0] = ptx::add_cc(E[1], O[0]);
  PSB_UNROLL
  for (int i = 1; i < PSB_TOP; i++) t[i] = ptx::addc_cc(E[i + 1], O[i]);
  t[PSB_TOP] = ptx::addc(0, O[PSB_TOP]);
  cond_sub_mod<FpT>(t);
}
__device__ PSB_INL void finish(uint32_t* r, const L12& E, const L12& O) {
  uint32_t t[PSB_NL];
  finish_rr(t, E, O);
  uint4* q = reinterpret_cast<uint4*>(r);   // every Fp is 16-byte aligned: three 128-bit stores
  q[0] = make_uint4(t[0], t[1], t[2], t[3]);
  q[1] = make_uint4(t[4], t[5], t[6], t[7]);
#if PSB_NL == 12
  q[2] = make_uint4(t[8], t[9], t[10], t[11]);
#endif
}
__device__ PSB_INL void load12(L12& d, const uint32_t* s) {
  const uint4* q = reinterpret_cast<const uint4*>(s);   // three 128-bit loads
  const uint4 v0 = q[0], v1 = q[1];
  d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w;
  d[4] = v1.x; d[5] = v1.y; d[6] = v1.z; d[7] = v1.w;
#if PSB_NL == 12
  const uint4 v2 = q[2];
  d[8] = v2.x; d[9] = v2.y; d[10] = v2.z; d[11] = v2.w;
#endif
}

// register-level cores: operands and result are register arrays (no memory operands), fully inlined into the
// fused tower functions (tower.cuh) so that sums / differences feed the multiplier without a stack round trip
__device__ PSB_INL void mul_rr(uint32_t* r, const L12& a, const L12& b) {
  L12 X, Y;
  first(X, Y, a, b[0]);
  reduce(X, Y);
  PSB_UNROLL
  for (int i = 1; i < PSB_NL; i += 2) {
    mac_shift(Y, X, a, b[i]);
    reduce(Y, X);
    if (i + 1 < PSB_NL) {
      mac_shift(X, Y, a, b[i + 1]);
      reduce(X, Y);
    }
  }
  finish_rr(r, Y, X);
}
// r = (a b + c d) / R mod p, one reduction per row
__device__ PSB_INL void dot2_rr(uint32_t* r, const L12& a, const L12& b, const L12& c, const L12& d) {
  L12 X, Y;
  first(X, Y, a, b[0]);
  mac(X, Y, c, d[0]);
  reduce(X, Y);
  PSB_UNROLL
  for (int i = 1; i < PSB_NL; i += 2) {
    mac_shift(Y, X, a, b[i]);
    mac(Y, X, c, d[i]);
    reduce(Y, X);
    if (i + 1 < PSB_NL) {
      mac_shift(X, Y, a, b[i + 1]);
      mac(X, Y, c, d[i + 1]);
      reduce(X, Y);
    }
  }
  finish_rr(r, Y, X);
}

// ---- rolled cores (EXPERIMENT, not on the product path) ------------------------------------------------
// Rolled loops keep the multiplicands in registers and stream the multiplier limbs from memory, two rows per
// iteration.  Measured (r1f): the per-row shift of the accumulator arrays is free register renaming in unrolled code
// but costs ~24 IMAD.MOV per array per iteration at the loop back edge (+45 % instructions) -> kept for reference only.
__device__ PSB_INL void zero12(L12& x) {
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) x[i] = 0;
}
__device__ PSB_INL uint2 ld2(const uint32_t* p) { return *reinterpret_cast<const uint2*>(p); }   // limb pairs are 8-byte aligned

// two independent products side by side:  r1 = a1 * b1,  r2 = a2 * b2   (b1, b2 in memory)
__device__ PSB_INL void mul2_loop(uint32_t* r1, uint32_t* r2, const L12& a1, const uint32_t* b1, const L12& a2, const uint32_t* b2) {
  L12 X1, Y1, X2, Y2;
  zero12(X1); zero12(Y1); zero12(X2); zero12(Y2);
#pragma unroll 1
  for (int i = 0; i < PSB_NL; i += 2) {
    const uint2 p = ld2(b1 + i), q = ld2(b2 + i);
    mac_shift(X1, Y1, a1, p.x); reduce(X1, Y1);
    mac_shift(X2, Y2, a2, q.x); reduce(X2, Y2);
    mac_shift(Y1, X1, a1, p.y); reduce(Y1, X1);
    mac_shift(Y2, X2, a2, q.y); reduce(Y2, X2);
  }
  finish_rr(r1, Y1, X1);
  finish_rr(r2, Y2, X2);
}
// Fp2 product with multiplicands xa, xb, nxb = -xb (mod p, any representative < 2p) in registers and the
// multiplier y = (ya, yb) in memory:  re = xa ya + nxb yb,  im = xa yb + xb ya   (four carry chains per row)
__device__ PSB_INL void fp2mul_loop(uint32_t* re, uint32_t* im, const L12& xa, const L12& xb, const L12& nxb,
                                    const uint32_t* ya, const uint32_t* yb) {
  L12 X1, Y1, X2, Y2;
  zero12(X1); zero12(Y1); zero12(X2); zero12(Y2);
#pragma unroll 1
  for (int i = 0; i < PSB_NL; i += 2) {
    const uint2 p = ld2(ya + i), q = ld2(yb + i);
    mac_shift(X1, Y1, xa, p.x); mac(X1, Y1, nxb, q.x); reduce(X1, Y1);
    mac_shift(X2, Y2, xa, q.x); mac(X2, Y2, xb, p.x); reduce(X2, Y2);
    mac_shift(Y1, X1, xa, p.y); mac(Y1, X1, nxb, q.y); reduce(Y1, X1);
    mac_shift(Y2, X2, xa, q.y); mac(Y2, X2, xb, p.y); reduce(Y2, X2);
  }
  finish_rr(re, Y1, X1);
  finish_rr(im, Y2, X2);
}
// single product / two-product dot with streamed multipliers
__device__ PSB_INL void mul_loop(uint32_t* r, const L12& a, const uint32_t* b) {
  L12 X, Y;
  zero12(X); zero12(Y);
#pragma unroll 1
  for (int i = 0; i < PSB_NL; i += 2) {
    const uint2 p = ld2(b + i);
    mac_shift(X, Y, a, p.x); reduce(X, Y);
    mac_shift(Y, X, a, p.y); reduce(Y, X);
  }
  finish_rr(r, Y, X);
}
__device__ PSB_INL void dot2_loop(uint32_t* r, const L12& a, const uint32_t* b, const L12& c, const uint32_t* d) {
  L12 X, Y;
  zero12(X); zero12(Y);
#pragma unroll 1
  for (int i = 0; i < PSB_NL; i += 2) {
    const uint2 p = ld2(b + i), q = ld2(d + i);
    mac_shift(X, Y, a, p.x); mac(X, Y, c, q.x); reduce(X, Y);
    mac_shift(Y, X, a, p.y); mac(Y, X, c, q.y); reduce(Y, X);
  }
  finish_rr(r, Y, X);
}

__device__ PSB_INL void store12(uint32_t* r, const L12& t) {
  uint4* q = reinterpret_cast<uint4*>(r);   // every Fp is 16-byte aligned: three 128-bit stores
  q[0] = make_uint4(t[0], t[1], t[2], t[3]);
  q[1] = make_uint4(t[4], t[5], t[6], t[7]);
#if PSB_NL == 12
  q[2] = make_uint4(t[8], t[9], t[10], t[11]);
#endif
}
// r = a * b / R mod p  (memory operands; r may alias a or b)
__device__ PSB_NOINL void mul(uint32_t* r, const uint32_t* a_, const uint32_t* b_) {
  L12 a, b, t;
  load12(a, a_);
  load12(b, b_);
  mul_rr(t, a, b);
  store12(r, t);
}
// r = (a b + c d) / R mod p  (memory operands)
__device__ PSB_NOINL void dot2(uint32_t* r, const uint32_t* a_, const uint32_t* b_, const uint32_t* c_, const uint32_t* d_) {
  L12 a, b, c, d, t;
  load12(a, a_);
  load12(c, c_);
  load12(b, b_);
  load12(d, d_);
  dot2_rr(t, a, b, c, d);
  store12(r, t);
}

}  // namespace cios
}  // namespace psb
