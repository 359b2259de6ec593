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

// ---- split form: wide product and stand-alone reduction (lazy reduction of the Fp2 engine, tower.cuh) -------------
// T (2N limbs) = a * b without reduction (mcl FpDbl::mulPre, fp_tower.hpp:13-178): the rows of the fused multiplier
// with the reduction steps left out -- after each row the low limb of the window is final and is peeled off before the
// window shifts.  Any N-limb operands (the Karatsuba middle term multiplies sums up to 4p).  N^2 wide MACs.
__device__ PSB_INL void mulpre_rr(uint32_t* T, const L12& a, const L12& b) {
  L12 X, Y;
  first(X, Y, a, b[0]);
  T[0] = X[0];
  PSB_UNROLL
  for (int i = 1; i < PSB_NL; i += 2) {
    mac_shift(Y, X, a, b[i]);
    T[i] = Y[0];
    if (i + 1 < PSB_NL) {
      mac_shift(X, Y, a, b[i + 1]);
      T[i + 1] = X[0];
    }
  }
  T[PSB_NL] = ptx::add_cc(Y[1], X[0]);                       // high half = (Y >> 32) + X
  PSB_UNROLL
  for (int i = 1; i < PSB_TOP; i++) T[PSB_NL + i] = ptx::addc_cc(Y[i + 1], X[i]);
  T[PSB_NL + PSB_TOP] = ptx::addc_cc(0, X[PSB_TOP]);         // (carry out is 0 by the value bound: see chain_after)
}
// Orders two independent products for ptxas: x picks up the (always zero) carry that the previous mulpre_rr left, so the
// next product cannot start before the previous one has finished.  Left to itself ptxas interleaves independent
// products for ILP the multiply pipe cannot use (one wide MAC per 4 cycles and warp) and runs out of predicate
// registers for their carry chains: 160 LOP3 of predicate spills in one Fp2 product.
__device__ PSB_INL void chain_after(uint32_t& x) { asm volatile("addc.u32 %0, %0, 0;" : "+r"(x)); }
// r = T / R mod p for a 2N-limb T < pR (mcl FpDbl::mod, low_func.hpp:511-547), canonical.  The window holds the low
// half only: r = (T_lo + M p) / R + T_hi < p + 1 + T_hi <= 2p.  The shift of the window is folded into the m p_odd
// chain of the next row (PSB_RED_ODD_RSHIFT), as the fused multiplier folds it into its a_odd b chain: N^2 wide MACs
// + N plain ones, like the reduction half of the fused form.
__device__ PSB_INL void redc_rr(uint32_t* r, const uint32_t* T) {
  L12 X, Y;
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) X[i] = T[i];
  uint32_t m = X[0] * PSB_FP_N0;
  PSB_X(PSB_RED_FIRST_ODD, PSB_ALL(Y), m);
  PSB_X(PSB_RED_EVEN, PSB_ALL(X), Y[PSB_TOP], m);
  PSB_UNROLL
  for (int i = 1; i < PSB_NL; i += 2) {
    PSB_X(PSB_RED_ODD_RSHIFT, m, Y[0], PSB_ALL(X));
    PSB_X(PSB_RED_EVEN, PSB_ALL(Y), X[PSB_TOP], m);
    if (i + 1 < PSB_NL) {
      PSB_X(PSB_RED_ODD_RSHIFT, m, X[0], PSB_ALL(Y));
      PSB_X(PSB_RED_EVEN, PSB_ALL(X), Y[PSB_TOP], m);
    }
  }
  uint32_t t[PSB_NL];
  t[0] = ptx::add_cc(Y[1], X[0]);                             // (Y >> 32) + X   (Y[0] == 0)
  PSB_UNROLL
  for (int i = 1; i < PSB_TOP; i++) t[i] = ptx::addc_cc(Y[i + 1], X[i]);
  t[PSB_TOP] = ptx::addc(0, X[PSB_TOP]);
  r[0] = ptx::add_cc(t[0], T[PSB_NL]);                        // + T_hi
  PSB_UNROLL
  for (int i = 1; i < PSB_TOP; i++) r[i] = ptx::addc_cc(t[i], T[PSB_NL + i]);
  r[PSB_TOP] = ptx::addc(t[PSB_TOP], T[PSB_NL + PSB_TOP]);
  cond_sub_mod<FpT>(r);
}

// T (2N limbs) = a^2: the cross products once (N (N - 1) / 2 wide MACs on the EV / OD accumulator arrays, generated rows
// PSB_SQR_CROSS), merged and doubled, plus the N diagonal products -- 78 wide MACs for N = 12 instead of 144
// (mcl sqrPre, low_func.hpp:554-652 / fp_generator.hpp gen_sqr).  Any N-limb operand.
__device__ PSB_INL void sqrpre_rr(uint32_t* T, const L12& a) {
  uint32_t EV[2 * PSB_NL], OD[2 * PSB_NL];
  PSB_UNROLL
  for (int i = 0; i < 2 * PSB_NL; i++) { EV[i] = 0; OD[i] = 0; }
  PSB_SQR_CROSS(a, EV, OD);
  T[0] = EV[0];                                               // T = EV + (OD << 32)
  T[1] = ptx::add_cc(EV[1], OD[0]);
  PSB_UNROLL
  for (int i = 2; i < 2 * PSB_NL - 1; i++) T[i] = ptx::addc_cc(EV[i], OD[i - 1]);
  T[2 * PSB_NL - 1] = ptx::addc(EV[2 * PSB_NL - 1], OD[2 * PSB_NL - 2]);
  T[0] = ptx::add_cc(T[0], T[0]);                             // T = 2 T  (cross sum < 2^(64 N - 1))
  PSB_UNROLL
  for (int i = 1; i < 2 * PSB_NL - 1; i++) T[i] = ptx::addc_cc(T[i], T[i]);
  T[2 * PSB_NL - 1] = ptx::addc(T[2 * PSB_NL - 1], T[2 * PSB_NL - 1]);
  PSB_SQR_DIAG(a, T);
}
// r = a^2 / R mod p, canonical, for a < 2p (T < 4 p^2 < pR): N (N + 1) / 2 + N^2 + N wide MACs (234 against the product's 300)
__device__ PSB_INL void sqr_rr(uint32_t* r, const L12& a) {
  uint32_t T[2 * PSB_NL];
  sqrpre_rr(T, a);
  redc_rr(r, T);
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
// r = a * a / R mod p  (memory operand; r may alias a): the dedicated squaring
__device__ PSB_NOINL void sqr(uint32_t* r, const uint32_t* a_) {
  L12 a, t;
  load12(a, a_);
  sqr_rr(t, a);
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
