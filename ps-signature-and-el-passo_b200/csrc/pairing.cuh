// pairing.cuh -- optimal-ate pairing on BLS12-381 (M-type twist, loop over |z|) and BN254 (D-type twist, loop over
// |6z+2| plus the two Frobenius additions): a two-pairing ("multi") Miller loop and mcl's final exponentiation.
//
// Reference semantics (third-parties/mcl/include/mcl/bn.hpp): millerLoop :1660-1710,
// precomputeG2 :1719-1758, precomputedMillerLoop2mixed :1827-1897, finalExp :1643-1659 =
// mapToCyclotomic :1494-1502 + expHardPartBLS12 :1508-1555 (exponent 3(p^4-p^2+1)/r: SURVEY F3),
// pow_z :1150-1176, cyclotomic squaring :1075-1144.
//
// Only the final-exponentiation OUTPUT is canonical; the Miller value differs from mcl's by
// factors in proper subfields (our line functions are scaled differently: homogeneous projective
// coordinates after Costello-Lange-Naehrig instead of mcl's Jacobian), which the final
// exponentiation erases.  GT bytes therefore match mcl bit for bit.
#pragma once
#include "curve.cuh"

namespace psb {

// one precomputed line of the fixed G2 argument (affine slope form):
//   l(P) = c0 + (nl * xP) w^2 + yP w^3,   nl = -lambda,  c0 = lambda*xT - yT
struct FixedLine { Fp2 nl, c0; };
// doublings + additions of the loop parameter (BLS12-381: |z| = 0xd201000000010000 -> 63 + 5; BN254: |6z+2| =
// 0x18300000000000004 -> 64 + 4, plus the two closing additions of pi(Q) and -pi^2(Q), bn.hpp:1698-1709)
constexpr int kMillerSteps = PSB_ML_NBITS + PSB_ML_ADDS + (PSB_IS_BN ? 2 : 0);
// a fixed-line table holds one trailing slot whose first word says whether the lines were scaled to constant term 1
// (precompute_fixed_lines): any factor in Fp2 dies in the final exponentiation, and a line 1 + c2 w^2 + c3 w^3 costs
// 9 Fp2 products instead of 13
constexpr int kFixedLineSlots = kMillerSteps + 1;

struct G2H { Fp2 x, y, z; };  // homogeneous projective: (X/Z, Y/Z)

PSB_HD PSB_INL bool z_bit(int i) { return (PSB_Z_ABS >> i) & 1ull; }
// bit i of the Miller loop parameter (65 bits for BN254)
PSB_HD PSB_INL bool ml_bit(int i) { return i >= 64 ? ((PSB_ML_HI >> (i - 64)) & 1ull) : ((PSB_ML_LO >> i) & 1ull); }

// 3 b' c with b' the constant of the twist E'
#if PSB_TWIST_MTYPE
// b' = 4 xi:  (4 xi) * 3 * c = 12 xi c
PSB_HD PSB_INL void fp2_mul_3bt(Fp2& r, const Fp2& c) {
  Fp2 t, u;
  fp2_mul_xi(t, c);
  fp2_dbl(t, t); fp2_dbl(t, t);      // 4 xi c
  fp2_dbl(u, t); fp2_add(r, u, t);   // 12 xi c
}
#else
// b' = 2 / xi = 1 - i (mcl twist_b shortcut, bn.hpp:950-953):  (a + b i)(1 - i) = (a + b) + (b - a) i
PSB_HD PSB_INL void fp2_mul_3bt(Fp2& r, const Fp2& c) {
  Fp2 t, u;
  fp_add(t.a, c.a, c.b);
  fp_sub(t.b, c.b, c.a);
  fp2_dbl(u, t); fp2_add(r, u, t);
}
#endif

// f *= line, the line given as the three Fp2 values the step functions produce: c0 (constant term), c2 (the xP term),
// c3 (the yP term).  M-type twist: c0 + c2 w^2 + c3 w^3.  D-type twist: the untwist multiplies instead of dividing,
// l(P) = yP - lambda xP w + (lambda xT - yT) w^3, i.e. the same three values sit at w^3, w^1 and w^0.
PSB_HD PSB_INL void ml_mul_line(Fp12& f, const Fp2& c0, const Fp2& c2, const Fp2& c3) {
#if PSB_TWIST_MTYPE
  fp12_mul_line(f, c0, c2, c3);
#else
  fp12_mul_line_d(f, c3, c2, c0);
#endif
}

// f = line for f == 1 (all other coefficients of f are already 0): the same slots ml_mul_line multiplies by
PSB_HD PSB_INL void ml_set_line(Fp12& f, const Fp2& c0, const Fp2& c2, const Fp2& c3) {
#if PSB_TWIST_MTYPE
  f.a.a = c0; f.a.b = c2; f.b.b = c3;      // c0 + c2 w^2 + c3 w^3  (w^2 = v)
#else
  f.a.a = c3; f.b.a = c2; f.b.b = c0;      // c3 + c2 w + c0 w^3
#endif
}

// T <- 2T and the tangent line at T evaluated at P = (xP, yP); nyP = -yP:
//   line = (E - B) + (3 X^2 xP) w^2 + (-2YZ yP) w^3     (scaled by -2YZ in Fp2)
PSB_HD PSB_NOINL void ml_dbl_step(G2H& T, Fp2& c0, Fp2& c2, Fp2& c3, const Fp& xP, const Fp& nyP) {
  Fp2 A, B, C, E, F, H, J, t, u;
  fp2_mul(A, T.x, T.y);                 // XY
  fp2_sqr(B, T.y);
  fp2_sqr(C, T.z);
  fp2_sqr(J, T.x);
  fp2_sqr_sum(H, T.y, T.z); fp2_sub2(H, H, B, C);   // 2YZ
  fp2_mul_3bt(E, C);                    // 3 b' Z^2
  fp2_dbl(F, E); fp2_add(F, F, E);      // 3E
  // line
  fp2_sub(c0, E, B);
  fp2_dbl(t, J); fp2_add(t, t, J); fp2_mul_fp(c2, t, xP);
  fp2_mul_fp(c3, H, nyP);
  // point (scaled by 4): X3 = 2 XY (B - F), Y3 = (B + F)^2 - 12 E^2, Z3 = 4 B H
  // (a single "small multiple" function in place of the dbl / add chains was measured 3 % slower on this kernel, r1u)
  fp2_sub(t, B, F); fp2_mul(t, A, t); fp2_dbl(T.x, t);
  fp2_sqr_sum(t, B, F);
  fp2_dbl(u, E); fp2_sqr(u, u);         // 4 E^2
  fp2_sub2(t, t, u, u); fp2_sub(T.y, t, u);
  fp2_mul(t, B, H); fp2_dbl(t, t); fp2_dbl(T.z, t);
}

// T <- T + Q (Q homogeneous projective, Q != +-T, neither infinity) and the chord through them at P:
//   theta = Y1 Z2 - Y2 Z1, lam = X1 Z2 - X2 Z1
//   line = (theta X2 - lam Y2) + (-theta Z2 xP) w^2 + (lam Z2 yP) w^3   (scaled by lam Z2); nxP = -xP
PSB_HD PSB_NOINL void ml_add_step(G2H& T, Fp2& c0, Fp2& c2, Fp2& c3, const G2H& Q, const Fp& nxP, const Fp& yP) {
  Fp2 A, Y1Z2, th, lm, W, l2, l3, N, t, u;
  fp2_mul(A, T.x, Q.z);
  fp2_mul(t, Q.x, T.z); fp2_sub(lm, A, t);
  fp2_mul(Y1Z2, T.y, Q.z);
  fp2_mul(t, Q.y, T.z); fp2_sub(th, Y1Z2, t);
  // line
  fp2_mul(t, th, Q.x); fp2_mul(u, lm, Q.y); fp2_sub(c0, t, u);
  fp2_mul(t, th, Q.z); fp2_mul_fp(c2, t, nxP);
  fp2_mul(t, lm, Q.z); fp2_mul_fp(c3, t, yP);
  // point
  fp2_mul(W, T.z, Q.z);
  fp2_sqr(l2, lm);
  fp2_mul(l3, l2, lm);
  fp2_mul(l2, l2, A);                    // lam^2 A
  fp2_sqr(t, th); fp2_mul(t, t, W);      // theta^2 W
  fp2_sub(N, t, l2); fp2_sub(N, N, l2); fp2_add(N, N, l3);
  fp2_mul(T.x, lm, N);
  fp2_sub(t, l2, N); fp2_mul(t, th, t);
  fp2_mul(u, l3, Y1Z2);
  fp2_sub(T.y, t, u);
  fp2_mul(T.z, l3, W);
}

// f *= the precomputed line L of the fixed argument, evaluated at P2 = (x2, y2)
PSB_HD PSB_NOINL void ml_fixed_line(Fp12& f, const FixedLine& Lg, const Fp& x2, const Fp& y2, bool scaled) {
  const FixedLine L = Lg;   // global -> thread-local once: the multiplier engines keep their local-memory loads
  Fp2 c2, c3;
  fp2_mul_fp(c2, L.nl, x2);
#if PSB_TWIST_MTYPE
  if (scaled) {                                   // L = (nl / c0, 1 / c0)
    fp2_mul_fp(c3, L.c0, y2);
    fp12_mul_line_1(f, c2, c3);
    return;
  }
#endif
  (void)scaled;
  c3.a = y2; fp_set_zero(c3.b);
  ml_mul_line(f, L.c0, c2, c3);
}

#if PSB_IS_BN
// pi on the twist (mcl Frobenius(G2), bn.hpp:2155-2162): (x, y) -> (conj(x) cx, conj(y) cy); homogeneous: conj(Z)
PSB_HD PSB_INL void fp2_load_const(Fp2& c, const uint32_t* w) {
  for (int i = 0; i < PSB_NL; i++) { c.a.v[i] = w[i]; c.b.v[i] = w[PSB_NL + i]; }
}
PSB_HD PSB_NOINL void g2h_frobenius(G2H& Q) {
  Fp2 c;
  fp2_conj(Q.x, Q.x); fp2_conj(Q.y, Q.y); fp2_conj(Q.z, Q.z);
  fp2_load_const(c, PSB_K(PSI_CX)); fp2_mul(Q.x, Q.x, c);
  fp2_load_const(c, PSB_K(PSI_CY)); fp2_mul(Q.y, Q.y, c);
}
#endif

// f = ML(P1, Q1) * ML(P2, Q2_fixed), conjugated for z < 0.
//   P1 = (x1, y1), P2 = (x2, y2): affine Fp coordinates; an infinite P is passed as (0, 0) (its
//   lines fall into Fp2 and die in the final exponentiation, like mcl: bls12_test.cpp:288-296).
//   Q1: G2 Jacobian (any z); q1_zero => that pairing contributes 1 (bn.hpp:1666-1669).
//   lines2: kMillerSteps precomputed lines of the fixed Q2 (per key); use2 = false skips them.
// `block_sync` (here and in final_exp): the caller guarantees that EVERY thread of the block is inside this call (no early
// exit); the block is then re-aligned at a few points of the straight-line schedule so that its warps keep sharing
// instruction fetches (kernels.cuh, launch shape).  Measured on B200, 2^20 lanes (profiles/r1zb_ab_block_sync.txt).
#ifndef PSB_FE_SYNC
#define PSB_FE_SYNC 1
#endif
#ifndef PSB_ML_SYNC
#define PSB_ML_SYNC 16     // barrier every PSB_ML_SYNC Miller iterations (0 = never; 1: -3 %, 16: +0.5 %, 32: 0)
#endif
#ifndef PSB_ML_SYNC_ADD
#define PSB_ML_SYNC_ADD 0  // barrier before each addition step of the Miller loop (measured: no gain on top of the rest)
#endif
#ifndef PSB_ML_FIRST_LINE
#define PSB_ML_FIRST_LINE 1  // first Miller iteration: f = line instead of f = 1 * line (13 Fp2 products; Miller loop 393.8 -> 389.5 ms)
#endif
#ifndef PSB_POWZ_SYNC
#define PSB_POWZ_SYNC 0    // barrier every PSB_POWZ_SYNC compressed squarings inside pow_z (0 = never; 4/8/16/32 lose to MID alone)
#endif
#ifndef PSB_POWZ_SYNC_TAIL
#define PSB_POWZ_SYNC_TAIL 0 // barrier before each decompression + product of pow_z
#endif
#ifndef PSB_FE_SYNC_EASY
#define PSB_FE_SYNC_EASY 0   // barrier after the inversion of the easy part
#endif
#ifndef PSB_POWZ_SYNC_MID
#define PSB_POWZ_SYNC_MID 1  // barrier between the squaring chain and the decompressions of pow_z
#endif
#if defined(__CUDA_ARCH__) && PSB_FE_SYNC
#define PSB_FE_BARRIER(on) do { if (on) __syncthreads(); } while (0)
#else
#define PSB_FE_BARRIER(on) do { (void)(on); } while (0)
#endif
#if defined(__CUDA_ARCH__) && PSB_ML_SYNC
#define PSB_ML_BARRIER(on) do { if (on) __syncthreads(); } while (0)
#else
#define PSB_ML_BARRIER(on) do { (void)(on); } while (0)
#endif
PSB_HD PSB_NOINL void miller_loop2(Fp12& f, const Fp& x1, const Fp& y1, const G2J& Q1, const Fp& x2, const Fp& y2,
                                   const FixedLine* lines2, bool use2, bool block_sync = false) {
  const bool use1 = !pt_is_zero(Q1);
  const bool scaled2 = use2 && lines2[kMillerSteps].nl.a.v[0] != 0;
  Fp nx1, ny1;
  fp_neg(nx1, x1); fp_neg(ny1, y1);
  G2H Q, T;
  // Jacobian (X, Y, Z) -> homogeneous (X Z, Y, Z^3)
  fp2_mul(Q.x, Q1.x, Q1.z);
  Q.y = Q1.y;
  fp2_sqr(Q.z, Q1.z); fp2_mul(Q.z, Q.z, Q1.z);
  T = Q;
  fp12_set_one(f);
  Fp2 c0, c2, c3;
  int li = 0;
  PSB_ROLL
  for (int i = PSB_ML_NBITS - 1; i >= 0; i--) {
#if PSB_ML_SYNC
    if (i % PSB_ML_SYNC == 0) PSB_ML_BARRIER(block_sync);
#endif
    if (i != PSB_ML_NBITS - 1) fp12_sqr(f, f);
    if (use1) {
      ml_dbl_step(T, c0, c2, c3, x1, ny1);
#if PSB_ML_FIRST_LINE
      if (i == PSB_ML_NBITS - 1) ml_set_line(f, c0, c2, c3);   // f is still 1: the product is the line itself
      else
#endif
      ml_mul_line(f, c0, c2, c3);
    }
    if (use2) {
      ml_fixed_line(f, lines2[li], x2, y2, scaled2);
    }
    li++;
    if (ml_bit(i)) {
#if PSB_ML_SYNC_ADD
      PSB_ML_BARRIER(block_sync);
#endif
      if (use1) {
        ml_add_step(T, c0, c2, c3, Q, nx1, y1);
        ml_mul_line(f, c0, c2, c3);
      }
      if (use2) {
        ml_fixed_line(f, lines2[li], x2, y2, scaled2);
      }
      li++;
    }
  }
  fp6_neg(f.b, f.b);  // z < 0  (bn.hpp:1695-1697)
#if PSB_IS_BN
  // BN tail (bn.hpp:1698-1709): T <- -T (z < 0), then the chords through pi(Q) and -pi^2(Q)
  if (use1) {
    fp2_neg(T.y, T.y);
    g2h_frobenius(Q);
    ml_add_step(T, c0, c2, c3, Q, nx1, y1);
    ml_mul_line(f, c0, c2, c3);
    g2h_frobenius(Q);
    fp2_neg(Q.y, Q.y);
    ml_add_step(T, c0, c2, c3, Q, nx1, y1);
    ml_mul_line(f, c0, c2, c3);
  }
  if (use2) {
    for (int t = 0; t < 2; t++) {
      ml_fixed_line(f, lines2[li++], x2, y2, scaled2);
    }
  }
#endif
}

// precompute the kMillerSteps affine lines of a fixed Q (affine, not infinity) into out[kFixedLineSlots].  One thread, once per key.
PSB_HD PSB_NOINL void precompute_fixed_lines(FixedLine* out, const G2A& Q) {
  Fp2 x = Q.x, y = Q.y, lam, t, u, x3;
  int li = 0;
  for (int i = PSB_ML_NBITS - 1; i >= 0; i--) {
    // tangent: lam = 3x^2 / 2y
    fp2_sqr(t, x); fp2_dbl(u, t); fp2_add(t, t, u);
    fp2_dbl(u, y); fp2_inv(u, u); fp2_mul(lam, t, u);
    fp2_neg(out[li].nl, lam);
    fp2_mul(t, lam, x); fp2_sub(out[li].c0, t, y);
    li++;
    fp2_sqr(x3, lam); fp2_sub(x3, x3, x); fp2_sub(x3, x3, x);
    fp2_sub(t, x, x3); fp2_mul(t, lam, t); fp2_sub(y, t, y);
    x = x3;
    if (ml_bit(i)) {
      // chord through T and Q: lam = (yQ - y)/(xQ - x)
      fp2_sub(t, Q.y, y); fp2_sub(u, Q.x, x); fp2_inv(u, u); fp2_mul(lam, t, u);
      fp2_neg(out[li].nl, lam);
      fp2_mul(t, lam, Q.x); fp2_sub(out[li].c0, t, Q.y);
      li++;
      fp2_sqr(x3, lam); fp2_sub(x3, x3, x); fp2_sub(x3, x3, Q.x);
      fp2_sub(t, x, x3); fp2_mul(t, lam, t); fp2_sub(y, t, y);
      x = x3;
    }
  }
#if PSB_IS_BN
  fp2_neg(y, y);                                        // T <- -T
  Fp2 qx = Q.x, qy = Q.y, c;
  for (int k = 0; k < 2; k++) {
    fp2_conj(qx, qx); fp2_conj(qy, qy);
    fp2_load_const(c, PSB_K(PSI_CX)); fp2_mul(qx, qx, c);
    fp2_load_const(c, PSB_K(PSI_CY)); fp2_mul(qy, qy, c);
    if (k == 1) fp2_neg(qy, qy);                        // pi(Q), then -pi^2(Q)
    fp2_sub(t, qy, y); fp2_sub(u, qx, x); fp2_inv(u, u); fp2_mul(lam, t, u);
    fp2_neg(out[li].nl, lam);
    fp2_mul(t, lam, qx); fp2_sub(out[li].c0, t, qy);
    li++;
    fp2_sqr(x3, lam); fp2_sub(x3, x3, x); fp2_sub(x3, x3, qx);
    fp2_sub(t, x, x3); fp2_mul(t, lam, t); fp2_sub(y, t, y);
    x = x3;
  }
#endif
  // scale every line to constant term 1 (M-type twist): (nl, c0) -> (nl / c0, 1 / c0); a zero c0 (a tangent or chord through
  // x = 0, y = 0: probability ~1/p^2) leaves the table unscaled
  fp2_set_zero(out[kMillerSteps].nl); fp2_set_zero(out[kMillerSteps].c0);
#if PSB_TWIST_MTYPE && !defined(PSB_LINES_UNSCALED)
  bool all = true;
  for (int k = 0; k < kMillerSteps; k++) all = all && !fp2_is_zero(out[k].c0);
  if (all) {
    for (int k = 0; k < kMillerSteps; k++) {
      Fp2 ic;
      fp2_inv(ic, out[k].c0);
      fp2_mul(out[k].nl, out[k].nl, ic);
      out[k].c0 = ic;
    }
    out[kMillerSteps].nl.a.v[0] = 1;
  }
#endif
}

// y = x^z (z < 0): x^|z| by square-and-multiply with cyclotomic squarings, then conjugate (mcl pow_z, bn.hpp:1150-1176)
PSB_HD PSB_NOINL void pow_z_gs(Fp12& y, const Fp12& x) {
  Fp12 acc = x;
  PSB_ROLL
  for (int i = PSB_Z_NBITS - 1; i >= 0; i--) {
    fp12_cyclo_sqr(acc, acc);
    if (z_bit(i)) fp12_mul(acc, acc, x);
  }
  fp12_conj(y, acc);
}

// The same value through Karabina's compressed squarings: x^|z| = prod over the set bits i of x^(2^i).  ONE chain of
// PSB_Z_PIVOT compressed squarings (6 Fp2 squarings each instead of 9) passes through every x^(2^i), i <= pivot; the
// compressed values at the set bits are kept (the bits above the pivot are finished on the full element, see below), their (g0, g1) coordinates are rebuilt with ONE shared inversion (Montgomery's trick
// over the denominators 4 g2), and the full elements are multiplied together.  x must lie in the cyclotomic subgroup.
// x = 1 (the only element of the odd-order cyclotomic subgroup with a power-of-two power equal to 1) is handled in line
// with unit denominators; a zero g2 on any other element (probability ~2^-380) falls back to pow_z_gs, so the result
// is the same field element on every input.
#ifdef PSB_Z_PIVOT_OVERRIDE   // A/B builds: -DPSB_Z_PIVOT_OVERRIDE=<bit> -DPSB_Z_SETBITS_OVERRIDE=<set bits in [1, bit]>
#undef PSB_Z_PIVOT
#undef PSB_Z_SETBITS_HI
#define PSB_Z_PIVOT PSB_Z_PIVOT_OVERRIDE
#define PSB_Z_SETBITS_HI PSB_Z_SETBITS_OVERRIDE
#endif
constexpr int kZSetBits = PSB_Z_SETBITS_HI;          // set bits of |z| in [1, pivot]: one kept compressed value each
PSB_HD PSB_NOINL void pow_z(Fp12& y, const Fp12& x, bool block_sync = false) {
#ifdef PSB_POWZ_GS
  pow_z_gs(y, x);
#else
  CycC keep[kZSetBits];
  Fp2 pre[kZSetBits];
  {
    CycC c;
    c.g2 = x.b.a; c.g3 = x.a.c; c.g4 = x.a.b; c.g5 = x.b.c;
    int k = 0;
    PSB_ROLL   // (left to itself the compiler unrolls all 63 trips: 40 KB of straight-line code through the instruction caches)
    for (int i = 1; i <= PSB_Z_PIVOT; i++) {
      cyclo_csqr(c, c);
      if (i == PSB_Z_PIVOT || z_bit(i)) keep[k++] = c;
#if PSB_POWZ_SYNC
      if (i % PSB_POWZ_SYNC == 0) PSB_FE_BARRIER(block_sync);
#endif
    }
  }
#if PSB_POWZ_SYNC_MID
  PSB_FE_BARRIER(block_sync);
#endif
  // prefix products of the denominators 4 g2.  x = 1 (a lane whose Miller value lies in a proper subfield, e.g. a
  // credential with sigma = 0) compresses to (0, 0, 0, 0): there the denominator is replaced by 1 -- the numerator is 0,
  // so g1 = 0 and g0 = 1 come out right with no other code path (a tampered lane must not slow down its warp).
  // g2 = 0 on any OTHER element (probability ~2^-380) takes the Granger-Scott path.
  bool rare = false;
  PSB_ROLL
  for (int k = 0; k < kZSetBits; k++) {
    Fp2 d, one;
    fp2_set_one(one);
    fp2_dbl(d, keep[k].g2); fp2_dbl(d, d);      // (no call under a lane-dependent branch: see below)
    const bool z2 = fp2_is_zero(keep[k].g2);
    rare = rare || (z2 && !(fp2_is_zero(keep[k].g3) && fp2_is_zero(keep[k].g4) && fp2_is_zero(keep[k].g5)));
    fp2_cmov(d, one, z2);
    if (k == 0) pre[0] = d; else fp2_mul(pre[k], pre[k - 1], d);
  }
#ifdef __CUDA_ARCH__
  // Both paths give the same field element, so the whole warp takes the fallback when any lane needs it: the call
  // stays warp-uniform.  (A divergent call here let the callee's use of uniform registers clobber stack addresses the
  // other lanes still held in uniform registers -- observed as an invalid local read on the BN254 build.)
  rare = __any_sync(__activemask(), rare);
#endif
  if (rare) { pow_z_gs(y, x); return; }
  Fp2 inv;
  fp2_inv(inv, pre[kZSetBits - 1]);
  Fp12 acc, t;
  PSB_ROLL
  for (int k = kZSetBits - 1; k >= 0; k--) {
    Fp2 dinv, num, g1;
#if PSB_POWZ_SYNC_TAIL
    PSB_FE_BARRIER(block_sync);
#endif
    if (k > 0) {
      Fp2 d, one;
      fp2_mul(dinv, inv, pre[k - 1]);            // 1 / den_k
      fp2_set_one(one);
      fp2_dbl(d, keep[k].g2); fp2_dbl(d, d);
      fp2_cmov(d, one, fp2_is_zero(keep[k].g2));
      fp2_mul(inv, inv, d);                      // 1 / (den_0 ... den_{k-1})
    } else {
      dinv = inv;
    }
    cyclo_decompress_num(num, keep[k]);
    fp2_mul(g1, num, dinv);
    if (k == kZSetBits - 1) {
      cyclo_decompress_fill(acc, keep[k], g1);
      // the set bits above the pivot lie within three squarings of each other: x^(2^pivot) raised to |z| >> pivot by
      // plain cyclotomic squarings costs less than a decompression per bit (BLS12-381: 105 = 1101001b; BN254: 1)
      if ((PSB_Z_ABS >> PSB_Z_PIVOT) == 105) {
        // 105 = (2^3 - 1)(2^4 - 1), and an inverse is a conjugation in the cyclotomic subgroup:
        // 7 squarings + 2 products instead of 6 + 3 for the binary chain
        PSB_ROLL
        for (int r = 0; r < 2; r++) {
          fp12_conj(t, acc);
          for (int i = 0; i < 3 + r; i++) fp12_cyclo_sqr(acc, acc);
          fp12_mul(acc, acc, t);                  // y^7, then (y^7)^15
        }
      } else if ((PSB_Z_ABS >> PSB_Z_PIVOT) > 1) {
        t = acc;
        PSB_ROLL
        for (int i = PSB_Z_NBITS - PSB_Z_PIVOT - 1; i >= 0; i--) {
          fp12_cyclo_sqr(acc, acc);
          if (z_bit(PSB_Z_PIVOT + i)) fp12_mul(acc, acc, t);
        }
      }
    } else {
      cyclo_decompress_fill(t, keep[k], g1);
      fp12_mul(acc, acc, t);
    }
  }
  if (z_bit(0)) fp12_mul(acc, acc, x);
  fp12_conj(y, acc);
#endif
}

// y = x^((p^12-1)/r * 3), structured as mcl's finalExp (bn.hpp:1643-1659)
// `block_sync`: see miller_loop2; the block is re-aligned before each pow_z
PSB_HD PSB_NOINL void final_exp(Fp12& y, const Fp12& x, bool block_sync = false) {
  Fp12 a0, a1, a2, a3, a4, a5, a7, t;
  // easy part: t = x^((p^6-1)(p^2+1))   (mapToCyclotomic, bn.hpp:1494-1502)
  fp12_frobenius(a0, x, 2);
  fp12_mul(a0, a0, x);          // x^(p^2+1)
  fp12_inv(a1, a0);
#if PSB_FE_SYNC_EASY
  PSB_FE_BARRIER(block_sync);
#endif
  fp12_conj(a0, a0);            // ^(p^6)
  fp12_mul(t, a1, a0);
#if PSB_IS_BN
  // hard part (expHardPartBN, bn.hpp:1576-1621): t^(d 2z(6z^2+3z+1)), d = (p^4 - p^2 + 1)/r  (Fuentes-Castaneda et al.)
  Fp12 a, b;
  PSB_FE_BARRIER(block_sync);   // block re-alignment as in the BLS12 branch below
  pow_z(b, t, block_sync);      // t^z
  fp12_cyclo_sqr(b, b);         // t^2z
  fp12_cyclo_sqr(a, b);         // t^4z
  fp12_mul(a, a, b);            // t^6z
  PSB_FE_BARRIER(block_sync);
  pow_z(a2, a, block_sync);     // t^(6z^2)
  fp12_mul(a, a, a2);
  fp12_cyclo_sqr(a3, a2);       // t^(12z^2)
  PSB_FE_BARRIER(block_sync);
  pow_z(a3, a3, block_sync);    // t^(12z^3)
  fp12_mul(a, a, a3);
  fp12_conj(b, b);
  fp12_mul(b, b, a);
  fp12_mul(a2, a2, a);
  fp12_frobenius(a, a, 2);
  fp12_mul(a, a, a2);
  fp12_mul(a, a, t);
  fp12_conj(a4, t);
  fp12_mul(a4, a4, b);
  fp12_frobenius(b, b, 1);
  fp12_mul(a, a, b);
  fp12_frobenius(a4, a4, 3);
  fp12_mul(y, a4, a);
  (void)a5; (void)a7;
#else
  // hard part: the SAME exponent as mcl's expHardPartBLS12 (bn.hpp:1508-1555), 3 (p^4 - p^2 + 1) / r, through the
  // factorisation of Hayashida, Hayasaka and Teruya (2020):  (z - 1)^2 (z + p) (z^2 + p^2 - 1) + 3   (identity checked in
  // tests/test_oracle.py) -- five pow_z like mcl, but 7 Fp12 products, 1 cyclotomic squaring and 2 Frobenius maps
  // instead of 12, 2 and 3; equal exponents give the same field element, so GT bytes are unchanged.
  PSB_FE_BARRIER(block_sync);
  pow_z(a0, t, block_sync); fp12_conj(a1, t); fp12_mul(a0, a0, a1);                 // t^(z-1)
  PSB_FE_BARRIER(block_sync);
  pow_z(a2, a0, block_sync); fp12_conj(a1, a0); fp12_mul(a2, a2, a1);               // ^(z-1)
  PSB_FE_BARRIER(block_sync);
  pow_z(a0, a2, block_sync); fp12_frobenius(a1, a2, 1); fp12_mul(a0, a0, a1);       // ^(z+p)            =: u
  PSB_FE_BARRIER(block_sync);
  pow_z(a2, a0, block_sync);
  PSB_FE_BARRIER(block_sync);
  pow_z(a2, a2, block_sync);                                                        // u^(z^2)
  fp12_frobenius(a1, a0, 2); fp12_mul(a2, a2, a1);                      // u^(z^2+p^2)
  fp12_conj(a1, a0); fp12_mul(a2, a2, a1);                              // u^(z^2+p^2-1)
  fp12_cyclo_sqr(a1, t); fp12_mul(a1, a1, t);                           // t^3
  fp12_mul(y, a2, a1);
  (void)a3; (void)a4; (void)a5; (void)a7;
#endif
}

}  // namespace psb
