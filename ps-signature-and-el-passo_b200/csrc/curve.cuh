// curve.cuh -- G1 = E(Fp): y^2 = x^3 + b and G2 = E'(Fp2): y^2 = x^3 + b', Jacobian coordinates
// (BLS12-381: b = 4, b' = 4(1+i); BN254: b = 2, b' = 2/(1+i) = 1 - i).
//
// Replaces for the batched path mcl's EcT (reference: third-parties/mcl/include/mcl/ec.hpp:138-284
// dblJacobi/addJacobi, :77-88 normalizeJacobi, :799 isZero <=> z == 0) and its scalar
// multiplications (ec.hpp:1124-1139 mulArray, :1466-1523 GLV1; bn.hpp:765-860 GLV2).  Raw layout
// = mcl's: (x, y, z) Montgomery limbs, z == 0 is the point at infinity, z == 1 after normalisation.
// Internal formulas and schedules are our own; outputs are compared in canonical (normalised or
// serialised) form only.
#pragma once
#include "tower.cuh"

namespace psb {

// ---- uniform field interface (overloads on Fp / Fp2) -------------------------------------------
PSB_HD PSB_INL void f_add(Fp& r, const Fp& a, const Fp& b) { fp_add(r, a, b); }
PSB_HD PSB_INL void f_sub(Fp& r, const Fp& a, const Fp& b) { fp_sub(r, a, b); }
PSB_HD PSB_INL void f_dbl(Fp& r, const Fp& a) { fp_dbl(r, a); }
PSB_HD PSB_INL void f_neg(Fp& r, const Fp& a) { fp_neg(r, a); }
PSB_HD PSB_INL void f_mul(Fp& r, const Fp& a, const Fp& b) { fp_mul(r, a, b); }
PSB_HD PSB_INL void f_sqr(Fp& r, const Fp& a) { fp_sqr(r, a); }
PSB_HD PSB_INL void f_inv(Fp& r, const Fp& a) { fp_inv(r, a); }
PSB_HD PSB_INL bool f_is_zero(const Fp& a) { return fp_is_zero(a); }
PSB_HD PSB_INL bool f_eq(const Fp& a, const Fp& b) { return fp_eq(a, b); }
PSB_HD PSB_INL void f_set_zero(Fp& r) { fp_set_zero(r); }
PSB_HD PSB_INL void f_set_one(Fp& r) { fp_set_one(r); }
PSB_HD PSB_INL void f_cmov(Fp& r, const Fp& a, bool c) { fp_cmov(r, a, c); }

PSB_HD PSB_INL void f_add(Fp2& r, const Fp2& a, const Fp2& b) { fp2_add(r, a, b); }
PSB_HD PSB_INL void f_sub(Fp2& r, const Fp2& a, const Fp2& b) { fp2_sub(r, a, b); }
PSB_HD PSB_INL void f_dbl(Fp2& r, const Fp2& a) { fp2_dbl(r, a); }
PSB_HD PSB_INL void f_neg(Fp2& r, const Fp2& a) { fp2_neg(r, a); }
PSB_HD PSB_INL void f_mul(Fp2& r, const Fp2& a, const Fp2& b) { fp2_mul(r, a, b); }
PSB_HD PSB_INL void f_sqr(Fp2& r, const Fp2& a) { fp2_sqr(r, a); }
PSB_HD PSB_INL void f_inv(Fp2& r, const Fp2& a) { fp2_inv(r, a); }
PSB_HD PSB_INL bool f_is_zero(const Fp2& a) { return fp2_is_zero(a); }
PSB_HD PSB_INL bool f_eq(const Fp2& a, const Fp2& b) { return fp2_eq(a, b); }
PSB_HD PSB_INL void f_set_zero(Fp2& r) { fp2_set_zero(r); }
PSB_HD PSB_INL void f_set_one(Fp2& r) { fp2_set_one(r); }
PSB_HD PSB_INL void f_cmov(Fp2& r, const Fp2& a, bool c) { fp2_cmov(r, a, c); }

template <class F> struct Jac { F x, y, z; };   // mcl G1 / G2 object layout
template <class F> struct Aff { F x, y; };      // table entry / normalised input (never infinity)
typedef Jac<Fp> G1J;
typedef Jac<Fp2> G2J;
typedef Aff<Fp> G1A;
typedef Aff<Fp2> G2A;

template <class F> PSB_HD PSB_INL bool pt_is_zero(const Jac<F>& P) { return f_is_zero(P.z); }
template <class F> PSB_HD PSB_INL void pt_set_zero(Jac<F>& P) { f_set_zero(P.x); f_set_zero(P.y); f_set_zero(P.z); }
template <class F> PSB_HD PSB_INL void pt_neg(Jac<F>& R, const Jac<F>& P) { R.x = P.x; f_neg(R.y, P.y); R.z = P.z; }
template <class F> PSB_HD PSB_INL void pt_from_aff(Jac<F>& R, const Aff<F>& P) { R.x = P.x; R.y = P.y; f_set_one(R.z); }

// 2P, a = 0 (dbl-2009-l: 2M + 5S)
template <class F>
PSB_HD PSB_NOINL void pt_dbl(Jac<F>& R, const Jac<F>& P) {
  F A, B, C, D, E, t;
  f_sqr(A, P.x);
  f_sqr(B, P.y);
  f_sqr(C, B);
  f_add(t, P.x, B); f_sqr(t, t); f_sub(t, t, A); f_sub(t, t, C); f_dbl(D, t);   // D = 2((X+B)^2 - A - C)
  f_dbl(E, A); f_add(E, E, A);                                                   // E = 3A
  f_mul(t, P.y, P.z); f_dbl(R.z, t);                                             // Z3 = 2YZ (0 stays 0)
  f_sqr(t, E); f_sub(t, t, D); f_sub(R.x, t, D);                                 // X3 = E^2 - 2D
  f_sub(t, D, R.x); f_mul(t, t, E);
  f_dbl(C, C); f_dbl(C, C); f_dbl(C, C);                                         // 8C
  f_sub(R.y, t, C);
}

// P + Q, both Jacobian (add-2007-bl with the exceptional cases of mcl addJacobi, ec.hpp:210-284)
template <class F>
PSB_HD PSB_NOINL void pt_add(Jac<F>& R, const Jac<F>& P, const Jac<F>& Q) {
  if (pt_is_zero(P)) { R = Q; return; }
  if (pt_is_zero(Q)) { R = P; return; }
  F Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, r, V, t;
  f_sqr(Z1Z1, P.z);
  f_sqr(Z2Z2, Q.z);
  f_mul(U1, P.x, Z2Z2);
  f_mul(U2, Q.x, Z1Z1);
  f_mul(S1, P.y, Q.z); f_mul(S1, S1, Z2Z2);
  f_mul(S2, Q.y, P.z); f_mul(S2, S2, Z1Z1);
  f_sub(H, U2, U1);
  f_sub(r, S2, S1);
  if (f_is_zero(H)) {
    if (f_is_zero(r)) { pt_dbl(R, P); } else { pt_set_zero(R); }
    return;
  }
  f_dbl(r, r);
  f_dbl(I, H); f_sqr(I, I);
  f_mul(J, H, I);
  f_mul(V, U1, I);
  f_add(t, P.z, Q.z); f_sqr(t, t); f_sub(t, t, Z1Z1); f_sub(t, t, Z2Z2); f_mul(R.z, t, H);
  f_sqr(t, r); f_sub(t, t, J); f_sub(t, t, V); f_sub(R.x, t, V);
  f_sub(t, V, R.x); f_mul(t, t, r);
  f_mul(S1, S1, J); f_dbl(S1, S1);
  f_sub(R.y, t, S1);
}

// P + Q, Q affine and never infinity (madd-2007-bl: 7M + 4S)
template <class F>
PSB_HD PSB_NOINL void pt_madd(Jac<F>& R, const Jac<F>& P, const Aff<F>& Q) {
  if (pt_is_zero(P)) { pt_from_aff(R, Q); return; }
  F Z1Z1, U2, S2, H, HH, I, J, r, V, t;
  f_sqr(Z1Z1, P.z);
  f_mul(U2, Q.x, Z1Z1);
  f_mul(S2, Q.y, P.z); f_mul(S2, S2, Z1Z1);
  f_sub(H, U2, P.x);
  f_sub(r, S2, P.y);
  if (f_is_zero(H)) {
    if (f_is_zero(r)) { pt_dbl(R, P); } else { pt_set_zero(R); }
    return;
  }
  f_dbl(r, r);
  f_sqr(HH, H);
  f_dbl(I, HH); f_dbl(I, I);
  f_mul(J, H, I);
  f_mul(V, P.x, I);
  f_add(t, P.z, H); f_sqr(t, t); f_sub(t, t, Z1Z1);
  F y1 = P.y;
  f_sub(R.z, t, HH);
  f_sqr(t, r); f_sub(t, t, J); f_sub(t, t, V); f_sub(R.x, t, V);
  f_sub(t, V, R.x); f_mul(t, t, r);
  f_mul(J, y1, J); f_dbl(J, J);
  f_sub(R.y, t, J);
}

// (x/z^2, y/z^3, 1); infinity -> all-zero (canonical zero, what mcl's clear() holds)
template <class F>
PSB_HD PSB_NOINL void pt_normalize(Jac<F>& R, const Jac<F>& P) {
  if (pt_is_zero(P)) { pt_set_zero(R); return; }
  F zi, zi2;
  f_inv(zi, P.z);
  f_sqr(zi2, zi);
  f_mul(R.x, P.x, zi2);
  f_mul(zi2, zi2, zi);
  f_mul(R.y, P.y, zi2);
  f_set_one(R.z);
}

// k*P, variable base, no endomorphism: fixed 4-bit windows over the 256-bit normal-form scalar (MSB first).
// k = 8 limbs, normal form (NOT Montgomery), k < 2^256.  Kept as the plain reference path of the GLV/GLS versions.
template <class F>
PSB_HD PSB_NOINL void pt_mul_window(Jac<F>& R, const Jac<F>& P, const uint32_t* k) {
  Jac<F> tbl[16];
  pt_set_zero(tbl[0]);
  tbl[1] = P;
  pt_dbl(tbl[2], P);
  for (int i = 3; i < 16; i++) pt_add(tbl[i], tbl[i - 1], P);
  Jac<F> acc;
  pt_set_zero(acc);
  int top = 63;                                 // leading zero windows are skipped (the one caller passes a constant: the
  while (top > 0 && ((k[top >> 3] >> ((top & 7) * 4)) & 0xF) == 0) top--;   // 126-bit G1 cofactor -- half the doublings)
  for (int i = top; i >= 0; i--) {
    if (i != top) { pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); }
    const uint32_t d = (k[i >> 3] >> ((i & 7) * 4)) & 0xF;
    if (d) pt_add(acc, acc, tbl[d]);
  }
  R = acc;
}

// ---- helpers of the variable-base multiplications: signed 4-bit windows over a table of AFFINE multiples -------------------
// Signed radix-16 digits of a magnitude v < 2^(4 nd) (nibbles through get(i)): d_i in [-8, 8], sum d_i 16^i = v, nd + 1 digits
// (the last one is the final carry, 0 or 1).  Halves the table (multiples 1..8) against unsigned nibbles.
template <class Get>
PSB_HD PSB_INL void signed_nibbles(int8_t* dig, int nd, Get get) {
  int carry = 0;
  for (int i = 0; i < nd; i++) {
    int d = (int)get(i) + carry;
    carry = d > 8;
    dig[i] = (int8_t)(carry ? d - 16 : d);
  }
  dig[nd] = (int8_t)carry;
}
// tbl[i] = i P in affine coordinates for i = 1..8 with ONE inversion (Montgomery's trick over the z coordinates), so that the
// 64 window additions of a multiplication are mixed additions (7M + 4S instead of 11M + 5S).  Returns a mask: bit i set <=> i P
// is the point at infinity (P of small order -- adversarial input only -- or P itself infinite): those entries are skipped.
// A lane whose P is at infinity (a malformed wire message, a zero commitment) builds the table from a fixed finite stand-in
// point instead and reports every multiple as infinite: otherwise it would leave pt_add early while its warp carries on, and
// on sm_100a the lanes of such a warp read garbage stack addresses in the NEXT call (compute-sanitizer "Invalid __local__
// read" in pt_add, psb_verify_id_ser with malformed lanes; ptxas keeps stack addresses in uniform registers across a call
// that only part of the warp enters -- the same hazard pow_z guards against with its warp-uniform fallback).
PSB_HD PSB_INL void pt_load_standin(Jac<Fp>& S) {
  for (int i = 0; i < PSB_NL; i++) { S.x.v[i] = PSB_K(G1_STANDIN)[i]; S.y.v[i] = PSB_K(G1_STANDIN)[PSB_NL + i]; }
  fp_set_one(S.z);
}
PSB_HD PSB_INL void pt_load_standin(Jac<Fp2>& S) {
  for (int i = 0; i < PSB_NL; i++) {
    S.x.a.v[i] = PSB_K(G2_STANDIN)[i]; S.x.b.v[i] = PSB_K(G2_STANDIN)[PSB_NL + i];
    S.y.a.v[i] = PSB_K(G2_STANDIN)[2 * PSB_NL + i]; S.y.b.v[i] = PSB_K(G2_STANDIN)[3 * PSB_NL + i];
  }
  fp2_set_one(S.z);
}
template <class F>
PSB_HD PSB_NOINL uint32_t pt_affine_multiples8(Aff<F>* tbl /*[9], [0] unused*/, const Jac<F>& P) {
  Jac<F> J[9];
  F pre[9];
  const bool pinf = pt_is_zero(P);
  J[1] = P;
  if (pinf) pt_load_standin(J[1]);            // (element-wise stores, no call under this branch)
  pt_dbl(J[2], J[1]);
  for (int i = 3; i <= 8; i++) pt_add(J[i], J[i - 1], J[1]);
  uint32_t inf = 0;
  F inv, one;
  f_set_one(one);
  // prefix products of the z coordinates (an infinite multiple contributes 1), written IN PLACE as in k_build_table.  (A first
  // version kept the running product in a local, `f_mul(run, run, z); pre[i] = run;`: correct on the host build and for Fp2,
  // but the sm_100a build for Fp copied a stale `run` into pre[i] for i >= 2 -- entries 2..8 of the table were wrong on the
  // device only.  Found by the parity test of test op T_G1_AFFMUL; every prefix chain in this file now writes its array directly.)
  for (int i = 1; i <= 8; i++) {
    if (f_is_zero(J[i].z)) { inf |= 1u << i; J[i].z = one; }
    if (i == 1) pre[1] = J[1].z; else f_mul(pre[i], pre[i - 1], J[i].z);
  }
  f_inv(inv, pre[8]);
  for (int i = 8; i >= 1; i--) {
    F zi, zi2;
    if (i > 1) f_mul(zi, inv, pre[i - 1]); else zi = inv;
    f_mul(inv, inv, J[i].z);
    f_sqr(zi2, zi);
    f_mul(tbl[i].x, J[i].x, zi2);
    f_mul(zi2, zi2, zi);
    f_mul(tbl[i].y, J[i].y, zi2);
  }
  return pinf ? 0x1FEu : inf;
}

// ---- GLV (G1) and GLS (G2) variable-base multiplication -----------------------------------------------------
// Same group element as mcl's G1::mul / G2::mul (which use the same endomorphisms with w-NAF: ec.hpp:1457-1523
// GLV1, bn.hpp:765-860 GLV2); only the schedule is ours: fixed 4-bit joint windows, one shared table of multiples
// of P, endomorphic images of table entries formed on the fly.
//   G1: phi(X, Y, Z) = (beta X, Y, Z) = [lambda] P, lambda = z^2 - 1;  k = k1 + k2 lambda with k1, k2 < 2^128
//       (r = lambda^2 + lambda + 1, so k2 = floor(k / lambda) <= lambda + 1).  128 doublings instead of 256.
//   G2: psi(X, Y, Z) = (conj(X) cx, conj(Y) cy, conj(Z)) = [z] Q (z < 0);  k = sum d_i |z|^i, d_i < 2^64, so
//       k Q = d0 Q - d1 psi(Q) + d2 psi^2(Q) - d3 psi^3(Q).  64 doublings instead of 256.

#if !PSB_IS_BN
// q = floor(k / lambda), rem = k mod lambda for k < r  (Barrett with mu = floor(2^256 / lambda), deficit <= 2)
PSB_HD PSB_INL void glv1_split(uint32_t k1[4], uint32_t k2[4], const uint32_t* k) {
  const uint32_t* mu = PSB_K(GLV_MU);
  const uint32_t* lam = PSB_K(GLV_LAMBDA);
  // top 5 limbs of k * mu (13 limbs): only columns >= 6 can influence limbs 8..12 through carries of 2 limbs;
  // compute the full product column by column with a 3-word accumulator
  uint32_t q[5];
  {
    uint64_t lo = 0, hi = 0;
    for (int col = 0; col < 13; col++) {
      for (int i = 0; i < 8; i++) {
        const int j = col - i;
        if (j < 0 || j > 4) continue;
        const uint64_t pr = (uint64_t)k[i] * mu[j];
        const uint64_t nl = lo + pr;
        hi += nl < lo;
        lo = nl;
      }
      if (col >= 8) q[col - 8] = (uint32_t)lo;
      lo = (lo >> 32) | (hi << 32);
      hi = 0;
    }
  }
  // rem = k - q * lambda  (mod 2^160: rem < 3 lambda < 2^130)
  uint32_t rem[5];
  {
    uint32_t ql[5];
    uint64_t lo = 0, hi = 0;
    for (int col = 0; col < 5; col++) {
      for (int i = 0; i <= col; i++) {
        const int j = col - i;
        if (i > 4 || j > 3) continue;
        const uint64_t pr = (uint64_t)q[i] * lam[j];
        const uint64_t nl = lo + pr;
        hi += nl < lo;
        lo = nl;
      }
      ql[col] = (uint32_t)lo;
      lo = (lo >> 32) | (hi << 32);
      hi = 0;
    }
    int64_t c = 0;
    for (int i = 0; i < 5; i++) { c += (int64_t)k[i] - ql[i]; rem[i] = (uint32_t)c; c >>= 32; }
  }
  for (int it = 0; it < 3; it++) {   // at most two corrections
    uint32_t t[5];
    int64_t c = 0;
    for (int i = 0; i < 5; i++) { c += (int64_t)rem[i] - (i < 4 ? lam[i] : 0u); t[i] = (uint32_t)c; c >>= 32; }
    if (c < 0) break;                // rem < lambda
    for (int i = 0; i < 5; i++) rem[i] = t[i];
    uint64_t cc = 1;
    for (int i = 0; i < 5; i++) { cc += q[i]; q[i] = (uint32_t)cc; cc >>= 32; }
  }
  for (int i = 0; i < 4; i++) { k1[i] = rem[i]; k2[i] = q[i]; }
}

PSB_HD PSB_NOINL void g1_mul_glv(G1J& R, const G1J& P, const uint32_t* k) {
  uint32_t k1[4], k2[4];
  glv1_split(k1, k2, k);
  int8_t e1[33], e2[33];
  signed_nibbles(e1, 32, [&](int i) { return (k1[i >> 3] >> ((i & 7) * 4)) & 0xFu; });
  signed_nibbles(e2, 32, [&](int i) { return (k2[i >> 3] >> ((i & 7) * 4)) & 0xFu; });
  G1A tbl[9], T;
  const uint32_t inf = pt_affine_multiples8(tbl, P);
  Fp beta;
  for (int i = 0; i < 12; i++) beta.v[i] = PSB_K(GLV_BETA)[i];
  G1J acc;
  pt_set_zero(acc);
  for (int i = 32; i >= 0; i--) {
    if (i < 32) { pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); }
    const int d1 = e1[i], d2 = e2[i];
    const int a1 = d1 < 0 ? -d1 : d1, a2 = d2 < 0 ? -d2 : d2;
    if (a1 && !((inf >> a1) & 1u)) {
      T = tbl[a1];
      if (d1 < 0) fp_neg(T.y, T.y);
      pt_madd(acc, acc, T);
    }
    if (a2 && !((inf >> a2) & 1u)) {
      T = tbl[a2];
      fp_mul(T.x, T.x, beta);          // phi(d2 P)
      if (d2 < 0) fp_neg(T.y, T.y);
      pt_madd(acc, acc, T);
    }
  }
  R = acc;
}

// d_i = base-|z| digits of k (k < 2^256 -> four digits below 2^64 for k < r)
PSB_HD PSB_INL void gls_split(uint64_t d[4], const uint32_t* k) {
  uint32_t q[8];
  for (int i = 0; i < 8; i++) q[i] = k[i];
  for (int j = 0; j < 3; j++) {
    unsigned __int128 rem = 0;
    for (int i = 7; i >= 0; i--) {
      const unsigned __int128 cur = (rem << 32) | q[i];
      q[i] = (uint32_t)(cur / PSB_Z_ABS);
      rem = cur % PSB_Z_ABS;
    }
    d[j] = (uint64_t)rem;
  }
  d[3] = ((uint64_t)q[1] << 32) | q[0];
}

// (-1)^j psi^j(T) for an AFFINE point, j = 1..3 (constants from tools/gen_constants.py)
PSB_HD PSB_NOINL void g2_psi_signed(G2A& T, int j) {
  if (j == 2) {
    Fp n;
    for (int i = 0; i < 12; i++) n.v[i] = PSB_K(PSI_NCX)[i];
    fp2_mul_fp(T.x, T.x, n);
    fp2_neg(T.y, T.y);                 // psi^2(x, y) = (x N(cx), -y); sign +
    return;
  }
  Fp2 c;
  const uint32_t* cxp = (j == 1) ? PSB_K(PSI_CX) : PSB_K(PSI_CX3);
  for (int i = 0; i < 12; i++) { c.a.v[i] = cxp[i]; c.b.v[i] = cxp[12 + i]; }
  fp2_conj(T.x, T.x); fp2_conj(T.y, T.y);
  fp2_mul(T.x, T.x, c);
  for (int i = 0; i < 12; i++) { c.a.v[i] = PSB_K(PSI_CY)[i]; c.b.v[i] = PSB_K(PSI_CY)[12 + i]; }
  fp2_mul(T.y, T.y, c);
  if (j == 1) fp2_neg(T.y, T.y);       // -psi(T);  -psi^3(T) = (conj(x) cx3, +conj(y) cy)
}

PSB_HD PSB_NOINL void g2_mul_gls(G2J& R, const G2J& P, const uint32_t* k) {
  uint64_t d[4];
  gls_split(d, k);
  int8_t e[4][17];
  for (int j = 0; j < 4; j++) signed_nibbles(e[j], 16, [&](int i) { return (uint32_t)(d[j] >> (4 * i)) & 0xFu; });
  G2A tbl[9], T;
  const uint32_t inf = pt_affine_multiples8(tbl, P);
  G2J acc;
  pt_set_zero(acc);
  for (int i = 16; i >= 0; i--) {
    if (i < 16) { pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); }
    for (int j = 0; j < 4; j++) {
      const int dj = e[j][i], aj = dj < 0 ? -dj : dj;
      if (!aj || ((inf >> aj) & 1u)) continue;
      T = tbl[aj];
      if (j) g2_psi_signed(T, j);
      if (dj < 0) fp2_neg(T.y, T.y);
      pt_madd(acc, acc, T);
    }
  }
  R = acc;
}

// variable-base multiplication used by the protocol kernels
PSB_HD PSB_INL void pt_mul(G1J& R, const G1J& P, const uint32_t* k) { g1_mul_glv(R, P, k); }
PSB_HD PSB_INL void pt_mul(G2J& R, const G2J& P, const uint32_t* k) { g2_mul_gls(R, P, k); }
#else
// ---- BN254: GLV in G1 (2 dimensions), GLS in G2 (4 dimensions) ---------------------------------------------------------
// Same group elements as mcl's GLV1 / GLV2 on BN curves (ec.hpp:1457-1523, bn.hpp:765-860); the lattices, the rounding and
// the schedule are ours (tools/gen_constants.py derives and checks the constants on the integer operations used here):
//   G1: phi(X, Y, Z) = (beta X, Y, Z) = [lambda] P, lambda = 36 z^4 - 1;  k = k1 + k2 lambda, |k1|, |k2| < 2^127 by Babai
//       rounding on a reduced basis of {(a, b): a + b lambda = 0 mod r}.  128 doublings instead of 256.
//   G2: psi = [mu], mu = p mod r = 6 z^2;  k = d0 + d1 mu + d2 mu^2 + d3 mu^3, |d_i| < 2^64 by Babai rounding on an
//       LLL-reduced basis of the 4-dimensional lattice.  64 doublings instead of 256.
// Rounded quotients are c = (k g + 2^(32 S - 1)) >> 32 S with g = |round(2^(32 S) coefficient)|, S = PSB_GLVBN_SHIFT_LIMBS;
// only their low 128 bits are needed: the remainders are computed mod 2^128 in two's complement and are small.
typedef unsigned __int128 u128;
// low 128 bits of (k * g + 2^(32 S - 1)) >> 32 S;  k: 8 limbs, g: gl limbs
PSB_HD PSB_INL u128 mul_round_shift(const uint32_t* k, const uint32_t* g, int gl) {
  constexpr int S = PSB_GLVBN_SHIFT_LIMBS;
  uint32_t out[4] = {0, 0, 0, 0};
  uint64_t lo = 0, hi = 0;
  for (int col = 0; col < S + 4; col++) {
    if (col == S - 1) { const uint64_t nl = lo + 0x80000000ull; hi += nl < lo; lo = nl; }   // + 2^(32 S - 1)
    for (int i = 0; i < 8; i++) {
      const int j = col - i;
      if (j < 0 || j >= gl) continue;
      const uint64_t pr = (uint64_t)k[i] * g[j];
      const uint64_t nl = lo + pr;
      hi += nl < lo;
      lo = nl;
    }
    if (col >= S) out[col - S] = (uint32_t)lo;
    lo = (lo >> 32) | (hi << 32);
    hi = 0;
  }
  return ((u128)out[3] << 96) | ((u128)out[2] << 64) | ((u128)out[1] << 32) | out[0];
}
PSB_HD PSB_INL u128 load_u128(const uint32_t* w) { return ((u128)w[3] << 96) | ((u128)w[2] << 64) | ((u128)w[1] << 32) | w[0]; }
// |v| of a two's-complement 128-bit value, neg = sign
PSB_HD PSB_INL u128 abs_s128(u128 v, bool& neg) { neg = (v >> 127) != 0; return neg ? (u128)0 - v : v; }
PSB_HD PSB_INL uint32_t nibble_u128(u128 v, int i) { return (uint32_t)(v >> (4 * i)) & 0xFu; }

PSB_HD PSB_NOINL void g1_mul_glv(G1J& R, const G1J& P, const uint32_t* k) {
  const u128 c1 = mul_round_shift(k, PSB_K(GLVBN_G), 6), c2 = mul_round_shift(k, PSB_K(GLVBN_G) + 6, 6);
  const u128 klo = load_u128(k);
  bool n1, n2;
  const u128 k1 = abs_s128(klo - c1 * load_u128(PSB_K(GLVBN_SA)) - c2 * load_u128(PSB_K(GLVBN_SA) + 4), n1);
  const u128 k2 = abs_s128((u128)0 - c1 * load_u128(PSB_K(GLVBN_SB)) - c2 * load_u128(PSB_K(GLVBN_SB) + 4), n2);
  int8_t e1[33], e2[33];
  signed_nibbles(e1, 32, [&](int i) { return nibble_u128(k1, i); });
  signed_nibbles(e2, 32, [&](int i) { return nibble_u128(k2, i); });
  G1A tbl[9], T;
  const uint32_t inf = pt_affine_multiples8(tbl, P);
  Fp beta;
  for (int i = 0; i < PSB_NL; i++) beta.v[i] = PSB_K(GLV_BETA)[i];
  G1J acc;
  pt_set_zero(acc);
  for (int i = 32; i >= 0; i--) {
    if (i < 32) { pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); }
    const int d1 = e1[i], d2 = e2[i];
    const int a1 = d1 < 0 ? -d1 : d1, a2 = d2 < 0 ? -d2 : d2;
    if (a1 && !((inf >> a1) & 1u)) {
      T = tbl[a1];
      if ((d1 < 0) != n1) fp_neg(T.y, T.y);
      pt_madd(acc, acc, T);
    }
    if (a2 && !((inf >> a2) & 1u)) {
      T = tbl[a2];
      fp_mul(T.x, T.x, beta);          // phi(d2 P)
      if ((d2 < 0) != n2) fp_neg(T.y, T.y);
      pt_madd(acc, acc, T);
    }
  }
  R = acc;
}

// psi^j(T) for an AFFINE point, j = 1..3 (D-type twist constants from tools/gen_constants.py):
//   psi(x, y) = (conj(x) cx, conj(y) cy), psi^2(x, y) = (x N(cx), -y), psi^3(x, y) = (conj(x) cx3, -conj(y) cy)
PSB_HD PSB_NOINL void g2_psi_pow(G2A& T, int j) {
  if (j == 2) {
    Fp n;
    for (int i = 0; i < PSB_NL; i++) n.v[i] = PSB_K(PSI_NCX)[i];
    fp2_mul_fp(T.x, T.x, n);
    fp2_neg(T.y, T.y);
    return;
  }
  Fp2 c;
  const uint32_t* cxp = (j == 1) ? PSB_K(PSI_CX) : PSB_K(PSI_CX3);
  for (int i = 0; i < PSB_NL; i++) { c.a.v[i] = cxp[i]; c.b.v[i] = cxp[PSB_NL + i]; }
  fp2_conj(T.x, T.x); fp2_conj(T.y, T.y);
  fp2_mul(T.x, T.x, c);
  for (int i = 0; i < PSB_NL; i++) { c.a.v[i] = PSB_K(PSI_CY)[i]; c.b.v[i] = PSB_K(PSI_CY)[PSB_NL + i]; }
  fp2_mul(T.y, T.y, c);
  if (j == 3) fp2_neg(T.y, T.y);
}

PSB_HD PSB_NOINL void g2_mul_gls(G2J& R, const G2J& P, const uint32_t* k) {
  u128 c[4], d[4];
  bool neg[4];
  for (int j = 0; j < 4; j++) c[j] = mul_round_shift(k, PSB_K(GLSBN_G) + 8 * j, 8);
  for (int i = 0; i < 4; i++) {
    u128 v = i == 0 ? load_u128(k) : (u128)0;
    for (int j = 0; j < 4; j++) v -= c[j] * load_u128(PSB_K(GLSBN_SB) + 4 * (4 * j + i));
    d[i] = abs_s128(v, neg[i]);
  }
  int8_t e[4][PSB_GLS_BN_WINDOWS + 1];
  for (int j = 0; j < 4; j++) signed_nibbles(e[j], PSB_GLS_BN_WINDOWS, [&](int i) { return nibble_u128(d[j], i); });
  G2A tbl[9], T;
  const uint32_t inf = pt_affine_multiples8(tbl, P);
  G2J acc;
  pt_set_zero(acc);
  for (int i = PSB_GLS_BN_WINDOWS; i >= 0; i--) {
    if (i < PSB_GLS_BN_WINDOWS) { pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); pt_dbl(acc, acc); }
    for (int j = 0; j < 4; j++) {
      const int dj = e[j][i], aj = dj < 0 ? -dj : dj;
      if (!aj || ((inf >> aj) & 1u)) continue;
      T = tbl[aj];
      if (j) g2_psi_pow(T, j);
      if ((dj < 0) != neg[j]) fp2_neg(T.y, T.y);
      pt_madd(acc, acc, T);
    }
  }
  R = acc;
}

PSB_HD PSB_INL void pt_mul(G1J& R, const G1J& P, const uint32_t* k) { g1_mul_glv(R, P, k); }
PSB_HD PSB_INL void pt_mul(G2J& R, const G2J& P, const uint32_t* k) { g2_mul_gls(R, P, k); }
#endif

// ---- fixed-base windows --------------------------------------------------------------------------
// Signed w-bit recoding of a 256-bit normal-form scalar (< 2^255): digits in [-2^(w-1), 2^(w-1)],
// nwin = ceil(256 / w).  dig[j] receives the signed digit of window j.
PSB_HD PSB_INL int fixed_nwin(int w) { return (256 + w - 1) / w; }
PSB_HD PSB_INL uint32_t scalar_bits(const uint32_t* k, int pos, int w) {
  // w <= 16 bits starting at bit `pos` (may straddle a limb; bits >= 256 read as zero)
  const int limb = pos >> 5, sh = pos & 31;
  uint64_t v = (limb < 8) ? k[limb] : 0u;
  if (limb + 1 < 8) v |= (uint64_t)k[limb + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << w) - 1u);
}

// acc += k * B using the table of base `b`: entry (win, d) = d * 2^(w*win) * B, d = 1..2^(w-1), affine.
// tbl points at the first entry of this base: index = win * 2^(w-1) + (d - 1).
template <class F>
PSB_HD PSB_INL void pt_fixed_mul_acc(Jac<F>& acc, const Aff<F>* tbl, const uint32_t* k, int w) {
  const int nwin = fixed_nwin(w);
  const uint32_t half = 1u << (w - 1);
  uint32_t carry = 0;
  for (int j = 0; j < nwin; j++) {
    uint32_t d = scalar_bits(k, j * w, w) + carry;
    carry = 0;
    bool neg = false;
    if (d > half) { d = (1u << w) - d; neg = true; carry = 1; }
    if (d != 0) {
      Aff<F> e = tbl[(size_t)j * half + (d - 1)];
      if (neg) f_neg(e.y, e.y);
      pt_madd(acc, acc, e);
    }
  }
}

// ---- fixed-base sums with batched affine additions ----------------------------------------------------------
// A multi-scalar fixed-base sum is a sum of TABLE ENTRIES (affine points), n * nwin of them per lane.  Adding two affine
// points costs 1 inversion + 2M + 1S; with Montgomery's trick the inversions of a whole batch of independent pairs become ONE
// inversion plus 3M per pair, so a pair sum is 5M + 1S against 7M + 4S for pushing one of the two entries through the mixed
// Jacobian addition.  AffBatch pairs the entries as they come (slot 2i with slot 2i+1), forms the pair sums in affine
// coordinates and feeds those to the Jacobian accumulator: per two entries 5M + 1S + (7M + 4S) + 1/G of an inversion instead
// of 2 (7M + 4S).  Every lane pushes the same number of slots (a zero digit pushes kAffNone), so the flushes are warp-uniform.
// Pairs the affine formula cannot take (an absent entry, P = +-Q) go through pt_madd entry by entry: same group element.
// Per-lane state: 2G slot descriptors (29-bit table index | table id << 29 | sign << 31) and G prefix products; a batch
// draws its entries from up to three tables (the per-key and per-batch tables of the EL PASSO kernels).
// Measured on B200 (profiles/r2s_ab_msm_batched_affine.txt): k_verify_msm 76.4 -> 64.6 ms per 2^20 lanes at 5 attributes.
constexpr int kAffG = 64;                     // pairs per inversion at most (the exception mask is one 64-bit word)
constexpr uint32_t kAffNone = 0xFFFFFFFFu;    // table id 3 is never used, so this is no entry
constexpr uint32_t kAffIdx = 0x1FFFFFFFu;
constexpr size_t kAffMaxEntries = (size_t)1 << 29;
// below this many pairs one inversion (~110 Fp products of time) is not repaid by 4 Fp2 products + 3 squarings per pair
constexpr int kAffMinPairs = 10;
// optional second level: the pair sums of a batch are themselves paired up (explicit points) before they reach the Jacobian
// accumulator -- per FOUR entries 3 (5M + 1S) + (7M + 4S) instead of 2 (5M + 1S) + 2 (7M + 4S).  12 KB of lane state in G2.
constexpr int kAffL2 = 64;                    // points per second-level flush
template <class F>
struct AffPts {
  Aff<F> pts[kAffL2];
  F pre[kAffL2 / 2];
  uint64_t absent;                            // bit i: slot i holds no point
  int cnt;
};
template <class F>
struct AffBatch {
  uint32_t desc[2 * kAffG];
  F pre[kAffG];
  const Aff<F>* tbl[3];
  AffPts<F>* l2;                              // nullptr: pair sums go straight to the accumulator
  int cnt, cap;                               // slots filled / slots per flush (even, <= 2 kAffG)
  bool plain;                                 // a table too large for the 29-bit slot index: plain mixed-addition chain
};
// slots per flush for `total` slots: as few flushes as kAffG allows, of equal size
PSB_HD PSB_INL int aff_batch_cap(int total) {
  const int pairs = total / 2;
  if (pairs <= 0) return 2;
  const int nb = (pairs + kAffG - 1) / kAffG;
  return 2 * ((pairs + nb - 1) / nb);
}
// `max_entries`: size of the largest table the batch draws from (entries); kept 32-bit on purpose -- 64-bit descriptors cost
// k_vid_g2 78 registers and the whole gain (profiles/r2z_ab_config_windows.txt)
template <class F>
PSB_HD PSB_INL void aff_init(AffBatch<F>& b, int total_slots, size_t max_entries, const Aff<F>* t0, const Aff<F>* t1 = nullptr,
                             const Aff<F>* t2 = nullptr, AffPts<F>* l2 = nullptr) {
  b.cnt = 0; b.cap = aff_batch_cap(total_slots);
  b.plain = max_entries >= kAffMaxEntries;
  b.tbl[0] = t0; b.tbl[1] = t1; b.tbl[2] = t2;
  b.l2 = total_slots >= 4 * kAffMinPairs ? l2 : nullptr;     // fewer than kAffMinPairs second-level pairs: not worth it
  if (b.l2) { b.l2->cnt = 0; b.l2->absent = 0; }
}
template <class F> PSB_HD PSB_INL const Aff<F>* aff_entry(const AffBatch<F>& b, uint32_t d) { return b.tbl[(d >> 29) & 3u] + (d & kAffIdx); }
template <class F> PSB_HD PSB_INL void aff_fetch(Aff<F>& e, const AffBatch<F>& b, uint32_t d) {
  e = *aff_entry(b, d);
  if (d >> 31) f_neg(e.y, e.y);
}

// S = P + Q for affine P != +-Q, given li = 1 / (Q.x - P.x): 2M + 1S
template <class F>
PSB_HD PSB_INL void aff_pair_sum(Aff<F>& S, const Aff<F>& P, const Aff<F>& Q, const F& li) {
  F lam, t;
  f_sub(t, Q.y, P.y); f_mul(lam, t, li);
  f_sqr(t, lam); f_sub(t, t, P.x); f_sub(S.x, t, Q.x);
  f_sub(t, P.x, S.x); f_mul(t, lam, t); f_sub(S.y, t, P.y);
}

// second level: flush the explicit points (pairs 2i / 2i+1) into the accumulator
template <class F>
PSB_HD PSB_NOINL void aff_flush_l2(Jac<F>& acc, AffPts<F>& l) {
  const int np = l.cnt >> 1;
  Aff<F> S;
  if (np >= kAffMinPairs) {
    uint64_t exc = 0;
    F run, d;
    for (int i = 0; i < np; i++) {
      bool bad = ((l.absent >> (2 * i)) & 3ull) != 0;
      if (!bad) {
        f_sub(d, l.pts[2 * i + 1].x, l.pts[2 * i].x);
        bad = f_is_zero(d);
      }
      if (bad) { f_set_one(d); exc |= 1ull << i; }
      if (i == 0) l.pre[0] = d; else f_mul(l.pre[i], l.pre[i - 1], d);     // (written in place: see pt_affine_multiples8)
    }
    f_inv(run, l.pre[np - 1]);
    for (int i = np - 1; i >= 0; i--) {
      if ((exc >> i) & 1ull) {
        if (!((l.absent >> (2 * i)) & 1ull)) pt_madd(acc, acc, l.pts[2 * i]);
        if (!((l.absent >> (2 * i + 1)) & 1ull)) pt_madd(acc, acc, l.pts[2 * i + 1]);
        continue;
      }
      F li;
      f_sub(d, l.pts[2 * i + 1].x, l.pts[2 * i].x);
      if (i) f_mul(li, run, l.pre[i - 1]); else li = run;
      f_mul(run, run, d);
      aff_pair_sum(S, l.pts[2 * i], l.pts[2 * i + 1], li);
      pt_madd(acc, acc, S);
    }
  } else {
    for (int i = 0; i < 2 * np; i++)
      if (!((l.absent >> i) & 1ull)) pt_madd(acc, acc, l.pts[i]);
  }
  if ((l.cnt & 1) && !((l.absent >> (l.cnt - 1)) & 1ull)) pt_madd(acc, acc, l.pts[l.cnt - 1]);
  l.cnt = 0;
  l.absent = 0;
}
// hand one first-level result (or "nothing") to the second level; every lane hands over the same number of slots
template <class F>
PSB_HD PSB_INL void aff_l2_push(Jac<F>& acc, AffPts<F>& l, const Aff<F>* S) {
  if (S) l.pts[l.cnt] = *S; else l.absent |= 1ull << l.cnt;
  l.cnt++;
  if (l.cnt == kAffL2) aff_flush_l2(acc, l);
}

// first level: flush the table entries named by the slot descriptors.  `last`: nothing follows, drain the second level too.
template <class F>
PSB_HD PSB_NOINL void aff_flush(Jac<F>& acc, AffBatch<F>& b, bool last = true) {
  const int np = b.cnt >> 1;
  Aff<F> P, Q, S;
  if (np >= kAffMinPairs) {
    uint64_t exc = 0;
    F run, d;
    for (int i = 0; i < np; i++) {
      const uint32_t da = b.desc[2 * i], db = b.desc[2 * i + 1];
      bool bad = da == kAffNone || db == kAffNone;
      if (!bad) {
        f_sub(d, aff_entry(b, db)->x, aff_entry(b, da)->x);
        bad = f_is_zero(d);
      }
      if (bad) { f_set_one(d); exc |= 1ull << i; }
      if (i == 0) b.pre[0] = d; else f_mul(b.pre[i], b.pre[i - 1], d);     // (written in place: see pt_affine_multiples8)
    }
    f_inv(run, b.pre[np - 1]);                // 1 / (d_0 ... d_{np-1})
    for (int i = np - 1; i >= 0; i--) {
      const uint32_t da = b.desc[2 * i], db = b.desc[2 * i + 1];
      bool have = true;                       // S holds the pair sum
      if ((exc >> i) & 1ull) {                // d_i = 1: `run` is already 1 / (d_0 ... d_{i-1})
        const bool ha = da != kAffNone, hb = db != kAffNone;
        if (ha && hb) {                       // P = +-Q: through the Jacobian formulas
          aff_fetch(P, b, da); pt_madd(acc, acc, P);
          aff_fetch(Q, b, db); pt_madd(acc, acc, Q);
          have = false;
        } else if (ha || hb) {                // one entry: it IS the pair sum
          aff_fetch(S, b, ha ? da : db);
        } else {
          have = false;
        }
      } else {
        aff_fetch(P, b, da);
        aff_fetch(Q, b, db);
        F li;
        f_sub(d, Q.x, P.x);
        if (i) f_mul(li, run, b.pre[i - 1]); else li = run;     // 1 / d_i
        f_mul(run, run, d);
        aff_pair_sum(S, P, Q, li);
      }
      // (the hand-over sits after the branches: every lane reaches a second-level flush at the same point)
      if (b.l2) aff_l2_push(acc, *b.l2, have ? &S : (const Aff<F>*)nullptr);
      else if (have) pt_madd(acc, acc, S);
    }
  } else {
    for (int i = 0; i < 2 * np; i++)
      if (b.desc[i] != kAffNone) { aff_fetch(P, b, b.desc[i]); pt_madd(acc, acc, P); }
  }
  if (b.cnt & 1) {
    const uint32_t dl = b.desc[b.cnt - 1];
    if (dl != kAffNone) { aff_fetch(P, b, dl); pt_madd(acc, acc, P); }
  }
  b.cnt = 0;
  if (last && b.l2) aff_flush_l2(acc, *b.l2);
}

// push the nwin table entries of k * B (the entries pt_fixed_mul_acc would add): `first` = index of the first entry of
// this base in table `tid` of the batch
template <class F>
PSB_HD PSB_INL void aff_push_fixed_mul(Jac<F>& acc, AffBatch<F>& b, int tid, size_t first, const uint32_t* k, int w) {
  if (b.plain) { pt_fixed_mul_acc(acc, b.tbl[tid] + first, k, w); return; }
  const int nwin = fixed_nwin(w);
  const uint32_t half = 1u << (w - 1);
  const uint32_t tag = (uint32_t)tid << 29;
  uint32_t carry = 0;
  for (int j = 0; j < nwin; j++) {
    uint32_t d = scalar_bits(k, j * w, w) + carry;
    carry = 0;
    uint32_t neg = 0;
    if (d > half) { d = (1u << w) - d; neg = 1u << 31; carry = 1; }
    b.desc[b.cnt++] = d ? ((uint32_t)(first + (size_t)j * half + (d - 1)) | tag | neg) : kAffNone;
    if (b.cnt == b.cap) aff_flush(acc, b, false);
  }
}

}  // namespace psb
