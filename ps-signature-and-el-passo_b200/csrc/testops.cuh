// testops.cuh -- element-wise dispatch of every arithmetic layer, for parity tests against the
// oracle.  Used by the psb_test_op() entry of libpsb.so (one GPU thread per element) and by the
// CPU "hostsim" test library (same source compiled for the host; tests only).
#pragma once
#include "pairing.cuh"
#include "sha256.cuh"

namespace psb {

enum TestOp {
  T_FP_ADD = 0, T_FP_SUB, T_FP_MUL, T_FP_SQR, T_FP_NEG, T_FP_INV,
  T_FP2_ADD = 10, T_FP2_SUB, T_FP2_MUL, T_FP2_SQR, T_FP2_NEG, T_FP2_INV,
  T_FP6_MUL = 20, T_FP6_INV,
  T_FP12_MUL = 30, T_FP12_SQR, T_FP12_INV, T_FP12_FROB1, T_FP12_FROB2, T_FP12_FROB3, T_FP12_CYCLO_SQR,
  T_FP12_MUL_LINE,  // a = Fp12, b = 3 Fp2: M-type twist c0 + c2 w^2 + c3 w^3, D-type twist c0 + c1 w + c3 w^3
  T_G1_ADD = 40, T_G1_DBL, T_G1_NORM, T_G1_MUL /* b = Fr Montgomery */, T_G1_MADD /* b = normalised G1 */, T_G1_AFFMUL /* b = Fr m in 1..8: entry m of pt_affine_multiples8 */,
  T_G2_ADD = 50, T_G2_DBL, T_G2_NORM, T_G2_MUL, T_G2_MADD, T_G2_AFFMUL,
  T_PAIRING = 60,      // a = G1, b = G2 -> GT = finalExp(millerLoop(a, b))
  T_MILLER_FE_ONLY,    // a = Fp12 -> finalExp(a)
  T_PAIRING_RATIO,     // a = G1 P1, b = G2 Q1, c = G1 P2 | G2 Q2 (normalised) -> e(P1,Q1) e(P2,Q2)^-1
  T_FR_FROM_MONT = 70, T_FR_MUL, T_FR_SUB, T_FR_ADD,
};

// sizes in u32 words: {a, b, c, out}; 0 = unused
PSB_HD inline bool test_op_shape(int op, int s[4]) {
  const int FP = PSB_NL, FP2 = 2 * FP, FP6 = 6 * FP, FP12 = 12 * FP, G1 = 3 * FP, G2 = 6 * FP, FR = 8;
  s[0] = s[1] = s[2] = s[3] = 0;
  switch (op) {
    case T_FP_ADD: case T_FP_SUB: case T_FP_MUL: s[0] = s[1] = s[3] = FP; return true;
    case T_FP_SQR: case T_FP_NEG: case T_FP_INV: s[0] = s[3] = FP; return true;
    case T_FP2_ADD: case T_FP2_SUB: case T_FP2_MUL: s[0] = s[1] = s[3] = FP2; return true;
    case T_FP2_SQR: case T_FP2_NEG: case T_FP2_INV: s[0] = s[3] = FP2; return true;
    case T_FP6_MUL: s[0] = s[1] = s[3] = FP6; return true;
    case T_FP6_INV: s[0] = s[3] = FP6; return true;
    case T_FP12_MUL: s[0] = s[1] = s[3] = FP12; return true;
    case T_FP12_SQR: case T_FP12_INV: case T_FP12_FROB1: case T_FP12_FROB2: case T_FP12_FROB3:
    case T_FP12_CYCLO_SQR: s[0] = s[3] = FP12; return true;
    case T_FP12_MUL_LINE: s[0] = s[3] = FP12; s[1] = 3 * FP2; return true;
    case T_G1_ADD: case T_G1_MADD: s[0] = s[1] = s[3] = G1; return true;
    case T_G1_DBL: case T_G1_NORM: s[0] = s[3] = G1; return true;
    case T_G1_MUL: case T_G1_AFFMUL: s[0] = s[3] = G1; s[1] = FR; return true;
    case T_G2_ADD: case T_G2_MADD: s[0] = s[1] = s[3] = G2; return true;
    case T_G2_DBL: case T_G2_NORM: s[0] = s[3] = G2; return true;
    case T_G2_MUL: case T_G2_AFFMUL: s[0] = s[3] = G2; s[1] = FR; return true;
    case T_PAIRING: s[0] = G1; s[1] = G2; s[3] = FP12; return true;
    case T_MILLER_FE_ONLY: s[0] = s[3] = FP12; return true;
    case T_PAIRING_RATIO: s[0] = G1; s[1] = G2; s[2] = G1 + G2; s[3] = FP12; return true;
    case T_FR_FROM_MONT: s[0] = s[3] = FR; return true;
    case T_FR_MUL: case T_FR_SUB: case T_FR_ADD: s[0] = s[1] = s[3] = FR; return true;
  }
  return false;
}

template <class T>
PSB_HD PSB_INL void ld(T& dst, const uint32_t* src) {
  uint32_t* d = (uint32_t*)&dst;
  for (unsigned i = 0; i < sizeof(T) / 4; i++) d[i] = src[i];
}
template <class T>
PSB_HD PSB_INL void st(uint32_t* dst, const T& src) {
  const uint32_t* s = (const uint32_t*)&src;
  for (unsigned i = 0; i < sizeof(T) / 4; i++) dst[i] = s[i];
}

// affine Fp coordinates of a G1 input for the Miller loop: z == 1 -> as is; z == 0 -> (0,0);
// otherwise normalise (mcl normalises P and Q first: bn.hpp:1664-1665)
PSB_HD PSB_NOINL void g1_affine_for_pairing(Fp& x, Fp& y, const G1J& P) {
  Fp one;
  fp_set_one(one);
  if (fp_is_zero(P.z)) { fp_set_zero(x); fp_set_zero(y); return; }
  if (fp_eq(P.z, one)) { x = P.x; y = P.y; return; }
  G1J n;
  pt_normalize(n, P);
  x = n.x; y = n.y;
}

PSB_HD inline void test_op_run(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out) {
  switch (op) {
    case T_FP_ADD: case T_FP_SUB: case T_FP_MUL: case T_FP_SQR: case T_FP_NEG: case T_FP_INV: {
      Fp x, y, r; ld(x, a); if (b) ld(y, b);
      if (op == T_FP_ADD) fp_add(r, x, y); else if (op == T_FP_SUB) fp_sub(r, x, y);
      else if (op == T_FP_MUL) fp_mul(r, x, y); else if (op == T_FP_SQR) fp_sqr(r, x);
      else if (op == T_FP_NEG) fp_neg(r, x); else fp_inv(r, x);
      st(out, r); break; }
    case T_FP2_ADD: case T_FP2_SUB: case T_FP2_MUL: case T_FP2_SQR: case T_FP2_NEG: case T_FP2_INV: {
      Fp2 x, y, r; ld(x, a); if (b) ld(y, b);
      if (op == T_FP2_ADD) fp2_add(r, x, y); else if (op == T_FP2_SUB) fp2_sub(r, x, y);
      else if (op == T_FP2_MUL) fp2_mul(r, x, y); else if (op == T_FP2_SQR) fp2_sqr(r, x);
      else if (op == T_FP2_NEG) fp2_neg(r, x); else fp2_inv(r, x);
      st(out, r); break; }
    case T_FP6_MUL: case T_FP6_INV: {
      Fp6 x, y, r; ld(x, a); if (b) ld(y, b);
      if (op == T_FP6_MUL) fp6_mul(r, x, y); else fp6_inv(r, x);
      st(out, r); break; }
    case T_FP12_MUL: case T_FP12_SQR: case T_FP12_INV: case T_FP12_FROB1: case T_FP12_FROB2:
    case T_FP12_FROB3: case T_FP12_CYCLO_SQR: case T_FP12_MUL_LINE: case T_MILLER_FE_ONLY: {
      Fp12 x, r; ld(x, a);
      if (op == T_FP12_MUL) { Fp12 y; ld(y, b); fp12_mul(r, x, y); }
      else if (op == T_FP12_SQR) fp12_sqr(r, x);
      else if (op == T_FP12_INV) fp12_inv(r, x);
      else if (op == T_FP12_FROB1) fp12_frobenius(r, x, 1);
      else if (op == T_FP12_FROB2) fp12_frobenius(r, x, 2);
      else if (op == T_FP12_FROB3) fp12_frobenius(r, x, 3);
      else if (op == T_FP12_CYCLO_SQR) fp12_cyclo_sqr(r, x);
      else if (op == T_MILLER_FE_ONLY) final_exp(r, x);
      else {
        Fp2 c0, c2, c3; ld(c0, b); ld(c2, b + 2 * PSB_NL); ld(c3, b + 4 * PSB_NL); r = x;
#if PSB_TWIST_MTYPE
        fp12_mul_line(r, c0, c2, c3);        // c0 + c2 w^2 + c3 w^3
#else
        fp12_mul_line_d(r, c0, c2, c3);      // c0 + c2 w + c3 w^3
#endif
      }
      st(out, r); break; }
    case T_G1_ADD: case T_G1_DBL: case T_G1_NORM: case T_G1_MUL: case T_G1_MADD: case T_G1_AFFMUL: {
      G1J x, y, r; ld(x, a);
      if (op == T_G1_AFFMUL) {
        Fr k, kn; ld(k, b); fr_from_mont(kn, k);
        G1A t[9];
        const uint32_t inf = pt_affine_multiples8(t, x);
        const uint32_t m = kn.v[0] & 15u;
        if (m < 1 || m > 8 || ((inf >> m) & 1u)) pt_set_zero(r); else pt_from_aff(r, t[m]);
      }
      else if (op == T_G1_ADD) { ld(y, b); pt_add(r, x, y); }
      else if (op == T_G1_DBL) pt_dbl(r, x);
      else if (op == T_G1_NORM) pt_normalize(r, x);
      else if (op == T_G1_MADD) { ld(y, b); G1A q; q.x = y.x; q.y = y.y; pt_madd(r, x, q); }
      else { Fr k, kn; ld(k, b); fr_from_mont(kn, k); pt_mul(r, x, kn.v); }
      st(out, r); break; }
    case T_G2_ADD: case T_G2_DBL: case T_G2_NORM: case T_G2_MUL: case T_G2_MADD: case T_G2_AFFMUL: {
      G2J x, y, r; ld(x, a);
      if (op == T_G2_AFFMUL) {
        Fr k, kn; ld(k, b); fr_from_mont(kn, k);
        G2A t[9];
        const uint32_t inf = pt_affine_multiples8(t, x);
        const uint32_t m = kn.v[0] & 15u;
        if (m < 1 || m > 8 || ((inf >> m) & 1u)) pt_set_zero(r); else pt_from_aff(r, t[m]);
      }
      else if (op == T_G2_ADD) { ld(y, b); pt_add(r, x, y); }
      else if (op == T_G2_DBL) pt_dbl(r, x);
      else if (op == T_G2_NORM) pt_normalize(r, x);
      else if (op == T_G2_MADD) { ld(y, b); G2A q; q.x = y.x; q.y = y.y; pt_madd(r, x, q); }
      else { Fr k, kn; ld(k, b); fr_from_mont(kn, k); pt_mul(r, x, kn.v); }
      st(out, r); break; }
    case T_PAIRING: {
      G1J P; G2J Q; ld(P, a); ld(Q, b);
      Fp x1, y1, zero; fp_set_zero(zero);
      g1_affine_for_pairing(x1, y1, P);
      Fp12 f, e;
      miller_loop2(f, x1, y1, Q, zero, zero, nullptr, false);
      final_exp(e, f);
      st(out, e); break; }
    case T_PAIRING_RATIO: {
      G1J P1, P2; G2J Q1, Q2; ld(P1, a); ld(Q1, b); ld(P2, c); ld(Q2, c + 3 * PSB_NL);
      Fp x1, y1, x2, y2;
      g1_affine_for_pairing(x1, y1, P1);
      g1_affine_for_pairing(x2, y2, P2);
      fp_neg(y2, y2);
      FixedLine lines[kFixedLineSlots];
      G2A q2; q2.x = Q2.x; q2.y = Q2.y;
      precompute_fixed_lines(lines, q2);
      Fp12 f, e;
      miller_loop2(f, x1, y1, Q1, x2, y2, lines, true);
      final_exp(e, f);
      st(out, e); break; }
    case T_FR_FROM_MONT: case T_FR_MUL: case T_FR_SUB: case T_FR_ADD: {
      Fr x, y, r; ld(x, a); if (b) ld(y, b);
      if (op == T_FR_FROM_MONT) fr_from_mont(r, x); else if (op == T_FR_MUL) fr_mul(r, x, y);
      else if (op == T_FR_SUB) fr_sub(r, x, y); else fr_add(r, x, y);
      st(out, r); break; }
  }
}

}  // namespace psb
