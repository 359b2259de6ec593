// tower.cuh -- Fp2 / Fp6 / Fp12 for BLS12-381.
//
// Tower (same as mcl so that raw bytes are interchangeable):
//   Fp2  = Fp[i]/(i^2+1)            (mcl/include/mcl/fp_tower.hpp:214-611)
//   Fp6  = Fp2[v]/(v^3 - xi), xi=1+i (fp_tower.hpp:786-1060)
//   Fp12 = Fp6[w]/(w^2 - v)          (fp_tower.hpp:1066-1372)
// Memory order of an Fp12: a.a a.b a.c b.a b.b b.c = coefficients of w^0 w^2 w^4 w^1 w^3 w^5.
//
// Lazy reduction happens INSIDE the Fp2 product: each Fp2 coefficient is one fused two-product
// Montgomery pass (fp_dot2: a0 b0 + a1 (-b1), a0 b1 + a1 b0) with a single reduction, everything in
// registers.  mcl additionally carries unreduced 768-bit values across Fp6/Fp12 (Fp6Dbl,
// fp_tower.hpp:978-1022); on the GPU those wide temporaries live in thread-local memory and the
// measured L1/LSU traffic of that scheme outweighed the saved reductions, so above
// Fp2 all values are canonical 48-byte elements.
#pragma once
#include "fp.cuh"

namespace psb {

struct Fp2 { Fp a, b; };
struct Fp6 { Fp2 a, b, c; };
struct Fp12 { Fp6 a, b; };

// ---- Fp2 ------------------------------------------------------------------------------------------
PSB_HD PSB_NOINL void fp2_add(Fp2& r, const Fp2& x, const Fp2& y) { fp_add(r.a, x.a, y.a); fp_add(r.b, x.b, y.b); }
PSB_HD PSB_NOINL void fp2_sub(Fp2& r, const Fp2& x, const Fp2& y) { fp_sub(r.a, x.a, y.a); fp_sub(r.b, x.b, y.b); }
PSB_HD PSB_NOINL void fp2_neg(Fp2& r, const Fp2& x) { fp_neg(r.a, x.a); fp_neg(r.b, x.b); }
PSB_HD PSB_NOINL void fp2_dbl(Fp2& r, const Fp2& x) { fp_dbl(r.a, x.a); fp_dbl(r.b, x.b); }
PSB_HD PSB_INL void fp2_conj(Fp2& r, const Fp2& x) { r.a = x.a; fp_neg(r.b, x.b); }
PSB_HD PSB_INL void fp2_set_zero(Fp2& r) { fp_set_zero(r.a); fp_set_zero(r.b); }
PSB_HD PSB_INL void fp2_set_one(Fp2& r) { fp_set_one(r.a); fp_set_zero(r.b); }
PSB_HD PSB_INL bool fp2_is_zero(const Fp2& x) { return fp_is_zero(x.a) & fp_is_zero(x.b); }
PSB_HD PSB_INL bool fp2_eq(const Fp2& x, const Fp2& y) { return fp_eq(x.a, y.a) & fp_eq(x.b, y.b); }
PSB_HD PSB_INL void fp2_cmov(Fp2& r, const Fp2& x, bool c) { fp_cmov(r.a, x.a, c); fp_cmov(r.b, x.b, c); }
// (a + b i)(1 + i) = (a - b) + (a + b) i      (fp_tower.hpp:584-592)
PSB_HD PSB_NOINL void fp2_mul_xi(Fp2& r, const Fp2& x) {
  Fp t;
  fp_sub(t, x.a, x.b);
  fp_add(r.b, x.a, x.b);
  r.a = t;
}
PSB_HD PSB_INL void fp2_mul_fp(Fp2& r, const Fp2& x, const Fp& k) {
#if defined(__CUDA_ARCH__) && defined(PSB_FP2_FUSED)
  cios::fp2_mul_fp(r.a.v, x.a.v, k.v);   // Fp2 is two contiguous Fp: 24 limbs
#else
  fp_mul(r.a, x.a, k); fp_mul(r.b, x.b, k);
#endif
}

// (a0 + a1 i)(b0 + b1 i) = (a0 b0 - a1 b1) + (a0 b1 + a1 b0) i : two fused two-product passes
// (value of mcl Fp2::mul, fp_tower.hpp:528-534)
#if defined(__CUDA_ARCH__) && defined(PSB_FP2_FUSED)
// two accumulator pairs side by side: more ILP but ~160 registers -> lower occupancy; measured
// slower on B200 (750k vs 832k verif/s, r1), kept for experiments only
__device__ PSB_INL void fp2_mul(Fp2& r, const Fp2& x, const Fp2& y) { cios::fp2_mul(r.a.v, x.a.v, y.a.v); }
__device__ PSB_INL void fp2_sqr(Fp2& r, const Fp2& x) { cios::fp2_sqr(r.a.v, x.a.v); }
#else
PSB_HD PSB_NOINL void fp2_mul(Fp2& r, const Fp2& x, const Fp2& y) {
  Fp nb, re;
  fp_neg(nb, y.b);
  fp_dot2(re, x.a, y.a, x.b, nb);
  fp_dot2(r.b, x.a, y.b, x.b, y.a);
  r.a = re;
}
// (a + b i)^2 = (a + b)(a - b) + 2ab i       (fp_tower.hpp:539-550)
PSB_HD PSB_NOINL void fp2_sqr(Fp2& r, const Fp2& x) {
  Fp s, d, t;
  fp_add_nr(s, x.a, x.b);  // < 2p: fine as a Montgomery multiplicand (result still < 2p before the final subtraction)
  fp_sub(d, x.a, x.b);
  fp_add_nr(t, x.a, x.a);
  fp_mul(r.b, t, x.b);
  fp_mul(r.a, s, d);
}
#endif
// x^-1 = conj(x) / (a^2 + b^2)   (fp_tower.hpp:597-611)
PSB_HD PSB_NOINL void fp2_inv(Fp2& r, const Fp2& x) {
  Fp n;
  fp_dot2(n, x.a, x.a, x.b, x.b);
  fp_inv(n, n);
  fp_mul(r.a, x.a, n);
  fp_mul(n, x.b, n);
  fp_neg(r.b, n);
}

// ---- Fp6 ------------------------------------------------------------------------------------------
PSB_HD PSB_INL void fp6_add(Fp6& r, const Fp6& x, const Fp6& y) { fp2_add(r.a, x.a, y.a); fp2_add(r.b, x.b, y.b); fp2_add(r.c, x.c, y.c); }
PSB_HD PSB_INL void fp6_sub(Fp6& r, const Fp6& x, const Fp6& y) { fp2_sub(r.a, x.a, y.a); fp2_sub(r.b, x.b, y.b); fp2_sub(r.c, x.c, y.c); }
PSB_HD PSB_INL void fp6_neg(Fp6& r, const Fp6& x) { fp2_neg(r.a, x.a); fp2_neg(r.b, x.b); fp2_neg(r.c, x.c); }
PSB_HD PSB_INL void fp6_dbl(Fp6& r, const Fp6& x) { fp2_dbl(r.a, x.a); fp2_dbl(r.b, x.b); fp2_dbl(r.c, x.c); }
// (a + b v + c v^2) v = xi c + a v + b v^2
PSB_HD PSB_INL void fp6_mul_v(Fp6& r, const Fp6& x) {
  Fp2 t;
  fp2_mul_xi(t, x.c);
  r.c = x.b;
  r.b = x.a;
  r.a = t;
}

// z = x*y: Karatsuba over v, 6 Fp2 products (value of mcl Fp6::mul, fp_tower.hpp:978-1022)
PSB_HD PSB_NOINL void fp6_mul(Fp6& z, const Fp6& x, const Fp6& y) {
  Fp2 v0, v1, v2, s, t, c0, c1;
  fp2_mul(v0, x.a, y.a);
  fp2_mul(v1, x.b, y.b);
  fp2_mul(v2, x.c, y.c);
  // c0 = v0 + xi((b+c)(b'+c') - v1 - v2)
  fp2_add(s, x.b, x.c); fp2_add(t, y.b, y.c);
  fp2_mul(c0, s, t);
  fp2_sub(c0, c0, v1); fp2_sub(c0, c0, v2);
  fp2_mul_xi(c0, c0);
  fp2_add(c0, c0, v0);
  // c1 = (a+b)(a'+b') - v0 - v1 + xi v2
  fp2_add(s, x.a, x.b); fp2_add(t, y.a, y.b);
  fp2_mul(c1, s, t);
  fp2_sub(c1, c1, v0); fp2_sub(c1, c1, v1);
  fp2_mul_xi(s, v2);
  fp2_add(c1, c1, s);
  // c2 = (a+c)(a'+c') - v0 - v2 + v1
  fp2_add(s, x.a, x.c); fp2_add(t, y.a, y.c);
  fp2_mul(s, s, t);
  fp2_sub(s, s, v0); fp2_sub(s, s, v2);
  fp2_add(z.c, s, v1);
  z.a = c0;
  z.b = c1;
}

// x * (a0 + a1 v), 5 Fp2 products (sparse operand; cf. mcl Fp6mul_01, bn.hpp:1298-1320)
PSB_HD PSB_NOINL void fp6_mul_01(Fp6& z, const Fp6& x, const Fp2& a0, const Fp2& a1) {
  Fp2 v0, v1, s, t, r1;
  fp2_mul(v0, x.a, a0);
  fp2_mul(v1, x.b, a1);
  fp2_add(s, x.a, x.b); fp2_add(t, a0, a1);
  fp2_mul(r1, s, t);
  fp2_sub(r1, r1, v0); fp2_sub(r1, r1, v1);     // x0 a1 + x1 a0
  fp2_mul(s, x.c, a1);
  fp2_mul_xi(s, s);
  fp2_mul(t, x.c, a0);
  fp2_add(z.a, s, v0);                          // x0 a0 + xi x2 a1
  fp2_add(z.c, t, v1);                          // x1 a1 + x2 a0
  z.b = r1;
}
// x * (b1 v), 3 Fp2 products
PSB_HD PSB_NOINL void fp6_mul_1(Fp6& z, const Fp6& x, const Fp2& b1) {
  Fp2 t0, t1;
  fp2_mul(t0, x.c, b1);
  fp2_mul_xi(t0, t0);
  fp2_mul(t1, x.a, b1);
  fp2_mul(z.c, x.b, b1);
  z.b = t1;
  z.a = t0;
}

// x^-1  (fp_tower.hpp:917-948)
PSB_HD PSB_NOINL void fp6_inv(Fp6& r, const Fp6& x) {
  Fp2 t0, t1, t2, u, d;
  fp2_sqr(t0, x.a); fp2_mul(u, x.b, x.c); fp2_mul_xi(u, u); fp2_sub(t0, t0, u);      // a^2 - xi b c
  fp2_sqr(t1, x.c); fp2_mul_xi(t1, t1); fp2_mul(u, x.a, x.b); fp2_sub(t1, t1, u);    // xi c^2 - a b
  fp2_sqr(t2, x.b); fp2_mul(u, x.a, x.c); fp2_sub(t2, t2, u);                        // b^2 - a c
  fp2_mul(d, x.c, t1); fp2_mul(u, x.b, t2); fp2_add(d, d, u); fp2_mul_xi(d, d);
  fp2_mul(u, x.a, t0); fp2_add(d, d, u);
  fp2_inv(d, d);
  fp2_mul(r.a, t0, d); fp2_mul(r.b, t1, d); fp2_mul(r.c, t2, d);
}

// ---- Fp12 -------------------------------------------------------------------------------------
PSB_HD PSB_INL void fp12_set_one(Fp12& r) {
  fp2_set_one(r.a.a); fp2_set_zero(r.a.b); fp2_set_zero(r.a.c);
  fp2_set_zero(r.b.a); fp2_set_zero(r.b.b); fp2_set_zero(r.b.c);
}
PSB_HD PSB_INL void fp12_conj(Fp12& r, const Fp12& x) { r.a = x.a; fp6_neg(r.b, x.b); }  // unitaryInv
PSB_HD PSB_INL bool fp12_is_one(const Fp12& x) {
  const uint32_t* p = (const uint32_t*)&x;
  uint32_t o = 0;
  for (int i = 0; i < 12; i++) o |= p[i] ^ PSB_K(FP_ONE)[i];
  for (int i = 12; i < 144; i++) o |= p[i];
  return o == 0;
}

// z = x*y: Karatsuba over w (value of mcl Fp12::mul, fp_tower.hpp:1131-1160)
PSB_HD PSB_NOINL void fp12_mul(Fp12& z, const Fp12& x, const Fp12& y) {
  Fp6 t0, t1, s, t;
  fp6_mul(t0, x.a, y.a);
  fp6_mul(t1, x.b, y.b);
  fp6_add(s, x.a, x.b);
  fp6_add(t, y.a, y.b);
  fp6_mul(s, s, t);
  fp6_sub(s, s, t0);
  fp6_sub(z.b, s, t1);
  fp6_mul_v(t1, t1);
  fp6_add(z.a, t0, t1);
}

// z = x^2 (complex squaring: 2 Fp6 products; cf. fp_tower.hpp:1166-1178)
PSB_HD PSB_NOINL void fp12_sqr(Fp12& z, const Fp12& x) {
  Fp6 t0, t1, t2;
  fp6_add(t0, x.a, x.b);
  fp6_mul_v(t1, x.b);
  fp6_add(t1, t1, x.a);
  fp6_mul(t2, x.a, x.b);   // ab
  fp6_mul(t0, t0, t1);     // (a+b)(a+vb) = a^2 + v b^2 + ab + v ab
  fp6_sub(t0, t0, t2);
  fp6_mul_v(t1, t2);
  fp6_sub(z.a, t0, t1);
  fp6_dbl(z.b, t2);
}

// f *= (c0 + c2 w^2 + c3 w^3): the sparse line of an M-type twist (cf. mcl mul_041, bn.hpp:1436-1468)
// with A = c0 + c2 v, B = c3 v:  f.a' = fa A + v fb B,  f.b' = (fa+fb)(A+B) - fa A - fb B.   13 Fp2 products.
PSB_HD PSB_NOINL void fp12_mul_line(Fp12& f, const Fp2& c0, const Fp2& c2, const Fp2& c3) {
  Fp6 t0, t1, s;
  Fp2 c23;
  fp6_mul_01(t0, f.a, c0, c2);
  fp6_mul_1(t1, f.b, c3);
  fp6_add(s, f.a, f.b);
  fp2_add(c23, c2, c3);
  fp6_mul_01(s, s, c0, c23);
  fp6_sub(s, s, t0);
  fp6_sub(f.b, s, t1);
  fp6_mul_v(t1, t1);
  fp6_add(f.a, t0, t1);
}

// x^-1 = (a - b w)/(a^2 - v b^2)   (fp_tower.hpp:1183-1198)
PSB_HD PSB_NOINL void fp12_inv(Fp12& r, const Fp12& x) {
  Fp6 t0, t1;
  fp6_mul(t0, x.a, x.a);
  fp6_mul(t1, x.b, x.b);
  fp6_mul_v(t1, t1);
  fp6_sub(t0, t0, t1);
  fp6_inv(t0, t0);
  fp6_mul(r.a, x.a, t0);
  fp6_mul(t1, x.b, t0);
  fp6_neg(r.b, t1);
}

// Frobenius x -> x^(p^j), j = 1,2,3: coefficient of w^k -> conj^j(c_k) * gamma_{j,k}
// (fp_tower.hpp:1225-1265; constants = xi^(k(p^j-1)/6), built by tools/gen_constants.py)
PSB_HD PSB_INL Fp2* fp12_coeff(Fp12& x, int k) {  // k-th power of w -> slot
  Fp2* base = &x.a.a;
  return base + ((k & 1) ? 3 : 0) + (k >> 1);
}
PSB_HD PSB_NOINL void fp12_frobenius(Fp12& r, const Fp12& x, int j) {
  const uint32_t* tbl = (j == 1) ? PSB_K(FROB_G1) : (j == 2) ? PSB_K(FROB_G2) : PSB_K(FROB_G3);
  Fp12 in = x;
  for (int k = 0; k < 6; k++) {
    Fp2 c = *fp12_coeff(in, k);
    if (j & 1) fp_neg(c.b, c.b);
    if (k > 0) {
      Fp2 g;
      for (int i = 0; i < 12; i++) { g.a.v[i] = tbl[(k - 1) * 24 + i]; g.b.v[i] = tbl[(k - 1) * 24 + 12 + i]; }
      fp2_mul(c, c, g);
    }
    *fp12_coeff(r, k) = c;
  }
}

// Granger-Scott squaring in the cyclotomic subgroup (cf. mcl fasterSqr/sqrFp4, bn.hpp:1075-1144)
PSB_HD PSB_NOINL void fp4_sqr(Fp2& z0, Fp2& z1, const Fp2& x0, const Fp2& x1) {
  Fp2 t0, t1, s;
  fp2_sqr(t0, x0);
  fp2_sqr(t1, x1);
  fp2_add(s, x0, x1);
  fp2_sqr(s, s);
  fp2_sub(s, s, t0);
  fp2_sub(z1, s, t1);           // 2 x0 x1
  fp2_mul_xi(t1, t1);
  fp2_add(z0, t1, t0);          // x0^2 + xi x1^2
}
PSB_HD PSB_NOINL void fp12_cyclo_sqr(Fp12& y, const Fp12& x) {
  // slots: x0=a.a x4=a.b x3=a.c x2=b.a x1=b.b x5=b.c
  Fp2 t0, t1, t2, t3, u;
  const Fp2 x0 = x.a.a, x4 = x.a.b, x3 = x.a.c, x2 = x.b.a, x1 = x.b.b, x5 = x.b.c;
  fp4_sqr(t0, t1, x0, x1);
  fp2_sub(u, t0, x0); fp2_dbl(u, u); fp2_add(y.a.a, u, t0);     // y0 = 3 t0 - 2 x0
  fp2_add(u, t1, x1); fp2_dbl(u, u); fp2_add(y.b.b, u, t1);     // y1 = 3 t1 + 2 x1
  fp4_sqr(t0, t1, x2, x3);
  fp4_sqr(t2, t3, x4, x5);
  fp2_sub(u, t0, x4); fp2_dbl(u, u); fp2_add(y.a.b, u, t0);     // y4 = 3 t0 - 2 x4
  fp2_add(u, t1, x5); fp2_dbl(u, u); fp2_add(y.b.c, u, t1);     // y5 = 3 t1 + 2 x5
  fp2_mul_xi(t0, t3);
  fp2_add(u, t0, x2); fp2_dbl(u, u); fp2_add(y.b.a, u, t0);     // y2 = 3 xi t3 + 2 x2
  fp2_sub(u, t2, x3); fp2_dbl(u, u); fp2_add(y.a.c, u, t2);     // y3 = 3 t2 - 2 x3
}

}  // namespace psb
