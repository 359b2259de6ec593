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
#include "modinv.cuh"

namespace psb {

struct Fp2 { Fp a, b; };
struct Fp6 { Fp2 a, b, c; };
struct Fp12 { Fp6 a, b; };

// ---- Fp2 ------------------------------------------------------------------------------------------
// additive ops over `cnt` consecutive Fp components (2 = Fp2, 6 = Fp6, 12 = Fp12): ONE rolled loop each, so an
// Fp6 addition is one call and a few dozen instructions of code (instruction-cache budget, see the fused kernels below)
PSB_HD PSB_NOINL void fpn_add(Fp* r, const Fp* x, const Fp* y, int cnt) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < cnt; k++) fp_add(r[k], x[k], y[k]);
}
PSB_HD PSB_NOINL void fpn_sub(Fp* r, const Fp* x, const Fp* y, int cnt) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < cnt; k++) fp_sub(r[k], x[k], y[k]);
}
PSB_HD PSB_NOINL void fpn_neg(Fp* r, const Fp* x, int cnt) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < cnt; k++) fp_neg(r[k], x[k]);
}
PSB_HD PSB_NOINL void fpn_dbl(Fp* r, const Fp* x, int cnt) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < cnt; k++) fp_dbl(r[k], x[k]);
}
// r = a - b - c over cnt components
PSB_HD PSB_NOINL void fpn_sub2(Fp* r, const Fp* a, const Fp* b, const Fp* c, int cnt) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int k = 0; k < cnt; k++) {
    uint32_t x[PSB_NL], y[PSB_NL];
    fp_ld(x, a[k]); fp_ld(y, b[k]);
    mod_sub<FpT>(x, x, y);
    fp_ld(y, c[k]);
    mod_sub<FpT>(x, x, y);
    fp_st(r[k], x);
  }
}
PSB_HD PSB_INL void fp2_add(Fp2& r, const Fp2& x, const Fp2& y) { fpn_add(&r.a, &x.a, &y.a, 2); }
PSB_HD PSB_INL void fp2_sub(Fp2& r, const Fp2& x, const Fp2& y) { fpn_sub(&r.a, &x.a, &y.a, 2); }
PSB_HD PSB_INL void fp2_neg(Fp2& r, const Fp2& x) { fpn_neg(&r.a, &x.a, 2); }
PSB_HD PSB_INL void fp2_dbl(Fp2& r, const Fp2& x) { fpn_dbl(&r.a, &x.a, 2); }
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
// ---- fused Fp2 kernels ---------------------------------------------------------------------------------
// Two constraints shape these (both measured, profiles/r1d, r1e):
//  * the additive glue of the tower as separate calls (each a stack round trip) took ~30 % of the instructions but
//    ~58 % of the stall samples of the final exponentiation -> sums / differences are formed in registers next to
//    the multiplier, and multi-term combinations are single functions;
//  * fully unrolled fused functions (13-20 KB of code each) overflow the SM's instruction caches (50 % of the stall
//    samples became instruction-fetch) -> multipliers are ROLLED loops that keep the multiplicands in registers and
//    stream the multiplier limbs from memory (cios.cuh), and component-wise combinations loop over the components.
#ifdef __CUDA_ARCH__
#define PSB_ROLL _Pragma("unroll 1")
#else
#define PSB_ROLL
#endif
PSB_HD PSB_INL Fp* fp2_comp(Fp2& x, int k) { return &x.a + k; }
PSB_HD PSB_INL const Fp* fp2_comp(const Fp2& x, int k) { return &x.a + k; }

#ifdef __CUDA_ARCH__
// Two multiplier ENGINES carry every Fp2 product of the tower (instruction-cache budget: each engine is one fully
// unrolled register-resident body; the variants differ only in a short operand-preparation prologue selected by
// warp-uniform null-pointer tests):
//   engine A  (re, im) = (xa ya + xb (p - yb), xa yb + xb ya)   x = x1 [+ x2],  y = y1 [+ y2]      fp2_mul, fp2_mul_sum
//   engine B  (re, im) = (m1 n1, m2 n2)                         squares (x = x1 [+ x2]) and Fp2 * Fp  fp2_sqr, fp2_sqr_sum, fp2_mul_fp
#ifndef PSB_LAZY_Y
#define PSB_LAZY_Y 1
#endif
#if PSB_LAZY_Y && !PSB_IS_BN
// unreduced sums on the MULTIPLIER side need (8 p^2 + R p) / R < 2p, i.e. 8 p < R = 2^(32 N): three spare bits in the limbs
// (tests/test_cios_model.py::test_fused_dot2_with_unreduced_operands runs the generated rows at that bound)
static_assert(PSB_FP_BITS + 3 <= 32 * PSB_NL, "PSB_LAZY_Y needs 8p < R: keep the multiplier-side sums canonical on this curve");
#endif
// Engine A comes in two forms (PSB_ENGINE_A_LAZY):
//   0 (default)  the fused two-product form: one reduction per row inside the multiplier, 4 N^2 + 2 (N^2 + N) wide MACs
//                (888 for N = 12), one multiplier body executed twice (11.5 KB of code).
//   1            Karatsuba over i with LAZY REDUCTION (the shape of mcl's Fp2Dbl::mulPre + FpDbl::mod, fp_tower.hpp:713-757):
//                three wide products T0 = xa ya, T1 = xb yb, T2 = (xa + xb)(ya + yb), then re = T0 - T1 (+ pR on
//                borrow), im = T2 - T0 - T1, and two stand-alone Montgomery reductions: 3 N^2 + 2 (N^2 + N) wide MACs (744).
//                Bit-identical results (74 GPU parity tests green) but SLOWER on B200 (r2a, profiles/r2a_ab_lazy_engine.txt):
//                Miller loop 381.6 -> 422.6 ms per 2^20 lanes.  The 144 saved wide MACs cost ~190 extra additions and 5.5 KB
//                more straight-line code (17 KB): instruction-fetch stalls 0.54 -> 2.40 cycles per issue, multiply-pipe
//                active cycles 81 % -> 67 %.  Rolling the three products back into one body costs as many SEL / MOV as
//                it saves (15.8 KB).  Kept as an A/B build only; what mcl gains on x86 (cheap adds, deep caches) does
//                not transfer.  Bounds of this form are checked on the interpreted rows in tests/test_cios_model.py.
#ifndef PSB_ENGINE_A_LAZY
#define PSB_ENGINE_A_LAZY 0
#endif
__device__ PSB_NOINL void fp2_engine_a(Fp2& r, const Fp2* x1, const Fp2* x2, const Fp2* y1, const Fp2* y2) {
  Fp xa, xb, ya, yb, u;
  fp_get(xa, x1->a); fp_get(xb, x1->b); fp_get(ya, y1->a); fp_get(yb, y1->b);
  if (x2) {
    fp_get(u, x2->a); fp_addnr_rr(xa, xa, u);     // multiplicand side: unreduced (< 2p)
    fp_get(u, x2->b); fp_addnr_rr(xb, xb, u);
#if PSB_IS_BN || !PSB_LAZY_Y
    fp_get(u, y2->a); fp_add_rr(ya, ya, u);       // multiplier side: canonical
    fp_get(u, y2->b); fp_add_rr(yb, yb, u);
#else
    // 381-bit p in a 384-bit radix: the multiplier side may stay unreduced too.  Fused form: the accumulated pair of
    // products is <= 8 p^2 and (8 p^2 + R p) / R = p (1 + 8 p / R) < 1.82 p, so the ONE conditional subtraction of the
    // reduction still lands in [0, p) (BN254: 8 p / R > 1, stays canonical)
    fp_get(u, y2->a); fp_addnr_rr(ya, ya, u);
    fp_get(u, y2->b); fp_addnr_rr(yb, yb, u);
#endif
  }
#if PSB_ENGINE_A_LAZY
  uint32_t T0[2 * PSB_NL], T1[2 * PSB_NL], T2[2 * PSB_NL];
  Fp sa, sb;
  add_n<PSB_NL>(sa.v, xa.v, xb.v);                // Karatsuba sums, < 4p < R
  add_n<PSB_NL>(sb.v, ya.v, yb.v);
  cios::mulpre_rr(T0, xa.v, ya.v);
  cios::chain_after(yb.v[0]);
  cios::mulpre_rr(T1, xb.v, yb.v);
  cios::chain_after(sb.v[0]);
  cios::mulpre_rr(T2, sa.v, sb.v);
  sub_n<2 * PSB_NL>(T2, T2, T0);
  sub_n<2 * PSB_NL>(T2, T2, T1);                  // im = xa yb + xb ya
  const uint32_t borrow = sub_n<2 * PSB_NL>(T0, T0, T1);   // re = xa ya - xb yb, + pR when negative
  add_mod_masked_n<FpT>(T0 + PSB_NL, T0 + PSB_NL, 0u - borrow);
  // one reduction body, executed twice (instruction-cache budget, and the two reductions stay one after the other)
  PSB_ROLL
  for (int k = 0; k < 2; k++) {
    uint32_t T[2 * PSB_NL];
    PSB_UNROLL
    for (int i = 0; i < 2 * PSB_NL; i++) T[i] = k == 0 ? T0[i] : T2[i];
    cios::redc_rr(u.v, T);
    fp_put(*fp2_comp(r, k), u);
  }
#else
  Fp nb, re;
#if PSB_IS_BN || !PSB_LAZY_Y
  fp_pminus_rr(nb, yb);
#else
  if (x2) fp_2pminus_rr(nb, yb); else fp_pminus_rr(nb, yb);
#endif
  // one multiplier body, executed twice (real, imaginary): half the code of two unrolled copies -- a 20 KB
  // straight-line engine still stalled ~25 % of its samples on instruction fetch (r1g)
  PSB_ROLL
  for (int k = 0; k < 2; k++) {
    Fp q, t;
    if (k == 0) { q = ya; t = nb; } else { q = yb; t = ya; }
    fp_dot2_rr(re, xa, q, xb, t);
    fp_put(*fp2_comp(r, k), re);
  }
#endif
}
__device__ PSB_NOINL void fp2_engine_b(Fp2& r, const Fp2* x1, const Fp2* x2, const Fp* k) {
  Fp m1, n1, m2, n2, u, re;
  fp_get(m1, x1->a); fp_get(m2, x1->b);
  if (k) {                                         // (a + b i) k
    fp_get(n1, *k);
    n2 = n1;
  } else {                                         // (a + b i)^2 = (a + b)(a - b) + (2a) b i
    if (x2) {
      fp_get(u, x2->a); fp_add_rr(m1, m1, u);
      fp_get(u, x2->b); fp_add_rr(m2, m2, u);
    }
    n2 = m2;                                       // b
    fp_sub_rr(n1, m1, m2);                         // a - b
    fp_addnr_rr(m2, m1, m1);                       // 2a   (< 2p: multiplicand only)
    fp_addnr_rr(m1, m1, n2);                       // a + b
  }
  PSB_ROLL
  for (int k = 0; k < 2; k++) {
    Fp m, n;
    if (k == 0) { m = m1; n = n1; } else { m = m2; n = n2; }
    fp_mul_rr(re, m, n);
    fp_put(*fp2_comp(r, k), re);
  }
}
__device__ PSB_INL void fp2_mul(Fp2& r, const Fp2& x, const Fp2& y) { fp2_engine_a(r, &x, nullptr, &y, nullptr); }
__device__ PSB_INL void fp2_mul_sum(Fp2& r, const Fp2& x1, const Fp2& x2, const Fp2& y1, const Fp2& y2) { fp2_engine_a(r, &x1, &x2, &y1, &y2); }
__device__ PSB_INL void fp2_sqr(Fp2& r, const Fp2& x) { fp2_engine_b(r, &x, nullptr, nullptr); }
__device__ PSB_INL void fp2_sqr_sum(Fp2& r, const Fp2& x1, const Fp2& x2) { fp2_engine_b(r, &x1, &x2, nullptr); }
__device__ PSB_INL void fp2_mul_fp(Fp2& r, const Fp2& x, const Fp& k) { fp2_engine_b(r, &x, nullptr, &k); }
#else
// host build (tests/hostsim): same values through the generic code
PSB_HD inline void fp2_mul(Fp2& r, const Fp2& x, const Fp2& y) {
  Fp nb, re;
  fp_neg(nb, y.b);
  fp_dot2(re, x.a, y.a, x.b, nb);
  fp_dot2(r.b, x.a, y.b, x.b, y.a);
  r.a = re;
}
PSB_HD inline void fp2_sqr(Fp2& r, const Fp2& x) {
  Fp s, d, t;
  fp_add_nr(s, x.a, x.b);
  fp_sub(d, x.a, x.b);
  fp_add_nr(t, x.a, x.a);
  fp_mul(r.b, t, x.b);
  fp_mul(r.a, s, d);
}
PSB_HD inline void fp2_mul_sum(Fp2& r, const Fp2& x1, const Fp2& x2, const Fp2& y1, const Fp2& y2) {
  Fp2 s, t;
  fp2_add(s, x1, x2); fp2_add(t, y1, y2);
  fp2_mul(r, s, t);
}
PSB_HD inline void fp2_sqr_sum(Fp2& r, const Fp2& x1, const Fp2& x2) {
  Fp2 s;
  fp2_add(s, x1, x2);
  fp2_sqr(r, s);
}
PSB_HD inline void fp2_mul_fp(Fp2& r, const Fp2& x, const Fp& k) { fp_mul(r.a, x.a, k); fp_mul(r.b, x.b, k); }
#endif

// component-wise combinations: one call, a rolled loop over the two components
PSB_HD PSB_INL void fp2_sub2(Fp2& r, const Fp2& a, const Fp2& b, const Fp2& c) { fpn_sub2(&r.a, &a.a, &b.a, &c.a, 2); }   // a - b - c
// r = a - b - c + d
PSB_HD PSB_NOINL void fp2_sub2_add(Fp2& r, const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) {
  PSB_ROLL
  for (int k = 0; k < 2; k++) {
    Fp x, y;
    fp_get(x, *fp2_comp(a, k)); fp_get(y, *fp2_comp(b, k));
    fp_sub_rr(x, x, y);
    fp_get(y, *fp2_comp(c, k));
    fp_sub_rr(x, x, y);
    fp_get(y, *fp2_comp(d, k));
    fp_add_rr(x, x, y);
    fp_put(*fp2_comp(r, k), x);
  }
}
// r = 3a - 2b (minus = true) or 3a + 2b: the output combination of the Granger-Scott squaring
PSB_HD PSB_NOINL void fp2_3a2b(Fp2& r, const Fp2& a, const Fp2& b, bool minus) {
  PSB_ROLL
  for (int k = 0; k < 2; k++) {
    Fp x, y, u;
    fp_get(x, *fp2_comp(a, k)); fp_get(y, *fp2_comp(b, k));
    if (minus) fp_sub_rr(u, x, y); else fp_add_rr(u, x, y);
    fp_dbl_rr(u, u);
    fp_add_rr(u, u, x);
    fp_put(*fp2_comp(r, k), u);
  }
}
// r = b + xi a;   xi a = (a.a - a.b) + (a.a + a.b) i     (one variant only: instruction-cache budget)
PSB_HD PSB_NOINL void fp2_xi_add(Fp2& r, const Fp2& a, const Fp2& b) {
  Fp x, y, w0, w1;
  fp_get(x, a.a); fp_get(y, a.b);
  fp_sub_rr(w0, x, y);
  fp_add_rr(w1, x, y);
  fp_get(x, b.a); fp_get(y, b.b);
  fp_add_rr(x, x, w0); fp_add_rr(y, y, w1);
  fp_put(r.a, x); fp_put(r.b, y);
}

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
PSB_HD PSB_INL void fp6_add(Fp6& r, const Fp6& x, const Fp6& y) { fpn_add(&r.a.a, &x.a.a, &y.a.a, 6); }
PSB_HD PSB_INL void fp6_sub(Fp6& r, const Fp6& x, const Fp6& y) { fpn_sub(&r.a.a, &x.a.a, &y.a.a, 6); }
PSB_HD PSB_INL void fp6_neg(Fp6& r, const Fp6& x) { fpn_neg(&r.a.a, &x.a.a, 6); }
PSB_HD PSB_INL void fp6_dbl(Fp6& r, const Fp6& x) { fpn_dbl(&r.a.a, &x.a.a, 6); }
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
  Fp2 v0, v1, v2, m0, m1, m2;
  fp2_mul(v0, x.a, y.a);
  fp2_mul(v1, x.b, y.b);
  fp2_mul(v2, x.c, y.c);
  fp2_mul_sum(m0, x.b, x.c, y.b, y.c);
  fp2_mul_sum(m1, x.a, x.b, y.a, y.b);
  fp2_mul_sum(m2, x.a, x.c, y.a, y.c);
  fp2_sub2(m0, m0, v1, v2);
  fp2_xi_add(z.a, m0, v0);                // v0 + xi((b+c)(b'+c') - v1 - v2)
  fp2_sub2(m1, m1, v0, v1);
  fp2_xi_add(z.b, v2, m1);                // (a+b)(a'+b') - v0 - v1 + xi v2
  fp2_sub2_add(z.c, m2, v0, v2, v1);      // (a+c)(a'+c') - v0 - v2 + v1
}

// x * (a0 + a1 v), 5 Fp2 products (sparse operand; cf. mcl Fp6mul_01, bn.hpp:1298-1320)
PSB_HD PSB_NOINL void fp6_mul_01(Fp6& z, const Fp6& x, const Fp2& a0, const Fp2& a1) {
  Fp2 v0, v1, m, s, t;
  fp2_mul(v0, x.a, a0);
  fp2_mul(v1, x.b, a1);
  fp2_mul_sum(m, x.a, x.b, a0, a1);
  fp2_mul(s, x.c, a1);
  fp2_mul(t, x.c, a0);
  fp2_xi_add(z.a, s, v0);                       // x0 a0 + xi x2 a1
  fp2_add(z.c, t, v1);                          // x1 a1 + x2 a0
  fp2_sub2(z.b, m, v0, v1);                     // x0 a1 + x1 a0
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
  for (int i = 0; i < PSB_NL; i++) o |= p[i] ^ PSB_K(FP_ONE)[i];
  for (int i = PSB_NL; i < 12 * PSB_NL; i++) o |= p[i];
  return o == 0;
}

// r = a + v b  (Fp6)
PSB_HD PSB_INL void fp6_add_mulv(Fp6& r, const Fp6& a, const Fp6& b) {
  Fp2 t;
  fp2_xi_add(t, b.c, a.a);
  fp2_add(r.c, a.c, b.b);
  fp2_add(r.b, a.b, b.a);
  r.a = t;
}
// r = a - b - c  (Fp6)
PSB_HD PSB_INL void fp6_sub2(Fp6& r, const Fp6& a, const Fp6& b, const Fp6& c) { fpn_sub2(&r.a.a, &a.a.a, &b.a.a, &c.a.a, 6); }

// z = x*y: Karatsuba over w (value of mcl Fp12::mul, fp_tower.hpp:1131-1160)
PSB_HD PSB_NOINL void fp12_mul(Fp12& z, const Fp12& x, const Fp12& y) {
  Fp6 t0, t1, s, t;
  fp6_mul(t0, x.a, y.a);
  fp6_mul(t1, x.b, y.b);
  fp6_add(s, x.a, x.b);
  fp6_add(t, y.a, y.b);
  fp6_mul(s, s, t);
  fp6_sub2(z.b, s, t0, t1);
  fp6_add_mulv(z.a, t0, t1);
}

// z = x^2 (complex squaring: 2 Fp6 products; cf. fp_tower.hpp:1166-1178)
PSB_HD PSB_NOINL void fp12_sqr(Fp12& z, const Fp12& x) {
  Fp6 t0, t1, t2;
  fp6_add(t0, x.a, x.b);
  fp6_add_mulv(t1, x.a, x.b);
  fp6_mul(t2, x.a, x.b);   // ab
  fp6_mul(t0, t0, t1);     // (a+b)(a+vb) = a^2 + v b^2 + ab + v ab
  // z.a = t0 - t2 - v t2
  fp2_mul_xi(t1.a, t2.c);                 // (t1 is dead here)
  fp2_sub2(z.a.a, t0.a, t2.a, t1.a);
  fp2_sub2(z.a.b, t0.b, t2.b, t2.a);
  fp2_sub2(z.a.c, t0.c, t2.c, t2.b);
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
  fp6_sub2(f.b, s, t0, t1);
  fp6_add_mulv(f.a, t0, t1);
}

// f *= (1 + c2 w^2 + c3 w^3): a line whose constant term was scaled to 1 (fixed-argument lines, pairing.cuh).
// With A = 1 + c2 v, B = c3 v:  fa A = fa + fa (c2 v),  (fa + fb)(A + B) = fa + fb + (fa + fb)((c2 + c3) v)
//   f.a' = fa + fa (c2 v) + v fb (c3 v),   f.b' = fb + (fa + fb)((c2 + c3) v) - fa (c2 v) - fb (c3 v).    9 Fp2 products.
PSB_HD PSB_NOINL void fp12_mul_line_1(Fp12& f, const Fp2& c2, const Fp2& c3) {
  Fp6 t0, t1, s;
  Fp2 c23;
  fp6_mul_1(t0, f.a, c2);
  fp6_mul_1(t1, f.b, c3);
  fp6_add(s, f.a, f.b);
  fp2_add(c23, c2, c3);
  fp6_mul_1(s, s, c23);
  fp6_sub2(s, s, t0, t1);
  fp6_add(f.b, f.b, s);
  fp6_add_mulv(t0, t0, t1);
  fp6_add(f.a, f.a, t0);
}

// x * a0 with a0 in Fp2, 3 Fp2 products
PSB_HD PSB_NOINL void fp6_mul_fp2(Fp6& z, const Fp6& x, const Fp2& a0) {
  fp2_mul(z.a, x.a, a0);
  fp2_mul(z.b, x.b, a0);
  fp2_mul(z.c, x.c, a0);
}
// f *= (c0 + c1 w + c3 w^3): the sparse line of a D-type twist (cf. mcl mul_403, bn.hpp:1335-1419)
// with A = c0, B = c1 + c3 v:  f.a' = fa A + v fb B,  f.b' = (fa+fb)(A+B) - fa A - fb B.   13 Fp2 products.
PSB_HD PSB_NOINL void fp12_mul_line_d(Fp12& f, const Fp2& c0, const Fp2& c1, const Fp2& c3) {
  Fp6 t0, t1, s;
  Fp2 c01;
  fp6_mul_fp2(t0, f.a, c0);
  fp6_mul_01(t1, f.b, c1, c3);
  fp6_add(s, f.a, f.b);
  fp2_add(c01, c0, c1);
  fp6_mul_01(s, s, c01, c3);
  fp6_sub2(f.b, s, t0, t1);
  fp6_add_mulv(f.a, t0, t1);
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
  PSB_ROLL
  for (int k = 0; k < 6; k++) {
    Fp2 c = *fp12_coeff(in, k);
    if (j & 1) fp_neg(c.b, c.b);
    if (k > 0) {
      Fp2 g;
      for (int i = 0; i < PSB_NL; i++) { g.a.v[i] = tbl[(k - 1) * 2 * PSB_NL + i]; g.b.v[i] = tbl[(k - 1) * 2 * PSB_NL + PSB_NL + i]; }
      fp2_mul(c, c, g);
    }
    *fp12_coeff(r, k) = c;
  }
}

// Granger-Scott squaring in the cyclotomic subgroup (cf. mcl fasterSqr/sqrFp4, bn.hpp:1075-1144).
// With s = w^3 (s^2 = xi) an element is A + B w + C w^2 over Fp4 = Fp2[s]:  A = x0 + x1 s, B = x2 + x3 s, C = x4 + x5 s,
// and x^2 = (3 A^2 - 2 conj A) + (3 s C^2 + 2 conj B) w + (3 B^2 - 2 conj C) w^2.
// Per Fp4 square (xe + xo s)^2 = z0 + z1 s:  t0 = xe^2, t1 = xo^2, z0 = t0 + xi t1, z1 = (xe + xo)^2 - t0 - t1;
// outputs 3 z -+ 2 x are formed by the fused fp2_3a2b kernel (no intermediate z is stored).
//
// The B and C coordinates of the square depend on B and C only: that is Karabina's compressed squaring ("Squaring in
// cyclotomic subgroups", Math. Comp. 2013) -- 6 Fp2 squarings instead of 9 -- used by pow_z (pairing.cuh) for the
// long runs of squarings; (x0, x1) are recovered afterwards by cyclo_decompress_*.
struct CycC { Fp2 g2, g3, g4, g5; };   // (x2, x3, x4, x5) = slots b.a, a.c, a.b, b.c of an Fp12

// (y2, y3, y4, y5) = compressed square of (x2, x3, x4, x5); outputs may alias the inputs slot by slot
PSB_HD PSB_NOINL void cyclo_csqr4(Fp2& y2, Fp2& y3, Fp2& y4, Fp2& y5, const Fp2& x2, const Fp2& x3, const Fp2& x4, const Fp2& x5) {
  Fp2 t0, t1, s, v2, v3;
  // C^2 -> y2 = 3 xi z1 + 2 x2,  y3 = 3 z0 - 2 x3
  fp2_sqr(t0, x4); fp2_sqr(t1, x5); fp2_sqr_sum(s, x4, x5);
  fp2_sub2(s, s, t0, t1);
  fp2_mul_xi(s, s);
  fp2_xi_add(t0, t1, t0);
  fp2_3a2b(v2, s, x2, false);
  fp2_3a2b(v3, t0, x3, true);
  // B^2 -> y4 = 3 z0 - 2 x4,  y5 = 3 z1 + 2 x5
  fp2_sqr(t0, x2); fp2_sqr(t1, x3); fp2_sqr_sum(s, x2, x3);
  fp2_sub2(s, s, t0, t1);
  fp2_xi_add(t0, t1, t0);
  fp2_3a2b(y4, t0, x4, true);
  fp2_3a2b(y5, s, x5, false);
  y2 = v2; y3 = v3;
}
PSB_HD PSB_INL void cyclo_csqr(CycC& y, const CycC& x) { cyclo_csqr4(y.g2, y.g3, y.g4, y.g5, x.g2, x.g3, x.g4, x.g5); }

PSB_HD PSB_NOINL void fp12_cyclo_sqr(Fp12& y, const Fp12& x) {
  // slots: x0=a.a x4=a.b x3=a.c x2=b.a x1=b.b x5=b.c
  Fp2 t0, t1, s, u0, u1;
  // A^2 -> y0 = 3 z0 - 2 x0,  y1 = 3 z1 + 2 x1
  fp2_sqr(t0, x.a.a); fp2_sqr(t1, x.b.b); fp2_sqr_sum(s, x.a.a, x.b.b);
  fp2_sub2(s, s, t0, t1);                       // z1 = 2 x0 x1
  fp2_xi_add(t0, t1, t0);                       // z0 = x0^2 + xi x1^2
  fp2_3a2b(u0, t0, x.a.a, true);
  fp2_3a2b(u1, s, x.b.b, false);
  cyclo_csqr4(y.b.a, y.a.c, y.a.b, y.b.c, x.b.a, x.a.c, x.a.b, x.b.c);
  y.a.a = u0; y.b.b = u1;
}

// Decompression (Karabina, Thm 3.2 in the basis above), split around the one inversion so that several elements share it:
//   g1 = (xi g5^2 + 3 g4^2 - 2 g3) / (4 g2),   g0 = (2 g1^2 + g2 g5 - 3 g3 g4) xi + 1        (g2 != 0)
PSB_HD PSB_NOINL void cyclo_decompress_num(Fp2& num, const CycC& c) {
  Fp2 t0, t1;
  fp2_sqr(t0, c.g5);
  fp2_sqr(t1, c.g4);
  fp2_xi_add(t0, t0, t1);            // xi g5^2 + g4^2
  fp2_dbl(t1, t1);                   // 2 g4^2
  fp2_add(t0, t0, t1);
  fp2_dbl(t1, c.g3);
  fp2_sub(num, t0, t1);
}
// x = the full element with (g0, g1) rebuilt from g1 = num / (4 g2)
PSB_HD PSB_NOINL void cyclo_decompress_fill(Fp12& x, const CycC& c, const Fp2& g1) {
  Fp2 t0, t1, t2;
  fp2_sqr(t0, g1); fp2_dbl(t0, t0);             // 2 g1^2
  fp2_mul(t1, c.g2, c.g5);
  fp2_mul(t2, c.g3, c.g4);
  fp2_add(t0, t0, t1);
  fp2_dbl(t1, t2); fp2_add(t1, t1, t2);         // 3 g3 g4
  fp2_sub(t0, t0, t1);
  fp2_mul_xi(t0, t0);
  Fp2 one;
  fp2_set_one(one);
  fp2_add(x.a.a, t0, one);
  x.b.b = g1;
  x.b.a = c.g2; x.a.c = c.g3; x.a.b = c.g4; x.b.c = c.g5;
}

}  // namespace psb
