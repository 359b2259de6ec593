// fp.cuh -- prime-field arithmetic for BLS12-381 on 32-bit limbs (Fp: 12 limbs, Fr: 8 limbs).
//
// Replaces, for the batched path, mcl's field back-ends (reference: third-parties/mcl/src/fp.cpp,
// src/low_func.hpp:511-652 Montgomery, src/fp_generator.hpp Xbyak JIT, include/mcl/fp.hpp,
// include/mcl/fp_tower.hpp:13-178 FpDbl).  Representation is bit-identical to mcl's: little-endian
// limbs in Montgomery form with R = 2^(32 N), every Fp value canonical in [0, p).
//
// Device code: mad.lo.cc / madc.hi.cc / addc carry chains held in registers (product-scanning,
// one 3-word column accumulator), no tensor cores (nothing here is a dense contraction).
// The same source also compiles for the host (portable 64-bit arithmetic) -- that build exists
// ONLY for the CPU test-suite's "hostsim" library (tests/hostsim), never inside libpsb.so.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PSB_HD __host__ __device__
#define PSB_INL __forceinline__
#define PSB_NOINL __noinline__
#define PSB_CONST_ARRAY(name, n, ...)                                 \
  static __device__ __constant__ uint32_t name##_D[n] = {__VA_ARGS__}; \
  static const uint32_t name##_H[n] = {__VA_ARGS__};
#else
#define PSB_HD
#define PSB_INL inline __attribute__((always_inline))
#define PSB_NOINL __attribute__((noinline))
#define PSB_CONST_ARRAY(name, n, ...) static const uint32_t name##_H[n] = {__VA_ARGS__};
#endif
#ifdef __CUDA_ARCH__
#define PSB_K(name) name##_D
#define PSB_UNROLL _Pragma("unroll")
#else
#define PSB_K(name) name##_H
#define PSB_UNROLL
#endif

#ifdef PSB_BUILD_BN254
#include "constants_bn254.cuh"
#else
#include "constants.cuh"
#endif

namespace psb {

// ------------------------------------------------------------------------------------------------
// modulus traits: limbs as compile-time immediates (switch folds after unrolling)
// ------------------------------------------------------------------------------------------------
struct FpT {
  static constexpr int N = PSB_NL;
  static constexpr uint32_t N0 = PSB_FP_N0;
  PSB_HD static PSB_INL uint32_t p(int i) {
    switch (i) {
      case 0: return PSB_P0; case 1: return PSB_P1; case 2: return PSB_P2; case 3: return PSB_P3;
#if PSB_NL == 12
      case 4: return PSB_P4; case 5: return PSB_P5; case 6: return PSB_P6; case 7: return PSB_P7;
      case 8: return PSB_P8; case 9: return PSB_P9; case 10: return PSB_P10; default: return PSB_P11;
#else
      case 4: return PSB_P4; case 5: return PSB_P5; case 6: return PSB_P6; default: return PSB_P7;
#endif
    }
  }
};
struct FrT {
  static constexpr int N = 8;
  static constexpr uint32_t N0 = PSB_FR_N0;
  PSB_HD static PSB_INL uint32_t p(int i) {
    switch (i) {
      case 0: return PSB_R0; case 1: return PSB_R1; case 2: return PSB_R2; case 3: return PSB_R3;
      case 4: return PSB_R4; case 5: return PSB_R5; case 6: return PSB_R6; default: return PSB_R7;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// PTX carry-chain wrappers (device only).  Every statement that reads or writes CC.CF is
// `asm volatile`, so nvcc keeps their relative order; ptxas sees the true flag dataflow.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
namespace ptx {
__device__ PSB_INL uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ PSB_INL uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ PSB_INL uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ PSB_INL uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ PSB_INL uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ PSB_INL uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ PSB_INL uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ PSB_INL uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ PSB_INL uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ PSB_INL uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// one product accumulated into a 3-word column accumulator
__device__ PSB_INL void mac3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
  c0 = mad_lo_cc(a, b, c0);
  c1 = madc_hi_cc(a, b, c1);
  c2 = addc(c2, 0);
}
// 2*a*b accumulated (a*b added twice would double the multiplies; instead add the 64-bit product
// into a 3-word accumulator twice is avoided by the caller doubling the cross sum)
__device__ PSB_INL void acc3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t x) {
  c0 = add_cc(c0, x);
  c1 = addc_cc(c1, 0);
  c2 = addc(c2, 0);
}
}  // namespace ptx
#endif

// ------------------------------------------------------------------------------------------------
// limb-vector primitives (N compile-time)
// ------------------------------------------------------------------------------------------------
template <int N>
PSB_HD PSB_INL uint32_t add_n(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  r[0] = ptx::add_cc(a[0], b[0]);
  PSB_UNROLL
  for (int i = 1; i < N; i++) r[i] = ptx::addc_cc(a[i], b[i]);
  return ptx::addc(0, 0);
#else
  uint64_t c = 0;
  for (int i = 0; i < N; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)c;
#endif
}

// returns borrow (1 if a < b)
template <int N>
PSB_HD PSB_INL uint32_t sub_n(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  r[0] = ptx::sub_cc(a[0], b[0]);
  PSB_UNROLL
  for (int i = 1; i < N; i++) r[i] = ptx::subc_cc(a[i], b[i]);
  return ptx::subc(0, 0) & 1u;  // 0 - 0 - borrow = 0xffffffff when borrow
#else
  int64_t c = 0;
  for (int i = 0; i < N; i++) { c += (int64_t)a[i] - b[i]; r[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)(c & 1);
#endif
}

// r = a - modulus, returns borrow
template <class T>
PSB_HD PSB_INL uint32_t sub_mod_n(uint32_t* r, const uint32_t* a) {
  constexpr int N = T::N;
#ifdef __CUDA_ARCH__
  r[0] = ptx::sub_cc(a[0], T::p(0));
  PSB_UNROLL
  for (int i = 1; i < N; i++) r[i] = ptx::subc_cc(a[i], T::p(i));
  return ptx::subc(0, 0) & 1u;
#else
  int64_t c = 0;
  for (int i = 0; i < N; i++) { c += (int64_t)a[i] - T::p(i); r[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)(c & 1);
#endif
}

// r = a + (modulus & mask)   (mask = 0 or 0xffffffff), carry discarded
template <class T>
PSB_HD PSB_INL void add_mod_masked_n(uint32_t* r, const uint32_t* a, uint32_t mask) {
  constexpr int N = T::N;
#ifdef __CUDA_ARCH__
  r[0] = ptx::add_cc(a[0], T::p(0) & mask);
  PSB_UNROLL
  for (int i = 1; i < N - 1; i++) r[i] = ptx::addc_cc(a[i], T::p(i) & mask);
  r[N - 1] = ptx::addc(a[N - 1], T::p(N - 1) & mask);
#else
  uint64_t c = 0;
  for (int i = 0; i < N; i++) { c += (uint64_t)a[i] + (T::p(i) & mask); r[i] = (uint32_t)c; c >>= 32; }
#endif
}

// canonicalise x in [0, 2m) to [0, m)
template <class T>
PSB_HD PSB_INL void cond_sub_mod(uint32_t* x) {
  constexpr int N = T::N;
  uint32_t t[N];
  uint32_t borrow = sub_mod_n<T>(t, x);
  PSB_UNROLL
  for (int i = 0; i < N; i++) x[i] = borrow ? x[i] : t[i];
}

template <class T>
PSB_HD PSB_INL void mod_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = T::N;
  uint32_t s[N];
  add_n<N>(s, a, b);  // 2m < 2^(32N): no carry out
  cond_sub_mod<T>(s);
  PSB_UNROLL
  for (int i = 0; i < N; i++) r[i] = s[i];
}

template <class T>
PSB_HD PSB_INL void mod_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = T::N;
  uint32_t s[N];
  uint32_t borrow = sub_n<N>(s, a, b);
  add_mod_masked_n<T>(r, s, 0u - borrow);
}

template <class T>
PSB_HD PSB_INL bool is_zero_n(const uint32_t* a) {
  uint32_t o = 0;
  PSB_UNROLL
  for (int i = 0; i < T::N; i++) o |= a[i];
  return o == 0;
}

template <class T>
PSB_HD PSB_INL void mod_neg(uint32_t* r, const uint32_t* a) {
  constexpr int N = T::N;
  uint32_t z[N];
  PSB_UNROLL
  for (int i = 0; i < N; i++) z[i] = 0;
  // 0 - a mod m; mod_sub maps a = 0 to 0 (borrow = 0)
  mod_sub<T>(r, z, a);
}

// full 2N-limb product, product scanning
template <int N>
PSB_HD PSB_INL void mulw_n(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  PSB_UNROLL
  for (int k = 0; k < 2 * N - 1; k++) {
    PSB_UNROLL
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j >= 0 && j < N) ptx::mac3(c0, c1, c2, a[i], b[j]);
    }
    r[k] = c0; c0 = c1; c1 = c2; c2 = 0;
  }
  r[2 * N - 1] = c0;
#else
  uint64_t lo = 0; uint64_t hi = 0;  // 96-bit accumulator: lo (64) + hi (carry count << 64)
  for (int k = 0; k < 2 * N - 1; k++) {
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j >= 0 && j < N) {
        uint64_t pr = (uint64_t)a[i] * b[j];
        uint64_t nl = lo + pr;
        hi += nl < lo;
        lo = nl;
      }
    }
    r[k] = (uint32_t)lo;
    lo = (lo >> 32) | (hi << 32);
    hi = 0;
  }
  r[2 * N - 1] = (uint32_t)lo;
#endif
}

// full 2N-limb square: cross products once, doubled, plus the diagonal
template <int N>
PSB_HD PSB_INL void sqrw_n(uint32_t* r, const uint32_t* a) {
#ifdef __CUDA_ARCH__
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  r[0] = 0;
  PSB_UNROLL
  for (int k = 1; k < 2 * N - 1; k++) {
    PSB_UNROLL
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j > i && j < N) ptx::mac3(c0, c1, c2, a[i], a[j]);
    }
    r[k] = c0; c0 = c1; c1 = c2; c2 = 0;
  }
  r[2 * N - 1] = c0;
  // r = 2*r  (top bit of the cross sum is clear: cross sum < 2^(64N-1))
  r[0] = ptx::add_cc(r[0], r[0]);
  PSB_UNROLL
  for (int i = 1; i < 2 * N - 1; i++) r[i] = ptx::addc_cc(r[i], r[i]);
  r[2 * N - 1] = ptx::addc(r[2 * N - 1], r[2 * N - 1]);
  // r += sum a_i^2 * 2^(64 i)
  r[0] = ptx::mad_lo_cc(a[0], a[0], r[0]);
  r[1] = ptx::madc_hi_cc(a[0], a[0], r[1]);
  PSB_UNROLL
  for (int i = 1; i < N - 1; i++) {
    r[2 * i] = ptx::madc_lo_cc(a[i], a[i], r[2 * i]);
    r[2 * i + 1] = ptx::madc_hi_cc(a[i], a[i], r[2 * i + 1]);
  }
  r[2 * N - 2] = ptx::madc_lo_cc(a[N - 1], a[N - 1], r[2 * N - 2]);
  r[2 * N - 1] = ptx::madc_hi(a[N - 1], a[N - 1], r[2 * N - 1]);
#else
  mulw_n<N>(r, a, a);
#endif
}

// Montgomery reduction of t (2N limbs, value < m * 2^(32N)) -> r in [0, m)
template <class T>
PSB_HD PSB_INL void redc_n(uint32_t* r, const uint32_t* t) {
  constexpr int N = T::N;
  uint32_t m[N];
#ifdef __CUDA_ARCH__
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  PSB_UNROLL
  for (int k = 0; k < N; k++) {
    PSB_UNROLL
    for (int i = 0; i < N; i++) {
      if (i < k) ptx::mac3(c0, c1, c2, m[i], T::p(k - i));
    }
    ptx::acc3(c0, c1, c2, t[k]);
    m[k] = c0 * T::N0;
    ptx::mac3(c0, c1, c2, m[k], T::p(0));
    c0 = c1; c1 = c2; c2 = 0;
  }
  PSB_UNROLL
  for (int k = N; k < 2 * N; k++) {
    PSB_UNROLL
    for (int i = 0; i < N; i++) {
      if (i > k - N && k - i >= 0) ptx::mac3(c0, c1, c2, m[i], T::p(k - i));
    }
    ptx::acc3(c0, c1, c2, t[k]);
    r[k - N] = c0; c0 = c1; c1 = c2; c2 = 0;
  }
#else
  uint64_t lo = 0, hi = 0;
  auto mac = [&](uint32_t x, uint32_t y) { uint64_t pr = (uint64_t)x * y; uint64_t nl = lo + pr; hi += nl < lo; lo = nl; };
  auto acc = [&](uint32_t x) { uint64_t nl = lo + x; hi += nl < lo; lo = nl; };
  auto shift = [&]() { lo = (lo >> 32) | (hi << 32); hi = 0; };
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < k; i++) mac(m[i], T::p(k - i));
    acc(t[k]);
    m[k] = (uint32_t)lo * T::N0;
    mac(m[k], T::p(0));
    shift();
  }
  for (int k = N; k < 2 * N; k++) {
    for (int i = k - N + 1; i < N; i++) mac(m[i], T::p(k - i));
    acc(t[k]);
    r[k - N] = (uint32_t)lo;
    shift();
  }
#endif
  cond_sub_mod<T>(r);
}

// ------------------------------------------------------------------------------------------------
// typed wrappers
// ------------------------------------------------------------------------------------------------
struct alignas(16) Fp { uint32_t v[PSB_NL]; };    // canonical, Montgomery (mcl Fp: mcl/include/mcl/fp.hpp:76-106)
struct alignas(16) FpW { uint32_t v[2 * PSB_NL]; };   // unreduced double width, value in [0, p*R) (mcl FpDbl, fp_tower.hpp:13-178)
struct alignas(16) Fr { uint32_t v[8]; };

// 128-bit limb moves between an Fp in memory (16-byte aligned) and a register array
PSB_HD PSB_INL void fp_ld(uint32_t* d, const Fp& s) {
#ifdef __CUDA_ARCH__
  const uint4* q = reinterpret_cast<const uint4*>(s.v);
  const uint4 v0 = q[0], v1 = q[1];
  d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w; d[4] = v1.x; d[5] = v1.y; d[6] = v1.z; d[7] = v1.w;
#if PSB_NL == 12
  const uint4 v2 = q[2];
  d[8] = v2.x; d[9] = v2.y; d[10] = v2.z; d[11] = v2.w;
#endif
#else
  for (int i = 0; i < PSB_NL; i++) d[i] = s.v[i];
#endif
}
PSB_HD PSB_INL void fp_st(Fp& d, const uint32_t* t) {
#ifdef __CUDA_ARCH__
  uint4* q = reinterpret_cast<uint4*>(d.v);
  q[0] = make_uint4(t[0], t[1], t[2], t[3]);
  q[1] = make_uint4(t[4], t[5], t[6], t[7]);
#if PSB_NL == 12
  q[2] = make_uint4(t[8], t[9], t[10], t[11]);
#endif
#else
  for (int i = 0; i < PSB_NL; i++) d.v[i] = t[i];
#endif
}
PSB_HD PSB_INL void fp_add(Fp& r, const Fp& a, const Fp& b) { uint32_t x[PSB_NL], y[PSB_NL]; fp_ld(x, a); fp_ld(y, b); mod_add<FpT>(x, x, y); fp_st(r, x); }
PSB_HD PSB_INL void fp_sub(Fp& r, const Fp& a, const Fp& b) { uint32_t x[PSB_NL], y[PSB_NL]; fp_ld(x, a); fp_ld(y, b); mod_sub<FpT>(x, x, y); fp_st(r, x); }
PSB_HD PSB_INL void fp_neg(Fp& r, const Fp& a) { uint32_t x[PSB_NL]; fp_ld(x, a); mod_neg<FpT>(x, x); fp_st(r, x); }
PSB_HD PSB_INL void fp_dbl(Fp& r, const Fp& a) { uint32_t x[PSB_NL]; fp_ld(x, a); mod_add<FpT>(x, x, x); fp_st(r, x); }
// a + b without reduction (< 2p < 2^384): only as an operand of mulw (mcl Fp::addPre)
PSB_HD PSB_INL void fp_add_nr(Fp& r, const Fp& a, const Fp& b) { add_n<PSB_NL>(r.v, a.v, b.v); }
PSB_HD PSB_INL void fp_mulw(FpW& r, const Fp& a, const Fp& b) { mulw_n<PSB_NL>(r.v, a.v, b.v); }
PSB_HD PSB_INL void fp_sqrw(FpW& r, const Fp& a) { sqrw_n<PSB_NL>(r.v, a.v); }
PSB_HD PSB_INL void fp_redc(Fp& r, const FpW& t) { redc_n<FpT>(r.v, t.v); }
PSB_HD PSB_INL bool fp_is_zero(const Fp& a) { return is_zero_n<FpT>(a.v); }
PSB_HD PSB_INL bool fp_eq(const Fp& a, const Fp& b) {
  uint32_t o = 0;
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}
PSB_HD PSB_INL void fp_set_zero(Fp& r) {
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) r.v[i] = 0;
}
PSB_HD PSB_INL void fp_set_one(Fp& r) {
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) r.v[i] = PSB_K(FP_ONE)[i];
}
PSB_HD PSB_INL void fp_cmov(Fp& r, const Fp& a, bool c) {
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) r.v[i] = c ? a.v[i] : r.v[i];
}

// double-width add/sub modulo p*R (mcl FpDbl::add / ::sub, fp_tower.hpp:60-100): only the high
// half needs the correction because p*R has a zero low half.
PSB_HD PSB_INL void fpw_add(FpW& r, const FpW& a, const FpW& b) {
  uint32_t s[2 * PSB_NL];
  add_n<2 * PSB_NL>(s, a.v, b.v);  // < 2 p R < 2^768
  cond_sub_mod<FpT>(s + PSB_NL);
  PSB_UNROLL
  for (int i = 0; i < 2 * PSB_NL; i++) r.v[i] = s[i];
}
PSB_HD PSB_INL void fpw_sub(FpW& r, const FpW& a, const FpW& b) {
  uint32_t s[2 * PSB_NL];
  uint32_t borrow = sub_n<2 * PSB_NL>(s, a.v, b.v);
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) r.v[i] = s[i];
  add_mod_masked_n<FpT>(r.v + PSB_NL, s + PSB_NL, 0u - borrow);
}
// no-correction variants (caller guarantees 0 <= result < p*R)
PSB_HD PSB_INL void fpw_add_nr(FpW& r, const FpW& a, const FpW& b) { add_n<2 * PSB_NL>(r.v, a.v, b.v); }
PSB_HD PSB_INL void fpw_sub_nr(FpW& r, const FpW& a, const FpW& b) { sub_n<2 * PSB_NL>(r.v, a.v, b.v); }

// Montgomery product / square (mcl Fp::mul / Fp::sqr).  Device: even/odd CIOS kernels (cios.cuh).
}  // namespace psb
#ifdef __CUDA_ARCH__
#include "cios.cuh"
#endif
namespace psb {
#ifdef __CUDA_ARCH__
__device__ PSB_INL void fp_mul(Fp& r, const Fp& a, const Fp& b) { cios::mul(r.v, a.v, b.v); }
#ifndef PSB_FP_SQR_DEDICATED
#define PSB_FP_SQR_DEDICATED 1     // 0: squares through the general product (A/B builds)
#endif
#if PSB_FP_SQR_DEDICATED
__device__ PSB_INL void fp_sqr(Fp& r, const Fp& a) { cios::sqr(r.v, a.v); }
#else
__device__ PSB_INL void fp_sqr(Fp& r, const Fp& a) { cios::mul(r.v, a.v, a.v); }
#endif
__device__ PSB_INL void fp_dot2(Fp& r, const Fp& a, const Fp& b, const Fp& c, const Fp& d) { cios::dot2(r.v, a.v, b.v, c.v, d.v); }
#else
// host build (tests/hostsim only): same values through the generic product-scanning code.
// PSB_COUNT_OPS (tests/hostsim): count the calls -- on the device the same calls are one cios::mul (2 N^2 + N wide MACs),
// cios::sqr (N (N + 1) / 2 + N^2 + N) and cios::dot2 (3 N^2 + N) each, so the counts of a host run are the MACs a lane executes
// (bench.py `executed_mac32_per_lane`, pinned by tests/test_hostsim.py::test_executed_mac_counts)
#ifdef PSB_COUNT_OPS
namespace opcount { enum { MUL, SQR, DOT2, INV, N_ }; inline unsigned long long c[N_]; }
#define PSB_OPCOUNT(k) (++::psb::opcount::c[::psb::opcount::k])
#else
#define PSB_OPCOUNT(k) ((void)0)
#endif
PSB_HD inline void fp_mul(Fp& r, const Fp& a, const Fp& b) { PSB_OPCOUNT(MUL); FpW t; fp_mulw(t, a, b); fp_redc(r, t); }
PSB_HD inline void fp_sqr(Fp& r, const Fp& a) { PSB_OPCOUNT(SQR); FpW t; fp_sqrw(t, a); fp_redc(r, t); }
PSB_HD inline void fp_dot2(Fp& r, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
  PSB_OPCOUNT(DOT2);
  FpW t, u; fp_mulw(t, a, b); fp_mulw(u, c, d); fpw_add_nr(t, t, u); fp_redc(r, t);
}
#endif
#ifndef PSB_OPCOUNT
#define PSB_OPCOUNT(k) ((void)0)
#endif
// register-resident variants for the fused tower functions: operands are LOCAL Fp values that never touch memory
// (loaded once with fp_ld, stored once with fp_st); everything inlines.
#ifdef __CUDA_ARCH__
__device__ PSB_INL void fp_mul_rr(Fp& r, const Fp& a, const Fp& b) { cios::mul_rr(r.v, a.v, b.v); }
__device__ PSB_INL void fp_dot2_rr(Fp& r, const Fp& a, const Fp& b, const Fp& c, const Fp& d) { cios::dot2_rr(r.v, a.v, b.v, c.v, d.v); }
#else
PSB_HD PSB_INL void fp_mul_rr(Fp& r, const Fp& a, const Fp& b) { PSB_OPCOUNT(MUL); FpW t; mulw_n<PSB_NL>(t.v, a.v, b.v); redc_n<FpT>(r.v, t.v); }
PSB_HD PSB_INL void fp_dot2_rr(Fp& r, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
  PSB_OPCOUNT(DOT2);
  FpW t, u; mulw_n<PSB_NL>(t.v, a.v, b.v); mulw_n<PSB_NL>(u.v, c.v, d.v); add_n<2 * PSB_NL>(t.v, t.v, u.v); redc_n<FpT>(r.v, t.v);
}
#endif
PSB_HD PSB_INL void fp_add_rr(Fp& r, const Fp& a, const Fp& b) { mod_add<FpT>(r.v, a.v, b.v); }
PSB_HD PSB_INL void fp_sub_rr(Fp& r, const Fp& a, const Fp& b) { mod_sub<FpT>(r.v, a.v, b.v); }
PSB_HD PSB_INL void fp_dbl_rr(Fp& r, const Fp& a) { mod_add<FpT>(r.v, a.v, a.v); }
PSB_HD PSB_INL void fp_addnr_rr(Fp& r, const Fp& a, const Fp& b) { add_n<PSB_NL>(r.v, a.v, b.v); }   // < 2p: multiplicand only
// p - a for a in [0, p]: in (0, p] (NOT canonical for a = 0) -- multiplicand only
PSB_HD PSB_INL void fp_pminus_rr(Fp& r, const Fp& a) {
#ifdef __CUDA_ARCH__
  r.v[0] = ptx::sub_cc(FpT::p(0), a.v[0]);
  PSB_UNROLL
  for (int i = 1; i < PSB_NL - 1; i++) r.v[i] = ptx::subc_cc(FpT::p(i), a.v[i]);
  r.v[PSB_NL - 1] = ptx::subc(FpT::p(PSB_NL - 1), a.v[PSB_NL - 1]);
#else
  int64_t c = 0;
  for (int i = 0; i < PSB_NL; i++) { c += (int64_t)FpT::p(i) - a.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
#endif
}
// 2p - a for a in [0, 2p): in (0, 2p] -- multiplier side of a product whose operands are unreduced sums (tower.cuh, engine A)
PSB_HD PSB_INL uint32_t fp_2p_limb(int i) { return i == 0 ? (FpT::p(0) << 1) : ((FpT::p(i) << 1) | (FpT::p(i - 1) >> 31)); }
PSB_HD PSB_INL void fp_2pminus_rr(Fp& r, const Fp& a) {
#ifdef __CUDA_ARCH__
  r.v[0] = ptx::sub_cc(fp_2p_limb(0), a.v[0]);
  PSB_UNROLL
  for (int i = 1; i < PSB_NL - 1; i++) r.v[i] = ptx::subc_cc(fp_2p_limb(i), a.v[i]);
  r.v[PSB_NL - 1] = ptx::subc(fp_2p_limb(PSB_NL - 1), a.v[PSB_NL - 1]);
#else
  int64_t c = 0;
  for (int i = 0; i < PSB_NL; i++) { c += (int64_t)fp_2p_limb(i) - a.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
#endif
}
PSB_HD PSB_INL void fp_get(Fp& d, const Fp& mem) { fp_ld(d.v, mem); }
PSB_HD PSB_INL void fp_put(Fp& mem, const Fp& s) { fp_st(mem, s.v); }
PSB_HD PSB_INL void fp_mul_inl(Fp& r, const Fp& a, const Fp& b) { fp_mul(r, a, b); }
PSB_HD PSB_INL void fp_sqr_inl(Fp& r, const Fp& a) { fp_sqr(r, a); }

// a^-1 = a^(p-2) (Fermat; mcl Fp::inv gives the same canonical value, mcl/src/fp.cpp:215-246).
// Fixed 4-bit windows: constant schedule, no divergence.  inv(0) = 0 like mcl's.
PSB_HD PSB_NOINL void fp_pow_nib(Fp& r, const Fp& a, const uint32_t* nib /*8 * PSB_NL LE nibbles*/) {
  Fp tbl[16];
  fp_set_one(tbl[0]);
  tbl[1] = a;
  for (int i = 2; i < 16; i++) fp_mul(tbl[i], tbl[i - 1], a);
  Fp acc = tbl[nib[8 * PSB_NL - 1]];
  for (int i = 8 * PSB_NL - 2; i >= 0; i--) {
    fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc);
    const uint32_t d = nib[i];
    if (d) fp_mul(acc, acc, tbl[d]);  // nib is a per-kernel constant: uniform branch
  }
  r = acc;
}
// (kept as the cross-check of modinv.cuh and for A/B builds with -DPSB_INV_FERMAT; the product path uses fp_inv there)
PSB_HD PSB_INL void fp_inv_fermat(Fp& r, const Fp& a) { fp_pow_nib(r, a, PSB_K(FP_PM2_NIB)); }

// normal form <-> Montgomery
PSB_HD PSB_INL void fp_from_mont(Fp& r, const Fp& a) {
  FpW t;
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) { t.v[i] = a.v[i]; t.v[i + PSB_NL] = 0; }
  fp_redc(r, t);
}
PSB_HD PSB_INL void fp_to_mont(Fp& r, const Fp& a) {
  Fp r2;
  PSB_UNROLL
  for (int i = 0; i < PSB_NL; i++) r2.v[i] = PSB_K(FP_R2)[i];
  fp_mul(r, a, r2);
}

// ---- Fr (scalars; mcl Fr = FpT<FrTag,256>, Montgomery R = 2^256) --------------------------------
PSB_HD PSB_INL void fr_add(Fr& r, const Fr& a, const Fr& b) { mod_add<FrT>(r.v, a.v, b.v); }
PSB_HD PSB_INL void fr_sub(Fr& r, const Fr& a, const Fr& b) { mod_sub<FrT>(r.v, a.v, b.v); }
PSB_HD PSB_INL void fr_mul(Fr& r, const Fr& a, const Fr& b) {
  uint32_t t[16];
  mulw_n<8>(t, a.v, b.v);
  redc_n<FrT>(r.v, t);
}
PSB_HD PSB_INL void fr_from_mont(Fr& r, const Fr& a) {
  uint32_t t[16];
  PSB_UNROLL
  for (int i = 0; i < 8; i++) { t[i] = a.v[i]; t[i + 8] = 0; }
  redc_n<FrT>(r.v, t);
}
PSB_HD PSB_INL void fr_to_mont(Fr& r, const Fr& a) {
  Fr r2;
  PSB_UNROLL
  for (int i = 0; i < 8; i++) r2.v[i] = PSB_K(FR_R2)[i];
  fr_mul(r, a, r2);
}
PSB_HD PSB_INL bool fr_eq(const Fr& a, const Fr& b) {
  uint32_t o = 0;
  PSB_UNROLL
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}
// x >= r ?   (normal-form limbs)
PSB_HD PSB_INL bool fr_geq_modulus(const uint32_t* x) {
  uint32_t t[8];
  return sub_mod_n<FrT>(t, x) == 0;
}

}  // namespace psb
