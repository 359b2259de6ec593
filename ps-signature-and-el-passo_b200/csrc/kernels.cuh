// kernels.cuh -- __global__ kernels of the batch engine: one GPU thread per batch lane.
//
// Lanes are independent, so every kernel is a plain 1-D grid over lanes; the big
// per-lane state (an Fp12 is 576 B, a G2 point 288 B) lives in thread-local memory that the
// hardware interleaves per warp (coalesced, L1/L2 resident), while the limb arithmetic runs in
// registers.  Work is split into phase kernels (scalars -> fixed-base MSM -> Miller loop -> final
// exponentiation) so that each phase gets its own register allocation; the hand-over state goes
// through HBM once per phase (< 1 KB per lane against millions of integer MACs: negligible).
#pragma once
#include "testops.cuh"
#include "protocol.cuh"
#include "prover.cuh"
#include "hash_to_curve.cuh"
#include "wire.cuh"

namespace psb {

#ifndef PSB_BLOCK
#define PSB_BLOCK 128
#endif
constexpr int kBlock = PSB_BLOCK;
// The pairing-pipeline kernels (MSM, Miller loop, final exponentiation) are compiled for 128 registers/thread
// (__launch_bounds__(512, 1)) = 16 resident warps per SM, launched as ONE 512-thread block per SM for large batches and
// as 256- or 128-thread blocks for small ones (pair_block() in psb_api.cu).  Measured on B200 (r1s, 2^20 lanes): the warps
// of a block run the same straight-line schedule nearly in step and share instruction fetches -- the final
// exponentiation, whose stall samples are 18 % instruction fetch, takes 363 / 345 / 337 ms with 128 / 256 / 512-thread
// blocks at the same occupancy; the Miller loop and the MSM do not care.  More warps per SM (5/6/8 x 128 threads)
// only lowered the SM clock at the 1000 W cap (1942 -> 1837 -> 1762 -> 1702 MHz) and the throughput (r1).
#ifndef PSB_PAIR_BLOCK
#define PSB_PAIR_BLOCK 512
#endif
constexpr int kPairBlock = PSB_PAIR_BLOCK;
#ifndef PSB_PAIR_MINB
#define PSB_PAIR_MINB (512 / PSB_PAIR_BLOCK)   // 128 registers per thread at any block size; 1 with 256-thread blocks lets ptxas use up to 255
#endif
#define PSB_PAIR_BOUNDS __launch_bounds__(kPairBlock, PSB_PAIR_MINB)
// protocol kernels (EL PASSO NIZK steps, issuance, prover side): 128-thread blocks
#ifndef PSB_PROTO_MINB
#define PSB_PROTO_BOUNDS __launch_bounds__(kBlock)
#else
#define PSB_PROTO_BOUNDS __launch_bounds__(kBlock, PSB_PROTO_MINB)
#endif
// the two G2 NIZK kernels hold their register count at 168 (three blocks = 12 warps per SM): left to itself ptxas settles on
// 168 or ~248 registers depending on small changes elsewhere in the lane function, and at 248 (8 warps per SM) sign-on
// verification loses the whole gain of the batched affine sums (profiles/r2z_ab_config_windows.txt)
#define PSB_PROTO_G2_BOUNDS __launch_bounds__(kBlock, 3)

// ---- steering ptxas's pipe balancer ---------------------------------------------------------------------------------
// ptxas balances the integer-add (ALU) and multiply (FMA) pipes from the STATIC instruction counts of a whole kernel: in a kernel
// with much additive glue it rewrites carry absorptions and register moves INSIDE the multiplier rows as IMAD.X / IMAD.MOV
// (25 + 13 per engine pass in k_verify_miller), i.e. it puts them on the one pipe that is saturated at run time -- the static
// count cannot know that the 11 KB multiplier body executes 2 800 times per lane and the glue once.  A block of plain IMADs
// that never executes (its condition is a lane count no batch can have) tips the static balance: with it the same rows come
// out with IADD3.X / MOV only (checked with cuobjdump: tools/sass_regions.py; the engine compiled on its own gets the same
// code).  Cold code at the end of the kernel: no instruction-cache footprint on the hot path.
#ifndef PSB_FMA_BALLAST
#define PSB_FMA_BALLAST 3000
#endif
template <int COUNT>
__device__ __forceinline__ void ptxas_fma_ballast(bool never, void* sink) {
#if PSB_FMA_BALLAST > 0
  if (never) {
    uint32_t u = threadIdx.x;
    const uint32_t v = blockIdx.x;
#pragma unroll
    for (int i = 0; i < COUNT; i++) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(u) : "r"(v));
    *reinterpret_cast<volatile uint32_t*>(sink) = u;
  }
#else
  (void)never; (void)sink;
#endif
}
#define PSB_BALLAST(N, sink) ptxas_fma_ballast<PSB_FMA_BALLAST>((N) == ~(size_t)0, (void*)(sink))

// ---- parity probe ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_test_op(int op, size_t n, int s0, int s1, int s2, int s3,
                                                    const uint32_t* a, const uint32_t* b, const uint32_t* c,
                                                    uint32_t* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  test_op_run(op, a + i * s0, b ? b + i * s1 : nullptr, c ? c + i * s2 : nullptr, out + i * s3);
}

// ---- micro-benchmarks -----------------------------------------------------------------------------
// keeps random limb patterns below p (top limb of p: 0x1a0111ea for BLS12-381, 0x25236482 for BN254)
constexpr uint32_t kTopMask = 0x0fffffffu;
__global__ void __launch_bounds__(256) k_bench_fp(int kind, int iters, const uint32_t* seed, uint32_t* sink) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kind == 0 || kind == 1) {
    Fp x, y;
    for (int j = 0; j < PSB_NL; j++) { x.v[j] = seed[j] ^ (uint32_t)i; y.v[j] = seed[PSB_NL + j]; }
    x.v[PSB_NL - 1] &= kTopMask; y.v[PSB_NL - 1] &= kTopMask;
    for (int it = 0; it < iters; it++) {
      if (kind == 0) fp_mul_inl(x, x, y); else fp_sqr_inl(x, x);
    }
    if (x.v[0] == 0xdeadbeefu) sink[0] = x.v[1];
  } else if (kind == 2) {
    Fp2 x, y;
    uint32_t* px = (uint32_t*)&x; uint32_t* py = (uint32_t*)&y;
    for (int j = 0; j < 2 * PSB_NL; j++) { px[j] = seed[j % PSB_NL] ^ (uint32_t)i; py[j] = seed[PSB_NL + (j % PSB_NL)]; }
    x.a.v[PSB_NL - 1] &= kTopMask; x.b.v[PSB_NL - 1] &= kTopMask; y.a.v[PSB_NL - 1] &= kTopMask; y.b.v[PSB_NL - 1] &= kTopMask;
    for (int it = 0; it < iters; it++) fp2_mul(x, x, y);
    if (x.a.v[0] == 0xdeadbeefu) sink[0] = x.a.v[1];
  } else if (kind == 3) {
    Fp12 x, y;
    uint32_t* px = (uint32_t*)&x; uint32_t* py = (uint32_t*)&y;
    for (int j = 0; j < 12 * PSB_NL; j++) { px[j] = seed[j % PSB_NL] ^ (uint32_t)i; py[j] = seed[PSB_NL + (j % PSB_NL)]; }
    for (int j = PSB_NL - 1; j < 12 * PSB_NL; j += PSB_NL) { px[j] &= kTopMask; py[j] &= kTopMask; }
    for (int it = 0; it < iters; it++) fp12_mul(x, x, y);
    if (x.a.a.a.v[0] == 0xdeadbeefu) sink[0] = x.a.a.a.v[1];
  }
}

// raw integer-multiply issue-rate probes: 8 independent accumulator pairs per thread
__global__ void __launch_bounds__(256) k_bench_mad(int kind, int iters, const uint32_t* seed, uint32_t* sink) {
  const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
  uint32_t a = seed[0] ^ t, b = seed[1] + t;
  if (kind == 4) {
    uint32_t lo[8], hi[8];
    for (int j = 0; j < 8; j++) { lo[j] = seed[j] + t; hi[j] = seed[8 + j] ^ t; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a), "r"(b));
      }
      a += lo[0];
    }
    uint32_t s = 0;
    for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j];
    if (s == 0xdeadbeefu) sink[0] = s;
  } else if (kind == 5) {
    unsigned long long acc[8];
    for (int j = 0; j < 8; j++) acc[j] = ((unsigned long long)seed[j] << 32) | (seed[8 + j] ^ t);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a), "r"(b));
      }
      a += (uint32_t)acc[0];
    }
    unsigned long long s = 0;
    for (int j = 0; j < 8; j++) s ^= acc[j];
    if (s == 0xdeadbeefull) sink[0] = (uint32_t)s;
  } else if (kind == 7) {
#ifdef __CUDA_ARCH__
    // the multiplier's own instruction form: carry-chained wide MACs (IMAD.WIDE.U32.X rows of six, as in cios.cuh),
    // three independent accumulator pairs = six independent carry chains per thread.  72 MAC32 per iteration.
    cios::L12 E1, O1, E2, O2, E3, O3, av;
    for (int j = 0; j < PSB_NL; j++) { E1[j] = seed[j] + t; O1[j] = seed[PSB_NL + j] ^ t; E2[j] = E1[j] + 1; O2[j] = O1[j] + 2; E3[j] = E1[j] ^ 5; O3[j] = O1[j] ^ 9; av[j] = seed[j] * (t | 1); }
    for (int it = 0; it < iters; it++) {
      cios::mac(E1, O1, av, b);
      cios::mac(E2, O2, av, a);
      cios::mac(E3, O3, av, b ^ a);
      a += E1[0];
    }
    uint32_t s = 0;
    for (int j = 0; j < PSB_NL; j++) s ^= E1[j] ^ O1[j] ^ E2[j] ^ O2[j] ^ E3[j] ^ O3[j];
    if (s == 0xdeadbeefu) sink[0] = s;
#endif
  } else if (kind == 12 || kind == 13 || kind == 14) {
#ifdef __CUDA_ARCH__
    // kind 7's carry-chained rows (36 MAC32 per iteration) with an FP64 stream beside them: 24 (kind 12) or 72 (kind 13)
    // independent DFMA per iteration, or the 72 DFMA alone (kind 14).  Question: does the idle FP64 pipe issue beside the
    // saturated integer-multiply pipe?  (72 DFMA = two per MAC32: an exact 32 x 32 product takes two DFMA with 16-bit split factors.)
    cios::L12 E1, O1, E2, O2, E3, O3, av;
    for (int j = 0; j < PSB_NL; j++) { E1[j] = seed[j] + t; O1[j] = seed[PSB_NL + j] ^ t; E2[j] = E1[j] + 1; O2[j] = O1[j] + 2; E3[j] = E1[j] ^ 5; O3[j] = O1[j] ^ 9; av[j] = seed[j] * (t | 1); }
    double f[12], x = 1.0 + (double)(seed[0] & 7u) * 1e-9, y = (double)(t | 1u);
    for (int j = 0; j < 12; j++) f[j] = (double)(seed[j] ^ t);
    const int nd = kind == 12 ? 2 : 6;
    for (int it = 0; it < iters; it++) {
      if (kind != 14) cios::mac(E1, O1, av, b);
      for (int r = 0; r < nd; r++)
#pragma unroll
        for (int j = 0; j < 4; j++) f[j] = __fma_rz(f[j], x, y);
      if (kind != 14) cios::mac(E2, O2, av, a);
      for (int r = 0; r < nd; r++)
#pragma unroll
        for (int j = 4; j < 8; j++) f[j] = __fma_rz(f[j], x, y);
      if (kind != 14) cios::mac(E3, O3, av, b ^ a);
      for (int r = 0; r < nd; r++)
#pragma unroll
        for (int j = 8; j < 12; j++) f[j] = __fma_rz(f[j], x, y);
      a += E1[0];
    }
    uint32_t s = 0;
    double fs = 0;
    for (int j = 0; j < PSB_NL; j++) s ^= E1[j] ^ O1[j] ^ E2[j] ^ O2[j] ^ E3[j] ^ O3[j];
    for (int j = 0; j < 12; j++) fs += f[j];
    if (s == 0xdeadbeefu || fs == 12345.678) sink[0] = s;
#endif
  } else if (kind == 8) {
    // FP64 fused multiply-add (round toward zero, the form of the double-precision big-number multipliers): 8 chains.
    // Orientation for a DFMA-based Montgomery multiplier: one DFMA carries a 52-bit x 52-bit partial product.
    double acc[8], x = (double)(a | 1u), y = 1.0 + (double)(b & 0xffu) * 1e-9;
    for (int j = 0; j < 8; j++) acc[j] = (double)(seed[j] ^ t);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = __fma_rz(acc[j], y, x);
    }
    double s = 0;
    for (int j = 0; j < 8; j++) s += acc[j];
    if (s == 12345.678) sink[0] = 1;
  } else if (kind == 9) {
    // one row of a carry-free radix-2^30 multiplier: 13 independent 64-bit accumulators, 13 distinct multiplicand limbs
    unsigned long long acc[13];
    uint32_t av[13];
    for (int j = 0; j < 13; j++) { acc[j] = ((unsigned long long)(seed[j] & 0xffff) << 32) | (seed[8 + j] ^ t); av[j] = (seed[j + 3] + t) & 0x3fffffffu; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 13; j++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(av[j]), "r"(b));
      }
      b = (b + (uint32_t)acc[0]) & 0x3fffffffu;   // (sums wrap: a rate probe, not arithmetic)
    }
    unsigned long long s = 0;
    for (int j = 0; j < 13; j++) s ^= acc[j];
    if (s == 0xdeadbeefull) sink[0] = (uint32_t)s;
  } else if (kind == 10 || kind == 11) {
    // carry-free wide MACs that ptxas cannot fold: 16 independent 64-bit accumulators, a DISTINCT multiplicand per
    // accumulator (kind 10: shared multiplier b, the operand pattern of a multiplier row; kind 11: distinct multiplier
    // too).  (Kind 5 shares BOTH factors across its accumulators and ptxas turns it into ONE IMAD.WIDE plus 64-bit
    // additions per trip -- checked in SASS -- so it measures the adder, not the multiplier: not a MAC peak.)
    unsigned long long acc[16];
    uint32_t av[16], cv[16];
    for (int j = 0; j < 16; j++) { acc[j] = ((unsigned long long)seed[j] << 32) | (seed[16 + j] ^ t); av[j] = seed[j + 5] * (t | 1) + j; cv[j] = seed[j + 9] ^ (t * 2654435761u) ^ j; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        if (kind == 10) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(av[j]), "r"(b));
        else asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(av[j]), "r"(cv[j]));
      }
      b += (uint32_t)acc[0];
    }
    unsigned long long s = 0;
    for (int j = 0; j < 16; j++) s ^= acc[j];
    if (s == 0xdeadbeefull) sink[0] = (uint32_t)s;
  } else {
    // plain 32-bit IMAD (lo only), 16 independent chains
    uint32_t lo[16];
    for (int j = 0; j < 16; j++) lo[j] = seed[j] + t;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 16; j++) lo[j] = lo[j] * a + b;
      a += lo[0];
    }
    uint32_t s = 0;
    for (int j = 0; j < 16; j++) s ^= lo[j];
    if (s == 0xdeadbeefu) sink[0] = s;
  }
}

// ---- key setup --------------------------------------------------------------------------------------
// normalise `count` points in place (any z -> z = 1)
template <class F>
__global__ void k_normalize_points(Jac<F>* pts, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Jac<F> p = pts[i], r;
  pt_normalize(r, p);
  pts[i] = r;
}

// window bases: wb[b * nwin + j] = 2^(w j) * base[b]  (affine).  One thread per base.
template <class F>
__global__ void k_window_bases(const Jac<F>* bases /*normalised*/, int nbases, int w, Aff<F>* wb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbases) return;
  const int nwin = fixed_nwin(w);
  Jac<F> cur = bases[b];
  for (int j = 0; j < nwin; j++) {
    Jac<F> nrm;
    pt_normalize(nrm, cur);
    wb[(size_t)b * nwin + j].x = nrm.x;
    wb[(size_t)b * nwin + j].y = nrm.y;
    for (int t = 0; t < w; t++) pt_dbl(cur, cur);
  }
}

// table entries: tbl[(b * nwin + j) * half + (d-1)] = d * wb[b * nwin + j], d = 1..half, affine.
// One thread per chunk of kTblChunk consecutive d; chunk-local Montgomery batch inversion.
constexpr int kTblChunk = 16;
template <class F>
__global__ void __launch_bounds__(kBlock) k_build_table(const Aff<F>* wb, int nbases, int w, Aff<F>* tbl) {
  const int nwin = fixed_nwin(w);
  const uint32_t half = 1u << (w - 1);
  const uint32_t chunks_per_win = (half + kTblChunk - 1) / kTblChunk;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)nbases * nwin * chunks_per_win;
  if (gid >= total) return;
  const size_t bw = gid / chunks_per_win;       // b * nwin + j
  const uint32_t ch = (uint32_t)(gid % chunks_per_win);
  const uint32_t d0 = ch * kTblChunk;           // entries d0+1 .. d0+kTblChunk
  const Aff<F> B = wb[bw];
  // S = d0 * B by double-and-add (d0 < 2^(w-1))
  Jac<F> S;
  pt_set_zero(S);
  for (int bit = w - 2; bit >= 0; bit--) {
    pt_dbl(S, S);
    if ((d0 >> bit) & 1u) pt_madd(S, S, B);
  }
  Jac<F> J[kTblChunk];
  F pre[kTblChunk];
  for (int t = 0; t < kTblChunk; t++) {
    pt_madd(S, S, B);
    J[t] = S;
    if (t == 0) pre[0] = S.z; else f_mul(pre[t], pre[t - 1], S.z);
  }
  F inv;
  f_inv(inv, pre[kTblChunk - 1]);
  for (int t = kTblChunk - 1; t >= 0; t--) {
    F zi, zi2;
    if (t == 0) zi = inv; else f_mul(zi, inv, pre[t - 1]);
    f_mul(inv, inv, J[t].z);
    const uint32_t d = d0 + t + 1;
    if (d <= half) {
      Aff<F> e;
      f_sqr(zi2, zi);
      f_mul(e.x, J[t].x, zi2);
      f_mul(zi2, zi2, zi);
      f_mul(e.y, J[t].y, zi2);
      tbl[bw * half + (d - 1)] = e;
    }
  }
}

__global__ void k_fixed_lines(const G2J* gg /*normalised*/, FixedLine* lines) {
  if (blockIdx.x * blockDim.x + threadIdx.x != 0) return;
  G2A q;
  q.x = gg->x; q.y = gg->y;
  precompute_fixed_lines(lines, q);
}

// ---- PS verification pipeline -------------------------------------------------------------------------
// phase 1: scalars m_i (SHA-256 of the attribute strings, or the caller's Fr) and
//          K = XX + sum_i m_i YY_i from the per-key window tables.
__global__ void PSB_PAIR_BOUNDS k_verify_msm(size_t N, size_t base, int n, int w, const uint8_t* blob, const uint64_t* off,
                                                        const Fr* m_mont, const G2J* XX, const G2A* tbl, G2J* Kout, int affine) {
  const size_t lane = base + (size_t)blockIdx.x * blockDim.x + threadIdx.x;    // lanes [base, N): psb_api.cu, for_waves
  if (lane >= N) return;
  const size_t per_base = (size_t)fixed_nwin(w) << (w - 1);
  G2J acc = *XX;
  AffBatch<Fp2> batch;
  AffPts<Fp2> level2;
  if (affine) aff_init(batch, n * fixed_nwin(w), (size_t)n * per_base, tbl, (const G2A*)nullptr, (const G2A*)nullptr, affine > 1 ? &level2 : nullptr);
  for (int i = 0; i < n; i++) {
    uint32_t k[8];
    if (blob) {
      const uint64_t b = off[lane * n + i], e = off[lane * n + i + 1];
      fr_set_hash_of(k, blob + b, (size_t)(e - b));
    } else {
      Fr t, tn;
      t = m_mont[lane * n + i];
      fr_from_mont(tn, t);
      for (int j = 0; j < 8; j++) k[j] = tn.v[j];
    }
    if (affine) aff_push_fixed_mul(acc, batch, 0, (size_t)i * per_base, k, w);   // pair sums in affine coordinates (curve.cuh)
    else pt_fixed_mul_acc(acc, tbl + (size_t)i * per_base, k, w);
  }
  if (affine) aff_flush(acc, batch);
  Kout[lane] = acc;
  PSB_BALLAST(N, Kout);
}

// phase 2: f = ML(sig1, K) * ML(-sig2, gg)  (one multi-Miller loop per lane)
//          ss = stride of the credential arrays in G1J units: 1 for separate sigma1 / sigma2 arrays, 2 for an array of
//          (sigma1, sigma2) pairs such as std::vector<PSCredential> (sig2 = sig1 + 1)
__global__ void PSB_PAIR_BOUNDS k_verify_miller(size_t N, size_t base, const G1J* sig1, const G1J* sig2, int ss, const G2J* K,
                                                           const FixedLine* lines, Fp12* fout) {
  const size_t lane0 = base + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = lane0 < N;
  const size_t lane = live ? lane0 : N - 1;        // as in k_verify_final: no early exit, the block may re-align itself
  Fp x1, y1, x2, y2;
  {
    G1J p = sig1[lane * ss];
    g1_affine_for_pairing(x1, y1, p);
    p = sig2[lane * ss];
    g1_affine_for_pairing(x2, y2, p);
    fp_neg(y2, y2);
  }
  G2J q = K[lane];
  Fp12 f;
  miller_loop2(f, x1, y1, q, x2, y2, lines, true, true);
  if (live) fout[lane] = f;
  PSB_BALLAST(N, fout);
}

// phase 3: final exponentiation, verdict = (sig1 != 0) && (f^e == 1), optional GT
//          reject_zero_sig1: PSVerifier::verify rejects sig1 == 0 (ps-verifier.cc:16-18), el_passo_verify_id does not;
//          pre (optional): per-lane verdict of an earlier step (the NIZK check) that is ANDed in.
__global__ void PSB_PAIR_BOUNDS k_verify_final(size_t N, size_t base, const G1J* sig1, int ss, const Fp12* fin, uint8_t* verdict,
                                                          Fp12* gt, const uint8_t* pre, int reject_zero_sig1) {
  const size_t lane0 = base + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = lane0 < N;
  const size_t lane = live ? lane0 : N - 1;        // surplus threads of the last block shadow the last lane: no early exit,
  Fp12 f = fin[lane], e;                           // so the block can re-align itself inside final_exp
  final_exp(e, f, true);
  if (!live) return;
  const bool s1zero = reject_zero_sig1 && fp_is_zero(sig1[lane * ss].z);
  const bool pre_ok = pre ? pre[lane] != 0 : true;
  verdict[lane] = (pre_ok && !s1zero && fp12_is_one(e)) ? 1 : 0;
  if (gt) gt[lane] = e;
  PSB_BALLAST(N, verdict);
}

// plain pairing e(P, Q) per lane (no fixed argument)
__global__ void PSB_PAIR_BOUNDS k_pairing_miller(size_t N, size_t base, const G1J* P, const G2J* Q, Fp12* fout) {
  const size_t lane0 = base + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = lane0 < N;
  const size_t lane = live ? lane0 : N - 1;
  Fp x1, y1, zero;
  fp_set_zero(zero);
  G1J p = P[lane];
  g1_affine_for_pairing(x1, y1, p);
  G2J q = Q[lane];
  Fp12 f;
  miller_loop2(f, x1, y1, q, zero, zero, nullptr, false, true);
  if (live) fout[lane] = f;
  PSB_BALLAST(N, fout);
}
__global__ void PSB_PAIR_BOUNDS k_final_exp(size_t N, size_t base, const Fp12* fin, Fp12* out) {
  const size_t lane0 = base + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = lane0 < N;
  const size_t lane = live ? lane0 : N - 1;
  Fp12 f = fin[lane], e;
  final_exp(e, f, true);
  if (live) out[lane] = e;
  PSB_BALLAST(N, out);
}


// ---- PSRequester::randomize_credential (src/ps-requester.cc:139-148) -----------------------------------
// out = (t sig1, t sig2) normalised; t host-supplied (Fr Montgomery)
__global__ void PSB_PROTO_BOUNDS k_randomize(size_t N, const G1J* sig1, const G1J* sig2, const Fr* t,
                                                       G1J* out1, G1J* out2, uint8_t* ser) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  Fr k, kn;
  k = t[lane];
  fr_from_mont(kn, k);
  G1J a = sig1[lane], b = sig2[lane], ra, rb;
  pt_mul(ra, a, kn.v);
  pt_mul(rb, b, kn.v);
  g1_normalize2(ra, rb);
  out1[lane] = ra;
  out2[lane] = rb;
  if (ser) {
    g1_serialize_norm(ser + lane * 2 * kFpBytes, ra);
    g1_serialize_norm(ser + lane * 2 * kFpBytes + kFpBytes, rb);
  }
  PSB_BALLAST(N, ser);
}

// out[j] = k[j] * P[j or 0], normalised (generic batched G1::mul, used to synthesise workloads)
__global__ void PSB_PROTO_BOUNDS k_g1_mul(size_t N, const G1J* P, int p_stride, const Fr* k, G1J* out) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  Fr kk, kn;
  kk = k[lane];
  fr_from_mont(kn, kk);
  G1J a = P[p_stride ? lane : 0], r, n;
  pt_mul(r, a, kn.v);
  pt_normalize(n, r);
  out[lane] = n;
  PSB_BALLAST(N, out);
}

// ---- batched G1::deserialize / G2::deserialize (point decompression) ----------------------
// element j is read at ser + j * stride (+ offset applied by the caller); ok[j] &= / = decode success
__global__ void __launch_bounds__(kBlock) k_g1_deserialize(size_t N, const uint8_t* ser, size_t stride, G1J* out, uint8_t* ok,
                                                            int accumulate) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J P;
  const bool r = g1_deserialize(P, ser + lane * stride);
  out[lane] = P;
  ok[lane] = (uint8_t)((accumulate ? ok[lane] : 1) & (r ? 1 : 0));
}
__global__ void __launch_bounds__(kBlock) k_g2_deserialize(size_t N, const uint8_t* ser, size_t stride, G2J* out, uint8_t* ok,
                                                            int accumulate) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G2J P;
  const bool r = g2_deserialize(P, ser + lane * stride);
  out[lane] = P;
  ok[lane] = (uint8_t)((accumulate ? ok[lane] : 1) & (r ? 1 : 0));
}

// ---- PSSigner::el_passo_provide_id (src/ps-signer.cc:63-146) -------------------------------------------
// Lane geometry shared by the EL PASSO kernels: lane j reads its responses at rs + j * rs_stride, `per` of them
// (per_lane[j] when the counts come from parsed wire buffers, else the batch's `per`), and its attribute offsets at
// off + j * off_stride (n entries + the end: stride n for a packed batch whose lanes share boundaries, n + 1 for
// parsed buffers).  pre (optional): verdict of an earlier step (the wire parser) that is ANDed in.
struct LaneGeom {
  int per, rs_stride, off_stride;
  const int* per_lane;
  const uint8_t* pre;
  PSB_HD int count(size_t lane) const { return per_lane ? per_lane[lane] : per; }
  PSB_HD bool pre_ok(size_t lane) const { return pre ? pre[lane] != 0 : true; }
};
__global__ void PSB_PROTO_BOUNDS k_provide_id(size_t N, int n, int w, const G1A* tblG1, const G1J* g1pts,
                                                        const G1J* A, const Fr* c, const Fr* rs, LaneGeom lg,
                                                        const uint8_t* blob, const uint64_t* off, const uint8_t* ad_blob,
                                                        const uint64_t* ad_off, const Fr* u, uint8_t* verdict, G1J* sig1,
                                                        G1J* sig2, uint8_t* ser) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J s1, s2;
  const TblGeom tg{w};
  bool ok = provide_id_lane(n, tg, tblG1, g1pts[1], A[lane], c + lane, rs + lane * lg.rs_stride, lg.count(lane), blob,
                            off + lane * lg.off_stride, ad_blob + ad_off[lane], (size_t)(ad_off[lane + 1] - ad_off[lane]),
                            u + lane, s1, s2);
  if (!lg.pre_ok(lane)) { ok = false; pt_set_zero(s1); pt_set_zero(s2); }
  verdict[lane] = ok ? 1 : 0;
  sig1[lane] = s1;
  sig2[lane] = s2;
  if (ser) {
    g1_serialize_norm(ser + lane * 2 * kFpBytes, s1);
    g1_serialize_norm(ser + lane * 2 * kFpBytes + kFpBytes, s2);
  }
  PSB_BALLAST(N, ser);
}

// ---- PSSigner::sign_commitment / sign_hybrid (src/ps-signer.cc:112-146), u host-supplied ------------------------
__global__ void PSB_PROTO_BOUNDS k_sign(size_t N, int na, int w, const G1A* tblG1, const G1J* g1pts, const G1J* C,
                                                  const uint8_t* blob, const uint64_t* off, const Fr* u, G1J* sig1, G1J* sig2,
                                                  uint8_t* ser) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J s1, s2;
  sign_lane(na, TblGeom{w}, tblG1, g1pts[1], C[lane], blob, off ? off + lane * na : nullptr, u + lane, s1, s2);
  sig1[lane] = s1;
  sig2[lane] = s2;
  if (ser) {
    g1_serialize_norm(ser + lane * 2 * kFpBytes, s1);
    g1_serialize_norm(ser + lane * 2 * kFpBytes + kFpBytes, s2);
  }
  PSB_BALLAST(N, ser);
}

// ---- PSVerifier::el_passo_verify_id (src/ps-verifier.cc:37-212): NIZK steps; the pairing check reuses
//      k_verify_miller / k_verify_final with K from step 1 ------------------------------------------------
__global__ void PSB_PROTO_G2_BOUNDS k_vid_g2(size_t N, int n, int w, const G2A* tblYY, const G2A* tblAux, const G2J* k,
                                                    const Fr* c, const Fr* rs, LaneGeom lg, int with_id, const uint8_t* blob,
                                                    const uint64_t* off, G2J* Vk, G2J* K, uint8_t* ok) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G2J vk, kk;
  const TblGeom tg{w};
  const bool r = verify_id_g2_lane(n, tg, tblYY, tblAux, k[lane], c + lane, rs + lane * lg.rs_stride, lg.count(lane), with_id,
                                   blob, off + lane * lg.off_stride, vk, kk);
  Vk[lane] = vk;
  K[lane] = kk;
  ok[lane] = (r && lg.pre_ok(lane)) ? 1 : 0;
  PSB_BALLAST(N, ok);
}
__global__ void PSB_PROTO_BOUNDS k_vid_g1(size_t N, int wb, const G1A* tblB, const G1J* phi, const G1J* E1,
                                                    const G1J* E2, const Fr* c, const Fr* rs, LaneGeom lg, int with_id,
                                                    G1J* V /*3 per lane*/) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J a, b, d;
  const TblGeom tb{wb};
  verify_id_g1_lane(tb, tblB, phi[lane], with_id ? E1 + lane : nullptr, with_id ? E2 + lane : nullptr, c + lane,
                    rs + lane * lg.rs_stride, lg.count(lane), with_id, a, b, d);
  V[3 * lane] = a; V[3 * lane + 1] = b; V[3 * lane + 2] = d;
  PSB_BALLAST(N, V);
}
__global__ void PSB_PROTO_BOUNDS k_vid_hash(size_t N, const G2J* k, const G1J* phi, const G1J* E1, const G1J* E2,
                                                      const G2J* Vk, const G1J* V, int with_id, const Fr* c,
                                                      const uint8_t* ad_blob, const uint64_t* ad_off, uint8_t* ok) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  if (!ok[lane]) return;
  const bool r = verify_id_hash_lane(k[lane], phi[lane], with_id ? E1 + lane : nullptr, with_id ? E2 + lane : nullptr,
                                     Vk[lane], V[3 * lane], V[3 * lane + 1], V[3 * lane + 2], with_id, c + lane,
                                     ad_blob + ad_off[lane], (size_t)(ad_off[lane + 1] - ad_off[lane]));
  ok[lane] = r ? 1 : 0;
}

// ---- wire-format ingest (csrc/wire.cuh) ------------------------------------------------------------
// lane j's text is in[off[j] .. off[j+1]); the decoded bytes go to out + off[j] (never longer than the text)
__global__ void __launch_bounds__(kBlock) k_wire_base64(size_t N, const uint8_t* in, const uint64_t* off, uint8_t* out, uint32_t* out_len) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  out_len[lane] = (uint32_t)base64_decode_lane(out + off[lane], in + off[lane], (size_t)(off[lane + 1] - off[lane]));
}
// kind 0: IdProof, 1: PSCredRequest.  len (optional) = decoded lengths after k_wire_base64.  Outputs per lane: pos[W_SLOTS],
// c, rs[n + 2], per, the attribute strings copied to attr_blob + off[j] with attr_off[j * (n + 1) ..] absolute offsets,
// parsed (structure ok) and has_e.
__global__ void __launch_bounds__(kBlock) k_wire_parse(size_t N, int n, int kind, const uint8_t* buf, const uint64_t* off,
                                                        const uint32_t* len, uint32_t* pos, Fr* c, Fr* rs, int* per,
                                                        uint8_t* attr_blob, uint64_t* attr_off, uint8_t* parsed, uint8_t* has_e) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  const size_t l = len ? len[lane] : (size_t)(off[lane + 1] - off[lane]);
  Fr cc;
  for (int i = 0; i < 8; i++) cc.v[i] = 0;
  int p = 0;
  bool he = false;
  const bool ok = kind == 0
      ? parse_idproof_lane(buf + off[lane], l, n, pos + lane * W_SLOTS, cc, rs + lane * (n + 2), p, attr_blob + off[lane],
                           attr_off + lane * (n + 1), off[lane], he)
      : parse_request_lane(buf + off[lane], l, n, pos + lane * W_SLOTS, cc, rs + lane * (n + 2), p, attr_blob + off[lane],
                           attr_off + lane * (n + 1), off[lane]);
  c[lane] = cc;
  per[lane] = p;
  parsed[lane] = ok ? 1 : 0;
  has_e[lane] = he ? 1 : 0;
}
// decompression of every (lane, slot) pair, one thread each, slot-major so that a warp works on one kind of point:
// slots 0..4 = G1 (sig1 / A, sig2, phi, E1, E2) into g1out[slot][lane], slot 5 = G2 (k).  An absent slot gives the zero point
// and does not count as a failure; a payload mcl's deserialize would reject clears parsed[lane] (atomic AND through a byte store
// of 0: every writer writes the same value).
__global__ void __launch_bounds__(kBlock) k_and_flags(size_t N, uint8_t* a, const uint8_t* b) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane < N) a[lane] = (uint8_t)(a[lane] & b[lane]);
}
struct WireG1Out { G1J* p[5]; };
__global__ void __launch_bounds__(kBlock) k_wire_points(size_t N, int nslots, const uint8_t* buf, const uint64_t* off, const uint32_t* pos,
                                                         WireG1Out g1out, G2J* kout, uint8_t* parsed) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * (size_t)nslots) return;
  const int slot = (int)(t / N);
  const size_t lane = t % N;
  const uint32_t at = pos[lane * W_SLOTS + slot];
  bool ok = true;
  if (slot == W_K) {
    G2J P;
    pt_set_zero(P);
    if (at != kWireAbsent) ok = g2_deserialize(P, buf + off[lane] + at);
    kout[lane] = P;
  } else {
    G1J P;
    pt_set_zero(P);
    if (at != kWireAbsent) ok = g1_deserialize(P, buf + off[lane] + at);
    g1out.p[slot][lane] = P;
  }
  if (!ok) parsed[lane] = 0;
}

// ---- wire-format OUTPUT (csrc/wire.cuh encode_message_lane): what a batched prover sends.  One thread per lane writes its
// message to out + out_off[lane]; out_off comes from the host (wire_message_size).  pts.p[3] / p[4] null = no E1 / E2.
__global__ void PSB_PROTO_BOUNDS k_wire_encode(size_t N, int kind, int n, int per, WireG1Out pts, const G2J* k, const Fr* c, const Fr* rs,
                                               const uint8_t* attr, const uint64_t* attr_off, uint8_t* out, const uint64_t* out_off) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  const G1J* p[5];
  for (int i = 0; i < 5; i++) p[i] = pts.p[i] ? pts.p[i] + lane : nullptr;
  encode_message_lane(out + out_off[lane], kind, n, p, k ? k + lane : nullptr, c[lane], rs + lane * (size_t)per, per, attr,
                      attr_off + lane * (size_t)n);
}
__global__ void __launch_bounds__(kBlock) k_wire_base64_encode(size_t N, const uint8_t* in, const uint64_t* in_off, uint8_t* out,
                                                                const uint64_t* out_off) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  base64_encode_lane(out + out_off[lane], in + in_off[lane], (size_t)(in_off[lane + 1] - in_off[lane]));
}

// ---- prover side: PSRequester::el_passo_request_id / unblind_credential / el_passo_prove_id ----
// hide: n flags shared by the batch; rnd: `per` host-supplied scalars per lane in the reference's draw order (prover.cuh)
__global__ void PSB_PROTO_BOUNDS k_request_id(size_t N, int n, int w, const G1A* tblG1, const uint8_t* hide, int h,
                                                        const uint8_t* blob, const uint64_t* off, const uint8_t* ad_blob,
                                                        const uint64_t* ad_off, const Fr* rnd, G1J* A, Fr* c, Fr* rs) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J a;
  Fr cc;
  request_id_lane(n, TblGeom{w}, tblG1, hide, blob, off + lane * n, ad_blob + ad_off[lane],
                  (size_t)(ad_off[lane + 1] - ad_off[lane]), rnd + lane * (h + 2), a, cc, rs + lane * (h + 1));
  A[lane] = a;
  c[lane] = cc;
  PSB_BALLAST(N, rs);
}
__global__ void PSB_PROTO_BOUNDS k_unblind(size_t N, const G1J* sig1, const G1J* sig2, const Fr* t1, G1J* out2) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J r;
  unblind_lane(r, sig1[lane], sig2[lane], t1 + lane);
  out2[lane] = r;
  PSB_BALLAST(N, out2);
}
__global__ void PSB_PROTO_G2_BOUNDS k_pid_g2(size_t N, int n, int w, const G2A* tblYY, const G2A* tblAux, const G2J* XX,
                                                    const uint8_t* hide, int h, int with_id, const uint8_t* blob,
                                                    const uint64_t* off, const Fr* rnd, G2J* k, G2J* Vk) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G2J a, b;
  prove_id_g2_lane(n, TblGeom{w}, tblYY, tblAux, *XX, hide, blob, off + lane * n, rnd + lane * prove_rnd_per_lane(h, with_id), h,
                   with_id, a, b);
  k[lane] = a;
  Vk[lane] = b;
  PSB_BALLAST(N, Vk);
}
__global__ void PSB_PROTO_BOUNDS k_pid_g1(size_t N, int n, int wb, const G1A* tblB, const G1J* sig1, const G1J* sig2,
                                                    const uint8_t* blob, const uint64_t* off, const Fr* rnd, int h, int with_id,
                                                    G1J* o_sig1, G1J* o_sig2, G1J* W /*6 per lane*/) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J s1, s2, w6[6];
  prove_id_g1_lane(TblGeom{wb}, tblB, sig1[lane], sig2[lane], blob, off + lane * n, rnd + lane * prove_rnd_per_lane(h, with_id), h,
                   with_id, s1, s2, w6);
  o_sig1[lane] = s1;
  o_sig2[lane] = s2;
  for (int i = 0; i < 6; i++) W[6 * lane + i] = w6[i];
  PSB_BALLAST(N, W);
}
__global__ void PSB_PROTO_BOUNDS k_pid_hash(size_t N, int n, const uint8_t* hide, int h, int with_id, const uint8_t* blob,
                                                      const uint64_t* off, const uint8_t* ad_blob, const uint64_t* ad_off,
                                                      const Fr* rnd, G2J* k, const G2J* Vk, const G1J* W, G1J* phi, G1J* E1,
                                                      G1J* E2, Fr* c, Fr* rs) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G2J kk = k[lane];
  G1J w6[6];
  for (int i = 0; i < 6; i++) w6[i] = W[6 * lane + i];
  Fr cc;
  const int per = h + (with_id ? 2 : 1);
  prove_id_hash_lane(n, hide, blob, off + lane * n, ad_blob + ad_off[lane], (size_t)(ad_off[lane + 1] - ad_off[lane]),
                     rnd + lane * prove_rnd_per_lane(h, with_id), h, with_id, kk, Vk[lane], w6, cc, rs + lane * per);
  k[lane] = kk;
  phi[lane] = w6[0];
  if (with_id) { E1[lane] = w6[2]; E2[lane] = w6[3]; }
  c[lane] = cc;
}

// ---- batched hashAndMapToG1 (mcl bn.hpp:2088-2097) ---------------------------------------------
__global__ void __launch_bounds__(kBlock) k_hash_to_g1(size_t N, const uint8_t* blob, const uint64_t* off, G1J* out, uint8_t* ok) {
  const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= N) return;
  G1J P;
  const bool r = hash_and_map_to_g1(P, blob + off[lane], (size_t)(off[lane + 1] - off[lane]));
  out[lane] = P;
  ok[lane] = r ? 1 : 0;
}

}  // namespace psb
