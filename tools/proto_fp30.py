#!/usr/bin/env python3
"""tools/proto_fp30.py -- PROTOTYPE / measurement only (not product code, not on any product path).

DESIGN.md "what still bounds it", candidate (i): a carry-free radix-2^30 Montgomery multiplier for the 381-bit
field -- 13 limbs of 30 bits in 32-bit registers, 64-bit column accumulators fed by carry-free `mad.wide.u32`
(2 x 169 + 13 per product) instead of 300 carry-chained IMAD.WIDE.U32.X, a carry sweep every few rows.  This script
writes a self-contained CUDA program (constants embedded), compiles it for sm_100a and

    python tools/proto_fp30.py build     -> tools/_proto/proto_fp30 (binary, git-ignored via build dir name)
    tools/_proto/proto_fp30 [iters]      -> (on a B200) dependent-chain throughput in G products/s at several
                                            launch shapes + the limbs of one result for the check below
    python tools/proto_fp30.py check "<hex limbs printed by the binary>" iters
                                         -> recomputes the same chain with Python integers

The engine's own multiplier, same kind of chain: tools/microbench.py "fp_mul" (28.6 G/s on B200 = the carry-chain
ceiling, profiles/r1s_microbench.json)."""
import os
import subprocess
import sys

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
W, L = 30, 13
MASK = (1 << W) - 1
R = 1 << (W * L)
PINV = (-pow(P, -1, 1 << W)) % (1 << W)
X0 = 0x0123456789abcdef0fedcba9876543210123456789abcdef0fedcba9876543210123456789abcdef0fedcba98765432 % P
Y0 = 0x0fedcba9876543210123456789abcdef0fedcba9876543210123456789abcdef0fedcba9876543210123456789abcde % P
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_proto")


def limbs(v):
    return [(v >> (W * i)) & MASK for i in range(L)]


SRC = r"""
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int L = 13;
constexpr uint32_t MASK = (1u << 30) - 1u;
__constant__ uint32_t cP[L] = {%(P)s};
constexpr uint32_t PINV = %(PINV)du;
static const uint32_t hX[L] = {%(X)s};
static const uint32_t hY[L] = {%(Y)s};

// r = a * b / 2^390 mod p, canonical limbs (< 2^30), a, b < p
__device__ __forceinline__ void mont30(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint64_t t[L + 1];
#pragma unroll
  for (int j = 0; j <= L; j++) t[j] = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const uint32_t bi = b[i];
#pragma unroll
    for (int j = 0; j < L; j++) t[j] += (uint64_t)a[j] * bi;
    const uint32_t q = ((uint32_t)t[0] * PINV) & MASK;
#pragma unroll
    for (int j = 0; j < L; j++) t[j] += (uint64_t)q * cP[j];
    const uint64_t c = t[0] >> 30;          // t[0] = 0 mod 2^30
    t[0] = t[1] + c;
#pragma unroll
    for (int j = 1; j < L - 1; j++) t[j] = t[j + 1];
    t[L - 1] = 0;
    if (i == 4 || i == 8) {                 // <= 10 products of < 2^60 + carries per column between sweeps
#pragma unroll
      for (int j = 0; j < L - 1; j++) { const uint64_t k = t[j] >> 30; t[j] &= MASK; t[j + 1] += k; }
    }
  }
#pragma unroll
  for (int j = 0; j < L - 1; j++) { const uint64_t k = t[j] >> 30; t[j] &= MASK; t[j + 1] += k; }
  // t < 2p: conditional subtraction in radix 2^30
  uint32_t d[L];
  int32_t bw = 0;
#pragma unroll
  for (int j = 0; j < L; j++) {
    const int32_t v = (int32_t)(uint32_t)t[j] - (int32_t)cP[j] + bw;
    d[j] = (uint32_t)v & MASK;
    bw = v >> 30;                           // 0 or -1
  }
#pragma unroll
  for (int j = 0; j < L; j++) r[j] = bw ? (uint32_t)t[j] : d[j];
}

__global__ void __launch_bounds__(512, 1) k_chain(const uint32_t* x0, const uint32_t* y0, int iters, uint32_t* out) {
  uint32_t x[L], y[L];
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int j = 0; j < L; j++) { x[j] = x0[j]; y[j] = y0[j]; }
  x[0] = (x[0] + (tid & 1023u)) & MASK;     // lanes differ a little (still < p: top limbs decide)
#pragma unroll 1
  for (int it = 0; it < iters; it++) mont30(x, x, y);
  if (out) for (int j = 0; j < L; j++) out[(size_t)tid * L + j] = x[j];
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  uint32_t *dx, *dy, *dout;
  cudaMalloc(&dx, sizeof(hX)); cudaMalloc(&dy, sizeof(hY));
  cudaMemcpy(dx, hX, sizeof(hX), cudaMemcpyHostToDevice); cudaMemcpy(dy, hY, sizeof(hY), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, (size_t)148 * 8 * 512 * L * 4);
  // correctness sample: 7 products, thread 0 and thread 5
  k_chain<<<1, 32>>>(dx, dy, 7, dout);
  uint32_t h[32 * L];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  for (int t : {0, 5}) { printf("check tid=%%d iters=7:", t); for (int j = 0; j < L; j++) printf(" %%08x", h[t * L + j]); printf("\n"); }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int shapes[][2] = {{148, 128}, {148, 256}, {148, 512}, {296, 256}, {592, 128}};
  for (auto& s : shapes) {
    k_chain<<<s[0], s[1]>>>(dx, dy, 50, nullptr);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_chain<<<s[0], s[1]>>>(dx, dy, iters, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"probe\": \"fp30_mul chain\", \"blocks\": %%d, \"threads\": %%d, \"iters\": %%d, \"ms\": %%.3f, \"Gops_per_s\": %%.2f}\n",
           s[0], s[1], iters, ms, (double)s[0] * s[1] * iters / ms / 1e6);
  }
  printf("cuda: %%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
"""


def build():
    os.makedirs(OUT, exist_ok=True)
    cu = os.path.join(OUT, "proto_fp30.cu")
    fmt = lambda v: ", ".join("0x%08xu" % x for x in limbs(v))
    open(cu, "w").write(SRC % {"P": fmt(P), "PINV": PINV, "X": fmt(X0), "Y": fmt(Y0)})
    exe = os.path.join(OUT, "proto_fp30")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                           "-std=c++17", "-Xptxas", "-v", cu, "-o", exe])
    return exe


def expected(tid, iters):
    x = (X0 & ~MASK) | ((X0 + (tid & 1023)) & MASK)
    rinv = pow(R, -1, P)
    for _ in range(iters):
        x = x * Y0 * rinv % P
    return " ".join("%08x" % v for v in limbs(x))


if __name__ == "__main__":
    if sys.argv[1:2] == ["build"]:
        print(build())
    elif sys.argv[1:2] == ["expected"]:
        for t in (0, 5):
            print("check tid=%d iters=7: %s" % (t, expected(t, 7)))
    else:
        print(__doc__)
