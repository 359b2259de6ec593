#!/usr/bin/env python3
"""tools/proto_pipes.py -- PROTOTYPE / measurement only (not product code).

Do the integer ALU instructions (IADD3 / LOP3 / SHF: the "glue" around the multiplier) overlap with the multiplier's
IMAD.WIDE stream on a B200 SM, or do they add to it?  The loop body is 12 multiply-accumulates (carry-free
`mad.wide.u32` on 12 independent 64-bit accumulators, or two carry chains of six `mad.lo.cc / madc.hi.cc` = the CIOS
row of csrc/cios.cuh) plus K independent 32-bit ALU instructions, K = 0, 3, 6, 12, 24; one 512-thread block per SM.

    python tools/proto_pipes.py build   -> tools/_proto/proto_pipes
    tools/_proto/proto_pipes            -> (on a B200) clocks per warp and loop trip for every (form, K)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_proto")

SRC = r"""
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K, int FORM>   // FORM 0: carry-free mad.wide; 1: carry chains of six (IMAD.WIDE.U32.X)
__global__ void __launch_bounds__(512, 1) k_mix(uint32_t seed, int iters, uint32_t* out) {
  uint32_t a = seed + threadIdx.x, b[12], lo[12], hi[12], s[8];
  uint64_t acc[12];
#pragma unroll
  for (int j = 0; j < 12; j++) { b[j] = seed * (j + 3) + threadIdx.x; acc[j] = j; lo[j] = j; hi[j] = j + 1; }
#pragma unroll
  for (int j = 0; j < 8; j++) s[j] = seed + j;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (FORM == 0) {
#pragma unroll
      for (int j = 0; j < 12; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a), "r"(b[j]));
    } else {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        asm volatile(
            "mad.lo.cc.u32 %0, %12, %13, %0;\n\tmadc.hi.cc.u32 %1, %12, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %12, %14, %2;\n\tmadc.hi.cc.u32 %3, %12, %14, %3;\n\t"
            "madc.lo.cc.u32 %4, %12, %15, %4;\n\tmadc.hi.cc.u32 %5, %12, %15, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %16, %6;\n\tmadc.hi.cc.u32 %7, %12, %16, %7;\n\t"
            "madc.lo.cc.u32 %8, %12, %17, %8;\n\tmadc.hi.cc.u32 %9, %12, %17, %9;\n\t"
            "madc.lo.cc.u32 %10, %12, %18, %10;\n\tmadc.hi.u32 %11, %12, %18, %11;"
            : "+r"(lo[6 * h]), "+r"(hi[6 * h]), "+r"(lo[6 * h + 1]), "+r"(hi[6 * h + 1]), "+r"(lo[6 * h + 2]),
              "+r"(hi[6 * h + 2]), "+r"(lo[6 * h + 3]), "+r"(hi[6 * h + 3]), "+r"(lo[6 * h + 4]), "+r"(hi[6 * h + 4]),
              "+r"(lo[6 * h + 5]), "+r"(hi[6 * h + 5])
            : "r"(a), "r"(b[6 * h]), "r"(b[6 * h + 1]), "r"(b[6 * h + 2]), "r"(b[6 * h + 3]), "r"(b[6 * h + 4]),
              "r"(b[6 * h + 5]));
      }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (k % 3 == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[k % 8]) : "r"(a));
      else if (k % 3 == 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(s[k % 8]) : "r"(a));
      else asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(s[k % 8]) : "r"(a));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int j = 0; j < 12; j++) r ^= (uint32_t)acc[j] ^ (uint32_t)(acc[j] >> 32) ^ lo[j] ^ hi[j];
#pragma unroll
  for (int j = 0; j < 8; j++) r ^= s[j];
  if (r == 0x12345u) out[0] = r;
}

template <int K, int FORM> void run(uint32_t* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_mix<K, FORM><<<148, 512>>>(7, 100, d);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_mix<K, FORM><<<148, 512>>>(7, iters, d);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int dev = 0, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  // 4 warps per scheduler: scheduler clocks per trip of ONE warp = ms * clk / iters / 4
  const double clk = ms * 1e-3 * khz * 1e3 / iters / 4.0;
  printf("{\"form\": \"%s\", \"mac\": 12, \"alu\": %d, \"ms\": %.3f, \"clk_per_warp_trip\": %.2f}\n",
         FORM ? "carry chains of six (IMAD.WIDE.U32.X)" : "carry-free mad.wide.u32", K, ms, clk);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4);
  const int iters = 200000;
  run<0, 0>(d, iters); run<3, 0>(d, iters); run<6, 0>(d, iters); run<12, 0>(d, iters); run<24, 0>(d, iters);
  run<0, 1>(d, iters); run<3, 1>(d, iters); run<6, 1>(d, iters); run<12, 1>(d, iters); run<24, 1>(d, iters);
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
"""


def build():
    os.makedirs(OUT, exist_ok=True)
    cu = os.path.join(OUT, "proto_pipes.cu")
    open(cu, "w").write(SRC)
    exe = os.path.join(OUT, "proto_pipes")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                           "-std=c++17", cu, "-o", exe])
    return exe


if __name__ == "__main__":
    if sys.argv[1:2] == ["build"]:
        print(build())
    else:
        print(__doc__)
