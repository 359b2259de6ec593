// tests/hostsim/hostsim.cc -- TEST INFRASTRUCTURE ONLY.
// Compiles the engine's __host__ __device__ arithmetic headers for the CPU (portable 64-bit
// arithmetic replaces the PTX carry chains) so that the CPU test-suite (`pytest -m "not gpu"`)
// can check the tower / curve / pairing LOGIC against the oracle without a GPU.  This library is
// never linked into, loaded by, or used as a fallback for libpsb.so.
#include <stddef.h>
#include "../../ps-signature-and-el-passo_b200/csrc/testops.cuh"

extern "C" {
int hostsim_op_shape(int op, int* s) { return psb::test_op_shape(op, s) ? 0 : -1; }
int hostsim_op(int op, size_t n, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out) {
  int s[4];
  if (!psb::test_op_shape(op, s)) return -1;
  for (size_t i = 0; i < n; i++)
    psb::test_op_run(op, a + i * s[0], b ? b + i * s[1] : nullptr, c ? c + i * s[2] : nullptr, out + i * s[3]);
  return 0;
}
void hostsim_set_hash_of(const uint8_t* msg, size_t len, uint32_t* k) { psb::fr_set_hash_of(k, msg, len); }
}
