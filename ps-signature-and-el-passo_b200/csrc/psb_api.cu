// psb_api.cu -- the extern "C" layer of libpsb.so (declared in include/psb.h).
//
// Host side of the batch engine: device list, per-key tables in HBM, lane sharding over the GPUs of
// one box (one host worker thread + one stream per device, no inter-device communication, verdict
// bytes written into disjoint slices of the caller's array -- SURVEY.md 8e), and kernel launches.
// There is deliberately NO CPU implementation behind these entry points.
#include <cuda_runtime.h>

#include <array>
#include <algorithm>
#include <atomic>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/psb.h"
#include "kernels.cuh"

using namespace psb;

namespace {

// sizes of mcl's objects in 64-bit words and of their serialisations in bytes (BLS12-381: 6/18/36/72 words, 48/96
// bytes; BN254: 4/12/24/48 words, 32/64 bytes)
constexpr size_t kFpW = PSB_NL / 2;       // Fp
constexpr size_t kG1W = 3 * kFpW;         // G1 Jacobian
constexpr size_t kG2W = 6 * kFpW;         // G2 Jacobian
constexpr size_t kGtW = 12 * kFpW;        // GT
constexpr size_t kG1Ser = 4 * PSB_NL;     // compressed G1
constexpr size_t kG2Ser = 8 * PSB_NL;     // compressed G2
constexpr size_t kCredSer = 2 * kG1Ser;   // sig1 || sig2

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};
struct Dev {
  int ordinal = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy = nullptr;           // host->device staging of the NEXT chunk while the current one computes (psb_verify)
  cudaEvent_t ev_in[8] = {};             // "inputs of chunk c are on the device"
  DevBuf in[8];   // staging for host-pointer entry points
  DevBuf ws;      // phase hand-over scratch
  DevBuf arena;   // staging + scratch of the EL PASSO entry points (carved by Arena)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // phase boundaries of the last verify (profiling)
  std::mutex mu;  // one batch at a time per device
};

std::vector<Dev*> g_devs;
bool g_init = false;
uint64_t g_epoch = 0;                  // bumped by every psb_init / psb_shutdown
std::mutex g_keys_mu;
std::vector<psb_key*> g_keys;          // live keys of the current context: psb_shutdown releases their device memory
thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
bool g_profile = false;
// fixed-base sums of psb_verify with batched affine pair additions (curve.cuh, AffBatch); PSB_MSM_AFFINE=0 keeps the plain
// mixed-addition chain (A/B switch, read once per process)
// (2, default: the pair sums are paired up once more before they reach the Jacobian accumulator; 1: one level)
int g_msm_affine = [] { const char* e = getenv("PSB_MSM_AFFINE"); return e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 2; }();

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  else snprintf(buf, sizeof(buf), "%s", what);
  g_err = buf;
  return code;
}
inline bool key_dead(const psb_key* key);
#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) return fail(PSB_ERR_CUDA, #call, e_);        \
  } while (0)
#define LAUNCHED() g_launches.fetch_add(1, std::memory_order_relaxed)

inline unsigned nblocks(size_t n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }
// Launch shape of the pairing-pipeline kernels: ONE 512-thread block per SM and wave (shared instruction fetches,
// kernels.cuh) for the whole waves of a batch, then the ragged rest as ONE thinner block per SM in a second launch.
// A wave lasts as long as its fullest scheduler has warps (measured: 14- and 16-warp blocks take the same time, the
// phase kernels run alike from 2 warps per scheduler up), so a rest of full 512-thread blocks on part of the SMs costs
// a whole wave, the same lanes spread over all SMs only their share of one: 2^18 lanes = 3.46 waves cost 3.5, not 4.
int g_sms = 148;   // SMs of device 0 (psb_init)
template <class Launch> inline void for_waves(size_t n, Launch&& launch) {   // launch(first lane, end lane, block size)
  const size_t wave = (size_t)g_sms * kPairBlock;
  const size_t full = n / wave * wave;
  if (full) launch((size_t)0, full, kPairBlock);
  if (n > full) {
    const size_t per_sm = (n - full + g_sms - 1) / g_sms;
    const size_t b = std::min<size_t>(std::max<size_t>((per_sm + 31) / 32 * 32, 32), kPairBlock);
    launch(full, n, (int)b);
  }
}
#define PSB_WAVES(n, kernel, ...)                                                              \
  for_waves((n), [&](size_t wb_, size_t we_, int blk_) {                                       \
    kernel<<<nblocks(we_ - wb_, blk_), blk_, 0, st>>>(we_, wb_, __VA_ARGS__);        \
    LAUNCHED();                                                                                \
  })

// Single-kernel entry points (randomisation, issuance) as a two-stream pipeline: the lanes of a device are cut into a few
// chunks, chunk c runs copy-in -> kernel -> copy-out in order on stream c & 1, so the copies of one chunk overlap the kernel
// of the other and the kernels themselves overlap at their ragged ends (no wave is lost at a chunk boundary).  The device
// buffers hold all lanes; the chunks are disjoint slices.  Small batches stay one chunk.
constexpr size_t kPipeMinLanes = 1 << 15;
inline size_t pipe_chunk(size_t L) { return L < 2 * kPipeMinLanes ? L : (L + 3) / 4; }

int ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return PSB_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.cap = 0;
  size_t cap = bytes + (bytes >> 3) + 256;
  cudaError_t e = cudaMalloc(&b.p, cap);
  if (e != cudaSuccess) return fail(PSB_ERR_NOMEM, "cudaMalloc", e);
  b.cap = cap;
  return PSB_OK;
}

struct KeyDev {
  G1J* g1pts = nullptr;   // [0] g, [1] X (or zero), [2..2+n) Y_i         (normalised)
  G2J* g2pts = nullptr;   // [0] gg, [1] XX, [2..2+n) YY_i                (normalised)
  G2A* tblYY = nullptr;   // n * nwin * 2^(w-1) affine entries
  G2A* tblAux = nullptr;  // [gg, XX] tables (el_passo_verify_id; built on first use)
  G1A* tblG1 = nullptr;   // [g, Y_0 .. Y_{n-1}] tables (el_passo_provide_id; built on first use)
  FixedLine* lines = nullptr;
};

// window tables of the per-batch G1 bases of el_passo_verify_id: H(service), g, y, h (one value per
// batch).  Cached per key by the points' bytes so that a relying party's steady stream of batches builds them once.
constexpr int kBatchW = 12;
struct BatchTbl {
  std::array<uint64_t, 4 * kG1W> id{};
  int nbases = 0;
  std::vector<int> ordinals;
  std::vector<G1A*> tbl;   // per device
  ~BatchTbl() {
    for (size_t i = 0; i < tbl.size(); i++) if (tbl[i]) { cudaSetDevice(ordinals[i]); cudaFree(tbl[i]); }
  }
};

}  // namespace

struct psb_key {
  size_t n = 0;
  int w = 16;
  bool hasX = false;
  size_t table_bytes = 0;
  std::vector<KeyDev> d;                           // one entry per device of the context the key was created in
  std::vector<int> ordinals;                       // CUDA ordinal of d[i] (the key does not index g_devs)
  uint64_t epoch = 0;                              // psb_init generation; a key of an earlier context is dead
  std::mutex mu;                                   // lazy tables + batch-table cache
  bool haveG1 = false, haveAux = false;
  std::vector<std::shared_ptr<BatchTbl>> batch;    // FIFO, at most kBatchCache entries
};

namespace {
inline bool key_dead(const psb_key* key) { return key->epoch != g_epoch || key->d.size() != g_devs.size(); }
#define KEY_ALIVE(key) do { if (key_dead(key)) return fail(PSB_ERR_ARG, "the key belongs to a context that was shut down (psb_shutdown / psb_init)"); } while (0)
// device memory of a key; the secret X = g^x (slot 1 of g1pts, and its window-table base) is overwritten before the free
void key_release_device(psb_key* key) {
  for (size_t i = 0; i < key->d.size(); i++) {
    if (cudaSetDevice(key->ordinals[i]) != cudaSuccess) continue;
    KeyDev& k = key->d[i];
    if (key->hasX && k.g1pts) cudaMemset(k.g1pts + 1, 0, sizeof(G1J));
    cudaFree(k.g1pts); cudaFree(k.g2pts); cudaFree(k.tblYY); cudaFree(k.tblAux); cudaFree(k.tblG1);
    cudaFree(k.lines);
    k = KeyDev();
  }
  key->batch.clear();
  key->d.clear();
}

// run f(dev_index, lane_begin, lane_end) on every device over a contiguous split of [0, N)
template <class Fn>
int shard(size_t N, Fn fn) {
  const int G = (int)g_devs.size();
  if (G == 1 || N < 2 * (size_t)G) return fn(0, (size_t)0, N);
  std::vector<int> rc(G, 0);
  std::vector<std::string> errs(G);
  std::vector<std::thread> th;
  for (int k = 0; k < G; k++) {
    size_t b, e;
    psb_shard_range(N, G, k, &b, &e);
    th.emplace_back([&, k, b, e]() { rc[k] = fn(k, b, e); if (rc[k]) errs[k] = g_err; });
  }
  for (auto& t : th) t.join();
  for (int k = 0; k < G; k++) if (rc[k]) { g_err = errs[k]; return rc[k]; }
  return PSB_OK;
}

constexpr size_t kBatchCache = 8;

// fixed-base window tables of `nbases` NORMALISED device points: tbl[(b * nwin + j) * 2^(w-1) + d - 1] = d 2^(w j) B_b
template <class F>
int build_tables(cudaStream_t st, const Jac<F>* d_pts, int nbases, int w, Aff<F>** out) {
  const int nwin = fixed_nwin(w);
  const size_t half = (size_t)1 << (w - 1);
  const size_t entries = (size_t)nbases * nwin * half;
  Aff<F>* wb = nullptr;
  Aff<F>* tbl = nullptr;
  cudaError_t e = cudaMalloc(&wb, ((size_t)nbases * nwin + 1) * sizeof(Aff<F>));
  if (e == cudaSuccess) e = cudaMalloc(&tbl, (entries + 1) * sizeof(Aff<F>));
  if (e != cudaSuccess) { cudaFree(wb); return fail(PSB_ERR_NOMEM, "table allocation", e); }
  k_window_bases<F><<<nblocks(nbases, 32), 32, 0, st>>>(d_pts, nbases, w, wb); LAUNCHED();
  const size_t chunks = (size_t)nbases * nwin * ((half + kTblChunk - 1) / kTblChunk);
  k_build_table<F><<<nblocks(chunks), kBlock, 0, st>>>(wb, nbases, w, tbl); LAUNCHED();
  e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(wb);
  if (e != cudaSuccess) { cudaFree(tbl); return fail(PSB_ERR_CUDA, "table build kernels", e); }
  *out = tbl;
  return PSB_OK;
}

// bump allocator over one device buffer: sizes are summed first (measure pass), then carved
struct Arena {
  char* base = nullptr;
  size_t used = 0;
  template <class T> T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + used) : nullptr;
    used += bytes;
    return p;
  }
};

bool is_zero_words(const uint64_t* p, size_t n) {
  uint64_t o = 0;
  for (size_t i = 0; i < n; i++) o |= p[i];
  return o == 0;
}

int verify_launch(const psb_key* key, int di, size_t N, const G1J* d_sig1, const G1J* d_sig2, const uint8_t* d_blob,
                  const uint64_t* d_off, const Fr* d_m, uint8_t* d_verdict, Fp12* d_gt, void* d_ws, cudaStream_t st,
                  const uint8_t* d_pre = nullptr, int ss = 1) {
  if (N == 0) return PSB_OK;
  const KeyDev& kd = key->d[di];
  G2J* dK = (G2J*)d_ws;
  Fp12* dF = (Fp12*)((char*)d_ws + N * sizeof(G2J));
  Dev* dv = g_devs[di];
  const bool prof = g_profile;
  if (prof) {
    for (auto& e : dv->ev) if (!e) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(dv->ev[0], st));
  }
  const int affine = g_msm_affine;   // (tables beyond 2^29 entries take the plain chain inside the kernel: AffBatch::plain)
  PSB_WAVES(N, k_verify_msm, (int)key->n, key->w, d_blob, d_off, d_m, kd.g2pts + 1, kd.tblYY, dK, affine);
  if (prof) CK(cudaEventRecord(dv->ev[1], st));
  PSB_WAVES(N, k_verify_miller, d_sig1, d_sig2, ss, dK, kd.lines, dF);
  if (prof) CK(cudaEventRecord(dv->ev[2], st));
  PSB_WAVES(N, k_verify_final, d_sig1, ss, dF, d_verdict, d_gt, d_pre, 1);
  if (prof) CK(cudaEventRecord(dv->ev[3], st));
  CK(cudaGetLastError());
  return PSB_OK;
}

}  // namespace

extern "C" {

const char* psb_last_error(void) { return g_err.c_str(); }
uint64_t psb_launch_count(void) { return g_launches.load(); }
int psb_num_devices(void) { return g_init ? (int)g_devs.size() : 0; }
int psb_shard_range(size_t N, int ndev, int k, size_t* begin, size_t* end) {
  if (ndev <= 0 || k < 0 || k >= ndev || !begin || !end) return fail(PSB_ERR_ARG, "bad shard index");
  // contiguous, balanced to within one lane, covers [0, N) exactly; 128-bit product: N can be 2^24 * ndev
  *begin = (size_t)((unsigned __int128)N * (unsigned)k / (unsigned)ndev);
  *end = (size_t)((unsigned __int128)N * (unsigned)(k + 1) / (unsigned)ndev);
  return PSB_OK;
}
void* psb_host_alloc(size_t bytes) {
  if (!g_init) { fail(PSB_ERR_NOT_INIT, "psb_init not called"); return nullptr; }
  void* p = nullptr;
  const cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) { fail(PSB_ERR_NOMEM, "cudaHostAlloc", e); return nullptr; }
  return p;
}
void psb_host_free(void* p) { if (p) cudaFreeHost(p); }
int psb_set_profiling(int on) { g_profile = on != 0; return PSB_OK; }
int psb_last_phase_ms(int dev_index, float* ms) {
  if (!g_init || dev_index < 0 || dev_index >= (int)g_devs.size() || !ms) return fail(PSB_ERR_ARG, "bad argument");
  Dev* dv = g_devs[dev_index];
  if (!dv->ev[3]) return fail(PSB_ERR_ARG, "no profiled verify yet");
  CK(cudaSetDevice(dv->ordinal));
  CK(cudaEventSynchronize(dv->ev[3]));
  for (int i = 0; i < 3; i++) CK(cudaEventElapsedTime(&ms[i], dv->ev[i], dv->ev[i + 1]));
  return PSB_OK;
}

int psb_init(int curve, const int* devices, int ndev) {
#if PSB_IS_BN
  if (curve != PSB_MCL_CURVE) return fail(PSB_ERR_UNSUPPORTED, "this library (libpsb_bn254.so) is built for BN254 (curve 0); BLS12-381 is libpsb.so");
#else
  if (curve != PSB_MCL_CURVE) return fail(PSB_ERR_UNSUPPORTED, "this library (libpsb.so) is built for BLS12-381 (curve 5); BN254 is libpsb_bn254.so");
#endif
  psb_shutdown();      // also clears a half-built device list of an earlier failed psb_init
  struct Guard { bool ok = false; ~Guard() { if (!ok) psb_shutdown(); } } guard;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(PSB_ERR_CUDA, "no CUDA device (this library has no CPU path)", e);
  std::vector<int> ords;
  if (devices && ndev > 0) ords.assign(devices, devices + ndev); else ords.push_back(0);
  for (int o : ords) {
    if (o < 0 || o >= count) return fail(PSB_ERR_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, o));
    if (prop.major < 10) return fail(PSB_ERR_UNSUPPORTED, "kernels are built for sm_100a only");
    if (g_devs.empty()) {
      g_sms = prop.multiProcessorCount;
      // probe only (tools/probe_l2_fit.py): shape the launches as if the device had fewer SMs, so that a batch of
      // PSB_SMS x 512 lanes runs one full block on that many SMs and its thread-local state fits the L2
      if (const char* e = getenv("PSB_SMS")) { const int v = atoi(e); if (v > 0 && v <= g_sms) g_sms = v; }
    }
    Dev* d = new Dev();
    d->ordinal = o;
    g_devs.push_back(d);                 // owned by the list from here on: the guard frees it if a later step fails
    CK(cudaSetDevice(o));
    CK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d->copy, cudaStreamNonBlocking));
    for (auto& e : d->ev_in) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // thread-local state (Fp12 temporaries, window tables of points) lives in local memory:
    // prefer L1 over shared memory, and give deep call chains enough stack
    cudaDeviceSetLimit(cudaLimitStackSize, 32 * 1024);
    {
      const int l1 = cudaSharedmemCarveoutMaxL1;   // the kernels use no shared memory (measured: the driver picks this anyway)
      cudaFuncSetAttribute(k_verify_msm, cudaFuncAttributePreferredSharedMemoryCarveout, l1);
      cudaFuncSetAttribute(k_verify_miller, cudaFuncAttributePreferredSharedMemoryCarveout, l1);
      cudaFuncSetAttribute(k_verify_final, cudaFuncAttributePreferredSharedMemoryCarveout, l1);
      cudaFuncSetAttribute(k_pairing_miller, cudaFuncAttributePreferredSharedMemoryCarveout, l1);
      cudaFuncSetAttribute(k_final_exp, cudaFuncAttributePreferredSharedMemoryCarveout, l1);
    }
  }
  g_init = true;
  guard.ok = true;
  return PSB_OK;
}

void psb_shutdown(void) {
  {
    // keys outlive the context only as dead handles: their device memory goes with the context, every later call with
    // such a key fails with PSB_ERR_ARG, psb_key_destroy only frees the host object
    std::lock_guard<std::mutex> lk(g_keys_mu);
    for (psb_key* k : g_keys) key_release_device(k);
    g_keys.clear();
    g_epoch++;
  }
  for (Dev* d : g_devs) {
    cudaSetDevice(d->ordinal);
    for (auto& b : d->in) if (b.p) cudaFree(b.p);
    if (d->ws.p) cudaFree(d->ws.p);
    if (d->arena.p) cudaFree(d->arena.p);
    for (auto& e : d->ev) if (e) cudaEventDestroy(e);
    for (auto& e : d->ev_in) if (e) cudaEventDestroy(e);
    if (d->copy) cudaStreamDestroy(d->copy);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
  }
  g_devs.clear();
  g_init = false;
}

size_t psb_key_num_attributes(const psb_key* key) { return key ? key->n : 0; }
size_t psb_key_table_bytes(const psb_key* key) { return key ? key->table_bytes : 0; }

void psb_key_destroy(psb_key* key) {
  if (!key) return;
  {
    std::lock_guard<std::mutex> lk(g_keys_mu);
    auto it = std::find(g_keys.begin(), g_keys.end(), key);
    if (it != g_keys.end()) { g_keys.erase(it); key_release_device(key); }   // else: psb_shutdown already released it
  }
  delete key;
}

psb_key* psb_key_create(const uint64_t* g, const uint64_t* gg, const uint64_t* XX, const uint64_t* Y,
                        const uint64_t* YY, size_t n, const uint64_t* X_secret, int window_bits) {
  if (!g_init) { fail(PSB_ERR_NOT_INIT, "psb_init not called"); return nullptr; }
  if (!g || !gg || !XX || (n && (!Y || !YY))) { fail(PSB_ERR_ARG, "null key component"); return nullptr; }
  int w = window_bits == 0 ? 16 : window_bits;
  if (w < 4 || w > 24) { fail(PSB_ERR_ARG, "window_bits must be 4..24"); return nullptr; }
  // fixed bases must be finite points: their window tables hold affine entries (z == 0 <=> infinity)
  bool inf = is_zero_words(g + 2 * kFpW, kFpW) || is_zero_words(gg + 4 * kFpW, 2 * kFpW) || is_zero_words(XX + 4 * kFpW, 2 * kFpW);
  for (size_t i = 0; i < n; i++) inf = inf || is_zero_words(Y + kG1W * i + 2 * kFpW, kFpW) || is_zero_words(YY + kG2W * i + 4 * kFpW, 2 * kFpW);
  if (inf) { fail(PSB_ERR_ARG, "key component is the point at infinity"); return nullptr; }
  psb_key* key = new psb_key();
  key->n = n; key->w = w; key->hasX = X_secret != nullptr;
  key->d.resize(g_devs.size());
  for (Dev* dv : g_devs) key->ordinals.push_back(dv->ordinal);
  key->epoch = g_epoch;
  { std::lock_guard<std::mutex> lk(g_keys_mu); g_keys.push_back(key); }
  const int nwin = fixed_nwin(w);
  const size_t half = (size_t)1 << (w - 1);
  key->table_bytes = n * nwin * half * sizeof(G2A);
  std::vector<G1J> h1(2 + n);
  std::vector<G2J> h2(2 + n);
  memcpy(&h1[0], g, sizeof(G1J));
  if (X_secret) memcpy(&h1[1], X_secret, sizeof(G1J)); else memset(&h1[1], 0, sizeof(G1J));
  if (n) memcpy(&h1[2], Y, n * sizeof(G1J));
  memcpy(&h2[0], gg, sizeof(G2J));
  memcpy(&h2[1], XX, sizeof(G2J));
  if (n) memcpy(&h2[2], YY, n * sizeof(G2J));
  for (size_t di = 0; di < g_devs.size(); di++) {
    Dev* dv = g_devs[di];
    KeyDev& k = key->d[di];
    cudaStream_t st = dv->stream;
    bool ok = cudaSetDevice(dv->ordinal) == cudaSuccess;
    ok = ok && cudaMalloc(&k.g1pts, h1.size() * sizeof(G1J)) == cudaSuccess;
    ok = ok && cudaMalloc(&k.g2pts, h2.size() * sizeof(G2J)) == cudaSuccess;
    ok = ok && cudaMalloc(&k.lines, kFixedLineSlots * sizeof(FixedLine)) == cudaSuccess;
    if (!ok) { fail(PSB_ERR_NOMEM, "key allocation", cudaGetLastError()); psb_key_destroy(key); return nullptr; }
    cudaMemcpyAsync(k.g1pts, h1.data(), h1.size() * sizeof(G1J), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(k.g2pts, h2.data(), h2.size() * sizeof(G2J), cudaMemcpyHostToDevice, st);
    k_normalize_points<Fp><<<nblocks(h1.size(), 32), 32, 0, st>>>(k.g1pts, (int)h1.size()); LAUNCHED();
    k_normalize_points<Fp2><<<nblocks(h2.size(), 32), 32, 0, st>>>(k.g2pts, (int)h2.size()); LAUNCHED();
    k_fixed_lines<<<1, 32, 0, st>>>(k.g2pts, k.lines); LAUNCHED();
    if (n && build_tables<Fp2>(st, k.g2pts + 2, (int)n, w, &k.tblYY) != PSB_OK) { psb_key_destroy(key); return nullptr; }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fail(PSB_ERR_CUDA, "key setup kernels", e); psb_key_destroy(key); return nullptr; }
  }
  if (X_secret) {      // the staging copy of the secret does not linger on the host heap
    volatile uint64_t* z = reinterpret_cast<volatile uint64_t*>(&h1[1]);
    for (size_t i = 0; i < kG1W; i++) z[i] = 0;
  }
  return key;
}

size_t psb_verify_ws_bytes(const psb_key* key, size_t N) {
  (void)key;
  return N * (sizeof(G2J) + sizeof(Fp12)) + 256;
}

int psb_verify_dev(psb_key* key, int dev_index, size_t N, const uint64_t* d_sig1, const uint64_t* d_sig2,
                   const uint8_t* d_attr_blob, const uint64_t* d_attr_off, const uint64_t* d_m,
                   uint8_t* d_verdict, uint64_t* d_gt, void* d_ws, void* stream) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || dev_index < 0 || dev_index >= (int)g_devs.size()) return fail(PSB_ERR_ARG, "bad key/device");
  KEY_ALIVE(key);
  if (!d_sig1 || !d_sig2 || !d_verdict || !d_ws || (!d_attr_blob && !d_m && key->n)) return fail(PSB_ERR_ARG, "null buffer");
  if (d_attr_blob && !d_attr_off) return fail(PSB_ERR_ARG, "attr_off missing");
  CK(cudaSetDevice(g_devs[dev_index]->ordinal));
  cudaStream_t st = stream ? (cudaStream_t)stream : g_devs[dev_index]->stream;
  return verify_launch(key, dev_index, N, (const G1J*)d_sig1, (const G1J*)d_sig2, d_attr_blob, d_attr_off,
                       (const Fr*)d_m, d_verdict, (Fp12*)d_gt, d_ws, st);
}

// sig2 == nullptr: sig1 is an array of (sigma1, sigma2) PAIRS (std::vector<PSCredential>::data()), copied as it lies
static int verify_host(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint8_t* attr_blob,
                       const uint64_t* attr_off, const uint64_t* m, uint8_t* verdict, uint64_t* gt);
int psb_verify(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint8_t* attr_blob,
               const uint64_t* attr_off, const uint64_t* m, uint8_t* verdict, uint64_t* gt) {
  if (!sig2) return fail(PSB_ERR_ARG, "null argument");
  return verify_host(key, N, sig1, sig2, attr_blob, attr_off, m, verdict, gt);
}
int psb_verify_aos(psb_key* key, size_t N, const uint64_t* cred, const uint8_t* attr_blob, const uint64_t* attr_off,
                   const uint64_t* m, uint8_t* verdict, uint64_t* gt) {
  return verify_host(key, N, cred, nullptr, attr_blob, attr_off, m, verdict, gt);
}
static int verify_host(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint8_t* attr_blob,
                       const uint64_t* attr_off, const uint64_t* m, uint8_t* verdict, uint64_t* gt) {
  const bool aos = sig2 == nullptr;
  if (aos) sig2 = sig1;   // (only for the null check below)
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !sig1 || !sig2 || !verdict) return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  if (key->n && !attr_blob && !m) return fail(PSB_ERR_ARG, "need attributes or scalars");
  if (attr_blob && !attr_off) return fail(PSB_ERR_ARG, "attr_off missing");
  const size_t n = key->n;
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    int rc;
    // Chunks of whole waves: the inputs of chunk c+1 travel on the copy stream while chunk c runs its three phase
    // kernels on the compute stream (one event per chunk); small batches are one chunk.
    const size_t wave = (size_t)g_sms * kPairBlock;
    size_t chunk = 4 * wave;
    if (L < 2 * chunk) chunk = L;
    const size_t nchunks = (L + chunk - 1) / chunk;
    if (nchunks > 8) chunk = (((L + 7) / 8 + wave - 1) / wave) * wave;
    if ((rc = ensure(dv->in[0], L * sizeof(G1J) * (aos ? 2 : 1)))) return rc;
    if (!aos && (rc = ensure(dv->in[1], L * sizeof(G1J)))) return rc;
    if ((rc = ensure(dv->in[4], L + 16))) return rc;
    if ((rc = ensure(dv->ws, psb_verify_ws_bytes(key, chunk)))) return rc;
    if (gt && (rc = ensure(dv->in[5], L * sizeof(Fp12)))) return rc;
    const uint8_t* d_blob = nullptr; const uint64_t* d_off = nullptr; const Fr* d_m = nullptr;
    uint64_t o0 = 0;
    if (attr_blob) {
      o0 = attr_off[b * n];
      const uint64_t o1 = attr_off[e * n];
      if ((rc = ensure(dv->in[2], (size_t)(o1 - o0) + 16))) return rc;
      if ((rc = ensure(dv->in[3], (L * n + 1) * sizeof(uint64_t)))) return rc;
      d_blob = (const uint8_t*)dv->in[2].p - o0;  // offsets stay absolute
      d_off = (const uint64_t*)dv->in[3].p;
    } else if (n) {
      if ((rc = ensure(dv->in[2], L * n * sizeof(Fr)))) return rc;
      d_m = (const Fr*)dv->in[2].p;
    }
    const int ss = aos ? 2 : 1;
    G1J* dS1 = (G1J*)dv->in[0].p; G1J* dS2 = aos ? dS1 + 1 : (G1J*)dv->in[1].p;
    int c = 0;
    for (size_t cb = 0; cb < L; cb += chunk, c++) {
      const size_t cl = std::min(chunk, L - cb);
      cudaStream_t cs = (c == 0) ? st : dv->copy;          // the first chunk has nothing to overlap with
      CK(cudaMemcpyAsync(dS1 + cb * ss, sig1 + (b + cb) * kG1W * ss, cl * sizeof(G1J) * ss, cudaMemcpyHostToDevice, cs));
      if (!aos) CK(cudaMemcpyAsync(dS2 + cb, sig2 + (b + cb) * kG1W, cl * sizeof(G1J), cudaMemcpyHostToDevice, cs));
      if (attr_blob) {
        const uint64_t c0 = attr_off[(b + cb) * n], c1 = attr_off[(b + cb + cl) * n];
        if (c1 > c0) CK(cudaMemcpyAsync((uint8_t*)dv->in[2].p + (c0 - o0), attr_blob + c0, (size_t)(c1 - c0), cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync((uint64_t*)dv->in[3].p + cb * n, attr_off + (b + cb) * n, (cl * n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
      } else if (n) {
        CK(cudaMemcpyAsync((Fr*)dv->in[2].p + cb * n, m + (b + cb) * n * 4, cl * n * sizeof(Fr), cudaMemcpyHostToDevice, cs));
      }
      if (c > 0) {
        CK(cudaEventRecord(dv->ev_in[c & 7], dv->copy));
        CK(cudaStreamWaitEvent(st, dv->ev_in[c & 7], 0));
      }
      rc = verify_launch(key, di, cl, dS1 + cb * ss, dS2 + cb * ss, d_blob, d_off ? d_off + cb * n : nullptr, d_m ? d_m + cb * n : nullptr,
                         (uint8_t*)dv->in[4].p + cb, gt ? (Fp12*)dv->in[5].p + cb : nullptr, dv->ws.p, st, nullptr, ss);
      if (rc) return rc;
    }
    CK(cudaMemcpyAsync(verdict + b, dv->in[4].p, L, cudaMemcpyDeviceToHost, st));
    if (gt) CK(cudaMemcpyAsync(gt + b * kGtW, dv->in[5].p, L * sizeof(Fp12), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(dv->copy));
    return PSB_OK;
  });
}

static int deserialize_impl(bool g2, size_t N, const uint8_t* ser, size_t stride, uint64_t* out, uint8_t* ok) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  const size_t esz = g2 ? kG2Ser : kG1Ser, words = g2 ? kG2W : kG1W;
  if (!ser || !out || !ok || stride < esz) return fail(PSB_ERR_ARG, "bad argument");
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const size_t span = (L - 1) * stride + esz;
    Arena ar;
    uint8_t *dser = nullptr, *dok = nullptr; uint64_t* dout = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dser = ar.take<uint8_t>(span); dok = ar.take<uint8_t>(L); dout = ar.take<uint64_t>(L * words);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    CK(cudaMemcpyAsync(dser, ser + b * stride, span, cudaMemcpyHostToDevice, st));
    if (g2) k_g2_deserialize<<<nblocks(L), kBlock, 0, st>>>(L, dser, stride, (G2J*)dout, dok, 0);
    else k_g1_deserialize<<<nblocks(L), kBlock, 0, st>>>(L, dser, stride, (G1J*)dout, dok, 0);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out + b * words, dout, L * words * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ok + b, dok, L, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}
int psb_g1_deserialize(size_t N, const uint8_t* ser, size_t stride, uint64_t* out, uint8_t* ok) {
  return deserialize_impl(false, N, ser, stride, out, ok);
}
int psb_g2_deserialize(size_t N, const uint8_t* ser, size_t stride, uint64_t* out, uint8_t* ok) {
  return deserialize_impl(true, N, ser, stride, out, ok);
}

int psb_verify_ser(psb_key* key, size_t N, const uint8_t* cred, size_t stride, size_t off1, size_t off2,
                   const uint8_t* attr_blob, const uint64_t* attr_off, uint8_t* verdict, uint8_t* decoded) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !cred || !verdict || (key->n && (!attr_blob || !attr_off))) return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  if (off1 + kG1Ser > stride || off2 + kG1Ser > stride) return fail(PSB_ERR_ARG, "offsets exceed the stride");
  const size_t n = key->n;
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t o0 = n ? attr_off[b * n] : 0, o1 = n ? attr_off[e * n] : 0;
    Arena ar;
    uint8_t *dcred = nullptr, *dblob = nullptr, *dok = nullptr, *dver = nullptr; uint64_t* doff = nullptr;
    G1J *dS1 = nullptr, *dS2 = nullptr; char* dws = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dcred = ar.take<uint8_t>(L * stride); dS1 = ar.take<G1J>(L); dS2 = ar.take<G1J>(L);
      dblob = ar.take<uint8_t>((size_t)(o1 - o0) + 16); doff = ar.take<uint64_t>(L * n + 1);
      dok = ar.take<uint8_t>(L); dver = ar.take<uint8_t>(L); dws = ar.take<char>(psb_verify_ws_bytes(key, L));
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    CK(cudaMemcpyAsync(dcred, cred + b * stride, L * stride, cudaMemcpyHostToDevice, st));
    if (o1 > o0) CK(cudaMemcpyAsync(dblob, attr_blob + o0, (size_t)(o1 - o0), cudaMemcpyHostToDevice, st));
    if (n) CK(cudaMemcpyAsync(doff, attr_off + b * n, (L * n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    k_g1_deserialize<<<nblocks(L), kBlock, 0, st>>>(L, dcred + off1, stride, dS1, dok, 0); LAUNCHED();
    k_g1_deserialize<<<nblocks(L), kBlock, 0, st>>>(L, dcred + off2, stride, dS2, dok, 1); LAUNCHED();
    int rc = verify_launch(key, di, L, dS1, dS2, dblob - o0, doff, nullptr, dver, nullptr, dws, st, dok);
    if (rc) return rc;
    CK(cudaMemcpyAsync(verdict + b, dver, L, cudaMemcpyDeviceToHost, st));
    if (decoded) CK(cudaMemcpyAsync(decoded + b, dok, L, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

int psb_pairing(size_t N, const uint64_t* P, const uint64_t* Q, uint64_t* out) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!P || !Q || !out) return fail(PSB_ERR_ARG, "null argument");
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    int rc;
    if ((rc = ensure(dv->in[0], L * sizeof(G1J)))) return rc;
    if ((rc = ensure(dv->in[1], L * sizeof(G2J)))) return rc;
    if ((rc = ensure(dv->in[5], L * sizeof(Fp12)))) return rc;
    if ((rc = ensure(dv->ws, L * sizeof(Fp12)))) return rc;
    CK(cudaMemcpyAsync(dv->in[0].p, P + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dv->in[1].p, Q + b * kG2W, L * sizeof(G2J), cudaMemcpyHostToDevice, st));
    PSB_WAVES(L, k_pairing_miller, (const G1J*)dv->in[0].p, (const G2J*)dv->in[1].p, (Fp12*)dv->ws.p);
    PSB_WAVES(L, k_final_exp, (const Fp12*)dv->ws.p, (Fp12*)dv->in[5].p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out + b * kGtW, dv->in[5].p, L * sizeof(Fp12), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

int psb_test_op_shape(int op, int shape[4]) { return test_op_shape(op, shape) ? PSB_OK : PSB_ERR_ARG; }

int psb_test_op(int op, size_t n, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  int s[4];
  if (!test_op_shape(op, s)) return fail(PSB_ERR_ARG, "unknown test op");
  if (!a || !out || (s[1] && !b) || (s[2] && !c)) return fail(PSB_ERR_ARG, "null operand");
  if (n == 0) return PSB_OK;
  Dev* dv = g_devs[0];
  std::lock_guard<std::mutex> lk(dv->mu);
  CK(cudaSetDevice(dv->ordinal));
  cudaStream_t st = dv->stream;
  int rc;
  const uint32_t* hp[3] = {a, b, c};
  for (int i = 0; i < 3; i++) {
    if (!s[i]) continue;
    if ((rc = ensure(dv->in[i], n * s[i] * 4))) return rc;
    CK(cudaMemcpyAsync(dv->in[i].p, hp[i], n * s[i] * 4, cudaMemcpyHostToDevice, st));
  }
  if ((rc = ensure(dv->in[3], n * s[3] * 4))) return rc;
  k_test_op<<<nblocks(n), kBlock, 0, st>>>(op, n, s[0], s[1], s[2], s[3], (const uint32_t*)dv->in[0].p,
                                           s[1] ? (const uint32_t*)dv->in[1].p : nullptr,
                                           s[2] ? (const uint32_t*)dv->in[2].p : nullptr, (uint32_t*)dv->in[3].p);
  LAUNCHED();
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dv->in[3].p, n * s[3] * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return PSB_OK;
}

double psb_microbench(int kind, int blocks, int threads, int iters) {
  if (!g_init) { fail(PSB_ERR_NOT_INIT, "psb_init not called"); return -1.0; }
  Dev* dv = g_devs[0];
  std::lock_guard<std::mutex> lk(dv->mu);
  if (cudaSetDevice(dv->ordinal) != cudaSuccess) return -1.0;
  cudaStream_t st = dv->stream;
  if (ensure(dv->in[0], 64 * 4) || ensure(dv->in[1], 64)) return -1.0;
  uint32_t seed[32];
  for (int i = 0; i < 32; i++) seed[i] = 0x9e3779b9u * (i + 1) + 0x7f4a7c15u;
  cudaMemcpyAsync(dv->in[0].p, seed, sizeof(seed), cudaMemcpyHostToDevice, st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = -1.f;
  for (int rep = 0; rep < 2; rep++) {  // first pass = warm-up
    cudaEventRecord(e0, st);
    if (kind <= 3) k_bench_fp<<<blocks, threads, 0, st>>>(kind, iters, (const uint32_t*)dv->in[0].p, (uint32_t*)dv->in[1].p);
    else k_bench_mad<<<blocks, threads, 0, st>>>(kind, iters, (const uint32_t*)dv->in[0].p, (uint32_t*)dv->in[1].p);
    LAUNCHED();
    cudaEventRecord(e1, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { ms = -1.f; break; }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return (double)ms;
}

int psb_randomize(size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint64_t* t, uint64_t* out1,
                  uint64_t* out2, uint8_t* ser) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!sig1 || !sig2 || !t || !out1 || !out2) return fail(PSB_ERR_ARG, "null argument");
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    int rc;
    for (int i = 0; i < 2; i++) if ((rc = ensure(dv->in[i], L * sizeof(G1J)))) return rc;
    if ((rc = ensure(dv->in[2], L * sizeof(Fr)))) return rc;
    for (int i = 3; i < 5; i++) if ((rc = ensure(dv->in[i], L * sizeof(G1J)))) return rc;
    if (ser && (rc = ensure(dv->in[5], L * kCredSer))) return rc;
    G1J *d1 = (G1J*)dv->in[0].p, *d2 = (G1J*)dv->in[1].p, *o1 = (G1J*)dv->in[3].p, *o2 = (G1J*)dv->in[4].p;
    Fr* dt = (Fr*)dv->in[2].p;
    uint8_t* dser = ser ? (uint8_t*)dv->in[5].p : nullptr;
    const size_t chunk = pipe_chunk(L);
    int c = 0;
    for (size_t cb = 0; cb < L; cb += chunk, c++) {
      const size_t cl = std::min(chunk, L - cb);
      cudaStream_t cs = (c & 1) ? dv->copy : st;
      CK(cudaMemcpyAsync(d1 + cb, sig1 + (b + cb) * kG1W, cl * sizeof(G1J), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(d2 + cb, sig2 + (b + cb) * kG1W, cl * sizeof(G1J), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(dt + cb, t + (b + cb) * 4, cl * sizeof(Fr), cudaMemcpyHostToDevice, cs));
      k_randomize<<<nblocks(cl), kBlock, 0, cs>>>(cl, d1 + cb, d2 + cb, dt + cb, o1 + cb, o2 + cb, dser ? dser + cb * kCredSer : nullptr);
      LAUNCHED();
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(out1 + (b + cb) * kG1W, o1 + cb, cl * sizeof(G1J), cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(out2 + (b + cb) * kG1W, o2 + cb, cl * sizeof(G1J), cudaMemcpyDeviceToHost, cs));
      if (ser) CK(cudaMemcpyAsync(ser + (b + cb) * kCredSer, dser + cb * kCredSer, cl * kCredSer, cudaMemcpyDeviceToHost, cs));
    }
    CK(cudaStreamSynchronize(st));
    if (c > 1) CK(cudaStreamSynchronize(dv->copy));
    return PSB_OK;
  });
}

int psb_g1_mul(size_t N, const uint64_t* P, int p_stride, const uint64_t* k, uint64_t* out) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!P || !k || !out) return fail(PSB_ERR_ARG, "null argument");
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    int rc;
    const size_t np = p_stride ? L : 1;
    if ((rc = ensure(dv->in[0], np * sizeof(G1J)))) return rc;
    if ((rc = ensure(dv->in[2], L * sizeof(Fr)))) return rc;
    if ((rc = ensure(dv->in[3], L * sizeof(G1J)))) return rc;
    CK(cudaMemcpyAsync(dv->in[0].p, P + (p_stride ? b * kG1W : 0), np * sizeof(G1J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dv->in[2].p, k + b * 4, L * sizeof(Fr), cudaMemcpyHostToDevice, st));
    k_g1_mul<<<nblocks(L), kBlock, 0, st>>>(L, (const G1J*)dv->in[0].p, p_stride, (const Fr*)dv->in[2].p, (G1J*)dv->in[3].p);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out + b * kG1W, dv->in[3].p, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

// lazily built per-key tables (under key->mu): G1 [g, Y_i] for issuance, G2 [gg, XX] for sign-on verification
static int ensure_issuer_tables(psb_key* key) {
  std::lock_guard<std::mutex> lk(key->mu);
  if (key->haveG1) return PSB_OK;
  for (size_t di = 0; di < g_devs.size(); di++) {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> dl(dv->mu);
    CK(cudaSetDevice(dv->ordinal));
    KeyDev& k = key->d[di];
    // bases [g, Y_0..] are not contiguous in g1pts ([g, X, Y..]): gather them
    G1J* tmp = nullptr;
    CK(cudaMalloc(&tmp, (1 + key->n) * sizeof(G1J)));
    cudaMemcpyAsync(tmp, k.g1pts, sizeof(G1J), cudaMemcpyDeviceToDevice, dv->stream);
    if (key->n) cudaMemcpyAsync(tmp + 1, k.g1pts + 2, key->n * sizeof(G1J), cudaMemcpyDeviceToDevice, dv->stream);
    const int rc = build_tables<Fp>(dv->stream, tmp, (int)(1 + key->n), key->w, &k.tblG1);
    cudaFree(tmp);
    if (rc) return rc;
  }
  key->haveG1 = true;
  return PSB_OK;
}
static int ensure_verifier_tables(psb_key* key) {
  std::lock_guard<std::mutex> lk(key->mu);
  if (key->haveAux) return PSB_OK;
  for (size_t di = 0; di < g_devs.size(); di++) {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> dl(dv->mu);
    CK(cudaSetDevice(dv->ordinal));
    const int rc = build_tables<Fp2>(dv->stream, key->d[di].g2pts, 2, key->w, &key->d[di].tblAux);
    if (rc) return rc;
  }
  key->haveAux = true;
  return PSB_OK;
}
// per-batch G1 bases [H(service), g, y, h] (nb = 1 without id retrieval) -> cached window tables
static int get_batch_tables(psb_key* key, const uint64_t* const pts[4], int nb, std::shared_ptr<BatchTbl>& out) {
  std::array<uint64_t, 4 * kG1W> id{};
  for (int i = 0; i < nb; i++) {
    if (is_zero_words(pts[i] + 2 * kFpW, kFpW)) return fail(PSB_ERR_ARG, "service/authority point is the point at infinity");
    memcpy(&id[kG1W * i], pts[i], kG1W * sizeof(uint64_t));
  }
  std::lock_guard<std::mutex> lk(key->mu);
  for (auto& b : key->batch) if (b->nbases == nb && b->id == id) { out = b; return PSB_OK; }
  auto bt = std::make_shared<BatchTbl>();
  bt->id = id; bt->nbases = nb;
  for (size_t di = 0; di < g_devs.size(); di++) {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> dl(dv->mu);
    CK(cudaSetDevice(dv->ordinal));
    G1J* tmp = nullptr;
    CK(cudaMalloc(&tmp, nb * sizeof(G1J)));
    cudaMemcpyAsync(tmp, id.data(), nb * sizeof(G1J), cudaMemcpyHostToDevice, dv->stream);
    k_normalize_points<Fp><<<1, 32, 0, dv->stream>>>(tmp, nb); LAUNCHED();
    G1A* t = nullptr;
    const int rc = build_tables<Fp>(dv->stream, tmp, nb, kBatchW, &t);
    cudaFree(tmp);
    bt->ordinals.push_back(dv->ordinal);
    bt->tbl.push_back(t);
    if (rc) return rc;
  }
  if (key->batch.size() >= kBatchCache) key->batch.erase(key->batch.begin());
  key->batch.push_back(bt);
  out = bt;
  return PSB_OK;
}

// ---- device-side cores shared by the object-array entry points and the wire-format ones --------------------------
// sign-on verification of L lanes whose inputs are all on device `di`: NIZK steps, then the fused Miller loop and the
// final exponentiation; dok carries an optional pre-verdict in and the NIZK verdict out; scratch from the arena.
struct VidDev {
  const G1J *S1, *S2, *phi, *E1, *E2; const G2J* k; const Fr *c, *rs;
  const uint8_t* blob; const uint64_t* off; const uint8_t* ad; const uint64_t* adoff;
  G2J *Vk, *K; G1J* V; Fp12* F; uint8_t *ok, *ver;
};
static int verify_id_core(psb_key* key, int di, size_t L, const VidDev& v, LaneGeom lg, int with_id, int strict,
                          const BatchTbl& bt, cudaStream_t st) {
  const KeyDev& kd = key->d[di];
  k_vid_g2<<<nblocks(L), kBlock, 0, st>>>(L, (int)key->n, key->w, kd.tblYY, kd.tblAux, v.k, v.c, v.rs, lg, with_id, v.blob, v.off,
                                          v.Vk, v.K, v.ok);
  LAUNCHED();
  k_vid_g1<<<nblocks(L), kBlock, 0, st>>>(L, kBatchW, bt.tbl[di], v.phi, v.E1, v.E2, v.c, v.rs, lg, with_id, v.V);
  LAUNCHED();
  k_vid_hash<<<nblocks(L), kBlock, 0, st>>>(L, v.k, v.phi, v.E1, v.E2, v.Vk, v.V, with_id, v.c, v.ad, v.adoff, v.ok);
  LAUNCHED();
  PSB_WAVES(L, k_verify_miller, v.S1, v.S2, 1, v.K, kd.lines, v.F);
  PSB_WAVES(L, k_verify_final, v.S1, 1, v.F, v.ver, nullptr, v.ok, strict);
  CK(cudaGetLastError());
  return PSB_OK;
}

int psb_provide_id(psb_key* key, size_t N, const uint64_t* A, const uint64_t* c, const uint64_t* rs, size_t per,
                   const uint8_t* attr_blob, const uint64_t* attr_off, const uint8_t* ad_blob, const uint64_t* ad_off,
                   const uint64_t* u, uint8_t* verdict, uint64_t* sig1, uint64_t* sig2, uint8_t* ser) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !A || !c || (per && !rs) || !attr_blob || !attr_off || !ad_blob || !ad_off || !u || !verdict || !sig1 || !sig2)
    return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  if (!key->hasX) return fail(PSB_ERR_ARG, "key was created without the signer secret X");
  int rc = ensure_issuer_tables(key);
  if (rc) return rc;
  const size_t n = key->n;
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t o0 = attr_off[b * n], o1 = attr_off[e * n], a0 = ad_off[b], a1 = ad_off[e];
    Arena ar;
    G1J *dA = nullptr, *dS1 = nullptr, *dS2 = nullptr; Fr *dc = nullptr, *drs = nullptr, *du = nullptr;
    uint8_t *dblob = nullptr, *dad = nullptr, *dver = nullptr, *dser = nullptr; uint64_t *doff = nullptr, *dadoff = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dA = ar.take<G1J>(L); dc = ar.take<Fr>(L); drs = ar.take<Fr>(L * per + 1); du = ar.take<Fr>(L);
      dblob = ar.take<uint8_t>((size_t)(o1 - o0) + 16); doff = ar.take<uint64_t>(L * n + 1);
      dad = ar.take<uint8_t>((size_t)(a1 - a0) + 16); dadoff = ar.take<uint64_t>(L + 1);
      dver = ar.take<uint8_t>(L); dS1 = ar.take<G1J>(L); dS2 = ar.take<G1J>(L); dser = ar.take<uint8_t>(L * kCredSer);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    const KeyDev& kd = key->d[di];
    const LaneGeom lg{(int)per, (int)per, (int)n, nullptr, nullptr};
    const size_t chunk = pipe_chunk(L);     // two-stream pipeline over chunks of lanes (see pipe_chunk)
    int ci = 0;
    for (size_t cb = 0; cb < L; cb += chunk, ci++) {
      const size_t cl = std::min(chunk, L - cb), g = b + cb;
      cudaStream_t cs = (ci & 1) ? dv->copy : st;
      const uint64_t c0 = attr_off[g * n], c1 = attr_off[(g + cl) * n], d0 = ad_off[g], d1 = ad_off[g + cl];
      CK(cudaMemcpyAsync(dA + cb, A + g * kG1W, cl * sizeof(G1J), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(dc + cb, c + g * 4, cl * sizeof(Fr), cudaMemcpyHostToDevice, cs));
      if (per) CK(cudaMemcpyAsync(drs + cb * per, rs + g * per * 4, cl * per * sizeof(Fr), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(du + cb, u + g * 4, cl * sizeof(Fr), cudaMemcpyHostToDevice, cs));
      if (c1 > c0) CK(cudaMemcpyAsync(dblob + (c0 - o0), attr_blob + c0, (size_t)(c1 - c0), cudaMemcpyHostToDevice, cs));
      // (offsets stay absolute; consecutive chunks write the shared boundary entry with the same value)
      CK(cudaMemcpyAsync(doff + cb * n, attr_off + g * n, (cl * n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
      if (d1 > d0) CK(cudaMemcpyAsync(dad + (d0 - a0), ad_blob + d0, (size_t)(d1 - d0), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(dadoff + cb, ad_off + g, (cl + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
      k_provide_id<<<nblocks(cl), kBlock, 0, cs>>>(cl, (int)n, key->w, kd.tblG1, kd.g1pts, dA + cb, dc + cb, drs + cb * per, lg, dblob - o0,
                                                   doff + cb * n, dad - a0, dadoff + cb, du + cb, dver + cb, dS1 + cb, dS2 + cb,
                                                   ser ? dser + cb * kCredSer : nullptr);
      LAUNCHED();
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(verdict + g, dver + cb, cl, cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(sig1 + g * kG1W, dS1 + cb, cl * sizeof(G1J), cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(sig2 + g * kG1W, dS2 + cb, cl * sizeof(G1J), cudaMemcpyDeviceToHost, cs));
      if (ser) CK(cudaMemcpyAsync(ser + g * kCredSer, dser + cb * kCredSer, cl * kCredSer, cudaMemcpyDeviceToHost, cs));
    }
    CK(cudaStreamSynchronize(st));
    if (ci > 1) CK(cudaStreamSynchronize(dv->copy));
    return PSB_OK;
  });
}

int psb_sign(psb_key* key, size_t N, const uint64_t* commitment, size_t n_attrs, const uint8_t* attr_blob,
             const uint64_t* attr_off, const uint64_t* u, uint64_t* sig1, uint64_t* sig2, uint8_t* ser) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !commitment || !u || !sig1 || !sig2 || (n_attrs && (!attr_blob || !attr_off))) return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  if (!key->hasX) return fail(PSB_ERR_ARG, "key was created without the signer secret X");
  if (n_attrs > key->n) return fail(PSB_ERR_ARG, "more attributes than the key has bases (undefined in the reference: m_pk.Yi[i])");
  int rc = ensure_issuer_tables(key);
  if (rc) return rc;
  const size_t na = n_attrs;
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t o0 = na ? attr_off[b * na] : 0, o1 = na ? attr_off[e * na] : 0;
    Arena ar;
    G1J *dC = nullptr, *dS1 = nullptr, *dS2 = nullptr; Fr* du = nullptr;
    uint8_t *dblob = nullptr, *dser = nullptr; uint64_t* doff = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dC = ar.take<G1J>(L); du = ar.take<Fr>(L); dblob = ar.take<uint8_t>((size_t)(o1 - o0) + 16); doff = ar.take<uint64_t>(L * na + 1);
      dS1 = ar.take<G1J>(L); dS2 = ar.take<G1J>(L); dser = ar.take<uint8_t>(L * kCredSer);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    CK(cudaMemcpyAsync(dC, commitment + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(du, u + b * 4, L * sizeof(Fr), cudaMemcpyHostToDevice, st));
    if (o1 > o0) CK(cudaMemcpyAsync(dblob, attr_blob + o0, (size_t)(o1 - o0), cudaMemcpyHostToDevice, st));
    if (na) CK(cudaMemcpyAsync(doff, attr_off + b * na, (L * na + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const KeyDev& kd = key->d[di];
    k_sign<<<nblocks(L), kBlock, 0, st>>>(L, (int)na, key->w, kd.tblG1, kd.g1pts, dC, dblob - o0, na ? doff : nullptr, du, dS1, dS2,
                                          ser ? dser : nullptr);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(sig1 + b * kG1W, dS1, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sig2 + b * kG1W, dS2, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    if (ser) CK(cudaMemcpyAsync(ser + b * kCredSer, dser, L * kCredSer, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

int psb_verify_id(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint64_t* k,
                  const uint64_t* phi, const uint64_t* E1, const uint64_t* E2, const uint64_t* c, const uint64_t* rs,
                  size_t per, const uint8_t* attr_blob, const uint64_t* attr_off, const uint8_t* ad_blob,
                  const uint64_t* ad_off, const uint64_t* service_pt, const uint64_t* y, const uint64_t* g,
                  const uint64_t* h, int flags, uint8_t* verdict) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !sig1 || !sig2 || !k || !phi || !c || (per && !rs) || !attr_blob || !attr_off || !ad_blob || !ad_off ||
      !service_pt || !verdict)
    return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  const int with_id = flags & PSB_VID_WITH_ID, strict = (flags & PSB_VID_REJECT_ZERO_SIGMA) ? 1 : 0;
  if (with_id && (!E1 || !E2 || !y || !g || !h)) return fail(PSB_ERR_ARG, "E1/E2/y/g/h are required with id retrieval");
  int rc = ensure_verifier_tables(key);
  if (rc) return rc;
  std::shared_ptr<BatchTbl> bt;
  const uint64_t* const pts[4] = {service_pt, g, y, h};
  if ((rc = get_batch_tables(key, pts, with_id ? 4 : 1, bt))) return rc;
  const size_t n = key->n;
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t o0 = attr_off[b * n], o1 = attr_off[e * n], a0 = ad_off[b], a1 = ad_off[e];
    Arena ar;
    G1J *dS1 = nullptr, *dS2 = nullptr, *dphi = nullptr, *dE1 = nullptr, *dE2 = nullptr, *dV = nullptr;
    G2J *dk = nullptr, *dVk = nullptr, *dK = nullptr; Fr *dc = nullptr, *drs = nullptr; Fp12* dF = nullptr;
    uint8_t *dblob = nullptr, *dad = nullptr, *dver = nullptr, *dok = nullptr; uint64_t *doff = nullptr, *dadoff = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dS1 = ar.take<G1J>(L); dS2 = ar.take<G1J>(L); dk = ar.take<G2J>(L); dphi = ar.take<G1J>(L);
      dE1 = ar.take<G1J>(with_id ? L : 1); dE2 = ar.take<G1J>(with_id ? L : 1);
      dc = ar.take<Fr>(L); drs = ar.take<Fr>(L * per + 1);
      dblob = ar.take<uint8_t>((size_t)(o1 - o0) + 16); doff = ar.take<uint64_t>(L * n + 1);
      dad = ar.take<uint8_t>((size_t)(a1 - a0) + 16); dadoff = ar.take<uint64_t>(L + 1);
      dVk = ar.take<G2J>(L); dK = ar.take<G2J>(L); dV = ar.take<G1J>(3 * L); dF = ar.take<Fp12>(L);
      dok = ar.take<uint8_t>(L); dver = ar.take<uint8_t>(L);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    CK(cudaMemcpyAsync(dS1, sig1 + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dS2, sig2 + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dk, k + b * kG2W, L * sizeof(G2J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dphi, phi + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    if (with_id) {
      CK(cudaMemcpyAsync(dE1, E1 + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(dE2, E2 + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(dc, c + b * 4, L * sizeof(Fr), cudaMemcpyHostToDevice, st));
    if (per) CK(cudaMemcpyAsync(drs, rs + b * per * 4, L * per * sizeof(Fr), cudaMemcpyHostToDevice, st));
    if (o1 > o0) CK(cudaMemcpyAsync(dblob, attr_blob + o0, (size_t)(o1 - o0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(doff, attr_off + b * n, (L * n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (a1 > a0) CK(cudaMemcpyAsync(dad, ad_blob + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dadoff, ad_off + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const VidDev v{dS1, dS2, dphi, dE1, dE2, dk, dc, drs, dblob - o0, doff, dad - a0, dadoff, dVk, dK, dV, dF, dok, dver};
    const LaneGeom lg{(int)per, (int)per, (int)n, nullptr, nullptr};
    int r = verify_id_core(key, di, L, v, lg, with_id ? 1 : 0, strict, *bt, st);
    if (r) return r;
    CK(cudaMemcpyAsync(verdict + b, dver, L, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

// ---- wire-format entry points: lane j = buf[buf_off[j] .. buf_off[j+1]), the bytes of
//      IdProof::toBufferString() / PSCredRequest::toBufferString(), or their base64 text (PSBuffer::toBase64) --------------
struct WireDev {
  uint8_t *raw, *attr; uint64_t *off, *aoff; uint32_t *len, *pos; Fr *c, *rs; int* per; uint8_t *parsed, *has_e;
};
// H2D of the buffers + base64 + TLV walk on device; carves from `ar` (called in both arena passes)
static void wire_carve(Arena& ar, WireDev& w, size_t L, size_t bytes, size_t n, uint8_t** text) {
  *text = ar.take<uint8_t>(bytes + 16);
  w.raw = ar.take<uint8_t>(bytes + 16); w.attr = ar.take<uint8_t>(bytes + 16);
  w.off = ar.take<uint64_t>(L + 1); w.aoff = ar.take<uint64_t>(L * (n + 1) + 1);
  w.len = ar.take<uint32_t>(L); w.pos = ar.take<uint32_t>(L * W_SLOTS);
  w.c = ar.take<Fr>(L); w.rs = ar.take<Fr>(L * (n + 2) + 1); w.per = ar.take<int>(L);
  w.parsed = ar.take<uint8_t>(L); w.has_e = ar.take<uint8_t>(L);
}
static int wire_ingest(const WireDev& w, uint8_t* dtext, size_t L, size_t n, int kind, int base64, const uint8_t* buf,
                       const uint64_t* buf_off, size_t b, cudaStream_t st) {
  const uint64_t w0 = buf_off[b], w1 = buf_off[b + L];
  uint8_t* dst = base64 ? dtext : w.raw;
  if (w1 > w0) CK(cudaMemcpyAsync(dst, buf + w0, (size_t)(w1 - w0), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(w.off, buf_off + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  // device offsets stay absolute (as the host gave them): the data pointers are biased by -w0 instead
  if (base64) { k_wire_base64<<<nblocks(L), kBlock, 0, st>>>(L, dtext - w0, w.off, w.raw - w0, w.len); LAUNCHED(); }
  k_wire_parse<<<nblocks(L), kBlock, 0, st>>>(L, (int)n, kind, w.raw - w0, w.off, base64 ? w.len : nullptr, w.pos, w.c, w.rs, w.per,
                                              w.attr - w0, w.aoff, w.parsed, w.has_e);
  LAUNCHED();
  CK(cudaGetLastError());
  return PSB_OK;
}

int psb_verify_id_ser(psb_key* key, size_t N, const uint8_t* buf, const uint64_t* buf_off, int base64, const uint8_t* ad_blob,
                      const uint64_t* ad_off, const uint64_t* service_pt, const uint64_t* y, const uint64_t* g,
                      const uint64_t* h, int flags, uint8_t* verdict, uint8_t* parsed) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !buf || !buf_off || !ad_blob || !ad_off || !service_pt || !verdict) return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  const int with_id = flags & PSB_VID_WITH_ID, strict = (flags & PSB_VID_REJECT_ZERO_SIGMA) ? 1 : 0;
  if (with_id && (!y || !g || !h)) return fail(PSB_ERR_ARG, "y/g/h are required with id retrieval");
  int rc = ensure_verifier_tables(key);
  if (rc) return rc;
  std::shared_ptr<BatchTbl> bt;
  const uint64_t* const pts[4] = {service_pt, g, y, h};
  if ((rc = get_batch_tables(key, pts, with_id ? 4 : 1, bt))) return rc;
  const size_t n = key->n;
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t w0 = buf_off[b], w1 = buf_off[e], a0 = ad_off[b], a1 = ad_off[e];
    Arena ar;
    WireDev w{};
    uint8_t* dtext = nullptr;
    G1J *dS1 = nullptr, *dS2 = nullptr, *dphi = nullptr, *dE1 = nullptr, *dE2 = nullptr, *dV = nullptr;
    G2J *dk = nullptr, *dVk = nullptr, *dK = nullptr; Fp12* dF = nullptr;
    uint8_t *dad = nullptr, *dver = nullptr, *dok = nullptr; uint64_t* dadoff = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      wire_carve(ar, w, L, (size_t)(w1 - w0), n, &dtext);
      dS1 = ar.take<G1J>(L); dS2 = ar.take<G1J>(L); dk = ar.take<G2J>(L); dphi = ar.take<G1J>(L); dE1 = ar.take<G1J>(L); dE2 = ar.take<G1J>(L);
      dad = ar.take<uint8_t>((size_t)(a1 - a0) + 16); dadoff = ar.take<uint64_t>(L + 1);
      dVk = ar.take<G2J>(L); dK = ar.take<G2J>(L); dV = ar.take<G1J>(3 * L); dF = ar.take<Fp12>(L);
      dok = ar.take<uint8_t>(L); dver = ar.take<uint8_t>(L);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    int r = wire_ingest(w, dtext, L, n, 0, base64, buf, buf_off, b, st);
    if (r) return r;
    if (a1 > a0) CK(cudaMemcpyAsync(dad, ad_blob + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dadoff, ad_off + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const WireG1Out g1o{{dS1, dS2, dphi, dE1, dE2}};
    k_wire_points<<<nblocks(L * W_SLOTS), kBlock, 0, st>>>(L, W_SLOTS, w.raw - w0, w.off, w.pos, g1o, dk, w.parsed);
    LAUNCHED();
    if (with_id) { k_and_flags<<<nblocks(L), kBlock, 0, st>>>(L, w.parsed, w.has_e); LAUNCHED(); }   // no E1 / E2: the reference returns false
    const VidDev v{dS1, dS2, dphi, dE1, dE2, dk, w.c, w.rs, w.attr - w0, w.aoff, dad - a0, dadoff, dVk, dK, dV, dF, dok, dver};
    const LaneGeom lg{0, (int)n + 2, (int)n + 1, w.per, w.parsed};
    if ((r = verify_id_core(key, di, L, v, lg, with_id ? 1 : 0, strict, *bt, st))) return r;
    CK(cudaMemcpyAsync(verdict + b, dver, L, cudaMemcpyDeviceToHost, st));
    if (parsed) CK(cudaMemcpyAsync(parsed + b, w.parsed, L, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

int psb_provide_id_ser(psb_key* key, size_t N, const uint8_t* buf, const uint64_t* buf_off, int base64, const uint8_t* ad_blob,
                       const uint64_t* ad_off, const uint64_t* u, uint8_t* verdict, uint64_t* sig1, uint64_t* sig2, uint8_t* ser,
                       uint8_t* parsed) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !buf || !buf_off || !ad_blob || !ad_off || !u || !verdict || !sig1 || !sig2) return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  if (!key->hasX) return fail(PSB_ERR_ARG, "key was created without the signer secret X");
  int rc = ensure_issuer_tables(key);
  if (rc) return rc;
  const size_t n = key->n;
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t w0 = buf_off[b], w1 = buf_off[e], a0 = ad_off[b], a1 = ad_off[e];
    Arena ar;
    WireDev w{};
    uint8_t* dtext = nullptr;
    G1J *dA = nullptr, *dS1 = nullptr, *dS2 = nullptr; Fr* du = nullptr;
    uint8_t *dad = nullptr, *dver = nullptr, *dser = nullptr; uint64_t* dadoff = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      wire_carve(ar, w, L, (size_t)(w1 - w0), n, &dtext);
      dA = ar.take<G1J>(L); du = ar.take<Fr>(L); dad = ar.take<uint8_t>((size_t)(a1 - a0) + 16); dadoff = ar.take<uint64_t>(L + 1);
      dver = ar.take<uint8_t>(L); dS1 = ar.take<G1J>(L); dS2 = ar.take<G1J>(L); dser = ar.take<uint8_t>(L * kCredSer);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    int r = wire_ingest(w, dtext, L, n, 1, base64, buf, buf_off, b, st);
    if (r) return r;
    CK(cudaMemcpyAsync(du, u + b * 4, L * sizeof(Fr), cudaMemcpyHostToDevice, st));
    if (a1 > a0) CK(cudaMemcpyAsync(dad, ad_blob + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dadoff, ad_off + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const WireG1Out g1o{{dA, nullptr, nullptr, nullptr, nullptr}};
    k_wire_points<<<nblocks(L), kBlock, 0, st>>>(L, 1, w.raw - w0, w.off, w.pos, g1o, nullptr, w.parsed);
    LAUNCHED();
    const KeyDev& kd = key->d[di];
    const LaneGeom lg{0, (int)n + 2, (int)n + 1, w.per, w.parsed};
    k_provide_id<<<nblocks(L), kBlock, 0, st>>>(L, (int)n, key->w, kd.tblG1, kd.g1pts, dA, w.c, w.rs, lg, w.attr - w0, w.aoff,
                                                dad - a0, dadoff, du, dver, dS1, dS2, ser ? dser : nullptr);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(verdict + b, dver, L, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sig1 + b * kG1W, dS1, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sig2 + b * kG1W, dS2, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    if (ser) CK(cudaMemcpyAsync(ser + b * kCredSer, dser, L * kCredSer, cudaMemcpyDeviceToHost, st));
    if (parsed) CK(cudaMemcpyAsync(parsed + b, w.parsed, L, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

// ---- prover side -----------------------------------------------------------------------------
static size_t count_hidden(const uint8_t* hide, size_t n) {
  size_t h = 0;
  for (size_t i = 0; i < n; i++) h += hide[i] ? 1 : 0;
  return h;
}

int psb_request_id(psb_key* key, size_t N, const uint8_t* attr_blob, const uint64_t* attr_off, const uint8_t* hide,
                   const uint8_t* ad_blob, const uint64_t* ad_off, const uint64_t* rnd, uint64_t* A, uint64_t* c,
                   uint64_t* rs) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !attr_blob || !attr_off || (key->n && !hide) || !ad_blob || !ad_off || !rnd || !A || !c || !rs)
    return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  int rc = ensure_issuer_tables(key);
  if (rc) return rc;
  const size_t n = key->n, h = count_hidden(hide, n);
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t o0 = attr_off[b * n], o1 = attr_off[e * n], a0 = ad_off[b], a1 = ad_off[e];
    Arena ar;
    G1J* dA = nullptr; Fr *dc = nullptr, *drs = nullptr, *drnd = nullptr;
    uint8_t *dblob = nullptr, *dad = nullptr, *dhide = nullptr; uint64_t *doff = nullptr, *dadoff = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dA = ar.take<G1J>(L); dc = ar.take<Fr>(L); drs = ar.take<Fr>(L * (h + 1)); drnd = ar.take<Fr>(L * (h + 2));
      dblob = ar.take<uint8_t>((size_t)(o1 - o0) + 16); doff = ar.take<uint64_t>(L * n + 1);
      dad = ar.take<uint8_t>((size_t)(a1 - a0) + 16); dadoff = ar.take<uint64_t>(L + 1); dhide = ar.take<uint8_t>(n + 16);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    if (n) {   // shared by every chunk: on the device before either stream starts
      CK(cudaMemcpyAsync(dhide, hide, n, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
    }
    const KeyDev& kd = key->d[di];
    const size_t chunk = pipe_chunk(L);     // two-stream pipeline over chunks of lanes (see pipe_chunk)
    int ci = 0;
    for (size_t cb = 0; cb < L; cb += chunk, ci++) {
      const size_t cl = std::min(chunk, L - cb), g = b + cb;
      cudaStream_t cs = (ci & 1) ? dv->copy : st;
      const uint64_t c0 = attr_off[g * n], c1 = attr_off[(g + cl) * n], d0 = ad_off[g], d1 = ad_off[g + cl];
      CK(cudaMemcpyAsync(drnd + cb * (h + 2), rnd + g * (h + 2) * 4, cl * (h + 2) * sizeof(Fr), cudaMemcpyHostToDevice, cs));
      if (c1 > c0) CK(cudaMemcpyAsync(dblob + (c0 - o0), attr_blob + c0, (size_t)(c1 - c0), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(doff + cb * n, attr_off + g * n, (cl * n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
      if (d1 > d0) CK(cudaMemcpyAsync(dad + (d0 - a0), ad_blob + d0, (size_t)(d1 - d0), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(dadoff + cb, ad_off + g, (cl + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
      k_request_id<<<nblocks(cl), kBlock, 0, cs>>>(cl, (int)n, key->w, kd.tblG1, dhide, (int)h, dblob - o0, doff + cb * n, dad - a0,
                                                   dadoff + cb, drnd + cb * (h + 2), dA + cb, dc + cb, drs + cb * (h + 1));
      LAUNCHED();
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(A + g * kG1W, dA + cb, cl * sizeof(G1J), cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(c + g * 4, dc + cb, cl * sizeof(Fr), cudaMemcpyDeviceToHost, cs));
      CK(cudaMemcpyAsync(rs + g * (h + 1) * 4, drs + cb * (h + 1), cl * (h + 1) * sizeof(Fr), cudaMemcpyDeviceToHost, cs));
    }
    CK(cudaStreamSynchronize(st));
    if (ci > 1) CK(cudaStreamSynchronize(dv->copy));
    return PSB_OK;
  });
}

int psb_wire_encode(int kind, size_t N, size_t n_attrs, const uint64_t* p0, const uint64_t* sig2, const uint64_t* k,
                    const uint64_t* phi, const uint64_t* E1, const uint64_t* E2, const uint64_t* c, const uint64_t* rs,
                    size_t rs_per, const uint8_t* attr_blob, const uint64_t* attr_off, int base64, uint8_t* out, size_t out_cap,
                    uint64_t* out_off) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (kind != PSB_WIRE_IDPROOF && kind != PSB_WIRE_REQUEST) return fail(PSB_ERR_ARG, "kind must be PSB_WIRE_IDPROOF or PSB_WIRE_REQUEST");
  if (!p0 || !c || !rs || !attr_off || !out_off || (kind == PSB_WIRE_IDPROOF && (!sig2 || !k || !phi)))
    return fail(PSB_ERR_ARG, "null argument");
  if ((E1 == nullptr) != (E2 == nullptr)) return fail(PSB_ERR_ARG, "E1 and E2 go together");
  if (rs_per > 0xFFFF || n_attrs > 0xFFFF) return fail(PSB_ERR_ARG, "list longer than appendVar can express");
  const bool has_e = kind == PSB_WIRE_IDPROOF && E1;
  const int n = (int)n_attrs, per = (int)rs_per;
  // message sizes on the host: raw_off (binary) and, with base64, the text offsets
  std::vector<uint64_t> raw_off(base64 ? N + 1 : 0);
  uint64_t* ro = base64 ? raw_off.data() : out_off;
  ro[0] = 0;
  for (size_t j = 0; j < N; j++) {
    const uint64_t* ao = attr_off + j * n_attrs;
    for (size_t i = 0; i < n_attrs; i++) {
      if (ao[i + 1] < ao[i]) return fail(PSB_ERR_ARG, "attr_off must be non-decreasing");
      if (ao[i + 1] - ao[i] > 0xFFFF) return fail(PSB_ERR_ARG, "attribute longer than appendVar can express");
    }
    ro[j + 1] = ro[j] + wire_message_size(kind, n, per, ao, has_e);
  }
  if (base64) {
    out_off[0] = 0;
    for (size_t j = 0; j < N; j++) out_off[j + 1] = out_off[j] + base64_encoded_size((size_t)(ro[j + 1] - ro[j]));
  }
  if (!out) return PSB_OK;                       // size query
  if (out_cap < out_off[N]) return fail(PSB_ERR_ARG, "out_cap is smaller than out_off[N]");
  if (N && n_attrs && !attr_blob && attr_off[N * n_attrs] != attr_off[0]) return fail(PSB_ERR_ARG, "null argument");
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t a0 = attr_off[b * n_attrs], a1 = attr_off[e * n_attrs], r0 = ro[b], r1 = ro[e], t0 = out_off[b], t1 = out_off[e];
    Arena ar;
    G1J* dP[5] = {}; G2J* dk = nullptr; Fr *dc = nullptr, *drs = nullptr;
    uint8_t *dattr = nullptr, *draw = nullptr, *dtext = nullptr; uint64_t *daoff = nullptr, *droff = nullptr, *dtoff = nullptr;
    const int npts = kind == PSB_WIRE_IDPROOF ? (has_e ? 5 : 3) : 1;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      for (int i = 0; i < npts; i++) dP[i] = ar.take<G1J>(L);
      if (kind == PSB_WIRE_IDPROOF) dk = ar.take<G2J>(L);
      dc = ar.take<Fr>(L); drs = ar.take<Fr>(L * rs_per + 1);
      dattr = ar.take<uint8_t>((size_t)(a1 - a0) + 16); daoff = ar.take<uint64_t>(L * n_attrs + 1);
      draw = ar.take<uint8_t>((size_t)(r1 - r0) + 16); droff = ar.take<uint64_t>(L + 1);
      if (base64) { dtext = ar.take<uint8_t>((size_t)(t1 - t0) + 16); dtoff = ar.take<uint64_t>(L + 1); }
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    const uint64_t* src[5] = {p0, sig2, phi, E1, E2};
    for (int i = 0; i < npts; i++) CK(cudaMemcpyAsync(dP[i], src[i] + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    if (dk) CK(cudaMemcpyAsync(dk, k + b * kG2W, L * sizeof(G2J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dc, c + b * 4, L * sizeof(Fr), cudaMemcpyHostToDevice, st));
    if (rs_per) CK(cudaMemcpyAsync(drs, rs + b * rs_per * 4, L * rs_per * sizeof(Fr), cudaMemcpyHostToDevice, st));
    if (a1 > a0) CK(cudaMemcpyAsync(dattr, attr_blob + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(daoff, attr_off + b * n_attrs, (L * n_attrs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(droff, ro + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    WireG1Out pts{};
    for (int i = 0; i < npts; i++) pts.p[i] = dP[i];
    // offsets on the device stay ABSOLUTE; the base pointers are shifted by the range's first byte instead
    k_wire_encode<<<nblocks(L), kBlock, 0, st>>>(L, kind, n, per, pts, dk, dc, drs, dattr - a0, daoff, draw - r0, droff);
    LAUNCHED();
    if (base64) {
      CK(cudaMemcpyAsync(dtoff, out_off + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
      k_wire_base64_encode<<<nblocks(L), kBlock, 0, st>>>(L, draw - r0, droff, dtext - t0, dtoff);
      LAUNCHED();
    }
    CK(cudaGetLastError());
    if (t1 > t0) CK(cudaMemcpyAsync(out + t0, base64 ? dtext : draw, (size_t)(t1 - t0), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

int psb_unblind(size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint64_t* t1, uint64_t* out2) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!sig1 || !sig2 || !t1 || !out2) return fail(PSB_ERR_ARG, "null argument");
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    int rc;
    for (int i = 0; i < 2; i++) if ((rc = ensure(dv->in[i], L * sizeof(G1J)))) return rc;
    if ((rc = ensure(dv->in[2], L * sizeof(Fr)))) return rc;
    if ((rc = ensure(dv->in[3], L * sizeof(G1J)))) return rc;
    G1J *d1 = (G1J*)dv->in[0].p, *d2 = (G1J*)dv->in[1].p, *o2 = (G1J*)dv->in[3].p;
    Fr* dt = (Fr*)dv->in[2].p;
    const size_t chunk = pipe_chunk(L);     // two-stream pipeline over chunks of lanes (see pipe_chunk)
    int ci = 0;
    for (size_t cb = 0; cb < L; cb += chunk, ci++) {
      const size_t cl = std::min(chunk, L - cb);
      cudaStream_t cs = (ci & 1) ? dv->copy : st;
      CK(cudaMemcpyAsync(d1 + cb, sig1 + (b + cb) * kG1W, cl * sizeof(G1J), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(d2 + cb, sig2 + (b + cb) * kG1W, cl * sizeof(G1J), cudaMemcpyHostToDevice, cs));
      CK(cudaMemcpyAsync(dt + cb, t1 + (b + cb) * 4, cl * sizeof(Fr), cudaMemcpyHostToDevice, cs));
      k_unblind<<<nblocks(cl), kBlock, 0, cs>>>(cl, d1 + cb, d2 + cb, dt + cb, o2 + cb);
      LAUNCHED();
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(out2 + (b + cb) * kG1W, o2 + cb, cl * sizeof(G1J), cudaMemcpyDeviceToHost, cs));
    }
    CK(cudaStreamSynchronize(st));
    if (ci > 1) CK(cudaStreamSynchronize(dv->copy));
    return PSB_OK;
  });
}

int psb_prove_id(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint8_t* attr_blob,
                 const uint64_t* attr_off, const uint8_t* hide, const uint8_t* ad_blob, const uint64_t* ad_off,
                 const uint64_t* service_pt, const uint64_t* y, const uint64_t* g, const uint64_t* h_pt, int with_id,
                 const uint64_t* rnd, uint64_t* o_sig1, uint64_t* o_sig2, uint64_t* o_k, uint64_t* o_phi, uint64_t* o_E1,
                 uint64_t* o_E2, uint64_t* o_c, uint64_t* o_rs) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!key || !sig1 || !sig2 || !attr_blob || !attr_off || !hide || !ad_blob || !ad_off || !service_pt || !rnd || !o_sig1 ||
      !o_sig2 || !o_k || !o_phi || !o_c || !o_rs)
    return fail(PSB_ERR_ARG, "null argument");
  KEY_ALIVE(key);
  if (with_id && (!y || !g || !h_pt || !o_E1 || !o_E2)) return fail(PSB_ERR_ARG, "y/g/h/E1/E2 are required with id retrieval");
  // the reference reads attributes[0] (and attributes[1] with id retrieval) unconditionally: ps-requester.cc:173,185
  if (key->n < (with_id ? 2u : 1u)) return fail(PSB_ERR_ARG, "the proof needs attribute 0 (and attribute 1 with id retrieval)");
  int rc = ensure_verifier_tables(key);
  if (rc) return rc;
  std::shared_ptr<BatchTbl> bt;
  const uint64_t* const pts[4] = {service_pt, g, y, h_pt};
  if ((rc = get_batch_tables(key, pts, with_id ? 4 : 1, bt))) return rc;
  const size_t n = key->n, h = count_hidden(hide, n);
  const size_t rper = h + (with_id ? 5 : 3), per = h + (with_id ? 2 : 1);
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t o0 = attr_off[b * n], o1 = attr_off[e * n], a0 = ad_off[b], a1 = ad_off[e];
    Arena ar;
    G1J *dS1 = nullptr, *dS2 = nullptr, *dO1 = nullptr, *dO2 = nullptr, *dW = nullptr, *dphi = nullptr, *dE1 = nullptr, *dE2 = nullptr;
    G2J *dk = nullptr, *dVk = nullptr; Fr *dc = nullptr, *drs = nullptr, *drnd = nullptr;
    uint8_t *dblob = nullptr, *dad = nullptr, *dhide = nullptr; uint64_t *doff = nullptr, *dadoff = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dS1 = ar.take<G1J>(L); dS2 = ar.take<G1J>(L); dO1 = ar.take<G1J>(L); dO2 = ar.take<G1J>(L); dW = ar.take<G1J>(6 * L);
      dphi = ar.take<G1J>(L); dE1 = ar.take<G1J>(L); dE2 = ar.take<G1J>(L); dk = ar.take<G2J>(L); dVk = ar.take<G2J>(L);
      dc = ar.take<Fr>(L); drs = ar.take<Fr>(L * per); drnd = ar.take<Fr>(L * rper);
      dblob = ar.take<uint8_t>((size_t)(o1 - o0) + 16); doff = ar.take<uint64_t>(L * n + 1);
      dad = ar.take<uint8_t>((size_t)(a1 - a0) + 16); dadoff = ar.take<uint64_t>(L + 1); dhide = ar.take<uint8_t>(n + 16);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    CK(cudaMemcpyAsync(dS1, sig1 + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dS2, sig2 + b * kG1W, L * sizeof(G1J), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(drnd, rnd + b * rper * 4, L * rper * sizeof(Fr), cudaMemcpyHostToDevice, st));
    if (o1 > o0) CK(cudaMemcpyAsync(dblob, attr_blob + o0, (size_t)(o1 - o0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(doff, attr_off + b * n, (L * n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (a1 > a0) CK(cudaMemcpyAsync(dad, ad_blob + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dadoff, ad_off + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dhide, hide, n, cudaMemcpyHostToDevice, st));
    const KeyDev& kd = key->d[di];
    k_pid_g2<<<nblocks(L), kBlock, 0, st>>>(L, (int)n, key->w, kd.tblYY, kd.tblAux, kd.g2pts + 1, dhide, (int)h, with_id, dblob - o0,
                                            doff, drnd, dk, dVk);
    LAUNCHED();
    k_pid_g1<<<nblocks(L), kBlock, 0, st>>>(L, (int)n, kBatchW, bt->tbl[di], dS1, dS2, dblob - o0, doff, drnd, (int)h, with_id, dO1,
                                            dO2, dW);
    LAUNCHED();
    k_pid_hash<<<nblocks(L), kBlock, 0, st>>>(L, (int)n, dhide, (int)h, with_id, dblob - o0, doff, dad - a0, dadoff, drnd, dk, dVk,
                                              dW, dphi, dE1, dE2, dc, drs);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(o_sig1 + b * kG1W, dO1, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o_sig2 + b * kG1W, dO2, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o_k + b * kG2W, dk, L * sizeof(G2J), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o_phi + b * kG1W, dphi, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    if (with_id) {
      CK(cudaMemcpyAsync(o_E1 + b * kG1W, dE1, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(o_E2 + b * kG1W, dE2, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(o_c + b * 4, dc, L * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o_rs + b * per * 4, drs, L * per * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

int psb_hash_to_g1(size_t N, const uint8_t* msg_blob, const uint64_t* msg_off, uint64_t* out, uint8_t* ok) {
  if (!g_init) return fail(PSB_ERR_NOT_INIT, "psb_init not called");
  if (!msg_blob || !msg_off || !out || !ok) return fail(PSB_ERR_ARG, "null argument");
  return shard(N, [&](int di, size_t b, size_t e) -> int {
    Dev* dv = g_devs[di];
    std::lock_guard<std::mutex> lk(dv->mu);
    const size_t L = e - b;
    if (L == 0) return PSB_OK;
    CK(cudaSetDevice(dv->ordinal));
    cudaStream_t st = dv->stream;
    const uint64_t o0 = msg_off[b], o1 = msg_off[e];
    Arena ar;
    uint8_t *dblob = nullptr, *dok = nullptr; uint64_t* doff = nullptr; G1J* dout = nullptr;
    for (int pass = 0; pass < 2; pass++) {
      ar.used = 0;
      dblob = ar.take<uint8_t>((size_t)(o1 - o0) + 16); doff = ar.take<uint64_t>(L + 1); dout = ar.take<G1J>(L); dok = ar.take<uint8_t>(L);
      if (pass == 0) { int r = ensure(dv->arena, ar.used); if (r) return r; ar.base = (char*)dv->arena.p; }
    }
    if (o1 > o0) CK(cudaMemcpyAsync(dblob, msg_blob + o0, (size_t)(o1 - o0), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(doff, msg_off + b, (L + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    k_hash_to_g1<<<nblocks(L), kBlock, 0, st>>>(L, dblob - o0, doff, dout, dok);
    LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out + b * kG1W, dout, L * sizeof(G1J), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ok + b, dok, L, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PSB_OK;
  });
}

}  // extern "C"
