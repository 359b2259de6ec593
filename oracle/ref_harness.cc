// oracle/ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" shim around the UNMODIFIED reference (PS-Signature-and-EL-PASSO src/*.cc and
// its vendored pairing library mcl), compiled from the sources where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libpsref.so.  Nothing from the reference is
// copied into this repository: this file only *calls* the reference's public API
// (PSSigner / PSRequester / PSVerifier, mcl::bls12::{Fp,Fp2,Fp12,Fr,G1,G2,pairing,...}).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (libpsb.so) never does.
//
// All element arrays use mcl's in-memory layout: little-endian 64-bit limbs in
// Montgomery form; Fp = 6 x u64, Fr = 4 x u64, G1 = 3 Fp (Jacobian x,y,z), G2 = 3 Fp2,
// GT = 12 Fp.  For BLS12-381 sizeof(G1)=144, sizeof(G2)=288, sizeof(Fp12)=576.
#include <ps-requester.h>
#include <ps-signer.h>
#include <ps-verifier.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>
#include <cybozu/sha2.hpp>

using namespace mcl::bls12;
typedef uint64_t u64;
// serialized sizes: Fp / compressed G1 = 48 bytes for BLS12-381, 32 for the BN254 build (oracle/Makefile target bn254)
#ifdef PSREF_BN254
static const size_t SZ1 = 32;
#else
static const size_t SZ1 = 48;
#endif
static const size_t SZ2 = 2 * SZ1;

// ---------------------------------------------------------------------------------------------
// deterministic, thread-local byte stream installed as mcl's RandGen (SURVEY.md 8c "DetRng")
// xorshift64*: s^=s>>12; s^=s<<25; s^=s>>27; byte=(s*0x2545F4914F6CDD1D)>>56
// ---------------------------------------------------------------------------------------------
static thread_local u64 g_rng_state = 0x0123456789abcdefull;
static uint32_t det_read(void*, void* buf, uint32_t n) {
  uint8_t* p = (uint8_t*)buf;
  for (uint32_t i = 0; i < n; i++) {
    u64 s = g_rng_state;
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    g_rng_state = s;
    p[i] = (uint8_t)((s * 0x2545F4914F6CDD1Dull) >> 56);
  }
  return n;
}

template <class F>
static void par_for(size_t n, int nthreads, F f) {
  if (nthreads <= 1 || n < 2) { for (size_t i = 0; i < n; i++) f(i, 0); return; }
  std::vector<std::thread> th;
  std::atomic<size_t> next(0);
  const size_t chunk = 16;
  for (int t = 0; t < nthreads; t++) {
    th.emplace_back([&, t]() {
      for (;;) {
        size_t b = next.fetch_add(chunk);
        if (b >= n) break;
        size_t e = b + chunk < n ? b + chunk : n;
        for (size_t i = b; i < e; i++) f(i, t);
      }
    });
  }
  for (auto& x : th) x.join();
}

struct RefKey {
  PSPubKey pk;
  G1 X;  // signer secret g^x (may be zero when not supplied)
  bool hasX;
};

// A PSSigner whose key material is injected (the reference's key_gen discards x, y_i: SURVEY F8).
// Layout-compatible access to the private members is avoided: we rebuild the signer through its
// public constructor and re-derive m_sk_X by replaying key_gen under a seeded RandGen instead.
// For injected keys we use the friend-free route below (SignerShim mirrors sign_* using only
// public reference calls on a PSSigner constructed with the same g, gg).

extern "C" {

int ref_init(int curve) {
  try {
    if (curve == 5) initPairing(mcl::BLS12_381);
    else if (curve == 0) initPairing(mcl::BN254);
    else return -1;
    mcl::fp::RandGen::setRandFunc(nullptr, det_read);
    // setRandFunc(0, f) with self==0 but f!=0 installs f
    return 0;
  } catch (...) { return -1; }
}

void ref_seed(u64 seed) { g_rng_state = seed ? seed : 0x0123456789abcdefull; }

int ref_jit_enabled() { return mcl::fp::isEnableJIT() ? 1 : 0; }
int ref_hw_threads() { return (int)std::thread::hardware_concurrency(); }

void ref_sizes(int* out) {
  out[0] = (int)sizeof(Fp); out[1] = (int)sizeof(Fr); out[2] = (int)sizeof(G1);
  out[3] = (int)sizeof(G2); out[4] = (int)sizeof(Fp12); out[5] = (int)sizeof(Fp2);
  out[6] = (int)sizeof(Fp6);
}

// ---- field ops on raw Montgomery limbs --------------------------------------------------------
// op: 0 add, 1 sub, 2 mul, 3 sqr, 4 neg, 5 inv
#define DEF_FIELD_OP(NAME, T)                                                        \
  void NAME(int op, size_t n, const T* a, const T* b, T* out) {                      \
    for (size_t i = 0; i < n; i++) {                                                 \
      switch (op) {                                                                  \
        case 0: T::add(out[i], a[i], b[i]); break;                                   \
        case 1: T::sub(out[i], a[i], b[i]); break;                                   \
        case 2: T::mul(out[i], a[i], b[i]); break;                                   \
        case 3: T::sqr(out[i], a[i]); break;                                         \
        case 4: T::neg(out[i], a[i]); break;                                         \
        case 5: T::inv(out[i], a[i]); break;                                         \
      }                                                                              \
    }                                                                                \
  }
DEF_FIELD_OP(ref_fp_op, Fp)
DEF_FIELD_OP(ref_fr_op, Fr)
DEF_FIELD_OP(ref_fp2_op, Fp2)
DEF_FIELD_OP(ref_fp6_op, Fp6)
DEF_FIELD_OP(ref_fp12_op, Fp12)

// Frobenius^k on Fp12 (k = 1,2,3)
void ref_fp12_frobenius(int k, size_t n, const Fp12* a, Fp12* out) {
  for (size_t i = 0; i < n; i++) {
    if (k == 1) Fp12::Frobenius(out[i], a[i]);
    else if (k == 2) Fp12::Frobenius2(out[i], a[i]);
    else Fp12::Frobenius3(out[i], a[i]);
  }
}

// integers <-> Montgomery: little-endian byte strings (48 / 32 bytes), value must be < modulus
void ref_fp_from_bytes(size_t n, const uint8_t* in, Fp* out) {
  for (size_t i = 0; i < n; i++) out[i].setArray(in + SZ1 * i, SZ1);
}
void ref_fp_to_bytes(size_t n, const Fp* in, uint8_t* out) {
  for (size_t i = 0; i < n; i++) in[i].serialize(out + SZ1 * i, SZ1);
}
void ref_fr_from_bytes(size_t n, const uint8_t* in, Fr* out) {
  for (size_t i = 0; i < n; i++) out[i].setArray(in + 32 * i, 32);
}
void ref_fr_to_bytes(size_t n, const Fr* in, uint8_t* out) {
  for (size_t i = 0; i < n; i++) in[i].serialize(out + 32 * i, 32);
}

void ref_fr_set_hash_of(const uint8_t* msg, size_t len, Fr* out) { out->setHashOf(msg, len); }
void ref_fr_set_hash_of_batch(size_t n, const uint8_t* blob, const u64* off, Fr* out) {
  for (size_t i = 0; i < n; i++) out[i].setHashOf(blob + off[i], off[i + 1] - off[i]);
}
void ref_fr_rand(size_t n, Fr* out) { for (size_t i = 0; i < n; i++) out[i].setByCSPRNG(); }
void ref_sha256(const uint8_t* msg, size_t len, uint8_t* out32) {
  cybozu::Sha256 h; std::string d = h.digest(msg, len); memcpy(out32, d.data(), 32);
}

// ---- group ops --------------------------------------------------------------------------------
// op: 0 add, 1 sub, 2 dbl, 3 neg, 4 normalize
void ref_g1_op(int op, size_t n, const G1* a, const G1* b, G1* out) {
  for (size_t i = 0; i < n; i++) {
    switch (op) {
      case 0: G1::add(out[i], a[i], b[i]); break;
      case 1: G1::sub(out[i], a[i], b[i]); break;
      case 2: G1::dbl(out[i], a[i]); break;
      case 3: G1::neg(out[i], a[i]); break;
      case 4: out[i] = a[i]; out[i].normalize(); break;
    }
  }
}
void ref_g2_op(int op, size_t n, const G2* a, const G2* b, G2* out) {
  for (size_t i = 0; i < n; i++) {
    switch (op) {
      case 0: G2::add(out[i], a[i], b[i]); break;
      case 1: G2::sub(out[i], a[i], b[i]); break;
      case 2: G2::dbl(out[i], a[i]); break;
      case 3: G2::neg(out[i], a[i]); break;
      case 4: out[i] = a[i]; out[i].normalize(); break;
    }
  }
}
// out[i] = k[i] * P[i]   (P broadcast when p_stride == 0)
void ref_g1_mul(size_t n, const G1* P, size_t p_stride, const Fr* k, G1* out, int nthreads) {
  par_for(n, nthreads, [&](size_t i, int) { G1::mul(out[i], P[p_stride ? i : 0], k[i]); });
}
void ref_g2_mul(size_t n, const G2* P, size_t p_stride, const Fr* k, G2* out, int nthreads) {
  par_for(n, nthreads, [&](size_t i, int) { G2::mul(out[i], P[p_stride ? i : 0], k[i]); });
}
int ref_g1_is_valid(const G1* P) { return P->isValid() ? 1 : 0; }
int ref_g2_is_valid(const G2* P) { return P->isValid() ? 1 : 0; }
int ref_g1_eq(const G1* a, const G1* b) { return *a == *b; }
int ref_g2_eq(const G2* a, const G2* b) { return *a == *b; }

// mcl compressed little-endian serialization (48 / 96 bytes) and its lowercase-hex form
void ref_g1_serialize(size_t n, const G1* P, uint8_t* out) {
  for (size_t i = 0; i < n; i++) P[i].serialize(out + SZ1 * i, SZ1);
}
void ref_g2_serialize(size_t n, const G2* P, uint8_t* out) {
  for (size_t i = 0; i < n; i++) P[i].serialize(out + SZ2 * i, SZ2);
}
int ref_g1_deserialize(size_t n, const uint8_t* in, G1* P) {
  int ok = 1;
  for (size_t i = 0; i < n; i++) ok &= (P[i].deserialize(in + SZ1 * i, SZ1) == SZ1);
  return ok;
}
int ref_g2_deserialize(size_t n, const uint8_t* in, G2* P) {
  int ok = 1;
  for (size_t i = 0; i < n; i++) ok &= (P[i].deserialize(in + SZ2 * i, SZ2) == SZ2);
  return ok;
}
void ref_hash_to_g1(const uint8_t* msg, size_t len, G1* out) { hashAndMapToG1(*out, msg, len); }
// mapToG1 on a given field element (calcBN + cofactor); returns 0 for mcl's exceptional inputs (t = 0, ...)
int ref_map_to_g1(const Fp* t, G1* out) { bool b; mapToG1(&b, *out, *t); if (!b) out->clear(); return b ? 1 : 0; }
void ref_fp_set_hash_of(const uint8_t* msg, size_t len, Fp* out) { out->setHashOf(msg, len); }
void ref_hash_to_g2(const uint8_t* msg, size_t len, G2* out) { hashAndMapToG2(*out, msg, len); }

// ---- pairing ----------------------------------------------------------------------------------
void ref_miller_loop(size_t n, const G1* P, const G2* Q, Fp12* out) {
  for (size_t i = 0; i < n; i++) millerLoop(out[i], P[i], Q[i]);
}
void ref_final_exp(size_t n, const Fp12* in, Fp12* out) {
  for (size_t i = 0; i < n; i++) finalExp(out[i], in[i]);
}
void ref_pairing(size_t n, const G1* P, const G2* Q, Fp12* out, int nthreads) {
  par_for(n, nthreads, [&](size_t i, int) { pairing(out[i], P[i], Q[i]); });
}
// the fused lane value: finalExp(millerLoop(P1,Q1) * millerLoop(-P2,Q2)) (SURVEY 8d parity check)
void ref_pairing_ratio(size_t n, const G1* P1, const G2* Q1, const G1* P2, const G2* Q2, Fp12* out,
                       int nthreads) {
  par_for(n, nthreads, [&](size_t i, int) {
    Fp12 a, b; pairing(a, P1[i], Q1[i]); pairing(b, P2[i], Q2[i]);
    Fp12::unitaryInv(b, b); Fp12::mul(out[i], a, b);
  });
}
// seconds for `iters` pairings per thread on `nthreads` threads (CPU baseline for pairings/s)
double ref_time_pairing(size_t iters, int nthreads) {
  G1 P; G2 Q; hashAndMapToG1(P, "abc", 3); hashAndMapToG2(Q, "edf", 3);
  auto t0 = std::chrono::steady_clock::now();
  par_for((size_t)nthreads * iters, nthreads, [&](size_t, int) { Fp12 e; pairing(e, P, Q); });
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- keys -------------------------------------------------------------------------------------
// Own key generation with known exponents: x, y_i drawn from the seeded stream.
// Outputs: g, gg, XX, Y[n], YY[n] (all normalized, z = 1 as after deserialization), X = g^x,
// and the exponents (Fr Montgomery) x_out[1], y_out[n].
void ref_keygen(size_t n, u64 seed, const char* g_label, const char* gg_label, G1* g, G2* gg,
                G2* XX, G1* Y, G2* YY, G1* X, Fr* x_out, Fr* y_out) {
  ref_seed(seed);
  hashAndMapToG1(*g, g_label, strlen(g_label));
  hashAndMapToG2(*gg, gg_label, strlen(gg_label));
  g->normalize(); gg->normalize();
  x_out->setByCSPRNG();
  G1::mul(*X, *g, *x_out); X->normalize();
  G2::mul(*XX, *gg, *x_out); XX->normalize();
  for (size_t i = 0; i < n; i++) {
    y_out[i].setByCSPRNG();
    G1::mul(Y[i], *g, y_out[i]); Y[i].normalize();
    G2::mul(YY[i], *gg, y_out[i]); YY[i].normalize();
  }
}

void* ref_key_create(const G1* g, const G2* gg, const G2* XX, const G1* Y, const G2* YY, size_t n,
                     const G1* X_or_null) {
  RefKey* k = new RefKey();
  k->pk.g = *g; k->pk.gg = *gg; k->pk.XX = *XX;
  k->pk.Yi.assign(Y, Y + n); k->pk.YYi.assign(YY, YY + n);
  k->hasX = X_or_null != nullptr;
  if (k->hasX) k->X = *X_or_null; else k->X.clear();
  return k;
}
void ref_key_destroy(void* k) { delete (RefKey*)k; }

// serialized public key through the reference's own TLV codec (round trip check / sizes)
size_t ref_key_encode(void* key, uint8_t* out, size_t cap) {
  RefKey* k = (RefKey*)key;
  PSBuffer b = k->pk.toBufferString();
  if (b.size() <= cap) memcpy(out, b.data(), b.size());
  return b.size();
}

// credentials through the reference's own TLV codec (PSCredential::toBufferString / fromBufferString,
// src/ps-encoding.cc:384-401): n credentials -> n * 100 bytes; and back (decode returns the parsed points)
size_t ref_cred_encode(size_t n, const G1* sig1, const G1* sig2, uint8_t* out, size_t cap) {
  size_t used = 0;
  for (size_t i = 0; i < n; i++) {
    PSCredential c; c.sig1 = sig1[i]; c.sig2 = sig2[i];
    PSBuffer b = c.toBufferString();
    if (used + b.size() > cap) return 0;
    memcpy(out + used, b.data(), b.size());
    used += b.size();
  }
  return used;
}
void ref_cred_decode(size_t n, const uint8_t* in, size_t stride, G1* sig1, G1* sig2) {
  for (size_t i = 0; i < n; i++) {
    PSBuffer b; b.assign(in + i * stride, in + (i + 1) * stride);
    PSCredential c = PSCredential::fromBufferString(b);
    sig1[i] = c.sig1; sig2[i] = c.sig2;
  }
}

static inline std::vector<std::string> lane_attrs(const uint8_t* blob, const u64* off, size_t lane,
                                                  size_t n) {
  std::vector<std::string> v; v.reserve(n);
  for (size_t i = 0; i < n; i++) {
    const u64 b = off[lane * n + i], e = off[lane * n + i + 1];
    v.emplace_back((const char*)blob + b, (size_t)(e - b));
  }
  return v;
}

// ---- protocol: the reference's own entry points, lane by lane ---------------------------------
// PSVerifier::verify  (src/ps-verifier.cc:13-35).  gt_or_null receives lhs * unitaryInv(rhs) of the
// reference computation, recomputed with the same mcl calls (pairing twice), 576 B per lane.
void ref_ps_verify(void* key, size_t N, size_t n, const G1* sig1, const G1* sig2,
                   const uint8_t* blob, const u64* off, uint8_t* verdict, Fp12* gt_or_null,
                   int nthreads) {
  RefKey* k = (RefKey*)key;
  PSVerifier v(k->pk);
  par_for(N, nthreads, [&](size_t i, int) {
    PSCredential c; c.sig1 = sig1[i]; c.sig2 = sig2[i];
    std::vector<std::string> attrs = lane_attrs(blob, off, i, n);
    verdict[i] = v.verify(c, attrs) ? 1 : 0;
    if (gt_or_null) {
      G2 K = k->pk.XX, t; Fr m;
      for (size_t j = 0; j < n; j++) {
        m.setHashOf(attrs[j]); G2::mul(t, k->pk.YYi[j], m); G2::add(K, K, t);
      }
      Fp12 a, b; pairing(a, c.sig1, K); pairing(b, c.sig2, k->pk.gg);
      Fp12::unitaryInv(b, b); Fp12::mul(gt_or_null[i], a, b);
    }
  });
}

// PSRequester::randomize_credential (src/ps-requester.cc:139-148) with host-supplied t:
// the scalar is injected by making the RandGen return exactly the bytes that setByCSPRNG would
// turn into t; simpler and equivalent: call the same two G1::mul the reference calls.
// We ALSO run the real method under a seeded stream in ref_randomize_seeded for cross-checking.
void ref_randomize(size_t N, const G1* sig1, const G1* sig2, const Fr* t, G1* out1, G1* out2,
                   uint8_t* ser_or_null, int nthreads) {
  par_for(N, nthreads, [&](size_t i, int) {
    G1::mul(out1[i], sig1[i], t[i]);
    G1::mul(out2[i], sig2[i], t[i]);
    if (ser_or_null) {
      out1[i].serialize(ser_or_null + SZ2 * i, SZ1);
      out2[i].serialize(ser_or_null + SZ2 * i + SZ1, SZ1);
    }
  });
}
// the real PSRequester::randomize_credential under the seeded stream; also returns the t it drew
void ref_randomize_seeded(void* key, u64 seed, const G1* sig1, const G1* sig2, G1* out1, G1* out2,
                          Fr* t_out) {
  RefKey* k = (RefKey*)key;
  PSRequester u(k->pk);
  ref_seed(seed); t_out->setByCSPRNG();
  ref_seed(seed);
  PSCredential c; c.sig1 = *sig1; c.sig2 = *sig2;
  PSCredential r = u.randomize_credential(c);
  *out1 = r.sig1; *out2 = r.sig2;
}

// ---- issuance (PSSigner::el_passo_provide_id, src/ps-signer.cc:63-146) ------------------------
// The reference keeps X = g^x private inside PSSigner and draws u from the CSPRNG.  To drive the
// REAL code with an injected key we construct PSSigner(n, g, gg) and run key_gen() under the
// same seed that ref_keygen used: it then derives the identical x, y_i (same draw order:
// x first, then y_1..y_n -- src/ps-signer.cc:29-55 vs ref_keygen above).
struct RefSigner { PSSigner* s; PSPubKey pk; };
void* ref_signer_create(size_t n, u64 seed, const char* g_label, const char* gg_label) {
  G1 g; G2 gg;
  hashAndMapToG1(g, g_label, strlen(g_label));
  hashAndMapToG2(gg, gg_label, strlen(gg_label));
  g.normalize(); gg.normalize();
  RefSigner* r = new RefSigner();
  r->s = new PSSigner(n, g, gg);
  ref_seed(seed);
  r->pk = r->s->key_gen();
  return r;
}
void ref_signer_destroy(void* s) { RefSigner* r = (RefSigner*)s; delete r->s; delete r; }
// public key of the signer, NOT normalized by the reference (raw Jacobian from G?::mul)
void ref_signer_pubkey(void* s, G1* g, G2* gg, G2* XX, G1* Y, G2* YY) {
  RefSigner* r = (RefSigner*)s;
  *g = r->pk.g; *gg = r->pk.gg; *XX = r->pk.XX;
  for (size_t i = 0; i < r->pk.Yi.size(); i++) { Y[i] = r->pk.Yi[i]; YY[i] = r->pk.YYi[i]; }
}

// Requests: produced by the reference's own PSRequester::el_passo_request_id under a per-lane seed
// (seed + lane).  `hidden` = n flags (shared by all lanes).  Outputs: A, c, rs[(h+1)] per lane and
// the blinding t1 is not exported (unblind is out of scope); attrs are the caller's.
void ref_request_id(void* key, size_t N, size_t n, const uint8_t* blob, const u64* off,
                    const uint8_t* hidden, const uint8_t* ad_blob, const u64* ad_off, u64 seed,
                    G1* A, Fr* c, Fr* rs, int nthreads) {
  RefKey* k = (RefKey*)key;
  size_t h = 0; for (size_t j = 0; j < n; j++) h += hidden[j] ? 1 : 0;
  par_for(N, nthreads, [&](size_t i, int) {
    PSRequester u(k->pk);
    std::vector<std::string> a = lane_attrs(blob, off, i, n);
    std::vector<std::tuple<std::string, bool>> attrs;
    for (size_t j = 0; j < n; j++) attrs.push_back(std::make_tuple(a[j], hidden[j] != 0));
    std::string ad((const char*)ad_blob + ad_off[i], (size_t)(ad_off[i + 1] - ad_off[i]));
    ref_seed(seed + i);
    PSCredRequest r = u.el_passo_request_id(attrs, ad);
    A[i] = r.A; c[i] = r.c;
    for (size_t j = 0; j < h + 1; j++) rs[i * (h + 1) + j] = r.rs[j];
  });
}

// PSRequester::unblind_credential (src/ps-requester.cc:99-113) uses the private blinding m_t1 that
// el_passo_request_id drew: the request is replayed under the same per-lane seed on the same
// PSRequester object, then the real unblind_credential is called on (sig1, sig2).
void ref_unblind(void* key, size_t N, size_t n, const uint8_t* blob, const u64* off, const uint8_t* hidden,
                 const uint8_t* ad_blob, const u64* ad_off, u64 seed, const G1* sig1, const G1* sig2,
                 G1* o_sig1, G1* o_sig2, int nthreads) {
  RefKey* k = (RefKey*)key;
  par_for(N, nthreads, [&](size_t i, int) {
    PSRequester u(k->pk);
    std::vector<std::string> a = lane_attrs(blob, off, i, n);
    std::vector<std::tuple<std::string, bool>> attrs;
    for (size_t j = 0; j < n; j++) attrs.push_back(std::make_tuple(a[j], hidden[j] != 0));
    std::string ad((const char*)ad_blob + ad_off[i], (size_t)(ad_off[i + 1] - ad_off[i]));
    ref_seed(seed + i);
    (void)u.el_passo_request_id(attrs, ad);
    PSCredential c; c.sig1 = sig1[i]; c.sig2 = sig2[i];
    PSCredential r = u.unblind_credential(c);
    o_sig1[i] = r.sig1; o_sig2[i] = r.sig2;
  });
}

// el_passo_provide_id on every lane with the u_i the caller supplies: the RandGen stream is
// primed so that the reference's `u.setByCSPRNG()` (src/ps-signer.cc:135-136) yields exactly u_i.
// setByCSPRNG reads 32 raw bytes and applies the SmallMask rule; feeding the little-endian bytes
// of a canonical u < r reproduces u (mask is the identity for values < r).
static thread_local const uint8_t* g_inject = nullptr;
static thread_local uint32_t g_inject_left = 0;
static uint32_t inject_read(void*, void* buf, uint32_t n) {
  if (g_inject_left >= n) { memcpy(buf, g_inject, n); g_inject += n; g_inject_left -= n; return n; }
  return det_read(nullptr, buf, n);
}
void ref_provide_id(void* signer, size_t N, size_t n, const G1* A, const Fr* c, const Fr* rs,
                    size_t rs_per_lane, const uint8_t* blob, const u64* off,
                    const uint8_t* ad_blob, const u64* ad_off, const Fr* u, uint8_t* verdict,
                    G1* sig1, G1* sig2, uint8_t* ser_or_null, int nthreads) {
  RefSigner* s = (RefSigner*)signer;
  mcl::fp::RandGen::setRandFunc(nullptr, inject_read);
  par_for(N, nthreads, [&](size_t i, int) {
    PSCredRequest r; r.A = A[i]; r.c = c[i];
    r.rs.assign(rs + i * rs_per_lane, rs + (i + 1) * rs_per_lane);
    r.attributes = lane_attrs(blob, off, i, n);
    std::string ad((const char*)ad_blob + ad_off[i], (size_t)(ad_off[i + 1] - ad_off[i]));
    uint8_t ub[32]; u[i].serialize(ub, 32);
    g_inject = ub; g_inject_left = 32;
    PSCredential out; out.sig1.clear(); out.sig2.clear();
    bool ok = s->s->el_passo_provide_id(r, ad, out);
    g_inject_left = 0;
    verdict[i] = ok ? 1 : 0;
    sig1[i] = out.sig1; sig2[i] = out.sig2;
    if (ser_or_null) {
      out.sig1.serialize(ser_or_null + SZ2 * i, SZ1);
      out.sig2.serialize(ser_or_null + SZ2 * i + SZ1, SZ1);
    }
  });
  mcl::fp::RandGen::setRandFunc(nullptr, det_read);
}

// ---- sign-on proofs (PSRequester::el_passo_prove_id / PSVerifier::el_passo_verify_id) ---------
// proofs for N lanes from the reference prover under per-lane seeds; sig = unblinded credential.
// rs_per_lane = h + 2.  with_id = 0 selects the *_without_id_retrieval variants (rs = h + 1).
void ref_prove_id(void* key, size_t N, size_t n, const G1* sig1, const G1* sig2,
                  const uint8_t* blob, const u64* off, const uint8_t* hidden,
                  const uint8_t* ad_blob, const u64* ad_off, const char* service, const G1* y,
                  const G1* g, const G1* h, u64 seed, int with_id, G1* o_sig1, G1* o_sig2, G2* o_k,
                  G1* o_phi, G1* o_E1, G1* o_E2, Fr* o_c, Fr* o_rs, int nthreads) {
  RefKey* k = (RefKey*)key;
  size_t hn = 0; for (size_t j = 0; j < n; j++) hn += hidden[j] ? 1 : 0;
  const size_t per = hn + (with_id ? 2 : 1);
  std::string svc(service);
  par_for(N, nthreads, [&](size_t i, int) {
    PSRequester u(k->pk);
    std::vector<std::string> a = lane_attrs(blob, off, i, n);
    std::vector<std::tuple<std::string, bool>> attrs;
    for (size_t j = 0; j < n; j++) attrs.push_back(std::make_tuple(a[j], hidden[j] != 0));
    std::string ad((const char*)ad_blob + ad_off[i], (size_t)(ad_off[i + 1] - ad_off[i]));
    PSCredential c; c.sig1 = sig1[i]; c.sig2 = sig2[i];
    ref_seed(seed + i);
    IdProof p = with_id ? u.el_passo_prove_id(c, attrs, ad, svc, *y, *g, *h)
                        : u.el_passo_prove_id_without_id_retrieval(c, attrs, ad, svc);
    o_sig1[i] = p.sig1; o_sig2[i] = p.sig2; o_k[i] = p.k; o_phi[i] = p.phi; o_c[i] = p.c;
    if (with_id) { o_E1[i] = p.E1.value(); o_E2[i] = p.E2.value(); }
    for (size_t j = 0; j < per; j++) o_rs[i * per + j] = p.rs[j];
  });
}

void ref_verify_id(void* key, size_t N, size_t n, const G1* sig1, const G1* sig2, const G2* kk,
                   const G1* phi, const G1* E1, const G1* E2, const Fr* c, const Fr* rs,
                   size_t rs_per_lane, const uint8_t* blob, const u64* off, const uint8_t* ad_blob,
                   const u64* ad_off, const char* service, const G1* y, const G1* g, const G1* h,
                   int with_id, uint8_t* verdict, int nthreads) {
  RefKey* k = (RefKey*)key;
  PSVerifier v(k->pk);
  std::string svc(service);
  par_for(N, nthreads, [&](size_t i, int) {
    IdProof p; p.sig1 = sig1[i]; p.sig2 = sig2[i]; p.k = kk[i]; p.phi = phi[i]; p.c = c[i];
    if (with_id) { p.E1 = E1[i]; p.E2 = E2[i]; }
    p.rs.assign(rs + i * rs_per_lane, rs + (i + 1) * rs_per_lane);
    p.attributes = lane_attrs(blob, off, i, n);
    std::string ad((const char*)ad_blob + ad_off[i], (size_t)(ad_off[i + 1] - ad_off[i]));
    verdict[i] = (with_id ? v.el_passo_verify_id(p, ad, svc, *y, *g, *h)
                          : v.el_passo_verify_id_without_id_retrieval(p, ad, svc)) ? 1 : 0;
  });
}

// ---- the reference's wire formats (src/ps-encoding.cc): IdProof / PSCredRequest -> toBufferString [-> toBase64] and back --
// Encoders return the bytes used (0 = cap too small); lane j's message is out[out_off[j] .. out_off[j+1]).
static size_t put_wire(PSBuffer& b, int base64, uint8_t* out, u64* out_off, size_t i, size_t used, size_t cap) {
  std::string t;
  const uint8_t* src = b.data();
  size_t len = b.size();
  if (base64) { t = b.toBase64(); src = (const uint8_t*)t.data(); len = t.size(); }
  if (used + len > cap) return 0;
  memcpy(out + used, src, len);
  out_off[i + 1] = used + len;
  return len;
}
size_t ref_idproof_encode(size_t N, size_t n, const G1* sig1, const G1* sig2, const G2* kk, const G1* phi, const G1* E1,
                          const G1* E2, const Fr* c, const Fr* rs, size_t rs_per_lane, const uint8_t* blob, const u64* off,
                          int with_e, int base64, uint8_t* out, u64* out_off, size_t cap) {
  size_t used = 0;
  out_off[0] = 0;
  for (size_t i = 0; i < N; i++) {
    IdProof p; p.sig1 = sig1[i]; p.sig2 = sig2[i]; p.k = kk[i]; p.phi = phi[i]; p.c = c[i];
    if (with_e) { p.E1 = E1[i]; p.E2 = E2[i]; }
    p.rs.assign(rs + i * rs_per_lane, rs + (i + 1) * rs_per_lane);
    p.attributes = lane_attrs(blob, off, i, n);
    PSBuffer b = p.toBufferString();
    const size_t len = put_wire(b, base64, out, out_off, i, used, cap);
    if (!len) return 0;
    used += len;
  }
  return used;
}
size_t ref_request_encode(size_t N, size_t n, const G1* A, const Fr* c, const Fr* rs, size_t rs_per_lane, const uint8_t* blob,
                          const u64* off, int base64, uint8_t* out, u64* out_off, size_t cap) {
  size_t used = 0;
  out_off[0] = 0;
  for (size_t i = 0; i < N; i++) {
    PSCredRequest r; r.A = A[i]; r.c = c[i];
    r.rs.assign(rs + i * rs_per_lane, rs + (i + 1) * rs_per_lane);
    r.attributes = lane_attrs(blob, off, i, n);
    PSBuffer b = r.toBufferString();
    const size_t len = put_wire(b, base64, out, out_off, i, used, cap);
    if (!len) return 0;
    used += len;
  }
  return used;
}
static PSBuffer take_wire(const uint8_t* buf, const u64* off, size_t i, int base64) {
  if (base64) return PSBuffer::fromBase64(std::string((const char*)buf + off[i], (size_t)(off[i + 1] - off[i])));
  PSBuffer b; b.assign(buf + off[i], buf + off[i + 1]);
  return b;
}
// IdProof::fromBufferString + el_passo_verify_id[_without_id_retrieval] on every lane.  status: 0 = the reference ran
// to completion, 1 = it threw (PSBuffer::at on a short buffer, std::length_error, bad optional access ...): verdict 0.
// NOTE the reference ignores failed deserializations and wrong type bytes: on such lanes it verifies an
// object that is partly uninitialised -- its verdict is recorded but only `false` can be relied on.
void ref_verify_id_wire(void* key, size_t N, const uint8_t* buf, const u64* off, int base64, const uint8_t* ad_blob,
                        const u64* ad_off, const char* service, const G1* y, const G1* g, const G1* h, int with_id,
                        uint8_t* verdict, uint8_t* status, int nthreads) {
  RefKey* k = (RefKey*)key;
  PSVerifier v(k->pk);
  std::string svc(service);
  par_for(N, nthreads, [&](size_t i, int) {
    verdict[i] = 0; status[i] = 0;
    try {
      PSBuffer b = take_wire(buf, off, i, base64);
      IdProof p = IdProof::fromBufferString(b);
      std::string ad((const char*)ad_blob + ad_off[i], (size_t)(ad_off[i + 1] - ad_off[i]));
      verdict[i] = (with_id ? v.el_passo_verify_id(p, ad, svc, *y, *g, *h)
                            : v.el_passo_verify_id_without_id_retrieval(p, ad, svc)) ? 1 : 0;
    } catch (...) { status[i] = 1; }
  });
}
// the parsed fields of well-formed IdProof buffers (checks our device parser field by field); E1/E2 zero when absent
void ref_idproof_decode(size_t N, size_t n, const uint8_t* buf, const u64* off, int base64, G1* sig1, G1* sig2, G2* kk, G1* phi,
                        G1* E1, G1* E2, Fr* c, Fr* rs, size_t rs_cap, int* per, uint8_t* has_e, uint8_t* status) {
  for (size_t i = 0; i < N; i++) {
    status[i] = 0; per[i] = 0; has_e[i] = 0;
    E1[i].clear(); E2[i].clear();
    try {
      IdProof p = IdProof::fromBufferString(take_wire(buf, off, i, base64));
      sig1[i] = p.sig1; sig2[i] = p.sig2; kk[i] = p.k; phi[i] = p.phi; c[i] = p.c;
      per[i] = (int)p.rs.size();
      for (size_t j = 0; j < p.rs.size() && j < rs_cap; j++) rs[i * rs_cap + j] = p.rs[j];
      if (p.E1.has_value() && p.E2.has_value()) { has_e[i] = 1; E1[i] = *p.E1; E2[i] = *p.E2; }
      if (p.attributes.size() != n) status[i] = 2;
    } catch (...) { status[i] = 1; }
  }
}
// PSCredRequest::fromBufferString + el_passo_provide_id with the caller's u (see ref_provide_id)
void ref_provide_id_wire(void* signer, size_t N, const uint8_t* buf, const u64* off, int base64, const uint8_t* ad_blob,
                         const u64* ad_off, const Fr* u, uint8_t* verdict, G1* sig1, G1* sig2, uint8_t* ser, uint8_t* status,
                         int nthreads) {
  RefSigner* s = (RefSigner*)signer;
  mcl::fp::RandGen::setRandFunc(nullptr, inject_read);
  par_for(N, nthreads, [&](size_t i, int) {
    verdict[i] = 0; status[i] = 0;
    PSCredential out; out.sig1.clear(); out.sig2.clear();
    try {
      PSCredRequest r = PSCredRequest::fromBufferString(take_wire(buf, off, i, base64));
      std::string ad((const char*)ad_blob + ad_off[i], (size_t)(ad_off[i + 1] - ad_off[i]));
      uint8_t ub[32]; u[i].serialize(ub, 32);
      g_inject = ub; g_inject_left = 32;
      verdict[i] = s->s->el_passo_provide_id(r, ad, out) ? 1 : 0;
    } catch (...) { status[i] = 1; }
    g_inject_left = 0;
    sig1[i] = out.sig1; sig2[i] = out.sig2;
    out.sig1.serialize(ser + SZ2 * i, SZ1);
    out.sig2.serialize(ser + SZ2 * i + SZ1, SZ1);
  });
  mcl::fp::RandGen::setRandFunc(nullptr, det_read);
}
// PSSigner::sign_hybrid (na attributes per lane; na = 0: sign_commitment) with the caller's u
void ref_sign(void* signer, size_t N, size_t na, const G1* commitment, const uint8_t* blob, const u64* off, const Fr* u,
              G1* sig1, G1* sig2, uint8_t* ser, int nthreads) {
  RefSigner* s = (RefSigner*)signer;
  mcl::fp::RandGen::setRandFunc(nullptr, inject_read);
  par_for(N, nthreads, [&](size_t i, int) {
    uint8_t ub[32]; u[i].serialize(ub, 32);
    g_inject = ub; g_inject_left = 32;
    PSCredential out = na ? s->s->sign_hybrid(commitment[i], lane_attrs(blob, off, i, na)) : s->s->sign_commitment(commitment[i]);
    g_inject_left = 0;
    sig1[i] = out.sig1; sig2[i] = out.sig2;
    out.sig1.serialize(ser + SZ2 * i, SZ1);
    out.sig2.serialize(ser + SZ2 * i + SZ1, SZ1);
  });
  mcl::fp::RandGen::setRandFunc(nullptr, det_read);
}

// timed loop of the reference's own PSVerifier::verify for the CPU baseline: runs lanes
// [0, N) once on nthreads threads, returns seconds.
double ref_time_ps_verify(void* key, size_t N, size_t n, const G1* sig1, const G1* sig2,
                          const uint8_t* blob, const u64* off, uint8_t* verdict, int nthreads) {
  auto t0 = std::chrono::steady_clock::now();
  ref_ps_verify(key, N, n, sig1, sig2, blob, off, verdict, nullptr, nthreads);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
