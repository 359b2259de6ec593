// ps_batch.hpp -- C++ host side of the B200 batch engine: the reference's three role classes with
// BATCHED overloads that marshal into flat arrays and call the extern "C" layer (include/psb.h).
//
// Drop-in use: include this header instead of ps-signer.h / ps-requester.h / ps-verifier.h and write
// `psb::PSVerifier` (or `using psb::PSVerifier;`).  The classes derive from the reference's own
// (src/ps-verifier.h:11-71, src/ps-requester.h:11-128, src/ps-signer.h:11-96), so every scalar method
// keeps its signature and keeps running on the host through mcl; the overloads below add the same
// methods over std::vector batches and run on the GPUs:
//
//   psb::PSVerifier::verify(sigs, all_attributes)                       -> psb_verify
//   psb::PSVerifier::el_passo_verify_id(proofs, ads, service, y, g, h)  -> psb_verify_id
//   psb::PSVerifier::el_passo_verify_id_without_id_retrieval(...)       -> psb_verify_id (with_id = 0)
//   psb::PSRequester::verify(sigs, all_attributes)                      -> psb_verify
//   psb::PSRequester::randomize_credential(sigs, t)                     -> psb_randomize   (t host-supplied)
//   psb::PSRequester::el_passo_request_id(attributes, ads, rnd)         -> psb_request_id  (rnd host-supplied, draw order)
//   psb::PSRequester::unblind_credential(sigs, t1)                      -> psb_unblind     (t1 = rnd[j][0] of the request)
//   psb::PSRequester::el_passo_prove_id[_without_id_retrieval](...)     -> psb_prove_id    (rnd host-supplied, draw order)
//   psb::PSSigner::el_passo_provide_id(requests, ads, u, sigs)          -> psb_provide_id  (u host-supplied)
//   psb::PSSigner::sign_commitment / sign_hybrid(commitments, [attrs], u) -> psb_sign       (u host-supplied)
//   psb::PSVerifier::el_passo_verify_id[_without_id_retrieval](wire, ...) -> psb_verify_id_ser  (base64 text or PSBuffer bytes
//   psb::PSSigner::el_passo_provide_id(wire, ads, u, sigs)               -> psb_provide_id_ser  of toBufferString(), parsed on the GPU)
//
// Per-lane problems are DATA, never exceptions: a lane whose attribute list has the wrong length (undefined behaviour in
// the reference: m_pk.YYi[counter] out of range, src/ps-verifier.cc:26), a proof without E1/E2, an undecodable wire buffer
// get verdict 0 and the rest of the batch is unaffected.  Lanes may carry different numbers of responses (they are
// grouped by rs.size()).  Exceptions are for the CALL: mismatching vector lengths, engine failures.
//
// mcl objects are passed WITHOUT conversion: G1/G2/Fr in memory are Montgomery limb arrays in exactly
// the layout psb.h takes.  One curve per build, like mcl's own bn256 / bn384 libraries: with
// <mcl/bls12_381.hpp> objects (384-bit Fp, the reference's headers) call mcl::bn::initPairing(mcl::BLS12_381) and
// link libpsb.so; with mcl's 256-bit configuration (<mcl/bn256.hpp> first, 4-word Fp) call initPairing() = BN254 --
// what the reference's shipped tests select -- and link libpsb_bn254.so.  psb::init() once per process.  Errors: std::runtime_error("attribute size does not match") as in
// src/ps-requester.cc:31-33 for size mismatches; psb::Error for engine failures (no GPU, CUDA error).
// There is no CPU fallback for the batched overloads.
#ifndef PSB_HOST_PS_BATCH_HPP_
#define PSB_HOST_PS_BATCH_HPP_

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#include "ps-requester.h"
#include "ps-signer.h"
#include "ps-verifier.h"
#include "psb.h"

namespace psb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what + ": " + psb_last_error()), code(c) {}
};
inline void check(int rc, const char* what) { if (rc != PSB_OK) throw Error(rc, what); }

constexpr size_t kFpWords = sizeof(mcl::bls12::Fp) / 8;   // 6: BLS12-381 objects (libpsb.so), 4: BN254 objects (libpsb_bn254.so)
constexpr size_t kG1Ser = 8 * kFpWords;                   // compressed G1 on the wire: 48 / 32 bytes
constexpr size_t kCredSer = 2 * (2 + kG1Ser);             // PSCredential::toBufferString: two TLV-framed G1 (100 / 68 bytes)

// once per process, after mcl::bn::initPairing(...); devices = CUDA ordinals (empty = device 0)
inline void init(const std::vector<int>& devices = {}) {
  static_assert(kFpWords == 6 || kFpWords == 4, "mcl object size must be 384 bit (bls12_381.hpp) or 256 bit (bn256.hpp)");
  static_assert(sizeof(mcl::bls12::G1) == 3 * kFpWords * 8 && sizeof(mcl::bls12::G2) == 6 * kFpWords * 8 &&
                sizeof(mcl::bls12::Fr) == 4 * 8 && sizeof(mcl::bls12::GT) == 12 * kFpWords * 8, "unexpected mcl object layout");
  if (mcl::bls12::Fp::getOp().N != kFpWords || !mcl::bls12::Fp::getOp().isMont)
    throw std::runtime_error(kFpWords == 6 ? "psb: call mcl::bn::initPairing(mcl::BLS12_381) first (381-bit Montgomery field)"
                                           : "psb: call mcl::bn::initPairing() (BN254) first (254-bit Montgomery field)");
  check(psb_init(kFpWords == 6 ? PSB_CURVE_BLS12_381 : PSB_CURVE_BN254, devices.empty() ? nullptr : devices.data(),
                 (int)devices.size()), "psb_init");
}

// std::allocator drop-in over psb_host_alloc: std::vector<T, psb::pinned_allocator<T>> keeps a batch array in page-locked
// memory, which the psb_* calls copy at the full PCIe rate (include/psb.h).  Needs psb::init() first.
template <class T> struct pinned_allocator {
  typedef T value_type;
  pinned_allocator() = default;
  template <class U> pinned_allocator(const pinned_allocator<U>&) {}
  T* allocate(size_t n) {
    void* p = psb_host_alloc(n * sizeof(T));
    if (!p) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) { psb_host_free(p); }
  template <class U> bool operator==(const pinned_allocator<U>&) const { return true; }
  template <class U> bool operator!=(const pinned_allocator<U>&) const { return false; }
};

namespace detail {
inline const uint64_t* u64(const void* p) { return reinterpret_cast<const uint64_t*>(p); }
inline uint64_t* u64(void* p) { return reinterpret_cast<uint64_t*>(p); }

// flat string arrays: blob + offsets[count + 1]
struct Strings {
  std::vector<uint8_t> blob;
  std::vector<uint64_t> off{0};
  void add(const std::string& s) {
    blob.insert(blob.end(), s.begin(), s.end());
    off.push_back(blob.size());
  }
  const uint8_t* data() { if (blob.empty()) blob.push_back(0); return blob.data(); }
};

// The reference keeps the signer's secret X = g^x and its public key in PRIVATE members (src/ps-signer.h:92-95) and
// offers no accessor.  Explicit template instantiation may name private members (access checks do not apply to its
// arguments, [temp.spec]), which gives the batch layer read / write access WITHOUT touching the reference's sources,
// the process-wide RandGen, or any lock.  A maintainer who prefers an in-class accessor finds the two-line patch in
// INTEGRATION.md; this header then works unchanged.
template <class Tag, typename Tag::type Member> struct MemberAccess {
  friend typename Tag::type member_of(Tag) { return Member; }
};
typedef ::PSSigner RefSigner;   // (the reference's class; psb::PSSigner below derives from it)
struct SignerSecretX { typedef mcl::bls12::G1 RefSigner::*type; friend type member_of(SignerSecretX); };
struct SignerPubKey { typedef PSPubKey RefSigner::*type; friend type member_of(SignerPubKey); };
template struct MemberAccess<SignerSecretX, &RefSigner::m_sk_X>;
template struct MemberAccess<SignerPubKey, &RefSigner::m_pk>;

inline unsigned marshal_threads(size_t lanes) {
  if (lanes < 8192) return 1;
  const unsigned hw = std::thread::hardware_concurrency();
  return std::max(1u, std::min(hw ? hw : 4u, 16u));
}
template <class F> inline void parallel_lanes(size_t lanes, F f) {   // f(begin, end) on disjoint lane ranges
  const unsigned T = marshal_threads(lanes);
  if (T == 1) { f((size_t)0, lanes); return; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; t++) th.emplace_back([=] { f(lanes * t / T, lanes * (t + 1) / T); });
  for (auto& x : th) x.join();
}
// N lanes x n strings -> blob + off[N * n + 1], in parallel (two passes: sizes, then copies).  get(j) returns lane j's
// list; a lane whose list is not n long is flattened as n empty strings and flagged in bad[j].
template <class Get>
inline void flatten_lanes(size_t N, size_t n, Get get, std::vector<uint8_t>& blob, std::vector<uint64_t>& off,
                          std::vector<uint8_t>& bad) {
  std::vector<uint64_t> lane_bytes(N + 1, 0);
  bad.assign(N, 0);
  parallel_lanes(N, [&](size_t b, size_t e) {
    for (size_t j = b; j < e; j++) {
      const auto& l = get(j);
      if (l.size() != n) { bad[j] = 1; continue; }
      uint64_t t = 0;
      for (const auto& a : l) t += a.size();
      lane_bytes[j + 1] = t;
    }
  });
  for (size_t j = 0; j < N; j++) lane_bytes[j + 1] += lane_bytes[j];
  blob.resize((size_t)lane_bytes[N] + 8);
  off.resize(N * n + 1);
  off[N * n] = lane_bytes[N];
  parallel_lanes(N, [&](size_t b, size_t e) {
    for (size_t j = b; j < e; j++) {
      uint64_t at = lane_bytes[j];
      if (bad[j]) { for (size_t i = 0; i < n; i++) off[j * n + i] = at; continue; }
      const auto& l = get(j);
      for (size_t i = 0; i < n; i++) {
        off[j * n + i] = at;
        if (!l[i].empty()) std::memcpy(&blob[(size_t)at], l[i].data(), l[i].size());
        at += l[i].size();
      }
    }
  });
}

struct KeyDeleter { void operator()(psb_key* k) const { psb_key_destroy(k); } };
using KeyHandle = std::shared_ptr<psb_key>;

inline KeyHandle make_key(const PSPubKey& pk, const mcl::bls12::G1* X_secret, int window_bits) {
  if (pk.Yi.size() != pk.YYi.size()) throw std::runtime_error("attribute size does not match");
  psb_key* k = psb_key_create(u64(&pk.g), u64(&pk.gg), u64(&pk.XX), u64(pk.Yi.data()), u64(pk.YYi.data()), pk.Yi.size(),
                              X_secret ? u64(X_secret) : nullptr, window_bits);
  if (!k) throw Error(PSB_ERR_ARG, "psb_key_create");
  return KeyHandle(k, KeyDeleter());
}

// PSCredential is two G1 members and nothing else (src/ps-encoding.h:89-109): std::vector<PSCredential>::data() IS the
// array of (sigma1, sigma2) pairs psb_verify_aos takes -- no per-credential copy
static_assert(sizeof(PSCredential) == 2 * sizeof(mcl::bls12::G1), "PSCredential must be exactly (sig1, sig2)");
inline std::vector<uint8_t> verify_batch(psb_key* key, size_t n, const std::vector<PSCredential>& sigs,
                                         const std::vector<std::vector<std::string>>& all_attributes) {
  const size_t N = sigs.size();
  if (all_attributes.size() != N) throw std::runtime_error("attribute size does not match");
  std::vector<uint8_t> verdict(N);
  if (N == 0) return verdict;
  std::vector<uint8_t> blob, bad;
  std::vector<uint64_t> off;
  flatten_lanes(N, n, [&](size_t j) -> const std::vector<std::string>& { return all_attributes[j]; }, blob, off, bad);
  check(psb_verify_aos(key, N, u64(sigs.data()), blob.data(), off.data(), nullptr, verdict.data(), nullptr), "psb_verify_aos");
  for (size_t j = 0; j < N; j++) if (bad[j]) verdict[j] = 0;   // wrong attribute count: undefined in the reference, rejected here
  return verdict;
}
// wire buffers -> blob + off[N + 1]
template <class Buf>
inline void flatten_wire(const std::vector<Buf>& bufs, std::vector<uint8_t>& blob, std::vector<uint64_t>& off) {
  off.assign(bufs.size() + 1, 0);
  for (size_t j = 0; j < bufs.size(); j++) off[j + 1] = off[j] + bufs[j].size();
  blob.resize((size_t)off.back() + 8);
  parallel_lanes(bufs.size(), [&](size_t b, size_t e) {
    for (size_t j = b; j < e; j++) if (bufs[j].size()) std::memcpy(&blob[(size_t)off[j]], bufs[j].data(), bufs[j].size());
  });
}
}  // namespace detail

// ---- PSVerifier ----------------------------------------------------------------------------------------
class PSVerifier : public ::PSVerifier {
public:
  // strict_sigma: the batched sign-on verification ALSO rejects sigma1 == 0 (PSB_VID_REJECT_ZERO_SIGMA).  The reference's
  // el_passo_verify_id lacks that check: a proof with sigma1 = sigma2 = 0 and an honestly built NIZK
  // passes it without any credential.  Default ON for the batch path; pass false for bit-exact reference verdicts.
  explicit PSVerifier(const PSPubKey& pk, int window_bits = 0, bool strict_sigma = true)
      : ::PSVerifier(pk), m_n(pk.Yi.size()), m_strict(strict_sigma), m_key(detail::make_key(pk, nullptr, window_bits)) {}

  using ::PSVerifier::verify;
  using ::PSVerifier::el_passo_verify_id;
  using ::PSVerifier::el_passo_verify_id_without_id_retrieval;

  // batched PSVerifier::verify (src/ps-verifier.cc:13-35): verdict[j] = verify(sigs[j], all_attributes[j])
  std::vector<uint8_t> verify(const std::vector<PSCredential>& sigs,
                              const std::vector<std::vector<std::string>>& all_attributes) const {
    return detail::verify_batch(m_key.get(), m_n, sigs, all_attributes);
  }

  // batched verify straight from the WIRE form: credentials[j] = PSCredential::toBufferString() bytes (two G1 TLVs,
  // src/ps-encoding.cc:384-401).  Points are decompressed on the GPU (psb_verify_ser); buffers that are not the
  // canonical 100-byte (BN254: 68-byte) layout go through the reference's own parser on the host first.  An undecodable point gives
  // verdict 0 (the reference's parser ignores deserialize failures).
  std::vector<uint8_t> verify(const std::vector<PSBuffer>& credentials,
                              const std::vector<std::vector<std::string>>& all_attributes) const {
    using namespace mcl::bls12;
    const size_t N = credentials.size();
    if (all_attributes.size() != N) throw std::runtime_error("attribute size does not match");
    constexpr size_t S = kG1Ser, T = 2 + kG1Ser;   // compressed G1, one TLV-framed G1
    std::vector<uint8_t> flat(N * 2 * S + 1), verdict(N);
    if (N == 0) return verdict;
    std::vector<uint8_t> blob, bad;
    std::vector<uint64_t> off;
    detail::flatten_lanes(N, m_n, [&](size_t j) -> const std::vector<std::string>& { return all_attributes[j]; }, blob, off, bad);
    detail::parallel_lanes(N, [&](size_t b0, size_t e0) {
      for (size_t j = b0; j < e0; j++) {
        const PSBuffer& b = credentials[j];
        if (b.size() == kCredSer && b[0] == 1 && b[1] == S && b[T] == 1 && b[T + 1] == S) {
          std::memcpy(&flat[2 * S * j], &b[2], S);
          std::memcpy(&flat[2 * S * j + S], &b[T + 2], S);
        } else {
          bad[j] = 1;                              // not the canonical layout: the lane is rejected (all-zero = infinity = reject)
          std::memset(&flat[2 * S * j], 0, 2 * S);
        }
      }
    });
    check(psb_verify_ser(m_key.get(), N, flat.data(), 2 * S, 0, S, blob.data(), off.data(), verdict.data(), nullptr),
          "psb_verify_ser");
    for (size_t j = 0; j < N; j++) if (bad[j]) verdict[j] = 0;
    return verdict;
  }

  // batched el_passo_verify_id (src/ps-verifier.cc:37-138); one associated_data per proof
  std::vector<uint8_t> el_passo_verify_id(const std::vector<IdProof>& proofs, const std::vector<std::string>& associated_data,
                                          const std::string& service_name, const mcl::bls12::G1& authority_pk,
                                          const mcl::bls12::G1& g, const mcl::bls12::G1& h) const {
    return verify_id_batch(proofs, associated_data, service_name, &authority_pk, &g, &h);
  }
  // batched el_passo_verify_id_without_id_retrieval (src/ps-verifier.cc:140-212)
  std::vector<uint8_t> el_passo_verify_id_without_id_retrieval(const std::vector<IdProof>& proofs,
                                                               const std::vector<std::string>& associated_data,
                                                               const std::string& service_name) const {
    return verify_id_batch(proofs, associated_data, service_name, nullptr, nullptr, nullptr);
  }

  // The same two calls straight from the WIRE: wire[j] = base64 text (PSBuffer::toBase64 of IdProof::toBufferString(),
  // what a relying party receives) when Buf is std::string, the raw bytes when Buf is PSBuffer.  base64, the TLV walk
  // and the decompression of all points run on the GPU (psb_verify_id_ser); a malformed buffer is verdict 0.
  template <class Buf>
  std::vector<uint8_t> el_passo_verify_id(const std::vector<Buf>& wire, const std::vector<std::string>& associated_data,
                                          const std::string& service_name, const mcl::bls12::G1& authority_pk,
                                          const mcl::bls12::G1& g, const mcl::bls12::G1& h,
                                          typename std::enable_if<!std::is_same<Buf, IdProof>::value>::type* = nullptr) const {
    return verify_id_wire(wire, associated_data, service_name, &authority_pk, &g, &h);
  }
  template <class Buf>
  std::vector<uint8_t> el_passo_verify_id_without_id_retrieval(const std::vector<Buf>& wire,
                                                               const std::vector<std::string>& associated_data,
                                                               const std::string& service_name,
                                                               typename std::enable_if<!std::is_same<Buf, IdProof>::value>::type* = nullptr) const {
    return verify_id_wire(wire, associated_data, service_name, nullptr, nullptr, nullptr);
  }

private:
  int flags(bool with_id) const { return (with_id ? PSB_VID_WITH_ID : 0) | (m_strict ? PSB_VID_REJECT_ZERO_SIGMA : 0); }

  template <class Buf>
  std::vector<uint8_t> verify_id_wire(const std::vector<Buf>& wire, const std::vector<std::string>& ads,
                                      const std::string& service_name, const mcl::bls12::G1* y, const mcl::bls12::G1* g,
                                      const mcl::bls12::G1* h) const {
    using namespace mcl::bls12;
    static_assert(std::is_same<Buf, std::string>::value || std::is_same<Buf, PSBuffer>::value,
                  "wire buffers are std::string (base64 text) or PSBuffer (bytes)");
    const size_t N = wire.size();
    if (ads.size() != N) throw std::runtime_error("associated data size does not match");
    std::vector<uint8_t> verdict(N);
    if (N == 0) return verdict;
    std::vector<uint8_t> blob, adb;
    std::vector<uint64_t> off, ado;
    detail::flatten_wire(wire, blob, off);
    detail::flatten_wire(ads, adb, ado);
    G1 svc;
    hashAndMapToG1(svc, service_name);   // one value per batch, on the host
    check(psb_verify_id_ser(m_key.get(), N, blob.data(), off.data(), std::is_same<Buf, std::string>::value ? 1 : 0, adb.data(),
                            ado.data(), detail::u64(&svc), y ? detail::u64(y) : nullptr, y ? detail::u64(g) : nullptr,
                            y ? detail::u64(h) : nullptr, flags(y != nullptr), verdict.data(), nullptr), "psb_verify_id_ser");
    return verdict;
  }

  std::vector<uint8_t> verify_id_batch(const std::vector<IdProof>& proofs, const std::vector<std::string>& ads,
                                       const std::string& service_name, const mcl::bls12::G1* y, const mcl::bls12::G1* g,
                                       const mcl::bls12::G1* h) const {
    using namespace mcl::bls12;
    const size_t N = proofs.size();
    const bool with_id = y != nullptr;
    if (ads.size() != N) throw std::runtime_error("associated data size does not match");
    std::vector<uint8_t> verdict(N, 0);
    if (N == 0) return verdict;
    // lanes are grouped by their number of responses (proofs may hide different numbers of attributes); a lane with a
    // wrong attribute count or without E1 / E2 (the reference returns false, ps-verifier.cc:68-70) is verdict 0 and
    // does not enter any group
    std::map<size_t, std::vector<size_t>> groups;
    for (size_t j = 0; j < N; j++) {
      const IdProof& p = proofs[j];
      if (p.attributes.size() != m_n) continue;
      if (with_id && !(p.E1.has_value() && p.E2.has_value())) continue;
      groups[p.rs.size()].push_back(j);
    }
    G1 svc;
    hashAndMapToG1(svc, service_name);   // one value per batch, on the host
    for (const auto& grp : groups) {
      const size_t per = grp.first, M = grp.second.size();
      const std::vector<size_t>& idx = grp.second;
      std::vector<G1> s1(M), s2(M), phi(M), E1(with_id ? M : 0), E2(with_id ? M : 0);
      std::vector<G2> k(M);
      std::vector<Fr> c(M), rs(M * per + 1);
      std::vector<uint8_t> blob, bad, adb, v(M);
      std::vector<uint64_t> off, ado;
      detail::flatten_lanes(M, m_n, [&](size_t t) -> const std::vector<std::string>& { return proofs[idx[t]].attributes; }, blob, off, bad);
      detail::parallel_lanes(M, [&](size_t b0, size_t e0) {
        for (size_t t = b0; t < e0; t++) {
          const IdProof& p = proofs[idx[t]];
          s1[t] = p.sig1; s2[t] = p.sig2; k[t] = p.k; phi[t] = p.phi; c[t] = p.c;
          for (size_t i = 0; i < per; i++) rs[t * per + i] = p.rs[i];
          if (with_id) { E1[t] = *p.E1; E2[t] = *p.E2; }
        }
      });
      ado.assign(M + 1, 0);
      for (size_t t = 0; t < M; t++) ado[t + 1] = ado[t] + ads[idx[t]].size();
      adb.resize((size_t)ado[M] + 8);
      for (size_t t = 0; t < M; t++) if (!ads[idx[t]].empty()) std::memcpy(&adb[(size_t)ado[t]], ads[idx[t]].data(), ads[idx[t]].size());
      check(psb_verify_id(m_key.get(), M, detail::u64(s1.data()), detail::u64(s2.data()), detail::u64(k.data()),
                          detail::u64(phi.data()), with_id ? detail::u64(E1.data()) : nullptr,
                          with_id ? detail::u64(E2.data()) : nullptr, detail::u64(c.data()), detail::u64(rs.data()), per,
                          blob.data(), off.data(), adb.data(), ado.data(), detail::u64(&svc),
                          with_id ? detail::u64(y) : nullptr, with_id ? detail::u64(g) : nullptr,
                          with_id ? detail::u64(h) : nullptr, flags(with_id), v.data()),
            "psb_verify_id");
      for (size_t t = 0; t < M; t++) verdict[idx[t]] = v[t];
    }
    return verdict;
  }

  size_t m_n;
  bool m_strict;
  detail::KeyHandle m_key;
};

// ---- PSRequester ---------------------------------------------------------------------------------------
class PSRequester : public ::PSRequester {
public:
  explicit PSRequester(const PSPubKey& pk, int window_bits = 0)
      : ::PSRequester(pk), m_n(pk.Yi.size()), m_key(detail::make_key(pk, nullptr, window_bits)) {}

  using ::PSRequester::verify;
  using ::PSRequester::randomize_credential;
  using ::PSRequester::el_passo_request_id;
  using ::PSRequester::unblind_credential;
  using ::PSRequester::el_passo_prove_id;
  using ::PSRequester::el_passo_prove_id_without_id_retrieval;
  typedef std::vector<std::tuple<std::string, bool>> AttrList;   // (value, hide) as in src/ps-requester.h:39

  // batched PSRequester::verify (src/ps-requester.cc:115-137)
  std::vector<uint8_t> verify(const std::vector<PSCredential>& sigs,
                              const std::vector<std::vector<std::string>>& all_attributes) const {
    return detail::verify_batch(m_key.get(), m_n, sigs, all_attributes);
  }

  // batched randomize_credential (src/ps-requester.cc:139-148): out[j] = (t[j] sig1, t[j] sig2) with the
  // randomisers t supplied by the caller (draw them with Fr::setByCSPRNG exactly as the reference does)
  std::vector<PSCredential> randomize_credential(const std::vector<PSCredential>& sigs,
                                                 const std::vector<mcl::bls12::Fr>& t) const {
    using namespace mcl::bls12;
    const size_t N = sigs.size();
    if (t.size() != N) throw std::runtime_error("psb: one randomiser per credential");
    std::vector<PSCredential> out(N);
    if (N == 0) return out;
    std::vector<G1> s1(N), s2(N), o1(N), o2(N);
    for (size_t j = 0; j < N; j++) { s1[j] = sigs[j].sig1; s2[j] = sigs[j].sig2; }
    check(psb_randomize(N, detail::u64(s1.data()), detail::u64(s2.data()), detail::u64(t.data()),
                                detail::u64(o1.data()), detail::u64(o2.data()), nullptr), "psb_randomize");
    for (size_t j = 0; j < N; j++) { out[j].sig1 = o1[j]; out[j].sig2 = o2[j]; }
    return out;
  }

  // batched el_passo_request_id (src/ps-requester.cc:19-97).  All lanes share one hide pattern.  rnd[j] = the h + 2
  // scalars the reference would draw with Fr::setByCSPRNG, in its draw order: t1 (the blinding -- keep it for
  // unblind_credential; the scalar method stores it in m_t1), the randomness of g, one per hidden attribute.
  std::vector<PSCredRequest> el_passo_request_id(const std::vector<AttrList>& attributes,
                                                 const std::vector<std::string>& associated_data,
                                                 const std::vector<std::vector<mcl::bls12::Fr>>& rnd) const {
    using namespace mcl::bls12;
    const size_t N = attributes.size();
    if (associated_data.size() != N || rnd.size() != N) throw std::runtime_error("psb: one associated_data and one rnd per lane");
    std::vector<PSCredRequest> out(N);
    if (N == 0) return out;
    std::vector<uint8_t> hide;
    detail::Strings at, ad;
    const size_t h = flatten(attributes, associated_data, hide, at, ad);
    std::vector<Fr> r(N * (h + 2)), c(N), rs(N * (h + 1));
    std::vector<G1> A(N);
    for (size_t j = 0; j < N; j++) {
      if (rnd[j].size() != h + 2) throw std::runtime_error("psb: el_passo_request_id needs h + 2 scalars per lane");
      for (size_t i = 0; i < h + 2; i++) r[j * (h + 2) + i] = rnd[j][i];
    }
    check(psb_request_id(m_key.get(), N, at.data(), at.off.data(), hide.data(), ad.data(), ad.off.data(), detail::u64(r.data()),
                         detail::u64(A.data()), detail::u64(c.data()), detail::u64(rs.data())), "psb_request_id");
    for (size_t j = 0; j < N; j++) {
      out[j].A = A[j]; out[j].c = c[j];
      out[j].rs.assign(rs.begin() + j * (h + 1), rs.begin() + (j + 1) * (h + 1));
      for (const auto& a : attributes[j]) out[j].attributes.push_back(std::get<1>(a) ? std::string() : std::get<0>(a));
    }
    return out;
  }

  // batched unblind_credential (src/ps-requester.cc:99-113): (sig1, sig2 - t1[j] sig1)
  std::vector<PSCredential> unblind_credential(const std::vector<PSCredential>& sigs, const std::vector<mcl::bls12::Fr>& t1) const {
    using namespace mcl::bls12;
    const size_t N = sigs.size();
    if (t1.size() != N) throw std::runtime_error("psb: one blinding factor per credential");
    std::vector<PSCredential> out(N);
    if (N == 0) return out;
    std::vector<G1> s1(N), s2(N), o2(N);
    for (size_t j = 0; j < N; j++) { s1[j] = sigs[j].sig1; s2[j] = sigs[j].sig2; }
    check(psb_unblind(N, detail::u64(s1.data()), detail::u64(s2.data()), detail::u64(t1.data()), detail::u64(o2.data())), "psb_unblind");
    for (size_t j = 0; j < N; j++) { out[j].sig1 = s1[j]; out[j].sig2 = o2[j]; }
    return out;
  }

  // batched el_passo_prove_id (src/ps-requester.cc:150-310); rnd[j] = t, r, epsilon, one per hidden attribute, random2, random3
  std::vector<IdProof> el_passo_prove_id(const std::vector<PSCredential>& sigs, const std::vector<AttrList>& attributes,
                                         const std::vector<std::string>& associated_data, const std::string& service_name,
                                         const mcl::bls12::G1& authority_pk, const mcl::bls12::G1& g, const mcl::bls12::G1& h,
                                         const std::vector<std::vector<mcl::bls12::Fr>>& rnd) const {
    return prove_batch(sigs, attributes, associated_data, service_name, &authority_pk, &g, &h, rnd);
  }
  // batched el_passo_prove_id_without_id_retrieval (:312-432); rnd[j] = t, r, one per hidden attribute, random2
  std::vector<IdProof> el_passo_prove_id_without_id_retrieval(const std::vector<PSCredential>& sigs,
                                                              const std::vector<AttrList>& attributes,
                                                              const std::vector<std::string>& associated_data,
                                                              const std::string& service_name,
                                                              const std::vector<std::vector<mcl::bls12::Fr>>& rnd) const {
    return prove_batch(sigs, attributes, associated_data, service_name, nullptr, nullptr, nullptr, rnd);
  }

private:
  // (value, hide) lists -> flat strings + the batch's hide pattern; returns the number of hidden attributes
  size_t flatten(const std::vector<AttrList>& attributes, const std::vector<std::string>& ads, std::vector<uint8_t>& hide,
                 detail::Strings& at, detail::Strings& ad) const {
    hide.assign(m_n + 1, 0);
    for (size_t j = 0; j < attributes.size(); j++) {
      if (attributes[j].size() != m_n) throw std::runtime_error("attribute size does not match");   // ps-requester.cc:31-33
      for (size_t i = 0; i < m_n; i++) {
        const uint8_t hd = std::get<1>(attributes[j][i]) ? 1 : 0;
        if (j == 0) hide[i] = hd;
        else if (hide[i] != hd) throw std::runtime_error("psb: all lanes of a batch must hide the same attributes");
        at.add(std::get<0>(attributes[j][i]));
      }
      ad.add(ads[j]);
    }
    size_t h = 0;
    for (size_t i = 0; i < m_n; i++) h += hide[i];
    return h;
  }
  std::vector<IdProof> prove_batch(const std::vector<PSCredential>& sigs, const std::vector<AttrList>& attributes,
                                   const std::vector<std::string>& ads, const std::string& service_name, const mcl::bls12::G1* y,
                                   const mcl::bls12::G1* g, const mcl::bls12::G1* hp,
                                   const std::vector<std::vector<mcl::bls12::Fr>>& rnd) const {
    using namespace mcl::bls12;
    const size_t N = sigs.size();
    const bool with_id = y != nullptr;
    if (attributes.size() != N || ads.size() != N || rnd.size() != N) throw std::runtime_error("psb: one attribute list, associated_data and rnd per lane");
    std::vector<IdProof> out(N);
    if (N == 0) return out;
    std::vector<uint8_t> hide;
    detail::Strings at, ad;
    const size_t h = flatten(attributes, ads, hide, at, ad);
    const size_t rper = h + (with_id ? 5 : 3), per = h + (with_id ? 2 : 1);
    std::vector<Fr> r(N * rper), c(N), rs(N * per);
    std::vector<G1> s1(N), s2(N), o1(N), o2(N), phi(N), E1(N), E2(N);
    std::vector<G2> k(N);
    for (size_t j = 0; j < N; j++) {
      if (rnd[j].size() != rper) throw std::runtime_error("psb: el_passo_prove_id needs h + 5 (h + 3 without id retrieval) scalars per lane");
      for (size_t i = 0; i < rper; i++) r[j * rper + i] = rnd[j][i];
      s1[j] = sigs[j].sig1; s2[j] = sigs[j].sig2;
    }
    G1 svc;
    hashAndMapToG1(svc, service_name);   // one value per batch, on the host
    check(psb_prove_id(m_key.get(), N, detail::u64(s1.data()), detail::u64(s2.data()), at.data(), at.off.data(), hide.data(), ad.data(),
                       ad.off.data(), detail::u64(&svc), with_id ? detail::u64(y) : nullptr, with_id ? detail::u64(g) : nullptr,
                       with_id ? detail::u64(hp) : nullptr, with_id ? 1 : 0, detail::u64(r.data()), detail::u64(o1.data()),
                       detail::u64(o2.data()), detail::u64(k.data()), detail::u64(phi.data()), detail::u64(E1.data()),
                       detail::u64(E2.data()), detail::u64(c.data()), detail::u64(rs.data())), "psb_prove_id");
    for (size_t j = 0; j < N; j++) {
      IdProof& p = out[j];
      p.sig1 = o1[j]; p.sig2 = o2[j]; p.k = k[j]; p.phi = phi[j]; p.c = c[j];
      if (with_id) { p.E1 = E1[j]; p.E2 = E2[j]; }
      p.rs.assign(rs.begin() + j * per, rs.begin() + (j + 1) * per);
      for (const auto& a : attributes[j]) p.attributes.push_back(std::get<1>(a) ? std::string() : std::get<0>(a));
    }
    return out;
  }

  size_t m_n;
  detail::KeyHandle m_key;
};

// ---- PSSigner ------------------------------------------------------------------------------------------
// The batch path needs the signer's secret X = g^x on the device.  The reference keeps it private and drops the
// exponents (src/ps-signer.h:92); detail::MemberAccess reads it in place -- no RandGen games, no locks,
// nothing process-wide.  A signer whose key was LOADED rather than generated uses the (pk, X) constructor.
class PSSigner : public ::PSSigner {
public:
  explicit PSSigner(size_t attribute_num, int window_bits = 0) : ::PSSigner(attribute_num), m_w(window_bits) {}
  PSSigner(size_t attribute_num, const mcl::bls12::G1& g, const mcl::bls12::G2& gg, int window_bits = 0)
      : ::PSSigner(attribute_num, g, gg), m_w(window_bits) {}
  // existing key material: the public key and the secret X = g^x (e.g. from a key store).  The scalar methods of the
  // base class work on the same key.
  PSSigner(const PSPubKey& pk, const mcl::bls12::G1& X, int window_bits = 0)
      : ::PSSigner(pk.Yi.size(), pk.g, pk.gg), m_w(window_bits) {
    this->*member_of(detail::SignerPubKey()) = pk;
    this->*member_of(detail::SignerSecretX()) = X;
    m_n = pk.Yi.size();
    m_key = detail::make_key(pk, &X, m_w);
  }

  using ::PSSigner::el_passo_provide_id;
  using ::PSSigner::sign_commitment;
  using ::PSSigner::sign_hybrid;

  PSPubKey key_gen() {   // hides ::PSSigner::key_gen (not virtual there); same return value
    PSPubKey pk = ::PSSigner::key_gen();
    const mcl::bls12::G1& X = this->*member_of(detail::SignerSecretX());
    m_n = pk.Yi.size();
    m_key = detail::make_key(pk, &X, m_w);
    return pk;
  }

  // batched el_passo_provide_id (src/ps-signer.cc:63-72): returns the NIZK verdicts; sigs[j] is assigned for
  // accepted requests (normalised points, equal as group elements and byte-identical once serialised to the
  // reference's with the same u[j]); u[j] are the issuance scalars the reference draws with Fr::setByCSPRNG.
  std::vector<uint8_t> el_passo_provide_id(const std::vector<PSCredRequest>& requests,
                                           const std::vector<std::string>& associated_data,
                                           const std::vector<mcl::bls12::Fr>& u, std::vector<PSCredential>& sigs) const {
    using namespace mcl::bls12;
    need_key();
    const size_t N = requests.size();
    if (associated_data.size() != N || u.size() != N) throw std::runtime_error("psb: one associated_data and one u per request");
    std::vector<uint8_t> verdict(N, 0);
    sigs.resize(N);
    if (N == 0) return verdict;
    std::map<size_t, std::vector<size_t>> groups;   // by number of responses; wrong attribute count = verdict 0
    for (size_t j = 0; j < N; j++) if (requests[j].attributes.size() == m_n) groups[requests[j].rs.size()].push_back(j);
    for (const auto& grp : groups) {
      const size_t per = grp.first, M = grp.second.size();
      const std::vector<size_t>& idx = grp.second;
      std::vector<G1> A(M), o1(M), o2(M);
      std::vector<Fr> c(M), rs(M * per + 1), uu(M);
      std::vector<uint8_t> blob, bad, adb, v(M);
      std::vector<uint64_t> off, ado(M + 1, 0);
      detail::flatten_lanes(M, m_n, [&](size_t t) -> const std::vector<std::string>& { return requests[idx[t]].attributes; }, blob, off, bad);
      for (size_t t = 0; t < M; t++) {
        const PSCredRequest& r = requests[idx[t]];
        A[t] = r.A; c[t] = r.c; uu[t] = u[idx[t]];
        for (size_t i = 0; i < per; i++) rs[t * per + i] = r.rs[i];
        ado[t + 1] = ado[t] + associated_data[idx[t]].size();
      }
      adb.resize((size_t)ado[M] + 8);
      for (size_t t = 0; t < M; t++)
        if (!associated_data[idx[t]].empty()) std::memcpy(&adb[(size_t)ado[t]], associated_data[idx[t]].data(), associated_data[idx[t]].size());
      check(psb_provide_id(m_key.get(), M, detail::u64(A.data()), detail::u64(c.data()), detail::u64(rs.data()), per,
                           blob.data(), off.data(), adb.data(), ado.data(), detail::u64(uu.data()), v.data(),
                           detail::u64(o1.data()), detail::u64(o2.data()), nullptr), "psb_provide_id");
      for (size_t t = 0; t < M; t++) {
        verdict[idx[t]] = v[t];
        if (v[t]) { sigs[idx[t]].sig1 = o1[t]; sigs[idx[t]].sig2 = o2[t]; }
      }
    }
    return verdict;
  }

  // the same from the WIRE: wire[j] = base64 text (std::string) or bytes (PSBuffer) of PSCredRequest::toBufferString()
  template <class Buf>
  std::vector<uint8_t> el_passo_provide_id(const std::vector<Buf>& wire, const std::vector<std::string>& associated_data,
                                           const std::vector<mcl::bls12::Fr>& u, std::vector<PSCredential>& sigs,
                                           typename std::enable_if<!std::is_same<Buf, PSCredRequest>::value>::type* = nullptr) const {
    using namespace mcl::bls12;
    static_assert(std::is_same<Buf, std::string>::value || std::is_same<Buf, PSBuffer>::value,
                  "wire buffers are std::string (base64 text) or PSBuffer (bytes)");
    need_key();
    const size_t N = wire.size();
    if (associated_data.size() != N || u.size() != N) throw std::runtime_error("psb: one associated_data and one u per request");
    std::vector<uint8_t> verdict(N, 0);
    sigs.resize(N);
    if (N == 0) return verdict;
    std::vector<uint8_t> blob, adb;
    std::vector<uint64_t> off, ado;
    detail::flatten_wire(wire, blob, off);
    detail::flatten_wire(associated_data, adb, ado);
    std::vector<G1> o1(N), o2(N);
    check(psb_provide_id_ser(m_key.get(), N, blob.data(), off.data(), std::is_same<Buf, std::string>::value ? 1 : 0, adb.data(),
                             ado.data(), detail::u64(u.data()), verdict.data(), detail::u64(o1.data()), detail::u64(o2.data()),
                             nullptr, nullptr), "psb_provide_id_ser");
    for (size_t j = 0; j < N; j++) if (verdict[j]) { sigs[j].sig1 = o1[j]; sigs[j].sig2 = o2[j]; }
    return verdict;
  }

  // batched sign_commitment (src/ps-signer.cc:132-146): (u[j] g, u[j] (X + commitments[j])), u host-supplied
  std::vector<PSCredential> sign_commitment(const std::vector<mcl::bls12::G1>& commitments,
                                            const std::vector<mcl::bls12::Fr>& u) const {
    return sign_batch(commitments, nullptr, u);
  }
  // batched sign_hybrid (src/ps-signer.cc:112-130): every lane carries the same number of attribute strings ("" = committed)
  std::vector<PSCredential> sign_hybrid(const std::vector<mcl::bls12::G1>& commitments,
                                        const std::vector<std::vector<std::string>>& attributes,
                                        const std::vector<mcl::bls12::Fr>& u) const {
    return sign_batch(commitments, &attributes, u);
  }

private:
  void need_key() const { if (!m_key) throw std::runtime_error("psb: call key_gen() first (or construct from (pk, X))"); }
  std::vector<PSCredential> sign_batch(const std::vector<mcl::bls12::G1>& cm, const std::vector<std::vector<std::string>>* attrs,
                                       const std::vector<mcl::bls12::Fr>& u) const {
    using namespace mcl::bls12;
    need_key();
    const size_t N = cm.size();
    if (u.size() != N || (attrs && attrs->size() != N)) throw std::runtime_error("psb: one u (and one attribute list) per commitment");
    std::vector<PSCredential> out(N);
    if (N == 0) return out;
    const size_t na = attrs ? (*attrs)[0].size() : 0;
    if (na > m_n) throw std::runtime_error("attribute size does not match");
    std::vector<uint8_t> blob, bad;
    std::vector<uint64_t> off;
    if (na) {
      detail::flatten_lanes(N, na, [&](size_t j) -> const std::vector<std::string>& { return (*attrs)[j]; }, blob, off, bad);
      for (size_t j = 0; j < N; j++) if (bad[j]) throw std::runtime_error("psb: all lanes of a sign_hybrid batch carry the same number of attributes");
    }
    std::vector<G1> o1(N), o2(N);
    check(psb_sign(m_key.get(), N, detail::u64(cm.data()), na, na ? blob.data() : nullptr, na ? off.data() : nullptr,
                   detail::u64(u.data()), detail::u64(o1.data()), detail::u64(o2.data()), nullptr), "psb_sign");
    for (size_t j = 0; j < N; j++) { out[j].sig1 = o1[j]; out[j].sig2 = o2[j]; }
    return out;
  }
  int m_w;
  size_t m_n = 0;
  detail::KeyHandle m_key;
};

// ---- wire-format OUTPUT on the GPU (psb_wire_encode) -------------------------------------------------------------------------
// to_wire_base64(msgs)[j] == msgs[j].toBufferString().toBase64(), to_wire(msgs)[j] == msgs[j].toBufferString()
// (src/ps-encoding.cc:14-54, :429-439, :452-468) for std::vector<IdProof> and std::vector<PSCredRequest>: normalisation of
// every point (one inversion each in mcl's serialize), TLV framing and base64 run on the device -- what a batched prover
// (el_passo_prove_id / el_passo_request_id above) sends.  Messages may differ in their numbers of responses / attributes and in
// carrying E1 / E2: lanes are grouped, one device call per group.
namespace detail {
template <class Msg> struct WireKind;
template <> struct WireKind<IdProof> { static constexpr int kind = PSB_WIRE_IDPROOF; };
template <> struct WireKind<PSCredRequest> { static constexpr int kind = PSB_WIRE_REQUEST; };
inline const mcl::bls12::G1& first_point(const IdProof& p) { return p.sig1; }
inline const mcl::bls12::G1& first_point(const PSCredRequest& r) { return r.A; }

// one homogeneous group of messages (same number of responses and attributes, E1 / E2 all present or all absent)
template <class Msg>
inline void wire_encode_group(const std::vector<Msg>& msgs, const std::vector<size_t>& idx, bool base64, std::vector<uint8_t>& out,
                              std::vector<uint64_t>& off) {
  using namespace mcl::bls12;
  constexpr bool proof = std::is_same<Msg, IdProof>::value;
  const size_t N = idx.size();
  off.assign(N + 1, 0);
  out.clear();
  if (N == 0) return;
  const Msg& m0 = msgs[idx[0]];
  const size_t per = m0.rs.size(), n = m0.attributes.size();
  bool has_e = false;
  if constexpr (proof) has_e = m0.E1.has_value() && m0.E2.has_value();
  std::vector<G1> p0(N), s2(proof ? N : 0), phi(proof ? N : 0), E1(has_e ? N : 0), E2(has_e ? N : 0);
  std::vector<G2> k(proof ? N : 0);
  std::vector<Fr> c(N), rs(N * per + 1);
  std::vector<uint8_t> blob, bad;
  std::vector<uint64_t> aoff;
  flatten_lanes(N, n, [&](size_t t) -> const std::vector<std::string>& { return msgs[idx[t]].attributes; }, blob, aoff, bad);
  parallel_lanes(N, [&](size_t b, size_t e) {
    for (size_t t = b; t < e; t++) {
      const Msg& m = msgs[idx[t]];
      p0[t] = first_point(m);
      c[t] = m.c;
      for (size_t i = 0; i < per; i++) rs[t * per + i] = m.rs[i];
      if constexpr (proof) {
        s2[t] = m.sig2; k[t] = m.k; phi[t] = m.phi;
        if (has_e) { E1[t] = *m.E1; E2[t] = *m.E2; }
      }
    }
  });
  auto call = [&](uint8_t* dst, size_t cap) {
    check(psb_wire_encode(WireKind<Msg>::kind, N, n, u64(p0.data()), proof ? u64(s2.data()) : nullptr, proof ? u64(k.data()) : nullptr,
                          proof ? u64(phi.data()) : nullptr, has_e ? u64(E1.data()) : nullptr, has_e ? u64(E2.data()) : nullptr,
                          u64(c.data()), u64(rs.data()), per, blob.data(), aoff.data(), base64 ? 1 : 0, dst, cap, off.data()),
          "psb_wire_encode");
  };
  call(nullptr, 0);                          // sizes
  out.resize((size_t)off[N] + 8);
  call(out.data(), (size_t)off[N]);
}
// any mix of messages: lanes are grouped by (responses, attributes, E1 / E2 present), one device call per group;
// put(j, bytes, len) receives lane j's message
template <class Msg, class Put>
inline void wire_encode(const std::vector<Msg>& msgs, bool base64, Put put) {
  std::map<std::tuple<size_t, size_t, bool>, std::vector<size_t>> groups;
  for (size_t j = 0; j < msgs.size(); j++) {
    bool has_e = false;
    if constexpr (std::is_same<Msg, IdProof>::value) has_e = msgs[j].E1.has_value() && msgs[j].E2.has_value();
    groups[std::make_tuple(msgs[j].rs.size(), msgs[j].attributes.size(), has_e)].push_back(j);
  }
  for (const auto& g : groups) {
    std::vector<uint8_t> out;
    std::vector<uint64_t> off;
    wire_encode_group(msgs, g.second, base64, out, off);
    parallel_lanes(g.second.size(), [&](size_t b, size_t e) {
      for (size_t t = b; t < e; t++) put(g.second[t], out.data() + off[t], (size_t)(off[t + 1] - off[t]));
    });
  }
}
}  // namespace detail

template <class Msg>
inline std::vector<std::string> to_wire_base64(const std::vector<Msg>& msgs) {
  std::vector<std::string> r(msgs.size());
  detail::wire_encode(msgs, true, [&](size_t j, const uint8_t* p, size_t len) { r[j].assign(reinterpret_cast<const char*>(p), len); });
  return r;
}
template <class Msg>
inline std::vector<PSBuffer> to_wire(const std::vector<Msg>& msgs) {
  std::vector<PSBuffer> r(msgs.size());
  detail::wire_encode(msgs, false, [&](size_t j, const uint8_t* p, size_t len) { r[j].insert(r[j].end(), p, p + len); });
  return r;
}

}  // namespace psb
#endif  // PSB_HOST_PS_BATCH_HPP_
