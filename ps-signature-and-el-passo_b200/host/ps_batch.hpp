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

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
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

struct KeyDeleter { void operator()(psb_key* k) const { psb_key_destroy(k); } };
using KeyHandle = std::shared_ptr<psb_key>;

inline KeyHandle make_key(const PSPubKey& pk, const mcl::bls12::G1* X_secret, int window_bits) {
  if (pk.Yi.size() != pk.YYi.size()) throw std::runtime_error("attribute size does not match");
  psb_key* k = psb_key_create(u64(&pk.g), u64(&pk.gg), u64(&pk.XX), u64(pk.Yi.data()), u64(pk.YYi.data()), pk.Yi.size(),
                              X_secret ? u64(X_secret) : nullptr, window_bits);
  if (!k) throw Error(PSB_ERR_ARG, "psb_key_create");
  return KeyHandle(k, KeyDeleter());
}

inline std::vector<uint8_t> verify_batch(psb_key* key, size_t n, const std::vector<PSCredential>& sigs,
                                         const std::vector<std::vector<std::string>>& all_attributes) {
  const size_t N = sigs.size();
  if (all_attributes.size() != N) throw std::runtime_error("attribute size does not match");
  std::vector<mcl::bls12::G1> s1(N), s2(N);
  Strings at;
  for (size_t j = 0; j < N; j++) {
    if (all_attributes[j].size() != n) throw std::runtime_error("attribute size does not match");
    s1[j] = sigs[j].sig1; s2[j] = sigs[j].sig2;
    for (const auto& a : all_attributes[j]) at.add(a);
  }
  std::vector<uint8_t> verdict(N);
  if (N == 0) return verdict;
  check(psb_verify(key, N, u64(s1.data()), u64(s2.data()), at.data(), at.off.data(), nullptr, verdict.data(), nullptr),
        "psb_verify");
  return verdict;
}
}  // namespace detail

// ---- PSVerifier ----------------------------------------------------------------------------------------
class PSVerifier : public ::PSVerifier {
public:
  explicit PSVerifier(const PSPubKey& pk, int window_bits = 0)
      : ::PSVerifier(pk), m_n(pk.Yi.size()), m_key(detail::make_key(pk, nullptr, window_bits)) {}

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
    detail::Strings at;
    for (size_t j = 0; j < N; j++) {
      if (all_attributes[j].size() != m_n) throw std::runtime_error("attribute size does not match");
      const PSBuffer& b = credentials[j];
      if (b.size() == kCredSer && b[0] == 1 && b[1] == S && b[T] == 1 && b[T + 1] == S) {
        std::memcpy(&flat[2 * S * j], &b[2], S);
        std::memcpy(&flat[2 * S * j + S], &b[T + 2], S);
      } else {
        PSCredential c = PSCredential::fromBufferString(b);
        c.sig1.serialize(&flat[2 * S * j], S);
        c.sig2.serialize(&flat[2 * S * j + S], S);
      }
      for (const auto& a : all_attributes[j]) at.add(a);
    }
    if (N == 0) return verdict;
    check(psb_verify_ser(m_key.get(), N, flat.data(), 2 * S, 0, S, at.data(), at.off.data(), verdict.data(), nullptr),
          "psb_verify_ser");
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

private:
  std::vector<uint8_t> verify_id_batch(const std::vector<IdProof>& proofs, const std::vector<std::string>& ads,
                                       const std::string& service_name, const mcl::bls12::G1* y, const mcl::bls12::G1* g,
                                       const mcl::bls12::G1* h) const {
    using namespace mcl::bls12;
    const size_t N = proofs.size();
    const bool with_id = y != nullptr;
    if (ads.size() != N) throw std::runtime_error("associated data size does not match");
    std::vector<uint8_t> verdict(N);
    if (N == 0) return verdict;
    const size_t per = proofs[0].rs.size();
    std::vector<G1> s1(N), s2(N), phi(N), E1(with_id ? N : 0), E2(with_id ? N : 0);
    std::vector<G2> k(N);
    std::vector<Fr> c(N), rs(N * per + 1);
    std::vector<uint8_t> missing(N, 0);
    detail::Strings at, ad;
    for (size_t j = 0; j < N; j++) {
      const IdProof& p = proofs[j];
      if (p.attributes.size() != m_n) throw std::runtime_error("attribute size does not match");
      if (p.rs.size() != per) throw std::runtime_error("psb: proofs of one batch must carry the same number of responses");
      s1[j] = p.sig1; s2[j] = p.sig2; k[j] = p.k; phi[j] = p.phi; c[j] = p.c;
      for (size_t i = 0; i < per; i++) rs[j * per + i] = p.rs[i];
      if (with_id) {
        if (p.E1.has_value() && p.E2.has_value()) { E1[j] = *p.E1; E2[j] = *p.E2; }
        else { E1[j].clear(); E2[j].clear(); missing[j] = 1; }   // reference returns false (ps-verifier.cc:68-70)
      }
      for (const auto& a : p.attributes) at.add(a);
      ad.add(ads[j]);
    }
    G1 svc;
    hashAndMapToG1(svc, service_name);   // one value per batch, on the host
    check(psb_verify_id(m_key.get(), N, detail::u64(s1.data()), detail::u64(s2.data()), detail::u64(k.data()),
                                detail::u64(phi.data()), with_id ? detail::u64(E1.data()) : nullptr,
                                with_id ? detail::u64(E2.data()) : nullptr, detail::u64(c.data()), detail::u64(rs.data()), per,
                                at.data(), at.off.data(), ad.data(), ad.off.data(), detail::u64(&svc),
                                with_id ? detail::u64(y) : nullptr, with_id ? detail::u64(g) : nullptr,
                                with_id ? detail::u64(h) : nullptr, with_id ? 1 : 0, verdict.data()),
                  "psb_verify_id");
    for (size_t j = 0; j < N; j++) if (missing[j]) verdict[j] = 0;
    return verdict;
  }

  size_t m_n;
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
// The reference keeps X = g^x private (src/ps-signer.h:92) and drops the exponents.  The
// batch path needs X on the device; key_gen() below recovers it through the PUBLIC API only: with the
// process RandGen momentarily yielding u = 1, sign_commitment(g) returns (g, X + g) (ps-signer.cc:132-146).
// A maintainer applying INTEGRATION.md's in-class patch reads m_sk_X directly instead.
class PSSigner : public ::PSSigner {
public:
  explicit PSSigner(size_t attribute_num, int window_bits = 0) : ::PSSigner(attribute_num), m_w(window_bits) {}
  PSSigner(size_t attribute_num, const mcl::bls12::G1& g, const mcl::bls12::G2& gg, int window_bits = 0)
      : ::PSSigner(attribute_num, g, gg), m_w(window_bits) {}

  using ::PSSigner::el_passo_provide_id;

  PSPubKey key_gen() {   // hides ::PSSigner::key_gen (not virtual there); same return value
    using namespace mcl::bls12;
    PSPubKey pk = ::PSSigner::key_gen();
    mcl::fp::RandGen saved = mcl::fp::RandGen::get();
    mcl::fp::RandGen::setRandFunc(nullptr, read_one);
    PSCredential s = ::PSSigner::sign_commitment(pk.g);
    mcl::fp::RandGen::setRandGen(saved);
    G1 X;
    G1::sub(X, s.sig2, pk.g);
    if (s.sig1 != pk.g) throw std::runtime_error("psb: could not recover the signer secret through sign_commitment");
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
    if (!m_key) throw std::runtime_error("psb: call key_gen() first");
    const size_t N = requests.size();
    if (associated_data.size() != N || u.size() != N) throw std::runtime_error("psb: one associated_data and one u per request");
    std::vector<uint8_t> verdict(N);
    sigs.resize(N);
    if (N == 0) return verdict;
    const size_t per = requests[0].rs.size();
    std::vector<G1> A(N), o1(N), o2(N);
    std::vector<Fr> c(N), rs(N * per + 1);
    detail::Strings at, ad;
    for (size_t j = 0; j < N; j++) {
      const PSCredRequest& r = requests[j];
      if (r.attributes.size() != m_n) throw std::runtime_error("attribute size does not match");
      if (r.rs.size() != per) throw std::runtime_error("psb: requests of one batch must carry the same number of responses");
      A[j] = r.A; c[j] = r.c;
      for (size_t i = 0; i < per; i++) rs[j * per + i] = r.rs[i];
      for (const auto& a : r.attributes) at.add(a);
      ad.add(associated_data[j]);
    }
    check(psb_provide_id(m_key.get(), N, detail::u64(A.data()), detail::u64(c.data()), detail::u64(rs.data()), per,
                                 at.data(), at.off.data(), ad.data(), ad.off.data(), detail::u64(u.data()), verdict.data(),
                                 detail::u64(o1.data()), detail::u64(o2.data()), nullptr), "psb_provide_id");
    for (size_t j = 0; j < N; j++) if (verdict[j]) { sigs[j].sig1 = o1[j]; sigs[j].sig2 = o2[j]; }
    return verdict;
  }

private:
  static uint32_t read_one(void*, void* buf, uint32_t n) {   // little-endian integer 1
    std::memset(buf, 0, n);
    if (n) static_cast<uint8_t*>(buf)[0] = 1;
    return n;
  }
  int m_w;
  size_t m_n = 0;
  detail::KeyHandle m_key;
};

}  // namespace psb
#endif  // PSB_HOST_PS_BATCH_HPP_
