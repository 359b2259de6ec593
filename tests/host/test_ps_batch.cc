// tests/host/test_ps_batch.cc -- TEST of the C++ drop-in (host/ps_batch.hpp): the reference's own EL PASSO
// flow (test/ps-tests.cc:53-137), with every RP / IdP / randomize step ALSO run through the batched
// overloads on the GPU and compared lane by lane with the reference's scalar methods (mcl, host).
// Linked against the UNMODIFIED reference objects (oracle/_ref/*.o) and libpsb.so.  Exit 0 = all equal.
// Without a usable GPU psb::init throws: exit code 3 (there is no CPU fallback).
#include <cstdio>
#include <iostream>
#include <tuple>

#include "ps_batch.hpp"

using namespace mcl::bls12;

static int fails = 0;
#define EXPECT(cond)                                                          \
  do {                                                                        \
    if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); fails++; } \
  } while (0)

static std::string ser(const G1& p) { return p.serializeToHexStr(); }

int main(int argc, char** argv) {
  const size_t N = argc > 1 ? std::stoul(argv[1]) : 24;   // lanes
  const size_t n = 5;                                      // attributes, first two hidden
#ifdef PSREF_BN254
  initPairing();                 // BN254, as the reference's own tests do (test/ps-tests.cc:142)
#else
  initPairing(mcl::BLS12_381);
#endif
  try {
    psb::init();
  } catch (const std::exception& e) {
    std::printf("psb::init failed (no CPU fallback): %s\n", e.what());
    return 3;
  }
  {  // page-locked batch arrays through the std::allocator drop-in (psb_host_alloc / psb_host_free)
    std::vector<uint64_t, psb::pinned_allocator<uint64_t>> pinned(1 << 16, 7);
    pinned.resize(1 << 17);
    EXPECT(pinned[0] == 7 && pinned[(1 << 16) - 1] == 7 && pinned[(1 << 17) - 1] == 0);
  }
  G1 g, authority_pk, h;
  G2 gg;
  hashAndMapToG1(g, "abc");
  hashAndMapToG2(gg, "edf");
  hashAndMapToG1(authority_pk, "ghi");
  hashAndMapToG1(h, "jkl");

  psb::PSSigner idp(n, g, gg, 8);
  PSPubKey pk = idp.key_gen();
  psb::PSVerifier rp(pk, 8);
  psb::PSRequester wallet(pk, 8);

  // users: one reference PSRequester per lane (it keeps the blinding factor between request and unblind)
  std::vector<::PSRequester> users(N, ::PSRequester(pk));
  std::vector<std::vector<std::tuple<std::string, bool>>> attrs(N);
  std::vector<std::vector<std::string>> plain(N);
  std::vector<PSCredRequest> requests;
  std::vector<std::string> ads;
  for (size_t j = 0; j < N; j++) {
    for (size_t i = 0; i < n; i++) {
      attrs[j].push_back(std::make_tuple("a" + std::to_string(i) + ":" + std::to_string(j), i < 2));
      plain[j].push_back(std::get<0>(attrs[j][i]));
    }
    ads.push_back("sess" + std::to_string(j));
    requests.push_back(users[j].el_passo_request_id(attrs[j], ads[j]));
  }
  if (N > 3) requests[3].c += 1;   // tampered request: must be rejected by both paths

  // IdP-ProvideID, batched on the GPU with host-supplied u
  std::vector<Fr> u(N);
  for (auto& x : u) x.setByCSPRNG();
  std::vector<PSCredential> issued;
  std::vector<uint8_t> ok = idp.el_passo_provide_id(requests, ads, u, issued);
  std::vector<PSCredential> creds(N);
  for (size_t j = 0; j < N; j++) {
    PSCredential scalar_sig;
    const bool scalar_ok = idp.el_passo_provide_id(requests[j], ads[j], scalar_sig);   // reference, own random u
    EXPECT(scalar_ok == (ok[j] != 0));
    if (!ok[j]) { creds[j] = scalar_sig; continue; }
    // same u => same credential: sigma1 = u g, sigma2 / sigma1 relation checked through unblind + verify
    G1 ug; G1::mul(ug, pk.g, u[j]);
    EXPECT(ser(issued[j].sig1) == ser(ug));
    creds[j] = users[j].unblind_credential(issued[j]);
    EXPECT(users[j].verify(creds[j], plain[j]));      // reference verifies the GPU-issued credential
  }
  if (N > 3) {   // lane 3's request was tampered: give it a valid credential for the later steps
    PSCredRequest r = users[3].el_passo_request_id(attrs[3], ads[3]);
    PSCredential s; EXPECT(idp.el_passo_provide_id(r, ads[3], s));
    creds[3] = users[3].unblind_credential(s);
  }

  // batched verify (requester and verifier side) vs the reference, incl. tampered lanes
  std::vector<PSCredential> vc = creds;
  if (N > 5) { vc[5].sig2 += pk.g; vc[1].sig1.clear(); }
  std::vector<uint8_t> v1 = rp.verify(vc, plain), v2 = wallet.verify(vc, plain);
  for (size_t j = 0; j < N; j++) {
    const bool want = static_cast<const ::PSVerifier&>(rp).verify(vc[j], plain[j]);
    EXPECT(want == (v1[j] != 0) && want == (v2[j] != 0));
  }

  // the same credentials in wire form (PSCredential::toBufferString), decompressed on the GPU
  {
    std::vector<PSBuffer> wire;
    for (auto& c : vc) wire.push_back(c.toBufferString());
    std::vector<uint8_t> v3 = rp.verify(wire, plain);
    for (size_t j = 0; j < N; j++) EXPECT(v3[j] == v1[j]);
  }

  // batched randomize_credential with host-supplied t vs t * sigma on the host
  std::vector<Fr> t(N);
  for (auto& x : t) x.setByCSPRNG();
  std::vector<PSCredential> rnd = wallet.randomize_credential(creds, t);
  for (size_t j = 0; j < N; j++) {
    G1 a, b;
    G1::mul(a, creds[j].sig1, t[j]);
    G1::mul(b, creds[j].sig2, t[j]);
    EXPECT(ser(rnd[j].sig1) == ser(a) && ser(rnd[j].sig2) == ser(b));
  }

  // User-ProveID on the host (reference), RP-VerifyID batched on the GPU vs the reference
  std::vector<IdProof> proofs, proofs2;
  for (size_t j = 0; j < N; j++) {
    proofs.push_back(users[j].el_passo_prove_id(creds[j], attrs[j], ads[j], "service", authority_pk, g, h));
    proofs2.push_back(users[j].el_passo_prove_id_without_id_retrieval(creds[j], attrs[j], ads[j], "service"));
  }
  if (N > 6) {
    proofs[2].c += 1; proofs[4].rs[1] += 1; proofs[6].sig2 += pk.g; proofs2[2].k += pk.gg; proofs2[4].attributes[4] += "x";
    proofs[0].E1.reset();
  }
  std::vector<uint8_t> w1 = rp.el_passo_verify_id(proofs, ads, "service", authority_pk, g, h);
  std::vector<uint8_t> w2 = rp.el_passo_verify_id_without_id_retrieval(proofs2, ads, "service");
  size_t accepted = 0;
  for (size_t j = 0; j < N; j++) {
    const ::PSVerifier& ref = rp;
    EXPECT(ref.el_passo_verify_id(proofs[j], ads[j], "service", authority_pk, g, h) == (w1[j] != 0));
    EXPECT(ref.el_passo_verify_id_without_id_retrieval(proofs2[j], ads[j], "service") == (w2[j] != 0));
    accepted += w1[j];
  }
  EXPECT(accepted > 0 && accepted < N);

  // ---- prover side on the GPU: the reference's methods under a replayed CSPRNG stream vs the batched
  //      overloads fed the same scalars; requests / proofs compared field by field in serialized form -------------
  {
    struct Replay {
      static std::vector<uint8_t>& buf() { static std::vector<uint8_t> b; return b; }
      static size_t& pos() { static size_t p = 0; return p; }
      static uint32_t read(void*, void* out, uint32_t n) {
        for (uint32_t i = 0; i < n; i++) static_cast<uint8_t*>(out)[i] = buf()[(pos()++) % buf().size()];
        return n;
      }
    };
    const size_t hcount = 2;
    std::vector<std::vector<Fr>> rq_rnd(N), pv_rnd(N), pv2_rnd(N);
    std::vector<PSCredRequest> ref_req;
    std::vector<IdProof> ref_pv, ref_pv2;
    std::vector<::PSRequester> u2(N, ::PSRequester(pk));
    mcl::fp::RandGen saved = mcl::fp::RandGen::get();
    for (size_t j = 0; j < N; j++) {
      // a per-lane byte stream; the scalars are what setByCSPRNG makes of consecutive 32-byte reads
      Replay::buf().resize(32 * 16);
      for (size_t i = 0; i < Replay::buf().size(); i++) Replay::buf()[i] = (uint8_t)(i * 131 + j * 17 + 7);
      auto draws = [&](size_t cnt) { std::vector<Fr> v(cnt); Replay::pos() = 0; for (auto& x : v) x.setByCSPRNG(); Replay::pos() = 0; return v; };
      mcl::fp::RandGen::setRandFunc(nullptr, Replay::read);
      rq_rnd[j] = draws(hcount + 2);
      ref_req.push_back(u2[j].el_passo_request_id(attrs[j], ads[j]));
      pv_rnd[j] = draws(hcount + 5);
      ref_pv.push_back(u2[j].el_passo_prove_id(creds[j], attrs[j], ads[j], "service", authority_pk, g, h));
      pv2_rnd[j] = draws(hcount + 3);
      ref_pv2.push_back(u2[j].el_passo_prove_id_without_id_retrieval(creds[j], attrs[j], ads[j], "service"));
      mcl::fp::RandGen::setRandGen(saved);
    }
    std::vector<PSCredRequest> got_req = wallet.el_passo_request_id(attrs, ads, rq_rnd);
    std::vector<IdProof> got_pv = wallet.el_passo_prove_id(creds, attrs, ads, "service", authority_pk, g, h, pv_rnd);
    std::vector<IdProof> got_pv2 = wallet.el_passo_prove_id_without_id_retrieval(creds, attrs, ads, "service", pv2_rnd);
    std::vector<Fr> t1(N);
    for (size_t j = 0; j < N; j++) {
      EXPECT(got_req[j].toBufferString() == ref_req[j].toBufferString());     // the wire form of the request, byte for byte
      EXPECT(got_pv[j].toBufferString() == ref_pv[j].toBufferString());
      EXPECT(got_pv2[j].toBufferString() == ref_pv2[j].toBufferString());
      t1[j] = rq_rnd[j][0];
    }
    // issue on the GPU for the GPU-made requests, unblind on the GPU, verify with the reference
    std::vector<PSCredential> iss2;
    std::vector<uint8_t> ok2 = idp.el_passo_provide_id(got_req, ads, u, iss2);
    std::vector<PSCredential> un2 = wallet.unblind_credential(iss2, t1);
    for (size_t j = 0; j < N; j++) {
      EXPECT(ok2[j]);
      PSCredential r = u2[j].unblind_credential(iss2[j]);                      // m_t1 of u2[j] is the replayed t1
      EXPECT(ser(un2[j].sig1) == ser(r.sig1) && ser(un2[j].sig2) == ser(r.sig2));
      EXPECT(static_cast<const ::PSVerifier&>(rp).verify(un2[j], plain[j]));
    }
  }

  // ---- sign-on and issuance straight from the WIRE (base64 text / PSBuffer bytes parsed on the GPU) ---------------
  {
    psb::PSVerifier exact(pk, 8, /*strict_sigma=*/false);      // bit-exact reference verdicts
    std::vector<std::string> b64, b64n;
    std::vector<PSBuffer> raw;
    for (size_t j = 0; j < N; j++) {
      PSBuffer b = proofs[j].toBufferString();
      raw.push_back(b);
      b64.push_back(b.toBase64());
      b64n.push_back(proofs2[j].toBufferString().toBase64());
    }
    if (N > 8) { b64[7] = b64[7].substr(0, 37); raw[8][0] = 2; }   // a truncated text, a wrong type byte
    std::vector<uint8_t> x1 = exact.el_passo_verify_id(b64, ads, "service", authority_pk, g, h);
    std::vector<uint8_t> x2 = exact.el_passo_verify_id(raw, ads, "service", authority_pk, g, h);
    std::vector<uint8_t> x3 = exact.el_passo_verify_id_without_id_retrieval(b64n, ads, "service");
    for (size_t j = 0; j < N; j++) {
      if (N > 8 && j == 7) { EXPECT(x1[j] == 0); EXPECT(x2[j] == w1[j]); }
      else if (N > 8 && j == 8) { EXPECT(x2[j] == 0); EXPECT(x1[j] == w1[j]); }
      else { EXPECT(x1[j] == w1[j]); EXPECT(x2[j] == w1[j]); }
      EXPECT(x3[j] == w2[j]);
    }
    // proofs hiding DIFFERENT numbers of attributes in one batch (lanes are grouped by rs.size())
    std::vector<IdProof> mixed = proofs;
    std::vector<size_t> three;
    for (size_t j = 9; j < N && j < 12; j++) {
      auto a3 = attrs[j];
      std::get<1>(a3[2]) = true;
      mixed[j] = users[j].el_passo_prove_id(creds[j], a3, ads[j], "service", authority_pk, g, h);
      three.push_back(j);
    }
    std::vector<uint8_t> m1 = exact.el_passo_verify_id(mixed, ads, "service", authority_pk, g, h);
    for (size_t j = 0; j < N; j++)
      EXPECT(static_cast<const ::PSVerifier&>(exact).el_passo_verify_id(mixed[j], ads[j], "service", authority_pk, g, h) == (m1[j] != 0));
    for (size_t j : three) EXPECT(mixed[j].rs.size() == 5 && m1[j] == 1);
    // strict default: sigma = (0, 0) with an honest NIZK passes the reference, the batch classes reject it unless told otherwise
    if (N > 12) {
      std::vector<IdProof> z = proofs;
      z[12].sig1.clear(); z[12].sig2.clear();
      EXPECT(static_cast<const ::PSVerifier&>(rp).el_passo_verify_id(z[12], ads[12], "service", authority_pk, g, h));
      EXPECT(exact.el_passo_verify_id(z, ads, "service", authority_pk, g, h)[12] == 1);
      EXPECT(rp.el_passo_verify_id(z, ads, "service", authority_pk, g, h)[12] == 0);
    }
    // the other direction: the device-side encoder (psb_wire_encode) against toBufferString() / toBase64(), byte for byte
    {
      std::vector<std::string> e64 = psb::to_wire_base64(proofs), e64n = psb::to_wire_base64(proofs2), r64 = psb::to_wire_base64(requests);
      std::vector<PSBuffer> eraw = psb::to_wire(proofs);
      for (size_t j = 0; j < N; j++) {
        EXPECT(e64[j] == proofs[j].toBufferString().toBase64());
        EXPECT(e64n[j] == proofs2[j].toBufferString().toBase64());
        EXPECT(r64[j] == requests[j].toBufferString().toBase64());
        EXPECT(eraw[j] == proofs[j].toBufferString());
      }
    }
    // issuance from wire requests
    std::vector<std::string> rq64;
    for (auto& r : requests) rq64.push_back(r.toBufferString().toBase64());
    std::vector<PSCredential> iss3;
    std::vector<uint8_t> ok3 = idp.el_passo_provide_id(rq64, ads, u, iss3);
    for (size_t j = 0; j < N; j++) {
      EXPECT(ok3[j] == ok[j]);
      if (ok[j]) EXPECT(ser(iss3[j].sig1) == ser(issued[j].sig1) && ser(iss3[j].sig2) == ser(issued[j].sig2));
    }
  }

  // ---- a signer built from LOADED key material (pk, X) + batched sign_commitment / sign_hybrid ------------------------
  {
    const G1 X = idp.*member_of(psb::detail::SignerSecretX());
    psb::PSSigner idp2(pk, X, 8);
    std::vector<PSCredential> iss4;
    std::vector<uint8_t> ok4 = idp2.el_passo_provide_id(requests, ads, u, iss4);
    for (size_t j = 0; j < N; j++) {
      EXPECT(ok4[j] == ok[j]);
      if (ok[j]) EXPECT(ser(iss4[j].sig2) == ser(issued[j].sig2));
    }
    PSCredential sc; EXPECT(static_cast<const ::PSSigner&>(idp2).el_passo_provide_id(requests[0], ads[0], sc));   // scalar methods share the key
    std::vector<G1> cm(N);
    for (size_t j = 0; j < N; j++) cm[j] = requests[j].A;
    std::vector<PSCredential> s5 = idp2.sign_commitment(cm, u);
    std::vector<std::vector<std::string>> hy(N);
    for (size_t j = 0; j < N; j++) for (size_t i = 0; i < n; i++) hy[j].push_back(i < 2 ? std::string() : plain[j][i]);
    std::vector<PSCredential> s6 = idp2.sign_hybrid(cm, hy, u);
    for (size_t j = 0; j < N; j++) {
      G1 ug, t2, full = cm[j];
      G1::mul(ug, pk.g, u[j]);
      G1::add(t2, X, cm[j]); G1::mul(t2, t2, u[j]);
      EXPECT(ser(s5[j].sig1) == ser(ug) && ser(s5[j].sig2) == ser(t2));
      for (size_t i = 2; i < n; i++) { Fr m; m.setHashOf(plain[j][i]); G1 y; G1::mul(y, pk.Yi[i], m); G1::add(full, full, y); }
      G1::add(t2, X, full); G1::mul(t2, t2, u[j]);
      EXPECT(ser(s6[j].sig1) == ser(ug) && ser(s6[j].sig2) == ser(t2));
      if (ok[j]) EXPECT(ser(s6[j].sig2) == ser(issued[j].sig2));   // provide_id = NIZK check + sign_hybrid
    }
  }

  // a lane with the wrong number of attributes is DATA (verdict 0), not an exception: the reference indexes m_pk.YYi out
  // of range there (src/ps-verifier.cc:26), and one malformed lane must not abort the honest ones
  {
    std::vector<std::vector<std::string>> p2 = plain;
    p2[0].pop_back();
    std::vector<uint8_t> v9 = rp.verify(vc, p2);
    EXPECT(v9[0] == 0);
    for (size_t j = 1; j < N; j++) EXPECT(v9[j] == v1[j]);
    std::vector<IdProof> pz = proofs;
    pz[1].attributes.pop_back();
    std::vector<uint8_t> w9 = rp.el_passo_verify_id(pz, ads, "service", authority_pk, g, h);
    EXPECT(w9[1] == 0);
    for (size_t j = 0; j < N; j++) if (j != 1) EXPECT(w9[j] == w1[j]);
    bool threw = false;       // the CALL is still checked: one attribute list per credential
    try { p2.pop_back(); rp.verify(vc, p2); } catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
  }

  std::printf("test_ps_batch: %zu lanes, %llu kernel launches, %d failures\n", N, (unsigned long long)psb_launch_count(), fails);
  return fails ? 1 : 0;
}
