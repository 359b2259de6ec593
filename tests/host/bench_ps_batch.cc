// tests/host/bench_ps_batch.cc -- MEASUREMENT of the C++ drop-in at the headline shape: psb::PSVerifier::verify over
// std::vector<PSCredential> + std::vector<std::vector<std::string>> (what a user of the reference's classes holds), 5
// attributes, N lanes (default 2^20), end to end from those containers to the verdict bytes, with the host-side
// marshalling (flattening the attribute strings; the credentials are passed in place) timed separately.
// Linked against the unmodified reference objects and libpsb.so (oracle/Makefile `hostbench`).  One JSON line.
#include <chrono>
#include <cstdio>
#include <tuple>

#include "ps_batch.hpp"

using namespace mcl::bls12;
typedef std::chrono::steady_clock Clock;
static double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int main(int argc, char** argv) {
  const size_t N = argc > 1 ? std::stoul(argv[1]) : (size_t)1 << 20;
  const int wbits = argc > 2 ? std::stoi(argv[2]) : 20;
  const size_t n = 5, D = 256;
  initPairing(mcl::BLS12_381);
  try { psb::init(); } catch (const std::exception& e) { std::printf("{\"error\": \"%s\"}\n", e.what()); return 3; }
  G1 g; G2 gg;
  hashAndMapToG1(g, "abc");
  hashAndMapToG2(gg, "edf");
  psb::PSSigner idp(n, g, gg, 8);
  PSPubKey pk = idp.key_gen();
  psb::PSVerifier rp(pk, wbits);
  psb::PSRequester wallet(pk, 8);
  // D honest credentials through the reference's own flow (request -> provide -> unblind), on the host
  std::vector<PSCredential> base(D);
  std::vector<std::vector<std::string>> base_attrs(D);
  for (size_t j = 0; j < D; j++) {
    ::PSRequester user(pk);
    std::vector<std::tuple<std::string, bool>> a;
    for (size_t i = 0; i < n; i++) { a.push_back(std::make_tuple("a" + std::to_string(i) + ":" + std::to_string(j), i < 2)); base_attrs[j].push_back(std::get<0>(a[i])); }
    PSCredRequest r = user.el_passo_request_id(a, "ad");
    PSCredential s;
    if (!static_cast<const ::PSSigner&>(idp).el_passo_provide_id(r, "ad", s)) return 2;
    base[j] = user.unblind_credential(s);
  }
  // N distinct lanes: tiled and re-randomised on the GPU (t sigma is a valid credential for the same attributes)
  std::vector<PSCredential> creds(N);
  std::vector<std::vector<std::string>> attrs(N);
  std::vector<Fr> t(N);
  for (size_t j = 0; j < N; j++) { creds[j] = base[j % D]; attrs[j] = base_attrs[j % D]; t[j].setByCSPRNG(); }
  creds = wallet.randomize_credential(creds, t);
  for (size_t j = 1023; j < N; j += 1024) creds[j].sig2 += pk.g;          // tampered lanes
  std::vector<uint8_t> v = rp.verify(creds, attrs);                       // warm-up: staging buffers sized, tables hot
  size_t bad = 0;
  for (size_t j = 0; j < N; j++) bad += (v[j] != ((j % 1024) == 1023 ? 0 : 1));
  const int reps = 3;
  double best_e2e = 1e30, best_marshal = 1e30, best_call = 1e30;
  for (int r = 0; r < reps; r++) {
    auto t0 = Clock::now();
    v = rp.verify(creds, attrs);
    auto t1 = Clock::now();
    best_e2e = std::min(best_e2e, secs(t0, t1));
    std::vector<uint8_t> blob, badl; std::vector<uint64_t> off;
    t0 = Clock::now();
    psb::detail::flatten_lanes(N, n, [&](size_t j) -> const std::vector<std::string>& { return attrs[j]; }, blob, off, badl);
    t1 = Clock::now();
    best_marshal = std::min(best_marshal, secs(t0, t1));
  }
  {   // the C-ABI call alone on pre-flattened inputs
    std::vector<uint8_t> blob, badl, vv(N); std::vector<uint64_t> off;
    psb::detail::flatten_lanes(N, n, [&](size_t j) -> const std::vector<std::string>& { return attrs[j]; }, blob, off, badl);
    psb::PSVerifier* p = &rp; (void)p;
    for (int r = 0; r < reps; r++) {
      auto t0 = Clock::now();
      std::vector<uint8_t> x = rp.verify(creds, attrs);
      auto t1 = Clock::now();
      (void)x;
      best_call = std::min(best_call, secs(t0, t1) - best_marshal);
    }
  }
  std::printf("{\"bench\": \"psb::PSVerifier::verify(std::vector<PSCredential>, attributes)\", \"lanes\": %zu, \"n_attrs\": %zu, "
              "\"window_bits\": %d, \"e2e_verifications_per_s\": %.1f, \"e2e_seconds\": %.4f, \"marshal_seconds\": %.4f, "
              "\"marshal_threads\": %u, \"marshal_share\": %.4f, \"verdict_mismatches\": %zu, \"gpu_launches\": %llu}\n",
              N, n, wbits, N / best_e2e, best_e2e, best_marshal, psb::detail::marshal_threads(N), best_marshal / best_e2e, bad,
              (unsigned long long)psb_launch_count());
  return bad ? 1 : 0;
}
