// tests/hostsim/hostsim.cc -- TEST INFRASTRUCTURE ONLY.
// Compiles the engine's __host__ __device__ arithmetic headers for the CPU (portable 64-bit
// arithmetic replaces the PTX carry chains) so that the CPU test-suite (`pytest -m "not gpu"`)
// can check the tower / curve / pairing LOGIC against the oracle without a GPU.  This library is
// never linked into, loaded by, or used as a fallback for libpsb.so.
#include <stddef.h>
#define PSB_COUNT_OPS 1   // host-side call counters of fp_mul / fp_sqr / fp_dot2 / fp_inv (csrc/fp.cuh): the MACs a lane executes
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

// ---- protocol lane functions (csrc/protocol.cuh) on the host: table building + one lane at a time ------
#include <cstring>
#include <vector>
#include "../../ps-signature-and-el-passo_b200/csrc/prover.cuh"
#include "../../ps-signature-and-el-passo_b200/csrc/hash_to_curve.cuh"
#include "../../ps-signature-and-el-passo_b200/csrc/wire.cuh"
namespace {
using namespace psb;
constexpr int kG1U = 3 * PSB_NL, kG2U = 6 * PSB_NL;   // u32 words of a Jacobian G1 / G2 point
// same table geometry as k_window_bases / k_build_table: entry (win, d) = d * 2^(w win) * B, affine
template <class F>
void host_table(std::vector<Aff<F>>& out, const Jac<F>& base, int w) {
  const int nwin = fixed_nwin(w);
  const uint32_t half = 1u << (w - 1);
  Jac<F> cur = base;
  for (int j = 0; j < nwin; j++) {
    Jac<F> acc, nrm;
    pt_set_zero(acc);
    for (uint32_t d = 1; d <= half; d++) {
      pt_add(acc, acc, cur);
      pt_normalize(nrm, acc);
      Aff<F> e; e.x = nrm.x; e.y = nrm.y;
      out.push_back(e);
    }
    for (int t = 0; t < w; t++) pt_dbl(cur, cur);
  }
}
// fixed-base sum probe: acc0 + sum_i k_i B_i over host-built tables, once through the plain mixed-addition chain
// (pt_fixed_mul_acc) and once through the batched affine pair additions (AffBatch, curve.cuh) -- both normalised
template <class F>
static void fixed_msm(int w, int nbases, const uint32_t* bases, const uint32_t* k_mont, const uint32_t* acc0, uint32_t* out_plain,
                      uint32_t* out_aff, int two_levels) {
  constexpr int U = sizeof(Jac<F>) / 4;
  std::vector<Aff<F>> t;
  for (int i = 0; i < nbases; i++) { Jac<F> b; ld(b, bases + U * i); host_table(t, b, w); }
  const size_t pb = (size_t)fixed_nwin(w) << (w - 1);
  Jac<F> a, c, n;
  ld(a, acc0); c = a;
  AffBatch<F> batch;
  AffPts<F> level2;
  aff_init(batch, nbases * fixed_nwin(w), two_levels == 2 ? kAffMaxEntries : t.size(), t.data(), (const Aff<F>*)nullptr, (const Aff<F>*)nullptr,
           two_levels == 1 ? &level2 : nullptr);   // two_levels == 2: pretend the table is too large (plain chain)
  for (int i = 0; i < nbases; i++) {
    uint32_t k[8]; Fr km; ld(km, k_mont + 8 * i); fr_load_normal(k, &km);
    pt_fixed_mul_acc(a, t.data() + i * pb, k, w);
    aff_push_fixed_mul(c, batch, 0, i * pb, k, w);
  }
  aff_flush(c, batch);
  pt_normalize(n, a); st(out_plain, n);
  pt_normalize(n, c); st(out_aff, n);
}
// second level of the batched affine sums on its own: acc0 + sum of the given normalised points (present[i] = 0: an empty slot)
template <class F>
static void aff_l2_probe(int npts, const uint32_t* pts, const uint8_t* present, const uint32_t* acc0, uint32_t* out) {
  constexpr int U = sizeof(Jac<F>) / 4;
  Jac<F> a, n;
  ld(a, acc0);
  AffPts<F> l;
  l.cnt = 0; l.absent = 0;
  for (int i = 0; i < npts; i++) {
    Jac<F> p; ld(p, pts + U * i);
    Aff<F> e; e.x = p.x; e.y = p.y;
    aff_l2_push(a, l, present[i] ? &e : (const Aff<F>*)nullptr);
  }
  aff_flush_l2(a, l);
  pt_normalize(n, a); st(out, n);
}
}  // namespace

extern "C" {
// fixed-base multiplication probe: out = acc0 + k * B via a host-built table (k: Fr Montgomery)
void hostsim_fixed_mul_g1(int w, const uint32_t* B, const uint32_t* k_mont, uint32_t* out) {
  G1J b; ld(b, B);
  std::vector<G1A> t; host_table(t, b, w);
  uint32_t k[8]; Fr km; ld(km, k_mont); fr_load_normal(k, &km);
  G1J acc; pt_set_zero(acc);
  pt_fixed_mul_acc(acc, t.data(), k, w);
  G1J n; pt_normalize(n, acc); st(out, n);
}

void hostsim_aff_l2(int is_g2, int npts, const uint32_t* pts, const uint8_t* present, const uint32_t* acc0, uint32_t* out) {
  if (is_g2) aff_l2_probe<Fp2>(npts, pts, present, acc0, out); else aff_l2_probe<Fp>(npts, pts, present, acc0, out);
}
// executed Fp-level operations of one psb_verify lane, phase by phase (k_verify_msm / k_verify_miller / k_verify_final run the
// same functions): out[phase * 4 + {mul, sqr, dot2, inv}].  The MSM phase is counted on host-built tables of window w (the
// count depends on n and nwin only); `levels` as PSB_MSM_AFFINE (0 plain chain, 1 / 2 batched affine levels).
void hostsim_count_verify_ops(int n, int w, int levels, const uint32_t* gg, const uint32_t* XX, const uint32_t* YY, const uint32_t* k_mont,
                              const uint32_t* sig1, const uint32_t* sig2, unsigned long long* out) {
  std::vector<G2A> t;
  G2J b2, ggj;
  for (int i = 0; i < n; i++) { ld(b2, YY + kG2U * i); host_table(t, b2, w); }
  ld(ggj, gg);
  std::vector<FixedLine> lines(kFixedLineSlots);
  G2A q; q.x = ggj.x; q.y = ggj.y;
  precompute_fixed_lines(lines.data(), q);
  const size_t pb = (size_t)fixed_nwin(w) << (w - 1);
  auto snap = [&](int phase) { for (int j = 0; j < 4; j++) { out[phase * 4 + j] = opcount::c[j]; opcount::c[j] = 0; } };
  for (int j = 0; j < 4; j++) opcount::c[j] = 0;
  G2J K; ld(K, XX);
  AffBatch<Fp2> batch;
  AffPts<Fp2> level2;
  if (levels) aff_init(batch, n * fixed_nwin(w), t.size(), t.data(), (const G2A*)nullptr, (const G2A*)nullptr, levels > 1 ? &level2 : nullptr);
  for (int i = 0; i < n; i++) {
    uint32_t k[8]; Fr km; ld(km, k_mont + 8 * i); fr_load_normal(k, &km);
    if (levels) aff_push_fixed_mul(K, batch, 0, i * pb, k, w); else pt_fixed_mul_acc(K, t.data() + i * pb, k, w);
  }
  if (levels) aff_flush(K, batch);
  snap(0);
  G1J s1, s2; ld(s1, sig1); ld(s2, sig2);
  Fp x1, y1, x2, y2;
  g1_affine_for_pairing(x1, y1, s1);
  g1_affine_for_pairing(x2, y2, s2);
  fp_neg(y2, y2);
  Fp12 f, e;
  miller_loop2(f, x1, y1, K, x2, y2, lines.data(), true);
  snap(1);
  final_exp(e, f);
  snap(2);
}
void hostsim_fixed_msm(int is_g2, int w, int nbases, const uint32_t* bases, const uint32_t* k_mont, const uint32_t* acc0,
                       uint32_t* out_plain, uint32_t* out_aff, int two_levels) {
  if (is_g2) fixed_msm<Fp2>(w, nbases, bases, k_mont, acc0, out_plain, out_aff, two_levels);
  else fixed_msm<Fp>(w, nbases, bases, k_mont, acc0, out_plain, out_aff, two_levels);
}

void hostsim_provide_id(int n, int w, const uint32_t* g, const uint32_t* X, const uint32_t* Y, size_t N,
                        const uint32_t* A, const uint32_t* c, const uint32_t* rs, int per, const uint8_t* blob,
                        const uint64_t* off, const uint8_t* ad_blob, const uint64_t* ad_off, const uint32_t* u,
                        uint8_t* verdict, uint32_t* sig1, uint32_t* sig2) {
  std::vector<G1A> tbl;
  G1J b; ld(b, g); host_table(tbl, b, w);
  for (int i = 0; i < n; i++) { ld(b, Y + kG1U * i); host_table(tbl, b, w); }
  G1J Xs; ld(Xs, X);
  for (size_t j = 0; j < N; j++) {
    G1J a, s1, s2; ld(a, A + kG1U * j);
    bool ok = provide_id_lane(n, TblGeom{w}, tbl.data(), Xs, a, (const Fr*)(c + 8 * j), (const Fr*)(rs + 8 * per * j), per,
                              blob, off + j * n, ad_blob + ad_off[j], (size_t)(ad_off[j + 1] - ad_off[j]),
                              (const Fr*)(u + 8 * j), s1, s2);
    verdict[j] = ok;
    st(sig1 + kG1U * j, s1); st(sig2 + kG1U * j, s2);
  }
}

void hostsim_verify_id(int n, int w, const uint32_t* gg, const uint32_t* XX, const uint32_t* YY, size_t N,
                       const uint32_t* sig1, const uint32_t* sig2, const uint32_t* k, const uint32_t* phi,
                       const uint32_t* E1, const uint32_t* E2, const uint32_t* c, const uint32_t* rs, int per,
                       const uint8_t* blob, const uint64_t* off, const uint8_t* ad_blob, const uint64_t* ad_off,
                       const uint32_t* service_pt, const uint32_t* y, const uint32_t* g, const uint32_t* h, int with_id,
                       uint8_t* nizk, uint8_t* verdict) {
  std::vector<G2A> tYY, tAux;
  std::vector<G1A> tB;
  G2J b2; G1J b1;
  for (int i = 0; i < n; i++) { ld(b2, YY + kG2U * i); host_table(tYY, b2, w); }
  G2J ggj; ld(ggj, gg); host_table(tAux, ggj, w);
  ld(b2, XX); host_table(tAux, b2, w);
  ld(b1, service_pt); host_table(tB, b1, w);
  if (with_id) { ld(b1, g); host_table(tB, b1, w); ld(b1, y); host_table(tB, b1, w); ld(b1, h); host_table(tB, b1, w); }
  std::vector<FixedLine> lines(kFixedLineSlots);
  G2A q; q.x = ggj.x; q.y = ggj.y;
  precompute_fixed_lines(lines.data(), q);
  for (size_t j = 0; j < N; j++) {
    G2J kj, Vk, K; ld(kj, k + kG2U * j);
    G1J ph, e1, e2, Vphi, VE1, VE2, s1, s2;
    ld(ph, phi + kG1U * j);
    if (with_id) { ld(e1, E1 + kG1U * j); ld(e2, E2 + kG1U * j); }
    const Fr* cj = (const Fr*)(c + 8 * j);
    const Fr* rj = (const Fr*)(rs + 8 * per * j);
    bool ok = verify_id_g2_lane(n, TblGeom{w}, tYY.data(), tAux.data(), kj, cj, rj, per, with_id, blob, off + j * n, Vk, K);
    verify_id_g1_lane(TblGeom{w}, tB.data(), ph, &e1, &e2, cj, rj, per, with_id, Vphi, VE1, VE2);
    ok = ok && verify_id_hash_lane(kj, ph, &e1, &e2, Vk, Vphi, VE1, VE2, with_id, cj, ad_blob + ad_off[j],
                                   (size_t)(ad_off[j + 1] - ad_off[j]));
    nizk[j] = ok;
    ld(s1, sig1 + kG1U * j); ld(s2, sig2 + kG1U * j);
    Fp x1, y1, x2, y2;
    g1_affine_for_pairing(x1, y1, s1);
    g1_affine_for_pairing(x2, y2, s2);
    fp_neg(y2, y2);
    Fp12 f, e;
    miller_loop2(f, x1, y1, K, x2, y2, lines.data(), true);
    final_exp(e, f);
    verdict[j] = ok && fp12_is_one(e);
  }
}

// point decompression probes (csrc/protocol.cuh): returns 1 on success
int hostsim_g1_deserialize(const uint8_t* b, uint32_t* out) { G1J P; const bool r = g1_deserialize(P, b); st(out, P); return r; }
int hostsim_g2_deserialize(const uint8_t* b, uint32_t* out) { G2J P; const bool r = g2_deserialize(P, b); st(out, P); return r; }
// ---- prover side (csrc/prover.cuh) ------------------------------------------------------------------------
void hostsim_request_id(int n, int w, const uint32_t* g, const uint32_t* Y, size_t N, const uint8_t* hide, const uint8_t* blob,
                        const uint64_t* off, const uint8_t* ad_blob, const uint64_t* ad_off, const uint32_t* rnd, uint32_t* A,
                        uint32_t* c, uint32_t* rs) {
  std::vector<G1A> tbl;
  G1J b; ld(b, g); host_table(tbl, b, w);
  for (int i = 0; i < n; i++) { ld(b, Y + kG1U * i); host_table(tbl, b, w); }
  int h = 0; for (int i = 0; i < n; i++) h += hide[i] ? 1 : 0;
  for (size_t j = 0; j < N; j++) {
    G1J a; Fr cc;
    request_id_lane(n, TblGeom{w}, tbl.data(), hide, blob, off + j * n, ad_blob + ad_off[j], (size_t)(ad_off[j + 1] - ad_off[j]),
                    (const Fr*)(rnd + 8 * (h + 2) * j), a, cc, (Fr*)(rs + 8 * (h + 1) * j));
    st(A + kG1U * j, a); st(c + 8 * j, cc);
  }
}
void hostsim_unblind(size_t N, const uint32_t* sig1, const uint32_t* sig2, const uint32_t* t1, uint32_t* out2) {
  for (size_t j = 0; j < N; j++) {
    G1J a, b, r; ld(a, sig1 + kG1U * j); ld(b, sig2 + kG1U * j);
    unblind_lane(r, a, b, (const Fr*)(t1 + 8 * j));
    st(out2 + kG1U * j, r);
  }
}
void hostsim_prove_id(int n, int w, const uint32_t* gg, const uint32_t* XX, const uint32_t* YY, size_t N, const uint32_t* sig1,
                      const uint32_t* sig2, const uint8_t* hide, const uint8_t* blob, const uint64_t* off, const uint8_t* ad_blob,
                      const uint64_t* ad_off, const uint32_t* service_pt, const uint32_t* y, const uint32_t* g, const uint32_t* h_pt,
                      int with_id, const uint32_t* rnd, uint32_t* o_sig1, uint32_t* o_sig2, uint32_t* o_k, uint32_t* o_phi,
                      uint32_t* o_E1, uint32_t* o_E2, uint32_t* o_c, uint32_t* o_rs) {
  std::vector<G2A> tYY, tAux;
  std::vector<G1A> tB;
  G2J b2, XXj; G1J b1;
  for (int i = 0; i < n; i++) { ld(b2, YY + kG2U * i); host_table(tYY, b2, w); }
  ld(b2, gg); host_table(tAux, b2, w);
  ld(XXj, XX); host_table(tAux, XXj, w);
  ld(b1, service_pt); host_table(tB, b1, w);
  if (with_id) { ld(b1, g); host_table(tB, b1, w); ld(b1, y); host_table(tB, b1, w); ld(b1, h_pt); host_table(tB, b1, w); }
  int h = 0; for (int i = 0; i < n; i++) h += hide[i] ? 1 : 0;
  const int rper = prove_rnd_per_lane(h, with_id), per = h + (with_id ? 2 : 1);
  for (size_t j = 0; j < N; j++) {
    const Fr* rj = (const Fr*)(rnd + 8 * rper * j);
    G2J k, Vk; G1J s1, s2, o1, o2, W[6]; Fr cc;
    ld(s1, sig1 + kG1U * j); ld(s2, sig2 + kG1U * j);
    prove_id_g2_lane(n, TblGeom{w}, tYY.data(), tAux.data(), XXj, hide, blob, off + j * n, rj, h, with_id, k, Vk);
    prove_id_g1_lane(TblGeom{w}, tB.data(), s1, s2, blob, off + j * n, rj, h, with_id, o1, o2, W);
    prove_id_hash_lane(n, hide, blob, off + j * n, ad_blob + ad_off[j], (size_t)(ad_off[j + 1] - ad_off[j]), rj, h, with_id, k, Vk,
                       W, cc, (Fr*)(o_rs + 8 * per * j));
    st(o_sig1 + kG1U * j, o1); st(o_sig2 + kG1U * j, o2); st(o_k + kG2U * j, k); st(o_phi + kG1U * j, W[0]); st(o_c + 8 * j, cc);
    if (with_id) { st(o_E1 + kG1U * j, W[2]); st(o_E2 + kG1U * j, W[3]); }
  }
}
// hashAndMapToG1 probes (csrc/hash_to_curve.cuh)
void hostsim_sha512(const uint8_t* msg, size_t len, uint8_t* out64) {
  uint64_t h[8]; sha512(h, msg, len);
  for (int i = 0; i < 64; i++) out64[i] = (uint8_t)(h[i >> 3] >> (56 - 8 * (i & 7)));
}
int hostsim_hash_to_g1(const uint8_t* msg, size_t len, uint32_t* out) { G1J P; const bool r = hash_and_map_to_g1(P, msg, len); st(out, P); return r; }
int hostsim_map_to_g1(const uint32_t* t, uint32_t* out) { Fp tt; ld(tt, t); G1J P, Q; const bool r = map_to_g1(P, tt); if (r) g1_clear_cofactor(Q, P); else pt_set_zero(Q); st(out, Q); return r; }

// ---- wire formats (csrc/wire.cuh): base64 + TLV walk + point decompression of ONE message, as the device kernels
//      k_wire_base64 / k_wire_parse / k_wire_points do it.  kind 0 = IdProof, 1 = PSCredRequest.  Returns `parsed`.
int hostsim_wire_parse(int kind, int n, const uint8_t* in, size_t len, int base64, uint32_t* g1pts /*5 x G1*/, uint32_t* g2pt,
                       uint32_t* c, uint32_t* rs /*(n + 2) x Fr*/, int* per, uint8_t* attr_out, uint64_t* attr_off /*n + 1*/,
                       int* has_e) {
  std::vector<uint8_t> raw(in, in + len);
  if (base64) { raw.resize(len + 4); raw.resize(base64_decode_lane(raw.data(), in, len)); }
  uint32_t pos[W_SLOTS];
  Fr cc;
  for (int i = 0; i < 8; i++) cc.v[i] = 0;
  std::vector<Fr> r(n + 2);
  bool he = false;
  bool ok = kind == 0 ? parse_idproof_lane(raw.data(), raw.size(), n, pos, cc, r.data(), *per, attr_out, attr_off, 0, he)
                      : parse_request_lane(raw.data(), raw.size(), n, pos, cc, r.data(), *per, attr_out, attr_off, 0);
  *has_e = he;
  for (int i = 0; i < 8; i++) c[i] = cc.v[i];
  for (int j = 0; j < *per; j++) for (int i = 0; i < 8; i++) rs[8 * j + i] = r[j].v[i];
  for (int slot = 0; slot < W_SLOTS; slot++) {
    if (slot == W_K) {
      G2J P; pt_set_zero(P);
      if (pos[slot] != kWireAbsent && !g2_deserialize(P, raw.data() + pos[slot])) ok = false;
      st(g2pt, P);
    } else {
      G1J P; pt_set_zero(P);
      if (pos[slot] != kWireAbsent && !g1_deserialize(P, raw.data() + pos[slot])) ok = false;
      st(g1pts + kG1U * slot, P);
    }
  }
  return ok ? 1 : 0;
}
// the encoder of ONE message (csrc/wire.cuh encode_message_lane / base64_encode_lane, kernels k_wire_encode / k_wire_base64_encode):
// g1pts = sig1 sig2 phi E1 E2 (kind 0) or A (kind 1); returns the bytes written to out
size_t hostsim_wire_encode(int kind, int n, const uint32_t* g1pts, const uint32_t* g2pt, int has_e, const uint32_t* c, const uint32_t* rs,
                           int per, const uint8_t* attr, const uint64_t* aoff, int base64, uint8_t* out) {
  G1J P[5]; G2J K; Fr cc; std::vector<Fr> r(per + 1);
  const int np = kind == 0 ? 5 : 1;
  for (int i = 0; i < np; i++) ld(P[i], g1pts + kG1U * i);
  if (kind == 0) ld(K, g2pt);
  for (int i = 0; i < 8; i++) cc.v[i] = c[i];
  for (int j = 0; j < per; j++) for (int i = 0; i < 8; i++) r[j].v[i] = rs[8 * j + i];
  const G1J* pts[5] = {&P[0], &P[1], &P[2], has_e ? &P[3] : nullptr, has_e ? &P[4] : nullptr};
  const size_t want = wire_message_size(kind, n, per, aoff, has_e != 0);
  std::vector<uint8_t> raw(want + 8);
  const size_t got = encode_message_lane(raw.data(), kind, n, pts, &K, cc, r.data(), per, attr, aoff);
  if (got != want) return 0;
  if (!base64) { memcpy(out, raw.data(), got); return got; }
  base64_encode_lane(out, raw.data(), got);
  return base64_encoded_size(got);
}
size_t hostsim_base64_decode(const uint8_t* in, size_t len, uint8_t* out) { return base64_decode_lane(out, in, len); }

// PSSigner::sign_hybrid / sign_commitment lanes (csrc/protocol.cuh sign_lane)
void hostsim_sign(int n, int na, int w, const uint32_t* g, const uint32_t* X, const uint32_t* Y, size_t N, const uint32_t* Cm,
                  const uint8_t* blob, const uint64_t* off, const uint32_t* u, uint32_t* sig1, uint32_t* sig2) {
  std::vector<G1A> tbl;
  G1J b; ld(b, g); host_table(tbl, b, w);
  for (int i = 0; i < n; i++) { ld(b, Y + kG1U * i); host_table(tbl, b, w); }
  G1J Xs; ld(Xs, X);
  for (size_t j = 0; j < N; j++) {
    G1J cm, s1, s2; ld(cm, Cm + kG1U * j);
    sign_lane(na, TblGeom{w}, tbl.data(), Xs, cm, blob, na ? off + j * na : nullptr, (const Fr*)(u + 8 * j), s1, s2);
    st(sig1 + kG1U * j, s1); st(sig2 + kG1U * j, s2);
  }
}
}
