// prover.cuh -- per-lane PROVER side of EL PASSO: the requester's blind-issuance request,
// credential unblinding and the sign-on proof, as __host__ __device__ lane functions (kernels.cuh wraps them one
// thread per lane; the CPU test library compiles the same functions for the host: tests/hostsim).
//
// Reference semantics:
//   src/ps-requester.cc:19-97    PSRequester::el_passo_request_id
//   src/ps-requester.cc:99-113   PSRequester::unblind_credential
//   src/ps-requester.cc:150-310  PSRequester::el_passo_prove_id
//   src/ps-requester.cc:312-432  PSRequester::el_passo_prove_id_without_id_retrieval
// The reference draws its blinding / commitment scalars from mcl's CSPRNG (Fr::setByCSPRNG); the batch API takes
// them from the host IN THE REFERENCE'S DRAW ORDER ("host-supplied deterministic scalars"), so that the same scalars
// give byte-identical requests and proofs.  Every group element is computed from fixed-base window tables (the
// reference runs a full GLV multiplication for each), the points that enter the Fiat-Shamir hash are normalised
// with one shared inversion, and the outputs are normalised (z = 1) -- the reference returns raw Jacobian
// coordinates, which its own serialisation normalises before anything leaves the process.
#pragma once
#include "protocol.cuh"

namespace psb {

PSB_HD PSB_INL void fr_from_normal(Fr& r, const uint32_t k[8]) {
  Fr n;
  for (int i = 0; i < 8; i++) n.v[i] = k[i];
  fr_to_mont(r, n);
}
// Schnorr response  rnd - secret * c   (all Montgomery; ps-requester.cc:80-88, 283-296)
PSB_HD PSB_INL void fr_response(Fr& out, const Fr& rnd, const Fr& secret, const Fr& c) {
  Fr t;
  fr_mul(t, secret, c);
  fr_sub(out, rnd, t);
}

// ---- PSRequester::el_passo_request_id, one lane ------------------------------------------------------------
//   tblG1: fixed-base tables of [g, Y_0 .. Y_{n-1}];  hide[i] != 0: attribute i is committed, else sent in clear
//   rnd: h + 2 scalars (Montgomery) in draw order: t1 (the blinding kept by the requester), the commitment
//        randomness of g, then one per hidden attribute (ps-requester.cc:38, 48-49, 62-63)
//   outputs: A normalised, c and rs[h + 1] in Montgomery form
PSB_HD PSB_NOINL void request_id_lane(int n, TblGeom tg, const G1A* tblG1, const uint8_t* hide, const uint8_t* blob,
                                       const uint64_t* off, const uint8_t* ad, size_t ad_len, const Fr* rnd, G1J& A_out,
                                       Fr& c_out, Fr* rs_out) {
  const size_t pb = tg.per_base();
  uint32_t k[8];
  G1J A, V;
  pt_set_zero(A);
  pt_set_zero(V);
  fr_load_normal(k, rnd); pt_fixed_mul_acc(A, tblG1, k, tg.w);          // A = t1 g
  fr_load_normal(k, rnd + 1); pt_fixed_mul_acc(V, tblG1, k, tg.w);      // V = r0 g
  int j = 0;
  for (int i = 0; i < n; i++) {
    if (!hide[i]) continue;
    fr_set_hash_of(k, blob + off[i], (size_t)(off[i + 1] - off[i]));
    fr_from_normal(rs_out[1 + j], k);                                   // m_i parked here until c is known
    pt_fixed_mul_acc(A, tblG1 + (size_t)(1 + i) * pb, k, tg.w);         // A += m_i Y_i
    fr_load_normal(k, rnd + 2 + j);
    pt_fixed_mul_acc(V, tblG1 + (size_t)(1 + i) * pb, k, tg.w);         // V += r_i Y_i
    j++;
  }
  g1_normalize2(A, V);
  Sha256 s;
  sha256_init(s);
  sha_put_g1_hex(s, A);
  sha_put_g1_hex(s, V);
  uint32_t cn[8];
  challenge_finish(cn, s, ad, ad_len);
  Fr c, t1 = rnd[0], r0 = rnd[1];
  fr_from_normal(c, cn);
  fr_response(rs_out[0], r0, t1, c);
  for (int q = 0; q < j; q++) {
    Fr m = rs_out[1 + q], r = rnd[2 + q];
    fr_response(rs_out[1 + q], r, m, c);
  }
  A_out = A;
  c_out = c;
}

// ---- PSRequester::unblind_credential, one lane: sig2' = sig2 - t1 sig1 (normalised) ---------------------------
PSB_HD PSB_NOINL void unblind_lane(G1J& out2, const G1J& sig1, const G1J& sig2, const Fr* t1_mont) {
  uint32_t k[8];
  fr_load_normal(k, t1_mont);
  G1J P = sig1, T, S = sig2;
  pt_mul(T, P, k);
  pt_neg(T, T);
  pt_add(T, S, T);
  pt_normalize(out2, T);
}

// ---- PSRequester::el_passo_prove_id[_without_id_retrieval], one lane, three steps -----------------------------
// Scalars of one lane, Montgomery, in the reference's draw order (ps-requester.cc:163-165, 172, 238-253, 261):
//   with id retrieval:    t, r, epsilon, hid_0 .. hid_{h-1}, random2, random3      (h + 5)
//   without id retrieval: t, r,          hid_0 .. hid_{h-1}, random2               (h + 3)
// `_randomnesses` of the reference = (hid_0 .., random2[, random3]) = rnd + (with_id ? 3 : 2).
PSB_HD PSB_INL int prove_rnd_per_lane(int h, int with_id) { return h + (with_id ? 5 : 3); }

// step 1 (G2): k = XX + sum_hidden H(attr_i) YY_i + t gg;   V_k = XX + sum_hidden hid_j YY_i + random2 gg
PSB_HD PSB_NOINL void prove_id_g2_lane(int n, TblGeom tg, const G2A* tblYY, const G2A* tblAux /*[gg, XX]*/, const G2J& XX,
                                        const uint8_t* hide, const uint8_t* blob, const uint64_t* off, const Fr* rnd, int h,
                                        int with_id, G2J& k_out, G2J& Vk_out) {
  const size_t pb = tg.per_base();
  const Fr* rr = rnd + (with_id ? 3 : 2);
  uint32_t s[8];
  G2J K = XX, V = XX;
  // the table entries of k and of V_k are summed pairwise in affine coordinates (AffBatch, curve.cuh)
  AffBatch<Fp2> bK, bV;
  aff_init(bK, (h + 1) * fixed_nwin(tg.w), (size_t)(n > 2 ? n : 2) * pb, tblYY, tblAux);
  aff_init(bV, (h + 1) * fixed_nwin(tg.w), (size_t)(n > 2 ? n : 2) * pb, tblYY, tblAux);
  int j = 0;
  for (int i = 0; i < n; i++) {
    if (!hide[i]) continue;
    fr_set_hash_of(s, blob + off[i], (size_t)(off[i + 1] - off[i]));
    aff_push_fixed_mul(K, bK, 0, (size_t)i * pb, s, tg.w);
    fr_load_normal(s, rr + j);
    aff_push_fixed_mul(V, bV, 0, (size_t)i * pb, s, tg.w);
    j++;
  }
  fr_load_normal(s, rnd);          // t
  aff_push_fixed_mul(K, bK, 1, 0, s, tg.w);
  fr_load_normal(s, rr + h);       // random2
  aff_push_fixed_mul(V, bV, 1, 0, s, tg.w);
  aff_flush(K, bK);
  aff_flush(V, bV);
  k_out = K;
  Vk_out = V;
}

// step 2 (G1): the re-randomised credential and the G1 statements / commitments.
//   sig' = (r sig1, r (t sig1 + sig2))  normalised (final output)
//   phi = H(attr_0) S, V_phi = rr[0] S;  with id: E1 = eps g, E2 = eps y + H(attr_1) h, V_E1 = random3 g,
//   V_E2 = random3 y + rr[1] h                      tblB = per-batch tables of [S = H(service), g, y, h]
PSB_HD PSB_NOINL void prove_id_g1_lane(TblGeom tb, const G1A* tblB, const G1J& sig1, const G1J& sig2, const uint8_t* blob,
                                        const uint64_t* off, const Fr* rnd, int h, int with_id, G1J& o_sig1, G1J& o_sig2,
                                        G1J* W /* phi, Vphi, E1, E2, VE1, VE2 (unnormalised) */) {
  const size_t pb = tb.per_base();
  const Fr* rr = rnd + (with_id ? 3 : 2);
  uint32_t kt[8], kr[8], s[8];
  fr_load_normal(kt, rnd);
  fr_load_normal(kr, rnd + 1);
  G1J P = sig1, Q = sig2, T, S1, S2;
  pt_mul(S1, P, kr);
  pt_mul(T, P, kt);
  pt_add(T, T, Q);
  pt_mul(S2, T, kr);
  g1_normalize2(S1, S2);
  o_sig1 = S1;
  o_sig2 = S2;
  for (int i = 0; i < 6; i++) pt_set_zero(W[i]);
  fr_set_hash_of(s, blob + off[0], (size_t)(off[1] - off[0]));       // _s = H(attributes[0])
  pt_fixed_mul_acc(W[0], tblB, s, tb.w);
  fr_load_normal(s, rr);                                             // _randomnesses[0]
  pt_fixed_mul_acc(W[1], tblB, s, tb.w);
  if (with_id) {
    fr_load_normal(s, rnd + 2);                                      // epsilon
    pt_fixed_mul_acc(W[2], tblB + pb, s, tb.w);
    pt_fixed_mul_acc(W[3], tblB + 2 * pb, s, tb.w);
    fr_set_hash_of(s, blob + off[1], (size_t)(off[2] - off[1]));     // _gamma = H(attributes[1])
    pt_fixed_mul_acc(W[3], tblB + 3 * pb, s, tb.w);
    fr_load_normal(s, rr + h + 1);                                   // random3
    pt_fixed_mul_acc(W[4], tblB + pb, s, tb.w);
    pt_fixed_mul_acc(W[5], tblB + 2 * pb, s, tb.w);
    fr_load_normal(s, rr + 1);                                       // _randomnesses[1]
    pt_fixed_mul_acc(W[5], tblB + 3 * pb, s, tb.w);
  }
}

// step 3: normalise (one inversion), c = H(H(hex(k) hex(phi) [hex(E1) hex(E2)] hex(V_k) hex(V_phi) [hex(V_E1) hex(V_E2)] ad)),
//         responses rs[h + 1 (+ 1)].  k, phi, E1, E2 are rewritten in normalised form.
PSB_HD PSB_NOINL void prove_id_hash_lane(int n, const uint8_t* hide, const uint8_t* blob, const uint64_t* off, const uint8_t* ad,
                                          size_t ad_len, const Fr* rnd, int h, int with_id, G2J& k, const G2J& Vk_in, G1J* W,
                                          Fr& c_out, Fr* rs_out) {
  G2J Vk = Vk_in;
  const int n1 = with_id ? 6 : 2;
  Fp z[8];
  g2_norm_of_z(z[0], k);
  g2_norm_of_z(z[1], Vk);
  for (int i = 0; i < n1; i++) z[2 + i] = W[i].z;
  fp_batch_inv(z, 2 + n1);
  g2_apply_ninv(k, z[0]);
  g2_apply_ninv(Vk, z[1]);
  for (int i = 0; i < n1; i++) g1_apply_zinv(W[i], z[2 + i]);
  Sha256 s;
  sha256_init(s);
  sha_put_g2_hex(s, k);
  sha_put_g1_hex(s, W[0]);
  if (with_id) { sha_put_g1_hex(s, W[2]); sha_put_g1_hex(s, W[3]); }
  sha_put_g2_hex(s, Vk);
  sha_put_g1_hex(s, W[1]);
  if (with_id) { sha_put_g1_hex(s, W[4]); sha_put_g1_hex(s, W[5]); }
  uint32_t cn[8], mk[8];
  challenge_finish(cn, s, ad, ad_len);
  Fr c;
  fr_from_normal(c, cn);
  const Fr* rr = rnd + (with_id ? 3 : 2);
  int j = 0;
  for (int i = 0; i < n; i++) {
    if (!hide[i]) continue;
    Fr m, r = rr[j];
    fr_set_hash_of(mk, blob + off[i], (size_t)(off[i + 1] - off[i]));
    fr_from_normal(m, mk);
    fr_response(rs_out[j], r, m, c);
    j++;
  }
  Fr t = rnd[0], r2 = rr[h];
  fr_response(rs_out[h], r2, t, c);
  if (with_id) {
    Fr eps = rnd[2], r3 = rr[h + 1];
    fr_response(rs_out[h + 1], r3, eps, c);
  }
  c_out = c;
}

}  // namespace psb
