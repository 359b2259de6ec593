// protocol.cuh -- per-lane EL PASSO protocol steps: blind issuance (PSSigner::el_passo_provide_id)
// and sign-on proof verification (PSVerifier::el_passo_verify_id), as __host__ __device__ lane
// functions.  The kernels in kernels.cuh are thin one-thread-per-lane wrappers; the CPU test
// library (tests/hostsim) compiles the same lane functions for the host to check the LOGIC against
// the reference without a GPU.
//
// Reference semantics:
//   src/ps-signer.cc:63-146    el_passo_provide_id = el_passo_nizk_verify_request + sign_hybrid + sign_commitment
//   src/ps-verifier.cc:37-138  el_passo_verify_id, :140-212 ..._without_id_retrieval, :214-229 prepare_hybrid_verification
//   Fiat-Shamir challenge: c = Fr::setHashOf( SHA256( hex(P_1) || ... || hex(P_k) || ad ) ), hex = lowercase
//   hex of mcl's compressed serialisation (operator.hpp:187-193, ec.hpp:849-896) -- SURVEY.md F5.
// What differs from the reference is only HOW group elements are computed: fixed bases (g, Y_i, gg,
// XX, YY_i, and the per-batch service/authority points) come from window tables, the points hashed
// are normalised with one shared field inversion, and the two pairings are one multi-Miller loop.
#pragma once
#include "pairing.cuh"
#include "sha256.cuh"

namespace psb {

// ---- serialisation (mcl compressed little-endian form, ec.hpp:849-896, non-ETH mode) -----------------
// kFpBytes = 48 (BLS12-381) or 32 (BN254).
// G1: x as kFpBytes LE bytes of the NORMAL form, bit 7 of the last byte = y odd; infinity = all-zero bytes.
constexpr int kFpBytes = 4 * PSB_NL;
PSB_HD PSB_NOINL void g1_serialize_norm(uint8_t* out, const G1J& P /*normalised or zero*/) {
  if (fp_is_zero(P.z)) { for (int i = 0; i < kFpBytes; i++) out[i] = 0; return; }
  Fp x, y;
  fp_from_mont(x, P.x);
  fp_from_mont(y, P.y);
  for (int i = 0; i < PSB_NL; i++) {
    out[4 * i] = (uint8_t)x.v[i]; out[4 * i + 1] = (uint8_t)(x.v[i] >> 8);
    out[4 * i + 2] = (uint8_t)(x.v[i] >> 16); out[4 * i + 3] = (uint8_t)(x.v[i] >> 24);
  }
  if (y.v[0] & 1u) out[kFpBytes - 1] |= 0x80;
}
// G2: x.a || x.b (96 bytes), parity of y.a (fp_tower.hpp:312) in bit 7 of the last byte.
PSB_HD PSB_NOINL void g2_serialize_norm(uint8_t* out, const G2J& P) {
  if (fp2_is_zero(P.z)) { for (int i = 0; i < 2 * kFpBytes; i++) out[i] = 0; return; }
  Fp xa, xb, ya;
  fp_from_mont(xa, P.x.a);
  fp_from_mont(xb, P.x.b);
  fp_from_mont(ya, P.y.a);
  for (int i = 0; i < PSB_NL; i++) {
    out[4 * i] = (uint8_t)xa.v[i]; out[4 * i + 1] = (uint8_t)(xa.v[i] >> 8);
    out[4 * i + 2] = (uint8_t)(xa.v[i] >> 16); out[4 * i + 3] = (uint8_t)(xa.v[i] >> 24);
    out[kFpBytes + 4 * i] = (uint8_t)xb.v[i]; out[kFpBytes + 4 * i + 1] = (uint8_t)(xb.v[i] >> 8);
    out[kFpBytes + 4 * i + 2] = (uint8_t)(xb.v[i] >> 16); out[kFpBytes + 4 * i + 3] = (uint8_t)(xb.v[i] >> 24);
  }
  if (ya.v[0] & 1u) out[2 * kFpBytes - 1] |= 0x80;
}

// ---- deserialisation = point decompression (mcl EcT::load, ec.hpp:924-1057, IoSerialize, non-ETH mode) ---------------
// relying parties receive compressed points (48 / 96 bytes), and decompressing them with mcl on
// the host (a square root each) would cap a GPU box long before the pairing kernels do.
// 48 little-endian bytes -> canonical Montgomery Fp; false if the value is >= p (Fp::setArray, NoMask: fp.hpp:342-345)
PSB_HD PSB_INL bool fp_from_le_bytes(Fp& r, const uint8_t* b, uint8_t last_mask) {
  Fp n;
  for (int i = 0; i < PSB_NL; i++)
    n.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
             ((uint32_t)(i == PSB_NL - 1 ? (b[kFpBytes - 1] & last_mask) : b[4 * i + 3]) << 24);
  uint32_t t[PSB_NL];
  if (sub_mod_n<FpT>(t, n.v) == 0) return false;   // n >= p
  fp_to_mont(r, n);
  return true;
}
// mcl Fp::squareRoot for p = 3 mod 4 (gmp_util.hpp:880-897): a = 0 -> 0; non-residue -> false; else a^((p+1)/4)
PSB_HD PSB_NOINL bool fp_sqrt(Fp& y, const Fp& a) {
  if (fp_is_zero(a)) { fp_set_zero(y); return true; }
  Fp r, t;
  fp_pow_nib(r, a, PSB_K(FP_PP1D4_NIB));
  fp_sqr(t, r);
  if (!fp_eq(t, a)) return false;
  y = r;
  return true;
}
PSB_HD PSB_INL bool fp_is_odd(const Fp& a) {       // parity of the NORMAL form (Fp::isOdd)
  Fp n;
  fp_from_mont(n, a);
  return (n.v[0] & 1u) != 0;
}
PSB_HD PSB_INL void fp_div2(Fp& r, const Fp& a) {   // Fp::divBy2 (works on the Montgomery representative)
  uint32_t t[PSB_NL + 1];
  for (int i = 0; i < PSB_NL; i++) t[i] = a.v[i];
  t[PSB_NL] = 0;
  if (a.v[0] & 1u) {
    uint64_t c = 0;
    for (int i = 0; i < PSB_NL; i++) { c += (uint64_t)t[i] + FpT::p(i); t[i] = (uint32_t)c; c >>= 32; }
    t[PSB_NL] = (uint32_t)c;
  }
  for (int i = 0; i < PSB_NL; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
}
// mcl Fp2::squareRoot (fp_tower.hpp:320-352), same root selection
PSB_HD PSB_NOINL bool fp2_sqrt(Fp2& y, const Fp2& x) {
  Fp t1, t2;
  if (fp_is_zero(x.b)) {
    if (fp_sqrt(t1, x.a)) { y.a = t1; fp_set_zero(y.b); }
    else { fp_neg(t2, x.a); fp_sqrt(t1, t2); fp_set_zero(y.a); y.b = t1; }
    return true;
  }
  fp_dot2(t1, x.a, x.a, x.b, x.b);             // c^2 + d^2
  if (!fp_sqrt(t1, t1)) return false;
  Fp u;
  fp_add(u, x.a, t1);
  fp_div2(u, u);
  if (!fp_sqrt(t2, u)) {
    fp_sub(u, x.a, t1);
    fp_div2(u, u);
    fp_sqrt(t2, u);
  }
  y.a = t2;
  fp_dbl(t2, t2);
  fp_inv(t2, t2);
  fp_mul(y.b, x.b, t2);
  return true;
}
// the curve constant b (4 or 2) as a Montgomery element
PSB_HD PSB_INL void fp_set_curve_b(Fp& r) {
  fp_set_one(r); fp_dbl(r, r);
#if PSB_CURVE_B == 4
  fp_dbl(r, r);
#endif
}
// G1::deserialize: all-zero = infinity; bit 7 of the last byte = y odd; x >= p or x^3 + 4 a non-residue -> false
PSB_HD PSB_NOINL bool g1_deserialize(G1J& P, const uint8_t* b) {
  uint8_t o = 0;
  for (int i = 0; i < kFpBytes; i++) o |= b[i];
  pt_set_zero(P);
  if (o == 0) return true;
  const bool y_odd = (b[kFpBytes - 1] >> 7) != 0;
  Fp x, y, t, cb;
  if (!fp_from_le_bytes(x, b, 0x7f)) return false;
  fp_sqr(t, x); fp_mul(t, t, x);
  fp_set_curve_b(cb);
  fp_add(t, t, cb);
  if (!fp_sqrt(y, t)) return false;
  if (fp_is_odd(y) != y_odd) fp_neg(y, y);
  P.x = x; P.y = y; fp_set_one(P.z);
  return true;
}
// G2::deserialize: x.a || x.b (96 bytes), flag in the last byte, parity of y.a (Fp2::isOdd, fp_tower.hpp:312)
PSB_HD PSB_NOINL bool g2_deserialize(G2J& P, const uint8_t* b) {
  uint8_t o = 0;
  for (int i = 0; i < 2 * kFpBytes; i++) o |= b[i];
  pt_set_zero(P);
  if (o == 0) return true;
  const bool y_odd = (b[2 * kFpBytes - 1] >> 7) != 0;
  Fp2 x, y, t, bb;
  if (!fp_from_le_bytes(x.a, b, 0xff)) return false;
  if (!fp_from_le_bytes(x.b, b + kFpBytes, 0x7f)) return false;
  fp2_sqr(t, x); fp2_mul(t, t, x);
#if PSB_TWIST_MTYPE
  fp_set_curve_b(bb.a);                                          // b' = 4 xi = 4 + 4i
  bb.b = bb.a;
#else
  fp_set_one(bb.a); fp_neg(bb.b, bb.a);                          // b' = 2 / xi = 1 - i
#endif
  fp2_add(t, t, bb);
  if (!fp2_sqrt(y, t)) return false;
  if (fp_is_odd(y.a) != y_odd) fp2_neg(y, y);
  P.x = x; P.y = y; fp2_set_one(P.z);
  return true;
}

// simultaneous inversion (Montgomery's trick): z[i] <- z[i]^-1 for the non-zero entries, ONE fp_inv.
// Zero entries stay zero.  cnt <= 8.
PSB_HD PSB_NOINL void fp_batch_inv(Fp* z, int cnt) {
  Fp pre[8], one, acc, inv, t;
  fp_set_one(one);
  acc = one;
  for (int i = 0; i < cnt; i++) {
    pre[i] = acc;                                   // product of the non-zero z[0..i)
    if (!fp_is_zero(z[i])) fp_mul(acc, acc, z[i]);
  }
  fp_inv(inv, acc);
  for (int i = cnt - 1; i >= 0; i--) {
    if (fp_is_zero(z[i])) continue;
    fp_mul(t, inv, pre[i]);                         // z[i]^-1
    fp_mul(inv, inv, z[i]);
    z[i] = t;
  }
}

// normalise with a given z^-1 (zi == 0 <=> infinity -> canonical zero)
PSB_HD PSB_INL void g1_apply_zinv(G1J& P, const Fp& zi) {
  if (fp_is_zero(P.z)) { pt_set_zero(P); return; }
  Fp t;
  fp_sqr(t, zi); fp_mul(P.x, P.x, t); fp_mul(t, t, zi); fp_mul(P.y, P.y, t); fp_set_one(P.z);
}
// G2: z^-1 = conj(z) * n^-1 with n = |z|^2 in Fp (fp_tower.hpp:597-611); ni = n^-1
PSB_HD PSB_INL void g2_norm_of_z(Fp& n, const G2J& P) { fp_dot2(n, P.z.a, P.z.a, P.z.b, P.z.b); }
PSB_HD PSB_INL void g2_apply_ninv(G2J& P, const Fp& ni) {
  if (fp2_is_zero(P.z)) { pt_set_zero(P); return; }
  Fp2 zi, t;
  fp_mul(zi.a, P.z.a, ni);
  fp_mul(zi.b, P.z.b, ni);
  fp_neg(zi.b, zi.b);
  fp2_sqr(t, zi); fp2_mul(P.x, P.x, t); fp2_mul(t, t, zi); fp2_mul(P.y, P.y, t); fp2_set_one(P.z);
}

// normalise two G1 points with ONE field inversion; zero points stay canonical zero
PSB_HD PSB_NOINL void g1_normalize2(G1J& A, G1J& B) {
  Fp z[2] = {A.z, B.z};
  fp_batch_inv(z, 2);
  g1_apply_zinv(A, z[0]);
  g1_apply_zinv(B, z[1]);
}

// ---- hex + SHA-256 streaming ----------------------------------------------------------------------
PSB_HD PSB_INL uint8_t hex_digit(uint32_t v) { return (uint8_t)(v < 10 ? '0' + v : 'a' + (v - 10)); }
PSB_HD PSB_INL void sha_put_hex(Sha256& s, const uint8_t* bytes, int n) {
  for (int i = 0; i < n; i++) {
    sha256_put(s, hex_digit(bytes[i] >> 4));
    sha256_put(s, hex_digit(bytes[i] & 15));
  }
}
PSB_HD PSB_NOINL void sha_put_g1_hex(Sha256& s, const G1J& P /*normalised*/) {
  uint8_t b[kFpBytes];
  g1_serialize_norm(b, P);
  sha_put_hex(s, b, kFpBytes);
}
PSB_HD PSB_NOINL void sha_put_g2_hex(Sha256& s, const G2J& P /*normalised*/) {
  uint8_t b[2 * kFpBytes];
  g2_serialize_norm(b, P);
  sha_put_hex(s, b, 2 * kFpBytes);
}
// c = Fr::setHashOf(digest_engine.digest(ad)): the 32 raw digest bytes are hashed AGAIN (double hash)
PSB_HD PSB_NOINL void challenge_finish(uint32_t c[8], Sha256& s, const uint8_t* ad, size_t ad_len) {
  sha256_update(s, ad, ad_len);
  uint32_t d[8];
  sha256_final(s, d);
  uint8_t raw[32];
  for (int i = 0; i < 8; i++) {
    raw[4 * i] = (uint8_t)(d[i] >> 24); raw[4 * i + 1] = (uint8_t)(d[i] >> 16);
    raw[4 * i + 2] = (uint8_t)(d[i] >> 8); raw[4 * i + 3] = (uint8_t)d[i];
  }
  fr_set_hash_of(c, raw, 32);
}

PSB_HD PSB_INL void fr_load_normal(uint32_t k[8], const Fr* p) {
  Fr t = *p, n;
  fr_from_mont(n, t);
  for (int i = 0; i < 8; i++) k[i] = n.v[i];
}
PSB_HD PSB_INL bool k8_eq(const uint32_t* a, const uint32_t* b) {
  uint32_t o = 0;
  for (int i = 0; i < 8; i++) o |= a[i] ^ b[i];
  return o == 0;
}

// shared table geometry of one key: every fixed base has nwin(w) * 2^(w-1) affine entries
struct TblGeom {
  int w;
  PSB_HD size_t per_base() const { return (size_t)fixed_nwin(w) << (w - 1); }
};

// ---- PSSigner::el_passo_provide_id, one lane (src/ps-signer.cc:63-146) --------------------------------
//   tblG1: fixed-base tables of [g, Y_0 .. Y_{n-1}] (geometry tg);  X = secret g^x (normalised)
//   attrs: n strings at blob[off[i] .. off[i+1]) ("" = hidden);  rs: `per` scalars (Montgomery)
//   returns the NIZK verdict; sig1/sig2 normalised (zero when the NIZK fails)
PSB_HD PSB_NOINL bool provide_id_lane(int n, TblGeom tg, const G1A* tblG1, const G1J& X, const G1J& A_in, const Fr* c_mont,
                                       const Fr* rs, int per, const uint8_t* blob, const uint64_t* off,
                                       const uint8_t* ad, size_t ad_len, const Fr* u_mont, G1J& sig1, G1J& sig2) {
  const size_t pb = tg.per_base();
  uint32_t k[8], cn[8];
  bool ok = per >= 1;
  // V = c A + rs[0] g + sum_hidden rs[j] Y_i        (ps-signer.cc:80-92)
  G1J V, Ap = A_in, A = A_in;
  fr_load_normal(cn, c_mont);
  pt_mul(V, A, cn);
  // (G1 sums stay on the plain mixed-addition chain: batched affine pair additions, which pay in G2, measured 2.4 % SLOWER
  // here -- an Fp product is too cheap against the inversion and the second table fetch, profiles/r2u_ab_affine_protocol.txt)
  if (ok) { fr_load_normal(k, rs); pt_fixed_mul_acc(V, tblG1, k, tg.w); }
  int j = 1;
  for (int i = 0; i < n; i++) {
    const uint64_t b = off[i], e = off[i + 1];
    if (e == b) {
      if (j < per) { fr_load_normal(k, rs + j); pt_fixed_mul_acc(V, tblG1 + (size_t)(1 + i) * pb, k, tg.w); }
      else ok = false;   // the reference reads past rs here (undefined behaviour): reject
      j++;
    } else if (n != 1) {  // sign_hybrid: a single-attribute request is signed as a bare commitment (:115-117)
      fr_set_hash_of(k, blob + b, (size_t)(e - b));
      pt_fixed_mul_acc(Ap, tblG1 + (size_t)(1 + i) * pb, k, tg.w);
    }
  }
  // sigma1 = u g, sigma2 = u (X + A')                 (ps-signer.cc:132-146)
  G1J S1, S2, T;
  fr_load_normal(k, u_mont);
  pt_set_zero(S1);
  pt_fixed_mul_acc(S1, tblG1, k, tg.w);
  pt_add(T, X, Ap);
  pt_mul(S2, T, k);
  // one inversion for A, V (hashed) and sigma1, sigma2 (returned)
  Fp z[4] = {A.z, V.z, S1.z, S2.z};
  fp_batch_inv(z, 4);
  g1_apply_zinv(A, z[0]); g1_apply_zinv(V, z[1]); g1_apply_zinv(S1, z[2]); g1_apply_zinv(S2, z[3]);
  // c' = H(H(hex(A) || hex(V) || ad))                 (ps-signer.cc:94-101)
  Sha256 s;
  sha256_init(s);
  sha_put_g1_hex(s, A);
  sha_put_g1_hex(s, V);
  uint32_t c2[8];
  challenge_finish(c2, s, ad, ad_len);
  ok = ok && k8_eq(c2, cn);
  if (ok) { sig1 = S1; sig2 = S2; } else { pt_set_zero(sig1); pt_set_zero(sig2); }
  return ok;
}

// ---- PSSigner::sign_hybrid / sign_commitment, one lane (src/ps-signer.cc:112-146) ------------------------
//   na = number of attribute strings of the call (0: sign_commitment; sign_hybrid signs a ONE-attribute list as a bare
//   commitment, :114-116); "" = a committed attribute, skipped.  sig = (u g, u (X + C + sum H(attr_i) Y_i)), normalised.
PSB_HD PSB_NOINL void sign_lane(int na, TblGeom tg, const G1A* tblG1, const G1J& X, const G1J& C, const uint8_t* blob,
                                const uint64_t* off, const Fr* u_mont, G1J& sig1, G1J& sig2) {
  const size_t pb = tg.per_base();
  uint32_t k[8];
  G1J Ap = C, T;
  if (na != 1) {
    for (int i = 0; i < na; i++) {
      const uint64_t b = off[i], e = off[i + 1];
      if (e == b) continue;
      fr_set_hash_of(k, blob + b, (size_t)(e - b));
      pt_fixed_mul_acc(Ap, tblG1 + (size_t)(1 + i) * pb, k, tg.w);
    }
  }
  fr_load_normal(k, u_mont);
  pt_set_zero(sig1);
  pt_fixed_mul_acc(sig1, tblG1, k, tg.w);
  pt_add(T, X, Ap);
  pt_mul(sig2, T, k);
  g1_normalize2(sig1, sig2);
}

// ---- PSVerifier::el_passo_verify_id, one lane, three steps ---------------------------------------------
// step 1 (G2): V_k = c k + sum_hidden rs[cnt] YY_i + rs[gg_idx] gg + (1-c) XX   (ps-verifier.cc:72-88 / :166-182)
//              K   = k + sum_plain H(attr_i) YY_i                                   (:214-229)
//   tblYY: tables of YY_0..; tblAux: tables of [gg, XX].  Returns false if rs is too short (UB in the reference).
PSB_HD PSB_NOINL bool verify_id_g2_lane(int n, TblGeom tg, const G2A* tblYY, const G2A* tblAux, const G2J& k_in,
                                         const Fr* c_mont, const Fr* rs, int per, int with_id, const uint8_t* blob,
                                         const uint64_t* off, G2J& Vk, G2J& K) {
  const size_t pb = tg.per_base();
  uint32_t k[8];
  bool ok = per >= (with_id ? 2 : 1);
  fr_load_normal(k, c_mont);
  G2J kk = k_in;
  pt_mul(Vk, kk, k);
  K = kk;
  // the table entries of V_k and of K are summed pairwise in affine coordinates (AffBatch, curve.cuh)
  AffBatch<Fp2> bV, bK;
  {
    int hid = 0;
    for (int i = 0; i < n; i++) hid += off[i + 1] == off[i];
    aff_init(bV, (hid + 2) * fixed_nwin(tg.w), (size_t)(n > 2 ? n : 2) * pb, tblYY, tblAux);
    aff_init(bK, (n - hid) * fixed_nwin(tg.w), (size_t)n * pb, tblYY);
  }
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    const uint64_t b = off[i], e = off[i + 1];
    if (e == b) {
      if (cnt < per) { fr_load_normal(k, rs + cnt); aff_push_fixed_mul(Vk, bV, 0, (size_t)i * pb, k, tg.w); }
      else ok = false;
      cnt++;
    } else {
      fr_set_hash_of(k, blob + b, (size_t)(e - b));
      aff_push_fixed_mul(K, bK, 0, (size_t)i * pb, k, tg.w);
    }
  }
  if (ok) {
    fr_load_normal(k, rs + (with_id ? per - 2 : per - 1));
    aff_push_fixed_mul(Vk, bV, 1, 0, k, tg.w);
  }
  Fr one, omc, cm = *c_mont;
  for (int i = 0; i < 8; i++) one.v[i] = PSB_K(FR_ONE)[i];
  fr_sub(omc, one, cm);
  fr_load_normal(k, &omc);
  aff_push_fixed_mul(Vk, bV, 1, pb, k, tg.w);
  aff_flush(Vk, bV);
  aff_flush(K, bK);
  return ok;
}

// step 2 (G1): V_phi = c phi + rs[0] H(service);  V_E1 = c E1 + rs[per-1] g;  V_E2 = c E2 + rs[per-1] y + rs[1] h
//   (ps-verifier.cc:91-108).  tblB: per-batch tables of [H(service), g, y, h] with geometry tb.
PSB_HD PSB_NOINL bool verify_id_g1_lane(TblGeom tb, const G1A* tblB, const G1J& phi, const G1J* E1, const G1J* E2,
                                         const Fr* c_mont, const Fr* rs, int per, int with_id, G1J& Vphi, G1J& VE1,
                                         G1J& VE2) {
  const size_t pb = tb.per_base();
  uint32_t cn[8], k[8];
  const bool ok = per >= (with_id ? 2 : 1);
  fr_load_normal(cn, c_mont);
  G1J P = phi;
  pt_mul(Vphi, P, cn);
  if (per >= 1) { fr_load_normal(k, rs); pt_fixed_mul_acc(Vphi, tblB, k, tb.w); }
  pt_set_zero(VE1);
  pt_set_zero(VE2);
  if (with_id && ok) {
    P = *E1;
    pt_mul(VE1, P, cn);
    P = *E2;
    pt_mul(VE2, P, cn);
    fr_load_normal(k, rs + per - 1);
    pt_fixed_mul_acc(VE1, tblB + pb, k, tb.w);
    pt_fixed_mul_acc(VE2, tblB + 2 * pb, k, tb.w);
    fr_load_normal(k, rs + 1);
    pt_fixed_mul_acc(VE2, tblB + 3 * pb, k, tb.w);
  }
  return ok;
}

// step 3: c' = H(H(hex(k) hex(phi) [hex(E1) hex(E2)] hex(V_k) hex(V_phi) [hex(V_E1) hex(V_E2)] ad)) == c
//   (ps-verifier.cc:110-130 / :192-204).  All eight points are normalised with ONE field inversion.
PSB_HD PSB_NOINL bool verify_id_hash_lane(const G2J& k_in, const G1J& phi_in, const G1J* E1_in, const G1J* E2_in,
                                           const G2J& Vk_in, const G1J& Vphi_in, const G1J& VE1_in, const G1J& VE2_in,
                                           int with_id, const Fr* c_mont, const uint8_t* ad, size_t ad_len) {
  G2J k = k_in, Vk = Vk_in;
  G1J g1[6];
  g1[0] = phi_in; g1[1] = Vphi_in;
  const int n1 = with_id ? 6 : 2;
  if (with_id) { g1[2] = *E1_in; g1[3] = *E2_in; g1[4] = VE1_in; g1[5] = VE2_in; }
  Fp z[8];
  g2_norm_of_z(z[0], k);
  g2_norm_of_z(z[1], Vk);
  for (int i = 0; i < n1; i++) z[2 + i] = g1[i].z;
  fp_batch_inv(z, 2 + n1);
  g2_apply_ninv(k, z[0]);
  g2_apply_ninv(Vk, z[1]);
  for (int i = 0; i < n1; i++) g1_apply_zinv(g1[i], z[2 + i]);
  Sha256 s;
  sha256_init(s);
  sha_put_g2_hex(s, k);
  sha_put_g1_hex(s, g1[0]);
  if (with_id) { sha_put_g1_hex(s, g1[2]); sha_put_g1_hex(s, g1[3]); }
  sha_put_g2_hex(s, Vk);
  sha_put_g1_hex(s, g1[1]);
  if (with_id) { sha_put_g1_hex(s, g1[4]); sha_put_g1_hex(s, g1[5]); }
  uint32_t c2[8], cn[8];
  challenge_finish(c2, s, ad, ad_len);
  fr_load_normal(cn, c_mont);
  return k8_eq(c2, cn);
}

}  // namespace psb
