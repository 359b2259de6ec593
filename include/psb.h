/* psb.h -- C ABI of the B200 batch engine for PS signatures / EL PASSO (libpsb.so).
 *
 * The reference (Zhiyi-Zhang/PS-Signature-and-EL-PASSO) has no plugin or FFI mechanism: its hot
 * path is reached through three C++ classes that call mcl directly.  This header is
 * the boundary a maintainer binds instead of those mcl calls for BATCHED work; each entry point
 * cites the reference interface it replaces.  Conventions are modelled on mcl's own C API
 * (third-parties/mcl/include/mcl/bn.h:78-110): PODs are arrays of uint64_t holding little-endian
 * limbs in MONTGOMERY form exactly as mcl keeps them in memory, so `std::vector<G1>::data()` etc.
 * can be passed without conversion:
 *     Fp  6 x u64   Fr  4 x u64   G1 18 x u64 (Jacobian x,y,z)   G2 36 x u64   GT 72 x u64
 * z == 0 is the point at infinity.  int returns: 0 = ok, < 0 = error (psb_last_error()).
 * Per-lane cryptographic failure is DATA (verdict byte 0), never an error.
 * There is no CPU fallback: every entry fails with PSB_ERR_CUDA when no usable GPU is present.
 */
#ifndef PSB_H_
#define PSB_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_CURVE_BLS12_381 5 /* = MCL_BLS12_381 (mcl/include/mcl/curve_type.h:91): libpsb.so */
#define PSB_CURVE_BN254 0     /* = MCL_BN254 (curve_type.h:85), what initPairing() selects: libpsb_bn254.so */

#define PSB_OK 0
#define PSB_ERR_ARG (-1)
#define PSB_ERR_CUDA (-2)
#define PSB_ERR_NOT_INIT (-3)
#define PSB_ERR_UNSUPPORTED (-4)
#define PSB_ERR_NOMEM (-5)

typedef struct psb_key psb_key;

/* Replaces mcl::bn::initPairing(mcl::BLS12_381) (mcl/include/mcl/bn.hpp:2208) for the batch path.
 * devices = CUDA ordinals to shard over (NULL/0 -> device 0 only).  One host worker thread and
 * one stream per device; no inter-device communication. */
int psb_init(int curve, const int* devices, int ndev);
/* Releases every device resource, INCLUDING the device memory of keys that are still alive: such keys become dead
 * handles (every call with one fails with PSB_ERR_ARG; psb_key_destroy stays valid and frees the host object).
 * psb_init on an initialised library is psb_shutdown + init; a failed psb_init leaves the library shut down. */
void psb_shutdown(void);
int psb_num_devices(void);
const char* psb_last_error(void);
/* lane range [begin, end) that device k of ndev processes in every batched call (contiguous split):
 * the host-side "gather" is each device writing its slice of the caller's output arrays.  Pure host arithmetic. */
int psb_shard_range(size_t N, int ndev, int k, size_t* begin, size_t* end);
/* number of kernels this library has launched since psb_init (for the bench's gpu_launches) */
uint64_t psb_launch_count(void);

/* Public key of the batch = PSPubKey (src/ps-encoding.h:111-133): g, gg, XX, Y[n], YY[n]; all
 * points may have any z (they are normalised once here).  X_secret = g^x, the PSSigner secret
 * m_sk_X (src/ps-signer.h), or NULL for verify-only keys.  Builds the fixed-base window tables
 * (idea: mcl/include/mcl/window_method.hpp:68-108) in HBM on every device.
 * window_bits: 0 = default (16), else 4..24 (w = 20: 1.3 GB per G2 base, w = 22: 4.8 GB, w = 24: 17.7 GB -- sized for 180 GB of HBM). */
psb_key* psb_key_create(const uint64_t* g, const uint64_t* gg, const uint64_t* XX, const uint64_t* Y,
                        const uint64_t* YY, size_t n, const uint64_t* X_secret, int window_bits);
void psb_key_destroy(psb_key* key);
size_t psb_key_num_attributes(const psb_key* key);
size_t psb_key_table_bytes(const psb_key* key);

/* Batched PSVerifier::verify (src/ps-verifier.cc:13-35) == PSRequester::verify
 * (src/ps-requester.cc:115-137).  Lane j: sig1[j], sig2[j] and n attributes given either as
 * strings (attr_blob + attr_off[N*n+1], attribute i of lane j = blob[off[j*n+i] .. off[j*n+i+1]))
 * hashed on the device like Fr::setHashOf, or as precomputed scalars m (N*n Fr, Montgomery; used
 * when attr_blob == NULL).  verdict[j] = 1 iff sig1 != 0 and e(sig1, XX + sum m_i YY_i) == e(sig2, gg).
 * gt (optional, N x 72 u64) receives e(sig1,K) * e(sig2,gg)^-1 = lhs * unitaryInv(rhs). */
int psb_verify(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2,
               const uint8_t* attr_blob, const uint64_t* attr_off, const uint64_t* m,
               uint8_t* verdict, uint64_t* gt);

/* psb_verify on an ARRAY OF CREDENTIALS: cred = N x (sigma1 || sigma2), 2 x 18 u64 per lane -- the bytes of
 * std::vector<PSCredential> (src/ps-encoding.h:89-109: two G1 members), passed without a copy. */
int psb_verify_aos(psb_key* key, size_t N, const uint64_t* cred, const uint8_t* attr_blob, const uint64_t* attr_off,
                   const uint64_t* m, uint8_t* verdict, uint64_t* gt);

/* Batched G1::deserialize / G2::deserialize = point decompression (mcl/include/mcl/ec.hpp:924-1057, IoSerialize,
 * mcl's default little-endian mode): element j is the 48 (G1) / 96 (G2) bytes at ser + j * stride.  out[j] = the
 * point with z = 1 (all-zero for the infinity encoding); ok[j] = 1 iff mcl's deserialize would succeed (x < p and
 * x^3 + b a square).  SURVEY.md 8f rank 1: wire-format ingest on the device. */
int psb_g1_deserialize(size_t N, const uint8_t* ser, size_t stride, uint64_t* out, uint8_t* ok);
int psb_g2_deserialize(size_t N, const uint8_t* ser, size_t stride, uint64_t* out, uint8_t* ok);

/* psb_verify on SERIALIZED credentials: lane j reads sigma1 at cred + j * stride + off1 and sigma2 at + off2
 * (48 bytes each).  PSCredential::toBufferString (src/ps-encoding.cc:384-391) is two TLVs of 1 + 1 + 48 bytes:
 * stride = 100, off1 = 2, off2 = 52; a bare 96-byte sigma1 || sigma2 is stride = 96, off1 = 0, off2 = 48.
 * Points are decompressed on the device; a lane whose encoding mcl would reject gets verdict 0 and decoded[j] = 0
 * (decoded is optional).  Attributes as strings only. */
int psb_verify_ser(psb_key* key, size_t N, const uint8_t* cred, size_t stride, size_t off1, size_t off2,
                   const uint8_t* attr_blob, const uint64_t* attr_off, uint8_t* verdict, uint8_t* decoded);

/* Same computation with every buffer already resident in the memory of device `dev_index`
 * (index into the psb_init list); launches on `stream` (a cudaStream_t, NULL = the library's own
 * stream) and does not synchronise.  `ws` = device scratch of psb_verify_ws_bytes(key, N) bytes. */
size_t psb_verify_ws_bytes(const psb_key* key, size_t N);
int psb_verify_dev(psb_key* key, int dev_index, size_t N, const uint64_t* d_sig1, const uint64_t* d_sig2,
                   const uint8_t* d_attr_blob, const uint64_t* d_attr_off, const uint64_t* d_m,
                   uint8_t* d_verdict, uint64_t* d_gt, void* d_ws, void* stream);

/* Page-locked host memory for the arrays handed to the psb_* calls (optional: every entry point takes any host pointer,
 * but copies from and to page-locked memory run at the full PCIe rate and asynchronously -- measured on the EL PASSO calls,
 * 2^18 lanes: issuance 2.33 -> 3.24 M/s, randomisation 3.8 -> 6.1 M/s, sign-on verification 716 -> 768 k/s).  Portable across
 * the devices of the context.  The reference has no counterpart (its objects live in std::vector); a caller that keeps its
 * batch arrays in this memory needs no other change.  psb_host_alloc returns NULL on failure (psb_last_error). */
void* psb_host_alloc(size_t bytes);
void psb_host_free(void* p);

/* Per-phase device timing of psb_verify / psb_verify_dev: when profiling is on, CUDA events are
 * recorded on the launching stream around the three phase kernels (fixed-base MSM, multi-Miller
 * loop, final exponentiation); psb_last_phase_ms waits for the last profiled batch on that device
 * and returns the three durations in milliseconds. */
int psb_set_profiling(int on);
int psb_last_phase_ms(int dev_index, float ms[3]);

/* Batched PSRequester::randomize_credential (src/ps-requester.cc:139-148) with host-supplied
 * scalars t (N Fr, Montgomery): out = (t*sig1, t*sig2), NORMALISED (z = 1; infinity = all zero).
 * ser (optional, N x 96 bytes) = mcl serialisation of out1 || out2 (ec.hpp:849-896). */
int psb_randomize(size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint64_t* t,
                  uint64_t* out1, uint64_t* out2, uint8_t* ser);

/* Batched G1::mul (mcl/include/mcl/ec.hpp:1124-1139): out[j] = k[j] * P[j] (p_stride = 1) or k[j] * P[0]
 * (p_stride = 0), normalised.  k: N Fr, Montgomery. */
int psb_g1_mul(size_t N, const uint64_t* P, int p_stride, const uint64_t* k, uint64_t* out);

/* Batched PSSigner::el_passo_provide_id (src/ps-signer.cc:63-146): NIZK check of the request
 * (A, c, rs[rs_per_lane], attributes with "" = hidden, associated data), then
 * sig = (u*g, u*(X + A + sum_plain H(attr_i) Y_i)) with host-supplied u (N Fr, Montgomery).
 * verdict[j] = NIZK result; sig1/sig2 normalised (all-zero when the NIZK fails); ser optional. */
int psb_provide_id(psb_key* key, size_t N, const uint64_t* A, const uint64_t* c, const uint64_t* rs,
                   size_t rs_per_lane, const uint8_t* attr_blob, const uint64_t* attr_off,
                   const uint8_t* ad_blob, const uint64_t* ad_off, const uint64_t* u,
                   uint8_t* verdict, uint64_t* sig1, uint64_t* sig2, uint8_t* ser);

/* Batched PSSigner::sign_commitment (src/ps-signer.cc:132-146; n_attrs = 0) and PSSigner::sign_hybrid (:112-130; n_attrs =
 * the length of the attribute list, <= the key's, "" = committed attribute; a ONE-entry list is signed as a bare
 * commitment exactly like the reference does, :114-116):  sig = (u*g, u*(X + commitment + sum H(attr_i) Y_i)) with
 * host-supplied u (N Fr, Montgomery), NORMALISED; ser optional (N x 2 compressed G1).  Needs a key created with X_secret. */
int psb_sign(psb_key* key, size_t N, const uint64_t* commitment, size_t n_attrs, const uint8_t* attr_blob,
             const uint64_t* attr_off, const uint64_t* u, uint64_t* sig1, uint64_t* sig2, uint8_t* ser);

/* flags of the sign-on verification entry points */
#define PSB_VID_WITH_ID 1            /* el_passo_verify_id; clear = el_passo_verify_id_without_id_retrieval */
#define PSB_VID_REJECT_ZERO_SIGMA 2  /* STRICT: also reject sigma1 == 0.  The reference's el_passo_verify_id has no such
                                      * check (unlike PSVerifier::verify, src/ps-verifier.cc:16-18): a proof with sigma1 =
                                      * sigma2 = 0 and an honestly built k and NIZK passes there because e(0, K) = e(0, gg) = 1
                                      * -- a sign-on without a credential.  Clear = bit-exact reference verdicts. */

/* Batched PSVerifier::el_passo_verify_id (src/ps-verifier.cc:37-138; PSB_VID_WITH_ID) and
 * el_passo_verify_id_without_id_retrieval (:140-212; E1/E2/y/g/h ignored).
 * service_pt = hashAndMapToG1(service_name), one value per batch computed by the host.
 * flags = PSB_VID_* bits (0 / 1 keep their round-1 meaning of with_id). */
int psb_verify_id(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint64_t* k,
                  const uint64_t* phi, const uint64_t* E1, const uint64_t* E2, const uint64_t* c,
                  const uint64_t* rs, size_t rs_per_lane, const uint8_t* attr_blob,
                  const uint64_t* attr_off, const uint8_t* ad_blob, const uint64_t* ad_off,
                  const uint64_t* service_pt, const uint64_t* y, const uint64_t* g, const uint64_t* h,
                  int flags, uint8_t* verdict);

/* ---- wire-format ingest on the device.  Lane j = buf[buf_off[j] .. buf_off[j+1]): the bytes of
 * IdProof::toBufferString() (src/ps-encoding.cc:452-468) resp. PSCredRequest::toBufferString() (:429-439), or with
 * base64 != 0 their base64 text (PSBuffer::toBase64 / fromBase64, :111-122, decoder :56-96).  The TLV walk (:124-162,
 * :178-256, :332-384, :441-450, :470-489), base64 decoding and the decompression of every point run on the GPU; the lanes
 * then go through the same kernels as psb_verify_id / psb_provide_id, so a lane may carry any number of responses.
 * parsed[j] (optional) = 1 iff the buffer is a well-formed message whose points and scalars mcl's deserialize accepts and
 * whose attribute list has the key's length; other lanes get verdict 0 -- the reference either throws (short buffer),
 * walks on with an uninitialised object (wrong type byte) or ignores the failed deserialize and verifies a point that
 * is not on the curve; none of these can end in `true`.  A proof without E1/E2 fails PSB_VID_WITH_ID. ---- */
int psb_verify_id_ser(psb_key* key, size_t N, const uint8_t* buf, const uint64_t* buf_off, int base64,
                      const uint8_t* ad_blob, const uint64_t* ad_off, const uint64_t* service_pt, const uint64_t* y,
                      const uint64_t* g, const uint64_t* h, int flags, uint8_t* verdict, uint8_t* parsed);
int psb_provide_id_ser(psb_key* key, size_t N, const uint8_t* buf, const uint64_t* buf_off, int base64,
                       const uint8_t* ad_blob, const uint64_t* ad_off, const uint64_t* u, uint8_t* verdict,
                       uint64_t* sig1, uint64_t* sig2, uint8_t* ser, uint8_t* parsed);

/* ---- wire-format OUTPUT: the bytes a batched prover sends.  kind = PSB_WIRE_IDPROOF: IdProof::toBufferString()
 * (src/ps-encoding.cc:452-468: sig1 sig2 k phi c rs attributes [E1 E2]; E1 = E2 = NULL for a proof without id retrieval),
 * kind = PSB_WIRE_REQUEST: PSCredRequest::toBufferString() (:429-439: A c rs attributes; p0 = A, sig2 = k = phi = NULL).
 * Points may be any Jacobian representative (they are normalised like mcl's serialize does, ec.hpp:849-896); rs holds rs_per
 * scalars per lane; lane j's n_attrs strings are attr_blob[attr_off[j*n_attrs + i] .. attr_off[j*n_attrs + i + 1]) ("" = hidden).
 * base64 != 0 writes PSBuffer::toBase64() text (:14-54, '=' padded) instead of the raw bytes.  out_off[N + 1] is always
 * written (lane j's message = out[out_off[j] .. out_off[j+1])); with out == NULL nothing else happens (size query), else
 * out_cap >= out_off[N] is required.  Lists or strings longer than 0xFFFF (which the reference's appendVar silently
 * drops, :138-146) are PSB_ERR_ARG. ---- */
#define PSB_WIRE_IDPROOF 0
#define PSB_WIRE_REQUEST 1
int psb_wire_encode(int kind, size_t N, size_t n_attrs, const uint64_t* p0, const uint64_t* sig2, const uint64_t* k,
                    const uint64_t* phi, const uint64_t* E1, const uint64_t* E2, const uint64_t* c, const uint64_t* rs,
                    size_t rs_per, const uint8_t* attr_blob, const uint64_t* attr_off, int base64, uint8_t* out, size_t out_cap,
                    uint64_t* out_off);

/* ---- prover side.  The reference draws its blinding / commitment scalars from mcl's CSPRNG;
 * the batch entries take them from the host, per lane, IN THE REFERENCE'S DRAW ORDER, so the same scalars reproduce the
 * reference's requests and proofs byte for byte.  hide = n flags shared by the batch (1 = attribute hidden), h = their
 * sum.  All attribute strings (hidden ones included) are given: the prover knows them.  Outputs are normalised. ---- */

/* Batched PSRequester::el_passo_request_id (src/ps-requester.cc:19-97).  rnd: N x (h + 2) Fr = t1 (the blinding the
 * requester keeps for psb_unblind), the commitment randomness of g, one per hidden attribute.
 * Out: A (N G1), c (N Fr), rs (N x (h + 1) Fr); the request's attribute list is the input with hidden entries emptied. */
int psb_request_id(psb_key* key, size_t N, const uint8_t* attr_blob, const uint64_t* attr_off, const uint8_t* hide,
                   const uint8_t* ad_blob, const uint64_t* ad_off, const uint64_t* rnd, uint64_t* A, uint64_t* c,
                   uint64_t* rs);

/* Batched PSRequester::unblind_credential (src/ps-requester.cc:99-113): out2 = sig2 - t1 * sig1 (sig1 is unchanged). */
int psb_unblind(size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint64_t* t1, uint64_t* out2);

/* Batched PSRequester::el_passo_prove_id (src/ps-requester.cc:150-310; with_id = 1) and
 * el_passo_prove_id_without_id_retrieval (:312-432; with_id = 0, y/g/h/E1/E2 ignored).
 * rnd per lane: with_id: t, r, epsilon, hid_0..hid_{h-1}, random2, random3 (h + 5); else t, r, hid.., random2 (h + 3).
 * Out: the IdProof fields sig1, sig2, k (G2), phi, E1, E2, c, rs (N x (h + 2) or (h + 1) Fr). */
int psb_prove_id(psb_key* key, size_t N, const uint64_t* sig1, const uint64_t* sig2, const uint8_t* attr_blob,
                 const uint64_t* attr_off, const uint8_t* hide, const uint8_t* ad_blob, const uint64_t* ad_off,
                 const uint64_t* service_pt, const uint64_t* y, const uint64_t* g, const uint64_t* h, int with_id,
                 const uint64_t* rnd, uint64_t* o_sig1, uint64_t* o_sig2, uint64_t* o_k, uint64_t* o_phi, uint64_t* o_E1,
                 uint64_t* o_E2, uint64_t* o_c, uint64_t* o_rs);

/* Batched mcl::bn::hashAndMapToG1 (bn.hpp:2088-2097): message j = blob[off[j] .. off[j+1]);
 * out[j] = the NORMALISED point mcl returns as a group element (SHA-512 -> Fp, Shallue-van de Woestijne map, cofactor
 * (z-1)^2/3); ok[j] = 0 only for the exceptional hash values mcl asserts against.  For relying parties whose batches
 * mix many service names; with one service name per batch the host computes it once (psb_verify_id's service_pt). */
int psb_hash_to_g1(size_t N, const uint8_t* msg_blob, const uint64_t* msg_off, uint64_t* out, uint8_t* ok);

/* Batched mcl::bn::pairing (bn.hpp:1711-1715): out[j] = e(P[j], Q[j]), N x 72 u64. */
int psb_pairing(size_t N, const uint64_t* P, const uint64_t* Q, uint64_t* out);

/* Element-wise arithmetic probe for parity tests (ops in csrc/testops.cuh); one GPU thread per
 * element.  a/b/c/out are host arrays of u32 words with the shapes psb_test_op_shape reports. */
int psb_test_op_shape(int op, int shape[4]);
int psb_test_op(int op, size_t n, const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out);

/* Micro-benchmarks (device-timed, CUDA events): returns milliseconds for `iters` dependent
 * operations per thread over blocks x threads; kind: 0 = Fp mul, 1 = Fp sqr, 2 = Fp2 mul,
 * 3 = Fp12 mul, 4 = raw mad.lo/mad.hi issue-rate probe, 5 = mad.wide probe. */
double psb_microbench(int kind, int blocks, int threads, int iters);

#ifdef __cplusplus
}
#endif
#endif /* PSB_H_ */
