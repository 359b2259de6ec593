// wire.cuh -- the reference's wire formats parsed per lane on the device.
//
// A relying party receives base64 text of IdProof::toBufferString(), an identity provider base64 text of
// PSCredRequest::toBufferString(); decoding them with the reference's parser on the host (base64, TLV walk, one
// square root per compressed point) would cap a GPU box long before the pairing kernels do.  These lane functions
// restate that parser (reference: src/ps-encoding.cc):
//   base64_decode            :56-96    stops at the first '=' or non-alphabet byte; a partial last group yields i - 1 bytes
//   parseVar                 :148-161  one byte below 253, or 253 followed by a big-endian u16; 254 / 255 are invalid
//   parseG1/G2/FrElement     :178-256  [type byte] var(size) payload; the element is mcl's deserialize of the payload
//   parseFrList / parseStrList :332-384  type byte, var(count), then var(size) payload per entry (no type bytes)
//   IdProof::fromBufferString :470-489  sig1 sig2 k phi c rs attributes [E1 E2 when bytes remain]
//   PSCredRequest::fromBufferString :441-450  A c rs attributes
// Point payloads are NOT decompressed here: the parser records where they lie and k_wire_points decompresses every
// (lane, slot) pair with one thread each (csrc/protocol.cuh g1_deserialize / g2_deserialize = mcl's EcT::load).
//
// Where the reference is undefined we reject (the lane's `parsed` flag and verdict are 0): a buffer shorter than its
// own length fields claim (PSBuffer::at throws / mcl reads past the end), a wrong type byte (the reference returns 0
// and keeps walking an uninitialised object), an element payload shorter than the element (deserialize fails and the
// reference IGNORES the failure -- SURVEY F9 -- leaving a point that is not on the curve), a scalar >= r, an attribute
// count different from the key's, more than n + 2 responses.  Payloads LONGER than the element are accepted like the
// reference accepts them (mcl reads the leading bytes, the walk skips the whole payload).
#pragma once
#include "protocol.cuh"

namespace psb {

constexpr uint32_t kWireAbsent = 0xffffffffu;
// slots of the point table one lane fills: byte offsets of the compressed payloads inside the lane's buffer
enum { W_SIG1 = 0, W_SIG2 = 1, W_PHI = 2, W_E1 = 3, W_E2 = 4, W_K = 5, W_SLOTS = 6 };   // PSCredRequest: A in W_SIG1
constexpr int kFrBytes = 32;

PSB_HD PSB_INL int b64_value(uint8_t c) {
  if (c >= 'A' && c <= 'Z') return c - 'A';
  if (c >= 'a' && c <= 'z') return c - 'a' + 26;
  if (c >= '0' && c <= '9') return c - '0' + 52;
  if (c == '+') return 62;
  if (c == '/') return 63;
  return -1;
}
// base64_decode (src/ps-encoding.cc:56-96): returns the number of bytes written (<= 3 * len / 4); out may alias in
PSB_HD PSB_NOINL size_t base64_decode_lane(uint8_t* out, const uint8_t* in, size_t len) {
  size_t o = 0;
  uint32_t acc = 0;
  int i = 0;
  for (size_t p = 0; p < len; p++) {
    const int v = b64_value(in[p]);
    if (v < 0) break;                       // '=' or any byte outside the alphabet ends the input
    acc = (acc << 6) | (uint32_t)v;
    if (++i == 4) {
      out[o] = (uint8_t)(acc >> 16); out[o + 1] = (uint8_t)(acc >> 8); out[o + 2] = (uint8_t)acc;
      o += 3; i = 0; acc = 0;
    }
  }
  if (i) {                                  // partial group of i sextets: i - 1 bytes
    acc <<= 6 * (4 - i);
    const uint8_t b3[3] = {(uint8_t)(acc >> 16), (uint8_t)(acc >> 8), (uint8_t)acc};
    for (int j = 0; j < i - 1; j++) out[o++] = b3[j];
  }
  return o;
}

struct WireCursor {
  const uint8_t* b;
  size_t len, at;
  bool ok;
};
PSB_HD PSB_INL uint8_t wire_byte(WireCursor& c) {
  if (c.at >= c.len) { c.ok = false; return 0; }
  return c.b[c.at++];
}
PSB_HD PSB_INL size_t wire_var(WireCursor& c) {          // parseVar
  const uint8_t f = wire_byte(c);
  if (f < 253) return f;
  if (f != 253) { c.ok = false; return 0; }
  const size_t hi = wire_byte(c), lo = wire_byte(c);
  return (hi << 8) | lo;
}
// one element payload of at least `need` bytes: returns its offset, advances past the whole payload
PSB_HD PSB_INL uint32_t wire_payload(WireCursor& c, size_t need) {
  const size_t size = wire_var(c);
  if (!c.ok || size < need || c.at + size > c.len) { c.ok = false; return kWireAbsent; }
  const size_t at = c.at;
  c.at += size;
  return (uint32_t)at;
}
PSB_HD PSB_INL uint32_t wire_typed(WireCursor& c, uint8_t type, size_t need) {
  if (wire_byte(c) != type) { c.ok = false; return kWireAbsent; }
  return wire_payload(c, need);
}
// Fr::deserialize: 32 little-endian bytes, value < r (mcl rejects the rest), stored in Montgomery form
PSB_HD PSB_INL bool fr_from_le_bytes(Fr& r, const uint8_t* b) {
  Fr n;
  for (int i = 0; i < 8; i++)
    n.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  if (fr_geq_modulus(n.v)) return false;
  fr_to_mont(r, n);
  return true;
}

// Common tail of both messages: c (typed Fr), rs (FrList), attributes (StrList).  rs_out has n + 2 slots; the attribute
// strings are copied to attr_out back to back and attr_off[0 .. n] receives their offsets RELATIVE TO attr_base
// (attr_off[i + 1] - attr_off[i] = length; "" = hidden attribute).
PSB_HD PSB_INL void wire_scalars_and_strings(WireCursor& c, int n, Fr& cc, Fr* rs_out, int& per, uint8_t* attr_out,
                                             uint64_t* attr_off, uint64_t attr_base) {
  const uint32_t pc = wire_typed(c, 3, kFrBytes);
  if (c.ok && !fr_from_le_bytes(cc, c.b + pc)) c.ok = false;
  if (wire_byte(c) != 6) c.ok = false;
  const size_t cnt = wire_var(c);
  if (cnt > (size_t)n + 2) c.ok = false;
  per = c.ok ? (int)cnt : 0;
  for (int i = 0; i < per && c.ok; i++) {
    const uint32_t p = wire_payload(c, kFrBytes);
    if (c.ok && !fr_from_le_bytes(rs_out[i], c.b + p)) c.ok = false;
  }
  if (wire_byte(c) != 7) c.ok = false;
  const size_t na = wire_var(c);
  if (na != (size_t)n) c.ok = false;
  uint64_t w = 0;
  for (int i = 0; i <= n; i++) attr_off[i] = attr_base;      // defined offsets even for a rejected lane
  for (int i = 0; i < n && c.ok; i++) {
    const size_t sl = wire_var(c);
    if (!c.ok || c.at + sl > c.len) { c.ok = false; break; }
    for (size_t t = 0; t < sl; t++) attr_out[w + t] = c.b[c.at + t];
    c.at += sl;
    w += sl;
    attr_off[i + 1] = attr_base + w;
  }
  if (!c.ok) for (int i = 0; i <= n; i++) attr_off[i] = attr_base;
}

// IdProof::fromBufferString.  pos[W_SLOTS]: payload offsets (kWireAbsent when the lane has no such point).
// Returns the structural verdict; has_e = E1 and E2 present.
PSB_HD PSB_NOINL bool parse_idproof_lane(const uint8_t* b, size_t len, int n, uint32_t* pos, Fr& cc, Fr* rs_out, int& per,
                                          uint8_t* attr_out, uint64_t* attr_off, uint64_t attr_base, bool& has_e) {
  WireCursor c{b, len, 0, true};
  for (int i = 0; i < W_SLOTS; i++) pos[i] = kWireAbsent;
  pos[W_SIG1] = wire_typed(c, 1, kFpBytes);
  pos[W_SIG2] = wire_typed(c, 1, kFpBytes);
  pos[W_K] = wire_typed(c, 2, 2 * kFpBytes);
  pos[W_PHI] = wire_typed(c, 1, kFpBytes);
  wire_scalars_and_strings(c, n, cc, rs_out, per, attr_out, attr_off, attr_base);
  has_e = false;
  if (c.ok && c.at < c.len) {               // `if (step < buf.size())`: E1 and E2 follow
    pos[W_E1] = wire_typed(c, 1, kFpBytes);
    pos[W_E2] = wire_typed(c, 1, kFpBytes);
    has_e = c.ok;
  }
  if (!c.ok) for (int i = 0; i < W_SLOTS; i++) pos[i] = kWireAbsent;
  return c.ok;
}
// PSCredRequest::fromBufferString: A (slot W_SIG1), c, rs, attributes
PSB_HD PSB_NOINL bool parse_request_lane(const uint8_t* b, size_t len, int n, uint32_t* pos, Fr& cc, Fr* rs_out, int& per,
                                          uint8_t* attr_out, uint64_t* attr_off, uint64_t attr_base) {
  WireCursor c{b, len, 0, true};
  for (int i = 0; i < W_SLOTS; i++) pos[i] = kWireAbsent;
  pos[W_SIG1] = wire_typed(c, 1, kFpBytes);
  wire_scalars_and_strings(c, n, cc, rs_out, per, attr_out, attr_off, attr_base);
  if (!c.ok) pos[W_SIG1] = kWireAbsent;
  return c.ok;
}

// ---- the other direction: IdProof::toBufferString / PSCredRequest::toBufferString (src/ps-encoding.cc:429-439, :452-468) ------
// What a batched prover (psb_prove_id, psb_request_id) puts on the wire.  appendVar (:138-146) writes one byte below 253,
// else 253 + big-endian u16 (lengths above 0xFFFF write NOTHING in the reference: psb_wire_encode refuses them); points go
// through mcl's serialize, i.e. they are normalised first (ec.hpp:849-896); scalars are 32 little-endian bytes.
PSB_HD PSB_INL size_t wire_var_size(size_t v) { return v < 253 ? 1 : 3; }
// bytes of one message: kind 0 = IdProof (has_e: E1 and E2 appended), 1 = PSCredRequest; aoff = n + 1 offsets of the lane's strings
PSB_HD PSB_INL size_t wire_message_size(int kind, int n, int per, const uint64_t* aoff, bool has_e) {
  size_t s = (kind == 0 ? 3 * (2 + (size_t)kFpBytes) + (2 + 2 * (size_t)kFpBytes) : 2 + (size_t)kFpBytes) + (2 + kFrBytes) + 1 +
             wire_var_size((size_t)per) + (size_t)per * (1 + kFrBytes) + 1 + wire_var_size((size_t)n);
  for (int i = 0; i < n; i++) { const size_t l = (size_t)(aoff[i + 1] - aoff[i]); s += wire_var_size(l) + l; }
  if (kind == 0 && has_e) s += 2 * (2 + (size_t)kFpBytes);
  return s;
}
PSB_HD PSB_INL void wire_put_var(uint8_t*& o, size_t v) {
  if (v < 253) { *o++ = (uint8_t)v; return; }
  *o++ = 253; *o++ = (uint8_t)(v >> 8); *o++ = (uint8_t)v;
}
PSB_HD PSB_NOINL void wire_put_g1(uint8_t*& o, const G1J& P) {
  G1J Q;
  pt_normalize(Q, P);
  *o++ = 1; *o++ = (uint8_t)kFpBytes;
  g1_serialize_norm(o, Q);
  o += kFpBytes;
}
PSB_HD PSB_NOINL void wire_put_g2(uint8_t*& o, const G2J& P) {
  G2J Q;
  pt_normalize(Q, P);
  *o++ = 2; *o++ = (uint8_t)(2 * kFpBytes);
  g2_serialize_norm(o, Q);
  o += 2 * kFpBytes;
}
PSB_HD PSB_INL void wire_put_fr(uint8_t*& o, const Fr& a, bool typed) {
  Fr n;
  fr_from_mont(n, a);
  if (typed) *o++ = 3;
  *o++ = (uint8_t)kFrBytes;
  for (int i = 0; i < 8; i++) { *o++ = (uint8_t)n.v[i]; *o++ = (uint8_t)(n.v[i] >> 8); *o++ = (uint8_t)(n.v[i] >> 16); *o++ = (uint8_t)(n.v[i] >> 24); }
}
// pts: IdProof = {sig1, sig2, phi, E1, E2} (E1 / E2 null when absent), PSCredRequest = {A}; returns the bytes written
PSB_HD PSB_NOINL size_t encode_message_lane(uint8_t* out, int kind, int n, const G1J* const* pts, const G2J* k, const Fr& cc,
                                            const Fr* rs, int per, const uint8_t* attr, const uint64_t* aoff) {
  uint8_t* o = out;
  wire_put_g1(o, *pts[0]);
  if (kind == 0) {
    wire_put_g1(o, *pts[1]);
    wire_put_g2(o, *k);
    wire_put_g1(o, *pts[2]);
  }
  wire_put_fr(o, cc, true);
  *o++ = 6;
  wire_put_var(o, (size_t)per);
  for (int i = 0; i < per; i++) wire_put_fr(o, rs[i], false);
  *o++ = 7;
  wire_put_var(o, (size_t)n);
  for (int i = 0; i < n; i++) {
    const size_t l = (size_t)(aoff[i + 1] - aoff[i]);
    wire_put_var(o, l);
    for (size_t t = 0; t < l; t++) *o++ = attr[aoff[i] + t];
  }
  if (kind == 0 && pts[3] && pts[4]) { wire_put_g1(o, *pts[3]); wire_put_g1(o, *pts[4]); }
  return (size_t)(o - out);
}
// base64_encode (src/ps-encoding.cc:14-54): '=' padded; 4 * ceil(len / 3) characters
PSB_HD PSB_INL size_t base64_encoded_size(size_t len) { return (len + 2) / 3 * 4; }
PSB_HD PSB_NOINL void base64_encode_lane(uint8_t* out, const uint8_t* in, size_t len) {
  const char* tab = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  size_t o = 0;
  for (size_t p = 0; p < len; p += 3) {
    const int m = (int)(len - p < 3 ? len - p : 3);
    const uint32_t v = ((uint32_t)in[p] << 16) | ((m > 1 ? (uint32_t)in[p + 1] : 0u) << 8) | (m > 2 ? (uint32_t)in[p + 2] : 0u);
    out[o++] = (uint8_t)tab[(v >> 18) & 63];
    out[o++] = (uint8_t)tab[(v >> 12) & 63];
    out[o++] = m > 1 ? (uint8_t)tab[(v >> 6) & 63] : (uint8_t)'=';
    out[o++] = m > 2 ? (uint8_t)tab[v & 63] : (uint8_t)'=';
  }
}

}  // namespace psb
