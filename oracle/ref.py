"""oracle/ref.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libpsref.so, the UNMODIFIED
reference (PS-Signature-and-EL-PASSO + mcl) compiled by oracle/Makefile behind ref_harness.cc.

Arrays are numpy uint64 in mcl's in-memory layout (Montgomery limbs): Fp (.., 6), Fr (.., 4),
G1 (.., 18), G2 (.., 36), GT (.., 72).  Only tests/, smoke() and bench.py's CPU legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

# One curve per process (mcl keeps the curve in static state): PSB_CURVE=bn254 selects the BN254 build of the
# reference (oracle/Makefile target `bn254`, packed 256-bit objects), anything else BLS12-381.
CURVE = os.environ.get("PSB_CURVE", "bls12_381").lower()
assert CURVE in ("bls12_381", "bn254"), CURVE
BN254 = CURVE == "bn254"
LIB_PATH = os.path.join(_HERE, "_ref", "libpsref_bn254.so" if BN254 else "libpsref.so")
MCL_CURVE = 0 if BN254 else 5            # mcl/include/mcl/curve_type.h
FP = 4 if BN254 else 6                   # u64 words of an Fp
FR, G1, G2, GT, FP2, FP6 = 4, 3 * FP, 6 * FP, 12 * FP, 2 * FP, 6 * FP
SZ1, SZ2 = 8 * FP, 16 * FP               # serialized Fp / compressed G1, compressed G2 (bytes)
CRED = 2 * (2 + SZ1)                     # PSCredential::toBufferString: two TLV-framed G1 (100 / 68 bytes)
OP_ADD, OP_SUB, OP_MUL, OP_SQR, OP_NEG, OP_INV = range(6)
G_ADD, G_SUB, G_DBL, G_NEG, G_NORM = range(5)

_lib = None


def build(force: bool = False) -> bool:
    """Compile the reference where it lies (only possible where /root/reference exists)."""
    if os.path.exists(LIB_PATH) and not force:
        return True
    if not os.path.isdir(REF_ROOT):
        return False
    subprocess.check_call(["make", "-C", _HERE, "-j8", "all"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libpsref.so missing: run `make -C oracle` where "
                               "/root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_time_pairing.restype = C.c_double
        _lib.ref_time_ps_verify.restype = C.c_double
        _lib.ref_key_create.restype = C.c_void_p
        _lib.ref_signer_create.restype = C.c_void_p
        _lib.ref_key_encode.restype = C.c_size_t
        if _lib.ref_init(MCL_CURVE) != 0:
            raise RuntimeError(f"ref_init({CURVE}) failed")
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def hw_threads() -> int:
    return int(lib().ref_hw_threads())


def jit_enabled() -> bool:
    return bool(lib().ref_jit_enabled())


def seed(s: int) -> None:
    lib().ref_seed(C.c_uint64(s))


def _field_op(fn, width, op, a, b=None):
    a = _u64(a).reshape(-1, width)
    b = a if b is None else _u64(b).reshape(-1, width)
    out = np.empty_like(a)
    fn(C.c_int(op), C.c_size_t(a.shape[0]), _p(a), _p(b), _p(out))
    return out


def fp_op(op, a, b=None):
    return _field_op(lib().ref_fp_op, FP, op, a, b)


def fr_op(op, a, b=None):
    return _field_op(lib().ref_fr_op, FR, op, a, b)


def fp2_op(op, a, b=None):
    return _field_op(lib().ref_fp2_op, FP2, op, a, b)


def fp6_op(op, a, b=None):
    return _field_op(lib().ref_fp6_op, FP6, op, a, b)


def fp12_op(op, a, b=None):
    return _field_op(lib().ref_fp12_op, GT, op, a, b)


def fp12_frobenius(k, a):
    a = _u64(a).reshape(-1, GT)
    out = np.empty_like(a)
    lib().ref_fp12_frobenius(C.c_int(k), C.c_size_t(a.shape[0]), _p(a), _p(out))
    return out


def fp_from_ints(vals):
    buf = np.frombuffer(b"".join(int(v).to_bytes(SZ1, "little") for v in vals), dtype=np.uint8).copy()
    out = np.empty((len(vals), FP), dtype=np.uint64)
    lib().ref_fp_from_bytes(C.c_size_t(len(vals)), _p(buf), _p(out))
    return out


def fr_from_ints(vals):
    buf = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).copy()
    out = np.empty((len(vals), FR), dtype=np.uint64)
    lib().ref_fr_from_bytes(C.c_size_t(len(vals)), _p(buf), _p(out))
    return out


def fr_to_ints(a):
    a = _u64(a).reshape(-1, FR)
    buf = np.empty(a.shape[0] * 32, dtype=np.uint8)
    lib().ref_fr_to_bytes(C.c_size_t(a.shape[0]), _p(a), _p(buf))
    b = buf.tobytes()
    return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def fr_rand(n):
    out = np.empty((n, FR), dtype=np.uint64)
    lib().ref_fr_rand(C.c_size_t(n), _p(out))
    return out


def fr_set_hash_of(msg: bytes):
    out = np.empty(FR, dtype=np.uint64)
    lib().ref_fr_set_hash_of(C.c_char_p(msg), C.c_size_t(len(msg)), _p(out))
    return out


def pack_strings(strs):
    """list of bytes -> (blob uint8, offsets uint64[len+1])"""
    off = np.zeros(len(strs) + 1, dtype=np.uint64)
    if len(strs):
        off[1:] = np.cumsum([len(s) for s in strs])
    blob = np.frombuffer(b"".join(strs) + b"\0", dtype=np.uint8).copy()
    return blob, off


def fr_set_hash_of_batch(strs):
    blob, off = pack_strings(strs)
    out = np.empty((len(strs), FR), dtype=np.uint64)
    lib().ref_fr_set_hash_of_batch(C.c_size_t(len(strs)), _p(blob), _p(off), _p(out))
    return out


def sha256(msg: bytes) -> bytes:
    out = np.empty(32, dtype=np.uint8)
    lib().ref_sha256(C.c_char_p(msg), C.c_size_t(len(msg)), _p(out))
    return out.tobytes()


def g1_op(op, a, b=None):
    return _field_op(lib().ref_g1_op, G1, op, a, b)


def g2_op(op, a, b=None):
    return _field_op(lib().ref_g2_op, G2, op, a, b)


def _mul(fn, width, P, k, nthreads):
    P = _u64(P).reshape(-1, width)
    k = _u64(k).reshape(-1, FR)
    n = k.shape[0]
    stride = 0 if P.shape[0] == 1 and n != 1 else 1
    out = np.empty((n, width), dtype=np.uint64)
    fn(C.c_size_t(n), _p(P), C.c_size_t(stride), _p(k), _p(out), C.c_int(nthreads))
    return out


def g1_mul(P, k, nthreads=1):
    return _mul(lib().ref_g1_mul, G1, P, k, nthreads)


def g2_mul(P, k, nthreads=1):
    return _mul(lib().ref_g2_mul, G2, P, k, nthreads)


def g1_serialize(P):
    P = _u64(P).reshape(-1, G1)
    out = np.empty((P.shape[0], SZ1), dtype=np.uint8)
    lib().ref_g1_serialize(C.c_size_t(P.shape[0]), _p(P), _p(out))
    return out


def g2_serialize(P):
    P = _u64(P).reshape(-1, G2)
    out = np.empty((P.shape[0], SZ2), dtype=np.uint8)
    lib().ref_g2_serialize(C.c_size_t(P.shape[0]), _p(P), _p(out))
    return out


def g1_deserialize(b):
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, SZ1)
    out = np.empty((b.shape[0], G1), dtype=np.uint64)
    ok = lib().ref_g1_deserialize(C.c_size_t(b.shape[0]), _p(b), _p(out))
    return out, bool(ok)


def g2_deserialize(b):
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, SZ2)
    out = np.empty((b.shape[0], G2), dtype=np.uint64)
    ok = lib().ref_g2_deserialize(C.c_size_t(b.shape[0]), _p(b), _p(out))
    return out, bool(ok)


def cred_encode(sig1, sig2):
    """PSCredential::toBufferString of every lane -> uint8 (N, 100)."""
    sig1 = _u64(sig1).reshape(-1, G1)
    sig2 = _u64(sig2).reshape(-1, G1)
    N = sig1.shape[0]
    out = np.zeros(N * CRED, dtype=np.uint8)
    lib().ref_cred_encode.restype = C.c_size_t
    used = lib().ref_cred_encode(C.c_size_t(N), _p(sig1), _p(sig2), _p(out), C.c_size_t(out.size))
    assert used == N * CRED, used
    return out.reshape(N, CRED)


def cred_decode(buf):
    """PSCredential::fromBufferString of every lane (uint8 (N, 100)) -> sig1, sig2."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    N, stride = buf.shape
    s1 = np.zeros((N, G1), dtype=np.uint64)
    s2 = np.zeros((N, G1), dtype=np.uint64)
    lib().ref_cred_decode(C.c_size_t(N), _p(buf), C.c_size_t(stride), _p(s1), _p(s2))
    return s1, s2


def hash_to_g1(msg: bytes):
    out = np.empty(G1, dtype=np.uint64)
    lib().ref_hash_to_g1(C.c_char_p(msg), C.c_size_t(len(msg)), _p(out))
    return out


def map_to_g1(t):
    """mcl mapToG1 (calcBN + cofactor clearing) of one Fp (6 u64) -> (point, ok)."""
    out = np.zeros(G1, dtype=np.uint64)
    ok = lib().ref_map_to_g1(_p(_u64(t)), _p(out))
    return out, int(ok)


def fp_set_hash_of(msg: bytes):
    out = np.empty(FP, dtype=np.uint64)
    lib().ref_fp_set_hash_of(C.c_char_p(msg), C.c_size_t(len(msg)), _p(out))
    return out


def hash_to_g2(msg: bytes):
    out = np.empty(G2, dtype=np.uint64)
    lib().ref_hash_to_g2(C.c_char_p(msg), C.c_size_t(len(msg)), _p(out))
    return out


def miller_loop(P, Q):
    P = _u64(P).reshape(-1, G1)
    Q = _u64(Q).reshape(-1, G2)
    out = np.empty((P.shape[0], GT), dtype=np.uint64)
    lib().ref_miller_loop(C.c_size_t(P.shape[0]), _p(P), _p(Q), _p(out))
    return out


def final_exp(f):
    f = _u64(f).reshape(-1, GT)
    out = np.empty_like(f)
    lib().ref_final_exp(C.c_size_t(f.shape[0]), _p(f), _p(out))
    return out


def pairing(P, Q, nthreads=1):
    P = _u64(P).reshape(-1, G1)
    Q = _u64(Q).reshape(-1, G2)
    out = np.empty((P.shape[0], GT), dtype=np.uint64)
    lib().ref_pairing(C.c_size_t(P.shape[0]), _p(P), _p(Q), _p(out), C.c_int(nthreads))
    return out


def pairing_ratio(P1, Q1, P2, Q2, nthreads=1):
    P1 = _u64(P1).reshape(-1, G1)
    out = np.empty((P1.shape[0], GT), dtype=np.uint64)
    lib().ref_pairing_ratio(C.c_size_t(P1.shape[0]), _p(P1), _p(_u64(Q1)), _p(_u64(P2)),
                            _p(_u64(Q2)), _p(out), C.c_int(nthreads))
    return out


def time_pairing(iters, nthreads):
    return float(lib().ref_time_pairing(C.c_size_t(iters), C.c_int(nthreads)))


class KeyMaterial:
    """own keygen with known exponents; all key points normalized (z = 1)."""

    def __init__(self, n, seed_=1, g_label=b"abc", gg_label=b"edf"):
        self.n = n
        self.g = np.empty(G1, dtype=np.uint64)
        self.gg = np.empty(G2, dtype=np.uint64)
        self.XX = np.empty(G2, dtype=np.uint64)
        self.Y = np.empty((n, G1), dtype=np.uint64)
        self.YY = np.empty((n, G2), dtype=np.uint64)
        self.X = np.empty(G1, dtype=np.uint64)
        self.x = np.empty(FR, dtype=np.uint64)
        self.y = np.empty((n, FR), dtype=np.uint64)
        self.seed = seed_
        self.labels = (g_label, gg_label)
        lib().ref_keygen(C.c_size_t(n), C.c_uint64(seed_), C.c_char_p(g_label), C.c_char_p(gg_label),
                         _p(self.g), _p(self.gg), _p(self.XX), _p(self.Y), _p(self.YY), _p(self.X),
                         _p(self.x), _p(self.y))
        self.handle = C.c_void_p(lib().ref_key_create(_p(self.g), _p(self.gg), _p(self.XX), _p(self.Y),
                                                      _p(self.YY), C.c_size_t(n), _p(self.X)))
        self._signer = None

    def signer(self):
        """the reference's own PSSigner holding the same key (key_gen replayed under the seed)."""
        if self._signer is None:
            self._signer = C.c_void_p(lib().ref_signer_create(
                C.c_size_t(self.n), C.c_uint64(self.seed), C.c_char_p(self.labels[0]),
                C.c_char_p(self.labels[1])))
        return self._signer

    def encode(self) -> bytes:
        buf = np.empty(1 << 16, dtype=np.uint8)
        n = lib().ref_key_encode(self.handle, _p(buf), C.c_size_t(buf.size))
        return buf[:n].tobytes()


def pack_attrs(attrs):
    """attrs: list (lanes) of list (n) of bytes -> blob, offsets[N*n+1]"""
    flat = [a for lane in attrs for a in lane]
    return pack_strings(flat)


def ps_verify(key: KeyMaterial, sig1, sig2, attrs, want_gt=False, nthreads=1):
    sig1 = _u64(sig1).reshape(-1, G1)
    sig2 = _u64(sig2).reshape(-1, G1)
    N = sig1.shape[0]
    blob, off = pack_attrs(attrs)
    verdict = np.empty(N, dtype=np.uint8)
    gt = np.empty((N, GT), dtype=np.uint64) if want_gt else None
    lib().ref_ps_verify(key.handle, C.c_size_t(N), C.c_size_t(key.n), _p(sig1), _p(sig2), _p(blob),
                        _p(off), _p(verdict), _p(gt), C.c_int(nthreads))
    return (verdict, gt) if want_gt else verdict


def ps_verify_packed(key, sig1, sig2, blob, off, nthreads=1, timed=False):
    N = sig1.shape[0]
    verdict = np.empty(N, dtype=np.uint8)
    if timed:
        t = lib().ref_time_ps_verify(key.handle, C.c_size_t(N), C.c_size_t(key.n), _p(sig1), _p(sig2),
                                     _p(blob), _p(off), _p(verdict), C.c_int(nthreads))
        return verdict, float(t)
    lib().ref_ps_verify(key.handle, C.c_size_t(N), C.c_size_t(key.n), _p(sig1), _p(sig2), _p(blob),
                        _p(off), _p(verdict), None, C.c_int(nthreads))
    return verdict


def randomize(sig1, sig2, t, nthreads=1):
    sig1 = _u64(sig1).reshape(-1, G1)
    sig2 = _u64(sig2).reshape(-1, G1)
    t = _u64(t).reshape(-1, FR)
    N = sig1.shape[0]
    o1 = np.empty_like(sig1)
    o2 = np.empty_like(sig2)
    ser = np.empty((N, SZ2), dtype=np.uint8)
    lib().ref_randomize(C.c_size_t(N), _p(sig1), _p(sig2), _p(t), _p(o1), _p(o2), _p(ser),
                        C.c_int(nthreads))
    return o1, o2, ser


def randomize_seeded(key, seed_, sig1, sig2):
    o1 = np.empty(G1, dtype=np.uint64)
    o2 = np.empty(G1, dtype=np.uint64)
    t = np.empty(FR, dtype=np.uint64)
    lib().ref_randomize_seeded(key.handle, C.c_uint64(seed_), _p(_u64(sig1)), _p(_u64(sig2)), _p(o1),
                               _p(o2), _p(t))
    return o1, o2, t


def request_id(key, attrs, hidden, ads, seed_, nthreads=1):
    N = len(attrs)
    blob, off = pack_attrs(attrs)
    ad_blob, ad_off = pack_strings(ads)
    hidden = np.ascontiguousarray(hidden, dtype=np.uint8)
    h = int(hidden.sum())
    A = np.empty((N, G1), dtype=np.uint64)
    c = np.empty((N, FR), dtype=np.uint64)
    rs = np.empty((N, h + 1, FR), dtype=np.uint64)
    lib().ref_request_id(key.handle, C.c_size_t(N), C.c_size_t(key.n), _p(blob), _p(off), _p(hidden),
                         _p(ad_blob), _p(ad_off), C.c_uint64(seed_), _p(A), _p(c), _p(rs),
                         C.c_int(nthreads))
    return A, c, rs


def lane_draws(seed_, lanes, count):
    """the first `count` Fr::setByCSPRNG draws of each lane's stream (seed_ + lane): exactly the scalars the
    reference's prover methods consume under ref_seed(seed_ + lane), in draw order -> (lanes, count, 4)."""
    out = np.empty((lanes, count, FR), dtype=np.uint64)
    for j in range(lanes):
        seed(seed_ + j)
        out[j] = fr_rand(count)
    return out


def unblind(key, attrs, hidden, ads, seed_, sig1, sig2, nthreads=1):
    """replays el_passo_request_id under (seed_ + lane) to set m_t1, then the real unblind_credential."""
    N = len(attrs)
    blob, off = pack_attrs(attrs)
    ad_blob, ad_off = pack_strings(ads)
    hidden = np.ascontiguousarray(hidden, dtype=np.uint8)
    o1 = np.empty((N, G1), dtype=np.uint64)
    o2 = np.empty((N, G1), dtype=np.uint64)
    lib().ref_unblind(key.handle, C.c_size_t(N), C.c_size_t(key.n), _p(blob), _p(off), _p(hidden), _p(ad_blob),
                      _p(ad_off), C.c_uint64(seed_), _p(_u64(sig1)), _p(_u64(sig2)), _p(o1), _p(o2),
                      C.c_int(nthreads))
    return o1, o2


def provide_id(key, A, c, rs, attrs, ads, u, nthreads=1):
    """attrs here are the REQUEST's attribute lists (b"" for hidden)."""
    N = A.shape[0]
    blob, off = pack_attrs(attrs)
    ad_blob, ad_off = pack_strings(ads)
    rs = _u64(rs)
    per = rs.shape[1]
    verdict = np.empty(N, dtype=np.uint8)
    s1 = np.empty((N, G1), dtype=np.uint64)
    s2 = np.empty((N, G1), dtype=np.uint64)
    ser = np.empty((N, SZ2), dtype=np.uint8)
    lib().ref_provide_id(key.signer(), C.c_size_t(N), C.c_size_t(key.n), _p(_u64(A)), _p(_u64(c)),
                         _p(rs), C.c_size_t(per), _p(blob), _p(off), _p(ad_blob), _p(ad_off),
                         _p(_u64(u)), _p(verdict), _p(s1), _p(s2), _p(ser), C.c_int(nthreads))
    return verdict, s1, s2, ser


def prove_id(key, sig1, sig2, attrs, hidden, ads, service: bytes, y, g, h, seed_, with_id=True,
             nthreads=1):
    N = len(attrs)
    blob, off = pack_attrs(attrs)
    ad_blob, ad_off = pack_strings(ads)
    hidden = np.ascontiguousarray(hidden, dtype=np.uint8)
    per = int(hidden.sum()) + (2 if with_id else 1)
    o = dict(sig1=np.empty((N, G1), dtype=np.uint64), sig2=np.empty((N, G1), dtype=np.uint64),
             k=np.empty((N, G2), dtype=np.uint64), phi=np.empty((N, G1), dtype=np.uint64),
             E1=np.zeros((N, G1), dtype=np.uint64), E2=np.zeros((N, G1), dtype=np.uint64),
             c=np.empty((N, FR), dtype=np.uint64), rs=np.empty((N, per, FR), dtype=np.uint64))
    lib().ref_prove_id(key.handle, C.c_size_t(N), C.c_size_t(key.n), _p(_u64(sig1)), _p(_u64(sig2)),
                       _p(blob), _p(off), _p(hidden), _p(ad_blob), _p(ad_off), C.c_char_p(service),
                       _p(_u64(y)), _p(_u64(g)), _p(_u64(h)), C.c_uint64(seed_), C.c_int(int(with_id)),
                       _p(o["sig1"]), _p(o["sig2"]), _p(o["k"]), _p(o["phi"]), _p(o["E1"]), _p(o["E2"]),
                       _p(o["c"]), _p(o["rs"]), C.c_int(nthreads))
    return o


def verify_id(key, proof, attrs, ads, service: bytes, y, g, h, with_id=True, nthreads=1):
    """attrs here are the PROOF's attribute lists (b"" for hidden)."""
    N = proof["sig1"].shape[0]
    blob, off = pack_attrs(attrs)
    ad_blob, ad_off = pack_strings(ads)
    rs = _u64(proof["rs"])
    verdict = np.empty(N, dtype=np.uint8)
    lib().ref_verify_id(key.handle, C.c_size_t(N), C.c_size_t(key.n), _p(_u64(proof["sig1"])),
                        _p(_u64(proof["sig2"])), _p(_u64(proof["k"])), _p(_u64(proof["phi"])),
                        _p(_u64(proof["E1"])), _p(_u64(proof["E2"])), _p(_u64(proof["c"])), _p(rs),
                        C.c_size_t(rs.shape[1]), _p(blob), _p(off), _p(ad_blob), _p(ad_off),
                        C.c_char_p(service), _p(_u64(y)), _p(_u64(g)), _p(_u64(h)),
                        C.c_int(int(with_id)), _p(verdict), C.c_int(nthreads))
    return verdict


# ---- wire formats (src/ps-encoding.cc): IdProof / PSCredRequest through the reference's own encoder and parser ----------
def idproof_encode(key, proof, attrs, with_e=True, base64=False):
    """IdProof::toBufferString (or its base64 text) of every lane -> packed (blob uint8, off uint64[N+1])."""
    N = proof["sig1"].shape[0]
    blob, off = pack_attrs(attrs)
    rs = _u64(proof["rs"]).reshape(N, -1, FR)
    cap = N * (16 * G1 * 8 + rs.shape[1] * 40 + 64) * 2 + int(off[-1]) * 2 + 4 * len(off) + 1024
    out = np.zeros(cap, dtype=np.uint8)
    ooff = np.zeros(N + 1, dtype=np.uint64)
    lib().ref_idproof_encode.restype = C.c_size_t
    used = lib().ref_idproof_encode(C.c_size_t(N), C.c_size_t(key.n), _p(_u64(proof["sig1"])), _p(_u64(proof["sig2"])),
                                    _p(_u64(proof["k"])), _p(_u64(proof["phi"])), _p(_u64(proof["E1"])), _p(_u64(proof["E2"])),
                                    _p(_u64(proof["c"])), _p(rs), C.c_size_t(rs.shape[1]), _p(blob), _p(off),
                                    C.c_int(int(with_e)), C.c_int(int(base64)), _p(out), _p(ooff), C.c_size_t(cap))
    assert used and used == int(ooff[-1])
    return out[:used + 8].copy(), ooff


def request_encode(key, A, c, rs, attrs, base64=False):
    """PSCredRequest::toBufferString (or base64) of every lane -> packed (blob, off)."""
    N = A.shape[0]
    blob, off = pack_attrs(attrs)
    rs = _u64(rs).reshape(N, -1, FR)
    cap = N * (4 * G1 * 8 + rs.shape[1] * 40 + 64) * 2 + int(off[-1]) * 2 + 4 * len(off) + 1024
    out = np.zeros(cap, dtype=np.uint8)
    ooff = np.zeros(N + 1, dtype=np.uint64)
    lib().ref_request_encode.restype = C.c_size_t
    used = lib().ref_request_encode(C.c_size_t(N), C.c_size_t(key.n), _p(_u64(A)), _p(_u64(c)), _p(rs), C.c_size_t(rs.shape[1]),
                                    _p(blob), _p(off), C.c_int(int(base64)), _p(out), _p(ooff), C.c_size_t(cap))
    assert used and used == int(ooff[-1])
    return out[:used + 8].copy(), ooff


def verify_id_wire(key, wire, ads, service: bytes, y, g, h, with_id=True, base64=False, nthreads=1):
    """IdProof::fromBufferString + el_passo_verify_id on every lane -> (verdict, status); status 1 = the reference threw."""
    blob, off = wire
    N = off.shape[0] - 1
    ad_blob, ad_off = pack_strings(ads)
    verdict = np.zeros(N, dtype=np.uint8)
    status = np.zeros(N, dtype=np.uint8)
    lib().ref_verify_id_wire(key.handle, C.c_size_t(N), _p(blob), _p(off), C.c_int(int(base64)), _p(ad_blob), _p(ad_off),
                             C.c_char_p(service), _p(_u64(y)), _p(_u64(g)), _p(_u64(h)), C.c_int(int(with_id)),
                             _p(verdict), _p(status), C.c_int(nthreads))
    return verdict, status


def idproof_decode(key, wire, rs_cap, base64=False):
    """the fields IdProof::fromBufferString produces (dict of arrays) + per, has_e, status per lane."""
    blob, off = wire
    N = off.shape[0] - 1
    o = dict(sig1=np.zeros((N, G1), np.uint64), sig2=np.zeros((N, G1), np.uint64), k=np.zeros((N, G2), np.uint64),
             phi=np.zeros((N, G1), np.uint64), E1=np.zeros((N, G1), np.uint64), E2=np.zeros((N, G1), np.uint64),
             c=np.zeros((N, FR), np.uint64), rs=np.zeros((N, rs_cap, FR), np.uint64))
    per = np.zeros(N, dtype=np.int32)
    has_e = np.zeros(N, dtype=np.uint8)
    status = np.zeros(N, dtype=np.uint8)
    lib().ref_idproof_decode(C.c_size_t(N), C.c_size_t(key.n), _p(blob), _p(off), C.c_int(int(base64)), _p(o["sig1"]), _p(o["sig2"]),
                             _p(o["k"]), _p(o["phi"]), _p(o["E1"]), _p(o["E2"]), _p(o["c"]), _p(o["rs"]), C.c_size_t(rs_cap),
                             _p(per), _p(has_e), _p(status))
    return o, per, has_e, status


def provide_id_wire(key, wire, ads, u, base64=False, nthreads=1):
    blob, off = wire
    N = off.shape[0] - 1
    ad_blob, ad_off = pack_strings(ads)
    verdict = np.zeros(N, dtype=np.uint8)
    status = np.zeros(N, dtype=np.uint8)
    s1 = np.zeros((N, G1), dtype=np.uint64)
    s2 = np.zeros((N, G1), dtype=np.uint64)
    ser = np.zeros((N, SZ2), dtype=np.uint8)
    lib().ref_provide_id_wire(key.signer(), C.c_size_t(N), _p(blob), _p(off), C.c_int(int(base64)), _p(ad_blob), _p(ad_off),
                              _p(_u64(u)), _p(verdict), _p(s1), _p(s2), _p(ser), _p(status), C.c_int(nthreads))
    return verdict, s1, s2, ser, status


def sign(key, commitment, attrs, u, nthreads=1):
    """PSSigner::sign_hybrid (attrs: per lane na strings) or sign_commitment (attrs None) with the RandGen primed to u."""
    Cm = _u64(commitment).reshape(-1, G1)
    N = Cm.shape[0]
    if attrs is None:
        blob, off, na = np.zeros(8, dtype=np.uint8), np.zeros(1, dtype=np.uint64), 0
    else:
        blob, off = pack_attrs(attrs)
        na = len(attrs[0]) if N else 0
    s1 = np.zeros((N, G1), dtype=np.uint64)
    s2 = np.zeros((N, G1), dtype=np.uint64)
    ser = np.zeros((N, SZ2), dtype=np.uint8)
    lib().ref_sign(key.signer(), C.c_size_t(N), C.c_size_t(na), _p(Cm), _p(blob), _p(off), _p(_u64(u)), _p(s1), _p(s2), _p(ser),
                   C.c_int(nthreads))
    return s1, s2, ser
