"""tests/workload.py -- TEST INFRASTRUCTURE: synthetic PS / EL PASSO workloads
generated with the reference oracle (oracle/_ref/libpsref.so), and the oracle's expected outputs.

Used by tests/, __graft_entry__.smoke() and bench.py (expected verdicts + cpu_baseline only).
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402


@dataclass
class VerifyWorkload:
    key: "ref.KeyMaterial"
    sig1: np.ndarray  # (N, 18) normalized G1
    sig2: np.ndarray
    attrs: List[List[bytes]]
    tampered: np.ndarray  # lane indices
    blob: np.ndarray = field(default=None)
    off: np.ndarray = field(default=None)


def attr_strings(n_attrs: int, lanes: int, first_lane: int = 0) -> List[List[bytes]]:
    return [[b"a%d:%d" % (i, j) for i in range(n_attrs)] for j in range(first_lane, first_lane + lanes)]


def sign_lanes(key, attrs, seed: int, nthreads: int = 0):
    """honest PS signatures with known exponents: sigma1 = u g, sigma2 = (x + sum y_i m_i) sigma1,
    both normalized (as after deserialization) -- SURVEY 8d config 2."""
    nthreads = nthreads or ref.hw_threads()
    N, n = len(attrs), key.n
    m = ref.fr_set_hash_of_batch([a for lane in attrs for a in lane]).reshape(N, n, ref.FR)
    s = np.repeat(key.x.reshape(1, ref.FR), N, axis=0)
    for i in range(n):
        yi = np.repeat(key.y[i].reshape(1, ref.FR), N, axis=0)
        s = ref.fr_op(ref.OP_ADD, s, ref.fr_op(ref.OP_MUL, yi, m[:, i, :]))
    ref.seed(seed)
    u = ref.fr_rand(N)
    sig1 = ref.g1_op(ref.G_NORM, ref.g1_mul(key.g, u, nthreads))
    sig2 = ref.g1_op(ref.G_NORM, ref.g1_mul(sig1, s, nthreads))
    return sig1, sig2


def make_verify_workload(n_attrs: int = 5, lanes: int = 64, seed: int = 2, tamper_every: int = 0,
                         key_seed: int = 1) -> VerifyWorkload:
    key = ref.KeyMaterial(n_attrs, seed_=key_seed)
    attrs = attr_strings(n_attrs, lanes)
    sig1, sig2 = sign_lanes(key, attrs, seed)
    tampered = []
    if tamper_every:
        kinds = 0
        for j in range(tamper_every - 1, lanes, tamper_every):
            kind = kinds % 4
            kinds += 1
            if kind == 0:  # sigma2 += g
                sig2[j] = ref.g1_op(ref.G_NORM, ref.g1_op(ref.G_ADD, sig2[j:j + 1], key.g.reshape(1, -1)))[0]
            elif kind == 1 and n_attrs:  # attribute byte flipped
                a = bytearray(attrs[j][0])
                a[0] ^= 1
                attrs[j][0] = bytes(a)
            elif kind == 2:  # sigma1 = 0
                sig1[j] = 0
            else:  # swap
                sig1[j], sig2[j] = sig2[j].copy(), sig1[j].copy()
            tampered.append(j)
    wl = VerifyWorkload(key, sig1, sig2, attrs, np.array(tampered, dtype=np.int64))
    wl.blob, wl.off = ref.pack_attrs(attrs)
    return wl


def expected_verify(wl: VerifyWorkload, want_gt: bool = False, nthreads: int = 0):
    """the reference's own PSVerifier::verify on every lane (+ lhs * unitaryInv(rhs) as GT)."""
    nthreads = nthreads or ref.hw_threads()
    return ref.ps_verify(wl.key, wl.sig1, wl.sig2, wl.attrs, want_gt=want_gt, nthreads=nthreads)


# ---- EL PASSO issuance and sign-on (config 3) workloads ---------------------------
@dataclass
class IssuanceWorkload:
    key: "ref.KeyMaterial"
    A: np.ndarray            # (N, 18)
    c: np.ndarray            # (N, 4)
    rs: np.ndarray           # (N, h+1, 4)
    req_attrs: List[List[bytes]]   # the REQUEST's attribute lists: b"" for hidden
    ads: List[bytes]
    u: np.ndarray            # (N, 4) host-supplied issuance scalars
    tampered: np.ndarray


def _hidden_mask(n_attrs: int, n_hidden: int) -> np.ndarray:
    hidden = np.zeros(n_attrs, dtype=np.uint8)
    hidden[:n_hidden] = 1
    return hidden


def make_issuance_workload(n_attrs: int = 5, lanes: int = 16, n_hidden: int = 2, seed: int = 4,
                           tamper_every: int = 0, key_seed: int = 1, nthreads: int = 0) -> IssuanceWorkload:
    """requests from the reference's own PSRequester::el_passo_request_id under per-lane seeds."""
    nthreads = nthreads or ref.hw_threads()
    key = ref.KeyMaterial(n_attrs, seed_=key_seed)
    attrs = attr_strings(n_attrs, lanes)
    hidden = _hidden_mask(n_attrs, n_hidden)
    ads = [b"sess%d" % j for j in range(lanes)]
    A, c, rs = ref.request_id(key, attrs, hidden, ads, seed * 1000003, nthreads)
    req_attrs = [[b"" if hidden[i] else lane[i] for i in range(n_attrs)] for lane in attrs]
    ref.seed(seed + 1)
    u = ref.fr_rand(lanes)
    tampered = []
    if tamper_every:
        one = ref.fr_from_ints([1])
        for t, j in enumerate(range(tamper_every - 1, lanes, tamper_every)):
            kind = t % 4
            if kind == 0:      # challenge off by one
                c[j] = ref.fr_op(ref.OP_ADD, c[j:j + 1], one)[0]
            elif kind == 1:    # a response off by one
                rs[j, -1] = ref.fr_op(ref.OP_ADD, rs[j, -1:].copy(), one)[0]
            elif kind == 2:    # associated data changed
                ads[j] = ads[j] + b"!"
            else:              # commitment replaced
                A[j] = ref.g1_op(ref.G_ADD, A[j:j + 1], key.g.reshape(1, -1))[0]
            tampered.append(j)
    return IssuanceWorkload(key, A, c, rs, req_attrs, ads, u, np.array(tampered, dtype=np.int64))


def expected_provide_id(wl: IssuanceWorkload, nthreads: int = 0):
    """the reference's own PSSigner::el_passo_provide_id with the RandGen primed to yield u_j."""
    return ref.provide_id(wl.key, wl.A, wl.c, wl.rs, wl.req_attrs, wl.ads, wl.u, nthreads or ref.hw_threads())


@dataclass
class SignonWorkload:
    key: "ref.KeyMaterial"
    proof: dict              # sig1 sig2 k phi E1 E2 c rs
    proof_attrs: List[List[bytes]]
    ads: List[bytes]
    service: bytes
    service_pt: np.ndarray
    y: np.ndarray
    g: np.ndarray
    h: np.ndarray
    with_id: bool
    tampered: np.ndarray


def make_signon_workload(n_attrs: int = 5, lanes: int = 16, n_hidden: int = 2, seed: int = 3, with_id: bool = True,
                         tamper_every: int = 0, key_seed: int = 1, nthreads: int = 0) -> SignonWorkload:
    """proofs from the reference's own PSRequester::el_passo_prove_id (test/ps-tests.cc:106-111 shape)."""
    nthreads = nthreads or ref.hw_threads()
    key = ref.KeyMaterial(n_attrs, seed_=key_seed)
    attrs = attr_strings(n_attrs, lanes)
    hidden = _hidden_mask(n_attrs, n_hidden)
    sig1, sig2 = sign_lanes(key, attrs, seed + 7, nthreads)
    ads = [b"sess%d" % j for j in range(lanes)]
    service = b"rp.example"
    y, g, h = ref.hash_to_g1(b"ghi"), ref.hash_to_g1(b"abc"), ref.hash_to_g1(b"jkl")
    proof = ref.prove_id(key, sig1, sig2, attrs, hidden, ads, service, y, g, h, seed * 1000003, with_id, nthreads)
    proof_attrs = [[b"" if hidden[i] else lane[i] for i in range(n_attrs)] for lane in attrs]
    tampered = []
    if tamper_every:
        one = ref.fr_from_ints([1])
        for t, j in enumerate(range(tamper_every - 1, lanes, tamper_every)):
            kind = t % 6
            if kind == 0:
                proof["c"][j] = ref.fr_op(ref.OP_ADD, proof["c"][j:j + 1], one)[0]
            elif kind == 1:
                proof["rs"][j, 0] = ref.fr_op(ref.OP_ADD, proof["rs"][j, :1].copy(), one)[0]
            elif kind == 2:
                proof["k"][j] = ref.g2_op(ref.G_ADD, proof["k"][j:j + 1], key.gg.reshape(1, -1))[0]
            elif kind == 3:    # NIZK still fine, pairing check fails
                proof["sig2"][j] = ref.g1_op(ref.G_ADD, proof["sig2"][j:j + 1], key.g.reshape(1, -1))[0]
            elif kind == 4:    # sigma = (0, 0): el_passo_verify_id has no zero check
                proof["sig1"][j] = 0
                proof["sig2"][j] = 0
            else:              # plaintext attribute changed: NIZK fine, K differs
                if n_hidden < n_attrs:
                    proof_attrs[j][n_attrs - 1] = proof_attrs[j][n_attrs - 1] + b"x"
                else:
                    ads[j] = ads[j] + b"!"
            tampered.append(j)
    return SignonWorkload(key, proof, proof_attrs, ads, service, ref.hash_to_g1(service), y, g, h, with_id,
                          np.array(tampered, dtype=np.int64))


def expected_verify_id(wl: SignonWorkload, nthreads: int = 0):
    return ref.verify_id(wl.key, wl.proof, wl.proof_attrs, wl.ads, wl.service, wl.y, wl.g, wl.h, wl.with_id,
                         nthreads or ref.hw_threads())


# ---- prover side: the reference's own requester methods under per-lane seeds, plus the scalars
#      those methods drew (replayed from the same streams) ------------------------------------------------------------
@dataclass
class ProverRequestWorkload:
    key: "ref.KeyMaterial"
    attrs: List[List[bytes]]   # ALL attribute values (the requester knows the hidden ones)
    hidden: np.ndarray
    ads: List[bytes]
    rnd: np.ndarray            # (N, h+2, 4): t1, r0, one per hidden attribute
    exp_A: np.ndarray          # reference outputs (A is raw Jacobian)
    exp_c: np.ndarray
    exp_rs: np.ndarray
    blind_sig1: np.ndarray     # a blinded credential for the request (issued by the reference signer)
    blind_sig2: np.ndarray
    exp_unblind2: np.ndarray


def make_prover_request_workload(n_attrs=5, lanes=8, n_hidden=2, seed=6, key_seed=1, nthreads=0) -> ProverRequestWorkload:
    nthreads = nthreads or ref.hw_threads()
    key = ref.KeyMaterial(n_attrs, seed_=key_seed)
    attrs = attr_strings(n_attrs, lanes)
    hidden = _hidden_mask(n_attrs, n_hidden)
    ads = [b"sess%d" % j for j in range(lanes)]
    s = seed * 1000003
    A, c, rs = ref.request_id(key, attrs, hidden, ads, s, nthreads)
    rnd = ref.lane_draws(s, lanes, n_hidden + 2)
    req_attrs = [[b"" if hidden[i] else lane[i] for i in range(n_attrs)] for lane in attrs]
    ref.seed(seed + 1)
    u = ref.fr_rand(lanes)
    v, b1, b2, _ = ref.provide_id(key, A, c, rs, req_attrs, ads, u, nthreads)
    assert v.all()
    _, un2 = ref.unblind(key, attrs, hidden, ads, s, b1, b2, nthreads)
    return ProverRequestWorkload(key, attrs, hidden, ads, rnd, A, c, rs, b1, b2, un2)


@dataclass
class ProverSignonWorkload:
    key: "ref.KeyMaterial"
    sig1: np.ndarray
    sig2: np.ndarray
    attrs: List[List[bytes]]
    hidden: np.ndarray
    ads: List[bytes]
    service: bytes
    service_pt: np.ndarray
    y: np.ndarray
    g: np.ndarray
    h: np.ndarray
    with_id: bool
    rnd: np.ndarray            # (N, h+5 | h+3, 4) in the reference's draw order
    exp: dict                  # the reference's IdProof fields (raw Jacobian points)


def make_prover_signon_workload(n_attrs=5, lanes=8, n_hidden=2, seed=8, with_id=True, key_seed=1, nthreads=0):
    nthreads = nthreads or ref.hw_threads()
    key = ref.KeyMaterial(n_attrs, seed_=key_seed)
    attrs = attr_strings(n_attrs, lanes)
    hidden = _hidden_mask(n_attrs, n_hidden)
    sig1, sig2 = sign_lanes(key, attrs, seed + 7, nthreads)
    ads = [b"sess%d" % j for j in range(lanes)]
    service = b"rp.example"
    y, g, h = ref.hash_to_g1(b"ghi"), ref.hash_to_g1(b"abc"), ref.hash_to_g1(b"jkl")
    s = seed * 1000003
    exp = ref.prove_id(key, sig1, sig2, attrs, hidden, ads, service, y, g, h, s, with_id, nthreads)
    rnd = ref.lane_draws(s, lanes, n_hidden + (5 if with_id else 3))
    return ProverSignonWorkload(key, sig1, sig2, attrs, hidden, ads, service, ref.hash_to_g1(service), y, g, h, with_id, rnd, exp)


def assert_proof_equal(got: dict, exp: dict, with_id: bool):
    """our (normalised) proof against the reference's raw-Jacobian IdProof: points after mcl's normalize, scalars raw."""
    for name in ("sig1", "sig2", "phi") + (("E1", "E2") if with_id else ()):
        assert np.array_equal(got[name], ref.g1_op(ref.G_NORM, exp[name])), name
    assert np.array_equal(got["k"], ref.g2_op(ref.G_NORM, exp["k"])), "k"
    assert np.array_equal(got["c"], exp["c"]), "c"
    assert np.array_equal(np.asarray(got["rs"]).reshape(exp["rs"].shape), exp["rs"]), "rs"


# ---- wire-format ingest ------------------------------------------------------------------
def deserialize_cases(g2: bool, count=24):
    """serialized points (valid, infinity, flag flipped, x >= p, non-residue x) + mcl's verdict and result."""
    from tests.conftest import FIELD_P
    FB = 8 * ref.FP                              # bytes of a serialized Fp (48 / 32)
    rng = np.random.default_rng(11)
    base = ref.hash_to_g2(b"edf") if g2 else ref.hash_to_g1(b"abc")
    if g2:
        mul, ser, deser, SZ = ref.g2_mul, ref.g2_serialize, ref.g2_deserialize, 2 * FB
    else:
        mul, ser, deser, SZ = ref.g1_mul, ref.g1_serialize, ref.g1_deserialize, FB
    ref.seed(17)
    pts = mul(base, ref.fr_rand(count))
    enc = ser(pts).copy()
    enc[1] = 0                                   # infinity
    enc[2, SZ - 1] ^= 0x80                       # other root
    enc[3, :] = 0xff                             # x >= p (and flag set)
    pbytes = np.frombuffer(FIELD_P.to_bytes(FB, "little"), dtype=np.uint8)
    enc[4, SZ - FB:] = pbytes                    # x (or x.b) == p exactly
    for j in range(5, count):                    # random x: about half are non-residues
        if j % 2:
            enc[j, :SZ - 1] = np.frombuffer(rng.bytes(SZ - 1), dtype=np.uint8)
            enc[j, SZ - 1] &= 0x99               # keep x below p most of the time
    if g2:
        enc[5, :] = 0; enc[5, 0] = 3             # x = 3 (x.b = 0): exercises the x.b == 0 branch of Fp2::squareRoot
        enc[7, :] = 0; enc[7, 0] = 2; enc[7, SZ - 1] = 0x80
    want, okv = [], []
    for j in range(count):
        out, ok = deser(enc[j:j + 1])
        want.append(out[0]); okv.append(int(ok))
    return enc, np.array(want), np.array(okv, dtype=np.uint8)
