"""engine.py -- ctypes binding of libpsb.so (include/psb.h) and a Python mirror of the reference's
role classes for the batched path.

Class and method names follow the reference (src/ps-verifier.h, src/ps-requester.h, src/ps-signer.h):
PSVerifier.verify / el_passo_verify_id, PSRequester.verify / randomize_credential,
PSSigner.el_passo_provide_id -- each taking a BATCH (numpy arrays in mcl's in-memory layout, see
include/psb.h) instead of one object.  The real drop-in for C++ callers is host/ps_batch.hpp; this
mirror exists so tests and bench.py read like the reference's own tests.

There is no CPU fallback: loading fails loudly if libpsb.so is missing, and every call fails if no
GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# One curve per library / per process, like mcl's own bn256 / bn384 builds: PSB_CURVE=bn254 selects
# libpsb_bn254.so (what the reference's shipped tests get from initPairing()), anything else
# libpsb.so (BLS12-381, the north star's curve).
CURVE = os.environ.get("PSB_CURVE", "bls12_381").lower()
assert CURVE in ("bls12_381", "bn254"), CURVE
BN254 = CURVE == "bn254"
LIB_PATH = os.environ.get("PSB_LIB", os.path.join(_HERE, "libpsb_bn254.so" if BN254 else "libpsb.so"))

FP = 4 if BN254 else 6                              # u64 words of an Fp
FR, G1, G2, GT = 4, 3 * FP, 6 * FP, 12 * FP         # u64 words
G1_SER, G2_SER = 8 * FP, 16 * FP                    # compressed point bytes (mcl serialize)
CRED_SER = 2 * G1_SER                               # bare serialized credential (sigma1 || sigma2)
CURVE_BLS12_381, CURVE_BN254 = 5, 0                 # mcl/include/mcl/curve_type.h
VID_WITH_ID, VID_REJECT_ZERO_SIGMA = 1, 2           # include/psb.h PSB_VID_*
MCL_CURVE = CURVE_BN254 if BN254 else CURVE_BLS12_381

_lib = None
_inited = False


class PsbError(RuntimeError):
    pass


def lib():
    """dlopen libpsb.so (never builds, never falls back)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PsbError(f"{LIB_PATH} missing: run `python __graft_entry__.py build` (nvcc) first; "
                           "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.psb_last_error.restype = C.c_char_p
        L.psb_launch_count.restype = C.c_uint64
        L.psb_key_create.restype = C.c_void_p
        L.psb_key_num_attributes.restype = C.c_size_t
        L.psb_key_table_bytes.restype = C.c_size_t
        L.psb_verify_ws_bytes.restype = C.c_size_t
        L.psb_microbench.restype = C.c_double
        L.psb_host_alloc.restype = C.c_void_p
        L.psb_host_alloc.argtypes = [C.c_size_t]
        L.psb_host_free.argtypes = [C.c_void_p]
        L.psb_key_destroy.argtypes = [C.c_void_p]
        L.psb_key_num_attributes.argtypes = [C.c_void_p]
        L.psb_key_table_bytes.argtypes = [C.c_void_p]
        _lib = L
    return _lib


EXPORTS = ["psb_init", "psb_shutdown", "psb_num_devices", "psb_last_error", "psb_shard_range", "psb_launch_count",
           "psb_key_create", "psb_key_destroy", "psb_key_num_attributes", "psb_key_table_bytes",
           "psb_verify", "psb_verify_aos", "psb_verify_ser", "psb_g1_deserialize", "psb_g2_deserialize", "psb_verify_ws_bytes", "psb_verify_dev", "psb_randomize", "psb_provide_id",
           "psb_verify_id", "psb_sign", "psb_verify_id_ser", "psb_provide_id_ser", "psb_wire_encode", "psb_request_id", "psb_unblind", "psb_prove_id", "psb_hash_to_g1", "psb_pairing", "psb_g1_mul", "psb_host_alloc", "psb_host_free", "psb_set_profiling", "psb_last_phase_ms", "psb_test_op_shape", "psb_test_op", "psb_microbench"]


def _check(rc: int, what: str):
    if rc != 0:
        raise PsbError(f"{what} failed ({rc}): {lib().psb_last_error().decode()}")


def init(devices: Optional[Sequence[int]] = None) -> None:
    """psb_init: replaces initPairing(mcl::BLS12_381) for the batch path."""
    global _inited
    L = lib()
    if devices is None:
        rc = L.psb_init(MCL_CURVE, None, 0)
    else:
        arr = (C.c_int * len(devices))(*devices)
        rc = L.psb_init(MCL_CURVE, arr, len(devices))
    _check(rc, "psb_init")
    _inited = True


def shutdown() -> None:
    """psb_shutdown: frees every device context; live keys become dead handles (every later use is PSB_ERR_ARG)."""
    global _inited
    lib().psb_shutdown()
    _inited = False


def ensure_init():
    if not _inited:
        init()


def shard_range(N: int, ndev: int, k: int):
    """lane range [b, e) of device k (pure host arithmetic, no GPU needed)."""
    b, e = C.c_size_t(), C.c_size_t()
    _check(lib().psb_shard_range(C.c_size_t(N), C.c_int(ndev), C.c_int(k), C.byref(b), C.byref(e)), "psb_shard_range")
    return int(b.value), int(e.value)


def launch_count() -> int:
    return int(lib().psb_launch_count())


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _u64(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, width)


def pack_strings(strs: Sequence[bytes]):
    """list of bytes -> (blob uint8, offsets uint64[len+1])"""
    off = np.zeros(len(strs) + 1, dtype=np.uint64)
    if len(strs):
        off[1:] = np.cumsum([len(s) for s in strs], dtype=np.uint64)
    blob = np.frombuffer(b"".join(strs) + b"\0" * 8, dtype=np.uint8).copy()
    return blob, off


def pack_attrs(attrs: Sequence[Sequence[bytes]]):
    return pack_strings([a for lane in attrs for a in lane])


def _out(out, i, shape, dtype):
    """the i-th caller-supplied output buffer (e.g. page-locked, reused across calls) or a fresh array"""
    if out is None or out[i] is None:
        return np.zeros(shape, dtype=dtype)
    a = out[i]
    if a.shape != tuple(shape) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise ValueError(f"out[{i}] must be a C-contiguous {np.dtype(dtype).name}{tuple(shape)} array")
    return a


def _packed(x, flat: bool):
    """accept pre-packed (blob, off) pairs wherever string lists are taken (keeps python out of timed regions)."""
    if isinstance(x, tuple):
        return x
    return pack_strings(list(x)) if flat else pack_attrs(x)


class PSPubKey:
    """PSPubKey (src/ps-encoding.h:111-133) + optional signer secret X = g^x, resident on the GPUs
    together with its fixed-base window tables."""

    def __init__(self, g, gg, XX, Yi, YYi, X_secret=None, window_bits: int = 0):
        ensure_init()
        self.g = _u64(g, G1).copy()
        self.gg = _u64(gg, G2).copy()
        self.XX = _u64(XX, G2).copy()
        self.Yi = _u64(Yi, G1).copy()
        self.YYi = _u64(YYi, G2).copy()
        self.n = self.Yi.shape[0]
        if self.YYi.shape[0] != self.n:
            raise ValueError("attribute size does not match")
        self.X = None if X_secret is None else _u64(X_secret, G1).copy()
        h = lib().psb_key_create(_p(self.g), _p(self.gg), _p(self.XX), _p(self.Yi), _p(self.YYi),
                                 C.c_size_t(self.n), _p(self.X), C.c_int(window_bits))
        if not h:
            raise PsbError("psb_key_create failed: " + lib().psb_last_error().decode())
        self.handle = C.c_void_p(h)

    @property
    def table_bytes(self) -> int:
        return int(lib().psb_key_table_bytes(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            lib().psb_key_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _attr_args(pk: PSPubKey, N: int, attributes, scalars):
    if attributes is not None:
        if isinstance(attributes, tuple):  # pre-packed (blob, off)
            blob, off = attributes
        else:
            for lane in attributes:
                if len(lane) != pk.n:
                    raise ValueError("attribute size does not match")
            blob, off = pack_attrs(attributes)
        if off.shape[0] != N * pk.n + 1:
            raise ValueError("attribute size does not match")
        return blob, off, None
    m = _u64(scalars, FR)
    if m.shape[0] != N * pk.n:
        raise ValueError("attribute size does not match")
    return None, None, m


class PSVerifier:
    """batched PSVerifier (src/ps-verifier.h:11-71)."""

    def __init__(self, pk: PSPubKey):
        self.m_pk = pk

    def verify(self, sig1, sig2, all_attributes=None, scalars=None, want_gt: bool = False, out=None):
        """batched PSVerifier::verify (src/ps-verifier.cc:13-35).  sig1/sig2: (N,18) u64;
        all_attributes: N lists of n byte strings (or a packed (blob, off) pair), or scalars (N*n,4).
        Returns verdict uint8[N] (and GT (N,72) if want_gt)."""
        s1 = _u64(sig1, G1)
        s2 = _u64(sig2, G1)
        N = s1.shape[0]
        if s2.shape[0] != N:
            raise ValueError("sig1/sig2 length mismatch")
        blob, off, m = _attr_args(self.m_pk, N, all_attributes, scalars)
        verdict = np.zeros(N, dtype=np.uint8) if out is None else out
        if verdict.shape != (N,) or verdict.dtype != np.uint8:
            raise ValueError("out must be uint8[N]")
        gt = np.zeros((N, GT), dtype=np.uint64) if want_gt else None
        _check(lib().psb_verify(self.m_pk.handle, C.c_size_t(N), _p(s1), _p(s2), _p(blob), _p(off), _p(m),
                                _p(verdict), _p(gt)), "psb_verify")
        return (verdict, gt) if want_gt else verdict

    def verify_serialized(self, cred, all_attributes, stride: int = 2 * (2 + G1_SER), off1: int = 2, off2: int = 4 + G1_SER):
        """batched verify of SERIALIZED credentials (psb_verify_ser): cred = uint8 (N, stride); defaults are the
        layout of PSCredential::toBufferString (src/ps-encoding.cc:384-391).  Returns (verdict, decoded)."""
        cred = np.ascontiguousarray(cred, dtype=np.uint8).reshape(-1, stride)
        N = cred.shape[0]
        blob, off = _packed(all_attributes, False)
        if off.shape[0] != N * self.m_pk.n + 1:
            raise ValueError("attribute size does not match")
        verdict = np.zeros(N, dtype=np.uint8)
        decoded = np.zeros(N, dtype=np.uint8)
        _check(lib().psb_verify_ser(self.m_pk.handle, C.c_size_t(N), _p(cred), C.c_size_t(stride), C.c_size_t(off1),
                                    C.c_size_t(off2), _p(blob), _p(off), _p(verdict), _p(decoded)), "psb_verify_ser")
        return verdict, decoded

    def el_passo_verify_id_wire(self, buffers, associated_data: Sequence[bytes], service_pt, authority_pk=None, g=None,
                                h=None, with_id: bool = True, base64: bool = False, strict: bool = False, out=None):
        """el_passo_verify_id straight from the WIRE (psb_verify_id_ser): buffers = per lane the bytes of
        IdProof::toBufferString() (or their base64 text), or a packed (blob, off) pair.  Returns (verdict, parsed)."""
        blob, off = _packed(buffers, True)
        ad_blob, ad_off = _packed(associated_data, True)
        N = off.shape[0] - 1
        if ad_off.shape[0] != N + 1:
            raise ValueError("one associated_data per proof")
        verdict = _out(out, 0, (N,), np.uint8)    # out = (verdict, parsed)
        parsed = _out(out, 1, (N,), np.uint8)
        flags = (VID_WITH_ID if with_id else 0) | (VID_REJECT_ZERO_SIGMA if strict else 0)
        _check(lib().psb_verify_id_ser(
            self.m_pk.handle, C.c_size_t(N), _p(blob), _p(off), C.c_int(int(base64)), _p(ad_blob), _p(ad_off),
            _p(_u64(service_pt, G1)), _p(_u64(authority_pk, G1)) if with_id else None, _p(_u64(g, G1)) if with_id else None,
            _p(_u64(h, G1)) if with_id else None, C.c_int(flags), _p(verdict), _p(parsed)), "psb_verify_id_ser")
        return verdict, parsed

    def el_passo_verify_id(self, proof: dict, attributes, associated_data: Sequence[bytes], service_pt,
                           authority_pk=None, g=None, h=None, with_id: bool = True, strict: bool = False, out=None):
        """batched el_passo_verify_id (src/ps-verifier.cc:37-138) / _without_id_retrieval (:140-212).
        proof: dict of arrays sig1, sig2, k, phi, E1, E2, c, rs (N, per, 4); attributes: per lane the
        proof's attribute list (b"" = hidden)."""
        s1 = _u64(proof["sig1"], G1)
        N = s1.shape[0]
        rs = np.ascontiguousarray(proof["rs"], dtype=np.uint64).reshape(N, -1, FR)
        blob, off = _packed(attributes, False)
        ad_blob, ad_off = _packed(associated_data, True)
        if off.shape[0] != N * self.m_pk.n + 1 or ad_off.shape[0] != N + 1:
            raise ValueError("attribute size does not match")
        verdict = np.zeros(N, dtype=np.uint8) if out is None else _out((out,), 0, (N,), np.uint8)
        z1 = np.zeros((1, G1), dtype=np.uint64)
        _check(lib().psb_verify_id(
            self.m_pk.handle, C.c_size_t(N), _p(s1), _p(_u64(proof["sig2"], G1)), _p(_u64(proof["k"], G2)),
            _p(_u64(proof["phi"], G1)), _p(_u64(proof["E1"], G1)) if with_id else None,
            _p(_u64(proof["E2"], G1)) if with_id else None, _p(_u64(proof["c"], FR)), _p(rs),
            C.c_size_t(rs.shape[1]), _p(blob), _p(off), _p(ad_blob), _p(ad_off), _p(_u64(service_pt, G1)),
            _p(_u64(authority_pk, G1)) if with_id else _p(z1), _p(_u64(g, G1)) if with_id else _p(z1),
            _p(_u64(h, G1)) if with_id else _p(z1),
            C.c_int((VID_WITH_ID if with_id else 0) | (VID_REJECT_ZERO_SIGMA if strict else 0)), _p(verdict)), "psb_verify_id")
        return verdict


class PSRequester:
    """batched PSRequester (src/ps-requester.h:11-128): verify, randomize_credential and the prover side
    (el_passo_request_id, unblind_credential, el_passo_prove_id) with host-supplied scalars."""

    def __init__(self, pk: PSPubKey):
        self.m_pk = pk

    def verify(self, sig1, sig2, all_attributes=None, scalars=None, want_gt: bool = False):
        return PSVerifier(self.m_pk).verify(sig1, sig2, all_attributes, scalars, want_gt)

    @staticmethod
    def randomize_credential(sig1, sig2, t, want_serialized: bool = False, out=None):
        """batched randomize_credential (src/ps-requester.cc:139-148) with host-supplied t (N,4).
        out = (sig1', sig2', ser): optional caller-owned output buffers (e.g. page-locked, reused across calls)."""
        ensure_init()
        s1 = _u64(sig1, G1)
        s2 = _u64(sig2, G1)
        tt = _u64(t, FR)
        N = s1.shape[0]
        o1 = _out(out, 0, (N, G1), np.uint64)
        o2 = _out(out, 1, (N, G1), np.uint64)
        ser = _out(out, 2, (N, CRED_SER), np.uint8) if want_serialized else None
        _check(lib().psb_randomize(C.c_size_t(N), _p(s1), _p(s2), _p(tt), _p(o1), _p(o2), _p(ser)),
               "psb_randomize")
        return (o1, o2, ser) if want_serialized else (o1, o2)

    def el_passo_request_id(self, attributes, hidden, associated_data, rnd, out=None):
        """batched el_passo_request_id (src/ps-requester.cc:19-97).  attributes: per lane ALL n values; hidden: n flags
        shared by the batch; rnd (N, h+2, 4) = t1, r0, one per hidden attribute (the reference's draw order).
        Returns A (N,18) normalised, c (N,4), rs (N,h+1,4)."""
        pk = self.m_pk
        hidden = np.ascontiguousarray(hidden, dtype=np.uint8)
        if hidden.shape != (pk.n,):
            raise ValueError("attribute size does not match")
        h = int((hidden != 0).sum())
        blob, off = _packed(attributes, False)
        ad_blob, ad_off = _packed(associated_data, True)
        N = ad_off.shape[0] - 1
        if off.shape[0] != N * pk.n + 1:
            raise ValueError("attribute size does not match")
        rnd = np.ascontiguousarray(rnd, dtype=np.uint64).reshape(N, h + 2, FR)
        A = _out(out, 0, (N, G1), np.uint64)       # out = (A, c, rs)
        c = _out(out, 1, (N, FR), np.uint64)
        rs = _out(out, 2, (N, h + 1, FR), np.uint64)
        _check(lib().psb_request_id(pk.handle, C.c_size_t(N), _p(blob), _p(off), _p(hidden), _p(ad_blob), _p(ad_off),
                                    _p(rnd), _p(A), _p(c), _p(rs)), "psb_request_id")
        return A, c, rs

    @staticmethod
    def unblind_credential(sig1, sig2, t1, out=None):
        """batched unblind_credential (src/ps-requester.cc:99-113): (sig1, sig2 - t1 sig1); sig2 normalised.
        out: optional caller-owned (N, G1) buffer for sig2'."""
        ensure_init()
        s1 = _u64(sig1, G1)
        s2 = _u64(sig2, G1)
        N = s1.shape[0]
        o2 = _out((out,), 0, (N, G1), np.uint64)
        _check(lib().psb_unblind(C.c_size_t(N), _p(s1), _p(s2), _p(_u64(t1, FR)), _p(o2)), "psb_unblind")
        return s1, o2

    def el_passo_prove_id(self, sig1, sig2, attributes, hidden, associated_data, service_pt, authority_pk=None, g=None,
                          h=None, rnd=None, with_id: bool = True):
        """batched el_passo_prove_id (src/ps-requester.cc:150-310) / _without_id_retrieval (:312-432).
        rnd (N, h+5 | h+3, 4) in the reference's draw order (include/psb.h).  Returns the IdProof fields as a dict."""
        pk = self.m_pk
        hidden = np.ascontiguousarray(hidden, dtype=np.uint8)
        if hidden.shape != (pk.n,):
            raise ValueError("attribute size does not match")
        hh = int((hidden != 0).sum())
        s1 = _u64(sig1, G1)
        N = s1.shape[0]
        blob, off = _packed(attributes, False)
        ad_blob, ad_off = _packed(associated_data, True)
        if off.shape[0] != N * pk.n + 1 or ad_off.shape[0] != N + 1:
            raise ValueError("attribute size does not match")
        rnd = np.ascontiguousarray(rnd, dtype=np.uint64).reshape(N, hh + (5 if with_id else 3), FR)
        per = hh + (2 if with_id else 1)
        o = dict(sig1=np.zeros((N, G1), np.uint64), sig2=np.zeros((N, G1), np.uint64), k=np.zeros((N, G2), np.uint64),
                 phi=np.zeros((N, G1), np.uint64), E1=np.zeros((N, G1), np.uint64), E2=np.zeros((N, G1), np.uint64),
                 c=np.zeros((N, FR), np.uint64), rs=np.zeros((N, per, FR), np.uint64))
        _check(lib().psb_prove_id(
            pk.handle, C.c_size_t(N), _p(s1), _p(_u64(sig2, G1)), _p(blob), _p(off), _p(hidden), _p(ad_blob), _p(ad_off),
            _p(_u64(service_pt, G1)), _p(_u64(authority_pk, G1)) if with_id else None, _p(_u64(g, G1)) if with_id else None,
            _p(_u64(h, G1)) if with_id else None, C.c_int(int(with_id)), _p(rnd), _p(o["sig1"]), _p(o["sig2"]), _p(o["k"]),
            _p(o["phi"]), _p(o["E1"]), _p(o["E2"]), _p(o["c"]), _p(o["rs"])), "psb_prove_id")
        return o


class PSSigner:
    """batched PSSigner (src/ps-signer.h:11-96): el_passo_provide_id with host-supplied u."""

    def __init__(self, pk: PSPubKey):
        if pk.X is None:
            raise ValueError("signer key needs X_secret")
        self.m_pk = pk

    def el_passo_provide_id(self, A, c, rs, attributes, associated_data, u, out=None):
        """out = (verdict, sig1, sig2, ser): optional caller-owned output buffers (e.g. page-locked, reused across calls)"""
        A = _u64(A, G1)
        N = A.shape[0]
        rs = np.ascontiguousarray(rs, dtype=np.uint64).reshape(N, -1, FR)
        blob, off = _packed(attributes, False)
        ad_blob, ad_off = _packed(associated_data, True)
        if off.shape[0] != N * self.m_pk.n + 1 or ad_off.shape[0] != N + 1:
            raise ValueError("attribute size does not match")
        verdict = _out(out, 0, (N,), np.uint8)
        s1 = _out(out, 1, (N, G1), np.uint64)
        s2 = _out(out, 2, (N, G1), np.uint64)
        ser = _out(out, 3, (N, CRED_SER), np.uint8)
        _check(lib().psb_provide_id(self.m_pk.handle, C.c_size_t(N), _p(A), _p(_u64(c, FR)), _p(rs),
                                    C.c_size_t(rs.shape[1]), _p(blob), _p(off), _p(ad_blob), _p(ad_off),
                                    _p(_u64(u, FR)), _p(verdict), _p(s1), _p(s2), _p(ser)), "psb_provide_id")
        return verdict, s1, s2, ser

    def el_passo_provide_id_wire(self, buffers, associated_data, u, base64: bool = False):
        """el_passo_provide_id straight from the WIRE (psb_provide_id_ser): buffers = per lane the bytes of
        PSCredRequest::toBufferString() (or their base64 text).  Returns (verdict, sig1, sig2, ser, parsed)."""
        blob, off = _packed(buffers, True)
        ad_blob, ad_off = _packed(associated_data, True)
        N = off.shape[0] - 1
        if ad_off.shape[0] != N + 1:
            raise ValueError("one associated_data per request")
        verdict = np.zeros(N, dtype=np.uint8)
        parsed = np.zeros(N, dtype=np.uint8)
        s1 = np.zeros((N, G1), dtype=np.uint64)
        s2 = np.zeros((N, G1), dtype=np.uint64)
        ser = np.zeros((N, CRED_SER), dtype=np.uint8)
        _check(lib().psb_provide_id_ser(self.m_pk.handle, C.c_size_t(N), _p(blob), _p(off), C.c_int(int(base64)), _p(ad_blob),
                                        _p(ad_off), _p(_u64(u, FR)), _p(verdict), _p(s1), _p(s2), _p(ser), _p(parsed)),
               "psb_provide_id_ser")
        return verdict, s1, s2, ser, parsed

    def sign_commitment(self, commitment, u):
        """batched PSSigner::sign_commitment (src/ps-signer.cc:132-146) with host-supplied u: (sig1, sig2, ser)."""
        return self.sign_hybrid(commitment, None, u)

    def sign_hybrid(self, commitment, attributes, u):
        """batched PSSigner::sign_hybrid (src/ps-signer.cc:112-130): attributes = per lane the same number of strings
        (b"" = committed attribute), or None for sign_commitment."""
        Cm = _u64(commitment, G1)
        N = Cm.shape[0]
        if attributes is None:
            blob, off, na = None, None, 0
        else:
            blob, off = _packed(attributes, False)
            na = (off.shape[0] - 1) // N if N else 0
            if off.shape[0] != N * na + 1:
                raise ValueError("attribute size does not match")
        s1 = np.zeros((N, G1), dtype=np.uint64)
        s2 = np.zeros((N, G1), dtype=np.uint64)
        ser = np.zeros((N, CRED_SER), dtype=np.uint8)
        _check(lib().psb_sign(self.m_pk.handle, C.c_size_t(N), _p(Cm), C.c_size_t(na), _p(blob), _p(off), _p(_u64(u, FR)),
                              _p(s1), _p(s2), _p(ser)), "psb_sign")
        return s1, s2, ser


WIRE_IDPROOF, WIRE_REQUEST = 0, 1                    # include/psb.h PSB_WIRE_*


def _wire_encode(kind, n_attrs, p0, sig2, k, phi, E1, E2, c, rs, attributes, base64):
    ensure_init()
    p0 = _u64(p0, G1)
    N = p0.shape[0]
    rs = np.ascontiguousarray(rs, dtype=np.uint64).reshape(N, -1, FR)
    blob, off = _packed(attributes, False)
    if off.shape[0] != N * n_attrs + 1:
        raise ValueError("attribute size does not match")
    opt = lambda a, w: None if a is None else _p(_u64(a, w))  # noqa: E731
    out_off = np.zeros(N + 1, dtype=np.uint64)
    args = [C.c_int(kind), C.c_size_t(N), C.c_size_t(n_attrs), _p(p0), opt(sig2, G1), opt(k, G2), opt(phi, G1), opt(E1, G1),
            opt(E2, G1), _p(_u64(c, FR)), _p(rs), C.c_size_t(rs.shape[1]), _p(blob), _p(off), C.c_int(int(base64))]
    _check(lib().psb_wire_encode(*args, None, C.c_size_t(0), _p(out_off)), "psb_wire_encode")
    out = np.zeros(int(out_off[-1]) + 8, dtype=np.uint8)
    _check(lib().psb_wire_encode(*args, _p(out), C.c_size_t(int(out_off[-1])), _p(out_off)), "psb_wire_encode")
    return out, out_off


def idproof_serialize(proof: dict, attributes, n_attrs: int, with_id: bool = True, base64: bool = False):
    """batched IdProof::toBufferString() (or its base64 text) on the device: packed (blob, off) of N messages."""
    return _wire_encode(WIRE_IDPROOF, n_attrs, proof["sig1"], proof["sig2"], proof["k"], proof["phi"],
                        proof["E1"] if with_id else None, proof["E2"] if with_id else None, proof["c"], proof["rs"], attributes, base64)


def request_serialize(A, c, rs, attributes, n_attrs: int, base64: bool = False):
    """batched PSCredRequest::toBufferString() (or its base64 text) on the device."""
    return _wire_encode(WIRE_REQUEST, n_attrs, A, None, None, None, None, None, c, rs, attributes, base64)


def pairing(P, Q):
    """batched mcl::bn::pairing (bn.hpp:1711-1715)."""
    ensure_init()
    P = _u64(P, G1)
    Q = _u64(Q, G2)
    out = np.zeros((P.shape[0], GT), dtype=np.uint64)
    _check(lib().psb_pairing(C.c_size_t(P.shape[0]), _p(P), _p(Q), _p(out)), "psb_pairing")
    return out


def hash_and_map_to_g1(msgs):
    """batched mcl::bn::hashAndMapToG1: list of byte strings (or packed (blob, off)) -> (points (N,18) normalised, ok)."""
    ensure_init()
    blob, off = _packed(msgs, True)
    N = off.shape[0] - 1
    out = np.zeros((N, G1), dtype=np.uint64)
    ok = np.zeros(N, dtype=np.uint8)
    _check(lib().psb_hash_to_g1(C.c_size_t(N), _p(blob), _p(off), _p(out), _p(ok)), "psb_hash_to_g1")
    return out, ok


def g1_deserialize(ser, stride: int = G1_SER):
    """batched G1::deserialize (point decompression): ser uint8 (N, stride) -> (points (N,18), ok uint8[N])."""
    ensure_init()
    ser = np.ascontiguousarray(ser, dtype=np.uint8).reshape(-1, stride)
    N = ser.shape[0]
    out = np.zeros((N, G1), dtype=np.uint64)
    ok = np.zeros(N, dtype=np.uint8)
    _check(lib().psb_g1_deserialize(C.c_size_t(N), _p(ser), C.c_size_t(stride), _p(out), _p(ok)), "psb_g1_deserialize")
    return out, ok


def g2_deserialize(ser, stride: int = G2_SER):
    ensure_init()
    ser = np.ascontiguousarray(ser, dtype=np.uint8).reshape(-1, stride)
    N = ser.shape[0]
    out = np.zeros((N, G2), dtype=np.uint64)
    ok = np.zeros(N, dtype=np.uint8)
    _check(lib().psb_g2_deserialize(C.c_size_t(N), _p(ser), C.c_size_t(stride), _p(out), _p(ok)), "psb_g2_deserialize")
    return out, ok


def g1_mul(P, k):
    """batched G1::mul, normalised output; P is (N,18) or a single point (broadcast)."""
    ensure_init()
    P = _u64(P, G1)
    k = _u64(k, FR)
    N = k.shape[0]
    stride = 0 if (P.shape[0] == 1 and N != 1) else 1
    out = np.zeros((N, G1), dtype=np.uint64)
    _check(lib().psb_g1_mul(C.c_size_t(N), _p(P), C.c_int(stride), _p(k), _p(out)), "psb_g1_mul")
    return out


def test_op(op: int, a, b=None, c=None):
    """element-wise arithmetic probe (csrc/testops.cuh) -- parity tests only."""
    ensure_init()
    s = (C.c_int * 4)()
    _check(lib().psb_test_op_shape(op, s), "psb_test_op_shape")
    a32 = np.ascontiguousarray(a).view(np.uint32).reshape(-1, s[0])
    n = a32.shape[0]
    b32 = None if b is None or s[1] == 0 else np.ascontiguousarray(b).view(np.uint32).reshape(n, s[1])
    c32 = None if c is None or s[2] == 0 else np.ascontiguousarray(c).view(np.uint32).reshape(n, s[2])
    out = np.zeros((n, s[3]), dtype=np.uint32)
    _check(lib().psb_test_op(op, C.c_size_t(n), _p(a32), _p(b32), _p(c32), _p(out)), "psb_test_op")
    return out.view(np.uint64)


def microbench(kind: int, blocks: int, threads: int, iters: int) -> float:
    ensure_init()
    ms = float(lib().psb_microbench(kind, blocks, threads, iters))
    if ms < 0:
        raise PsbError("psb_microbench failed: " + lib().psb_last_error().decode())
    return ms


def verify_ws_bytes(pk: PSPubKey, N: int) -> int:
    return int(lib().psb_verify_ws_bytes(pk.handle, C.c_size_t(N)))


def verify_dev(pk: PSPubKey, dev_index: int, N: int, d_sig1: int, d_sig2: int, d_blob: int, d_off: int,
               d_m: int, d_verdict: int, d_gt: int, d_ws: int, stream: int = 0) -> None:
    """psb_verify_dev with raw device pointers (e.g. torch tensors' data_ptr())."""
    vp = lambda x: C.c_void_p(x) if x else None  # noqa: E731
    _check(lib().psb_verify_dev(pk.handle, C.c_int(dev_index), C.c_size_t(N), vp(d_sig1), vp(d_sig2), vp(d_blob),
                                vp(d_off), vp(d_m), vp(d_verdict), vp(d_gt), vp(d_ws), vp(stream)),
           "psb_verify_dev")


class _Pinned:
    """owner of one psb_host_alloc block, exported through the buffer protocol (PEP 688): the arrays made from it keep
    it alive as their base, and the block is freed with the last of them"""

    def __init__(self, nbytes: int):
        self.ptr = lib().psb_host_alloc(C.c_size_t(nbytes))
        if not self.ptr:
            raise PsbError("psb_host_alloc failed: " + lib().psb_last_error().decode())
        self.buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr)

    def __buffer__(self, flags):
        return memoryview(self.buf)

    def __del__(self):
        try:
            if self.ptr:
                lib().psb_host_free(C.c_void_p(self.ptr))
        except Exception:  # interpreter shutdown
            pass


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """a page-locked host array (psb_host_alloc) for the inputs and `out=` buffers of the batched calls"""
    ensure_init()
    dt = np.dtype(dtype)
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    n = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    return np.frombuffer(_Pinned(n), dtype=np.uint8, count=n).view(dt).reshape(shape)


def pinned_copy(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def set_profiling(on: bool) -> None:
    lib().psb_set_profiling(C.c_int(int(on)))


def last_phase_ms(dev_index: int = 0):
    """(msm, miller, final_exp) device milliseconds of the last profiled verify on that device."""
    ms = (C.c_float * 3)()
    _check(lib().psb_last_phase_ms(C.c_int(dev_index), ms), "psb_last_phase_ms")
    return [float(x) for x in ms]
