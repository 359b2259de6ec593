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
