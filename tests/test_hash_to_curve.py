"""mcl::bn::hashAndMapToG1 on the batch engine: SHA-512 -> Fp (Fp::setHashOf), the Shallue-van de
Woestijne map of MapTo::calcBN and the cofactor multiplication, against mcl itself -- on the CPU through hostsim, on
the GPU through psb_hash_to_g1.  Message lengths cross the SHA-512 padding boundaries (111/112, 127/128 bytes)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from tests.conftest import FIELD_P, FP_BYTES, G1W

MSGS = [b"", b"abc", b"rp.example", b"service", b"a" * 111, b"b" * 112, b"c" * 127, b"d" * 128, b"e" * 129, b"f" * 239, b"g" * 240,
        bytes(range(256)) * 3] + [b"svc%d.example.org" % i for i in range(40)]


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


def test_sha512_on_hostsim(hostsim):
    for m in MSGS[:12]:
        out = np.zeros(64, dtype=np.uint8)
        hostsim.hostsim_sha512(C.c_char_p(m), C.c_size_t(len(m)), _p(out))
        assert out.tobytes() == hashlib.sha512(m).digest(), len(m)


def test_hash_to_g1_lanes_match_mcl(hostsim, ref):
    for m in MSGS[:20]:
        out = np.zeros(G1W, dtype=np.uint64)
        assert hostsim.hostsim_hash_to_g1(C.c_char_p(m), C.c_size_t(len(m)), _p(out)) == 1
        want = ref.g1_op(ref.G_NORM, ref.hash_to_g1(m).reshape(1, -1))[0]
        assert np.array_equal(out, want), m[:16]


def test_map_to_g1_exceptional_and_signs(hostsim, ref):
    """t = 0 is rejected like mcl; t and -t map to points with opposite y (Legendre sign rule); all three x-candidates occur."""
    rng = np.random.default_rng(5)
    vals = [0, 1, 2, FIELD_P - 1, FIELD_P - 2] + [int.from_bytes(rng.bytes(FP_BYTES), "little") % FIELD_P for _ in range(30)]
    ts = ref.fp_from_ints(vals)
    for j, t in enumerate(ts):
        out = np.zeros(G1W, dtype=np.uint64)
        r = hostsim.hostsim_map_to_g1(_p(t), _p(out))
        want, ok = ref.map_to_g1(t)
        assert r == ok, vals[j]
        if ok:
            assert np.array_equal(out, ref.g1_op(ref.G_NORM, want.reshape(1, -1))[0]), vals[j]
    assert ref.map_to_g1(ts[0])[1] == 0


@pytest.mark.gpu
def test_hash_to_g1_on_gpu(gpu_pkg, ref):
    out, ok = gpu_pkg.hash_and_map_to_g1(MSGS)
    assert ok.all()
    want = ref.g1_op(ref.G_NORM, np.stack([ref.hash_to_g1(m) for m in MSGS]))
    assert np.array_equal(out, want)
    # the device value feeds el_passo_verify_id exactly like the host's hashAndMapToG1(service_name)
    assert np.array_equal(out[2], ref.g1_op(ref.G_NORM, ref.hash_to_g1(b"rp.example").reshape(1, -1))[0])
