"""GPU parity of the credential-side batch entry points against the reference: randomize_credential
(psb_randomize), batched G1::mul (psb_g1_mul), and the committed golden fixtures."""
import json
import os

import numpy as np
import pytest

from tests import workload
from tests.conftest import G1_SER, G1W, bls_only

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_randomize_matches_reference(gpu_pkg, ref):
    wl = workload.make_verify_workload(n_attrs=2, lanes=70, seed=5)
    ref.seed(8)
    t = ref.fr_rand(70)
    wl.sig1[3] = 0  # infinity stays infinity
    r1, r2, rser = ref.randomize(wl.sig1, wl.sig2, t)
    o1, o2, ser = gpu_pkg.PSRequester.randomize_credential(wl.sig1, wl.sig2, t, want_serialized=True)
    assert np.array_equal(ser, rser)                                   # serialized credentials, byte for byte
    assert np.array_equal(o1[4:], ref.g1_op(ref.G_NORM, r1)[4:])       # normalized raw limbs
    assert np.array_equal(o2, ref.g1_op(ref.G_NORM, r2))
    assert not o1[3].any()
    # randomized credentials still verify (reference check of GPU output)
    assert ref.ps_verify(wl.key, o1[4:], o2[4:], wl.attrs[4:]).all()


def test_pinned_buffers_and_out_arguments(gpu_pkg, ref):
    """psb_host_alloc / psb_host_free through the mirror (pinned_empty / pinned_copy): page-locked inputs and caller-owned,
    reused `out=` buffers give the same bytes as fresh pageable arrays; views keep the block alive; wrong shapes are refused."""
    wl = workload.make_verify_workload(n_attrs=2, lanes=70, seed=5)
    ref.seed(8)
    t = ref.fr_rand(70)
    exp = gpu_pkg.PSRequester.randomize_credential(wl.sig1, wl.sig2, t, want_serialized=True)
    s1, s2, tp = (gpu_pkg.pinned_copy(a) for a in (wl.sig1, wl.sig2, t))
    out = (gpu_pkg.pinned_empty((70, G1W), np.uint64), gpu_pkg.pinned_empty((70, G1W), np.uint64), gpu_pkg.pinned_empty((70, 2 * G1_SER), np.uint8))
    for _ in range(2):                                    # reused across calls
        got = gpu_pkg.PSRequester.randomize_credential(s1, s2, tp, want_serialized=True, out=out)
        assert all(g is o for g, o in zip(got, out))
        assert all(np.array_equal(g, e) for g, e in zip(got, exp))
    row = out[2][5]                                       # a view outlives the arrays it came from
    del out, got
    import gc
    gc.collect()
    assert np.array_equal(row, exp[2][5])
    with pytest.raises(ValueError):
        gpu_pkg.PSRequester.randomize_credential(s1, s2, tp, want_serialized=True, out=(np.zeros((3, G1W), np.uint64), None, None))
    v = gpu_pkg.pinned_empty(70, np.uint8)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
    assert gpu_pkg.PSVerifier(pk).verify(s1, s2, wl.attrs, out=v) is v and v.all()
    pk.close()


def test_randomize_seeded_reference_method(gpu_pkg, ref):
    """the real PSRequester::randomize_credential under the seeded RandGen draws t; same t on the GPU."""
    wl = workload.make_verify_workload(n_attrs=1, lanes=2, seed=9)
    r1, r2, t = ref.randomize_seeded(wl.key, 4242, wl.sig1[0], wl.sig2[0])
    o1, o2, ser = gpu_pkg.PSRequester.randomize_credential(wl.sig1[:1], wl.sig2[:1], t.reshape(1, 4), want_serialized=True)
    assert ser[0, :G1_SER].tobytes() == ref.g1_serialize(r1)[0].tobytes()
    assert ser[0, G1_SER:].tobytes() == ref.g1_serialize(r2)[0].tobytes()


def test_g1_mul_batch(gpu_pkg, ref):
    ref.seed(10)
    k = ref.fr_rand(40)
    g = ref.hash_to_g1(b"abc")
    assert np.array_equal(gpu_pkg.g1_mul(g, k), ref.g1_op(ref.G_NORM, ref.g1_mul(g, k)))
    P = ref.g1_mul(g, ref.fr_rand(40))
    assert np.array_equal(gpu_pkg.g1_mul(P, k), ref.g1_op(ref.G_NORM, ref.g1_mul(P, k)))


@bls_only
def test_golden_fixtures_on_gpu(gpu_pkg):
    """committed reference outputs (tests/golden/protocol.json): verify verdict + GT, randomize bytes."""
    keys = json.load(open(os.path.join(G, "keys.json")))["keys"]["5"]
    p = json.load(open(os.path.join(G, "protocol.json")))
    arr = lambda h, w: np.frombuffer(bytes.fromhex(h), dtype=np.uint64).reshape(-1, w).copy()  # noqa: E731
    pk = gpu_pkg.PSPubKey(arr(keys["g"], 18), arr(keys["gg"], 36), arr(keys["XX"], 36), arr(keys["Y"], 18),
                          arr(keys["YY"], 36), window_bits=8)
    v = p["verify"]
    s1, s2 = arr(v["sig1"], 18), arr(v["sig2"], 18)
    attrs = [[a.encode() for a in lane] for lane in v["attrs"]]
    got_v, got_gt = gpu_pkg.PSVerifier(pk).verify(s1, s2, attrs, want_gt=True)
    assert got_v.tolist() == v["verdict"]
    live = s1[:, 12:].any(axis=1)
    assert np.array_equal(got_gt[live], arr(v["gt"], 72)[live])
    _, _, ser = gpu_pkg.PSRequester.randomize_credential(s1, s2, arr(p["randomize"]["t"], 4), want_serialized=True)
    assert ser.tobytes().hex() == p["randomize"]["ser"]
