"""GPU parity of the EL PASSO batch entry points against the reference's own methods:
psb_provide_id  vs PSSigner::el_passo_provide_id   (src/ps-signer.cc:63-146), host-supplied u
psb_verify_id   vs PSVerifier::el_passo_verify_id  (src/ps-verifier.cc:37-138) and
                   el_passo_verify_id_without_id_retrieval (:140-212)
Requests / proofs come from the reference's own prover code (PSRequester) under seeded RandGens;
tampered lanes exercise every reject path (expected verdicts are the oracle's, not the generator's)."""
import json
import os

import numpy as np
import pytest

from tests import workload
from tests.conftest import bls_only

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pk(gpu_pkg, key, w=8, with_secret=False):
    return gpu_pkg.PSPubKey(key.g, key.gg, key.XX, key.Y, key.YY, X_secret=key.X if with_secret else None, window_bits=w)


@pytest.mark.parametrize("n_attrs,n_hidden,lanes,w", [(5, 2, 96, 8), (20, 2, 40, 16), (1, 1, 8, 6), (1, 0, 8, 8), (3, 3, 9, 5)])
def test_provide_id_matches_reference(gpu_pkg, ref, n_attrs, n_hidden, lanes, w):
    wl = workload.make_issuance_workload(n_attrs, lanes, n_hidden, seed=4, tamper_every=5)
    ev, e1, e2, eser = workload.expected_provide_id(wl)
    pk = _pk(gpu_pkg, wl.key, w, with_secret=True)
    v, s1, s2, ser = gpu_pkg.PSSigner(pk).el_passo_provide_id(wl.A, wl.c, wl.rs, wl.req_attrs, wl.ads, wl.u)
    assert v.tolist() == ev.tolist()
    assert 0 < ev.sum() == lanes - len(wl.tampered)
    ok = ev.astype(bool)
    assert np.array_equal(ser[ok], eser[ok])                      # serialized credentials, byte for byte
    assert np.array_equal(s1[ok], ref.g1_op(ref.G_NORM, e1[ok]))  # raw limbs after normalisation
    assert np.array_equal(s2[ok], ref.g1_op(ref.G_NORM, e2[ok]))
    assert not s1[~ok].any() and not s2[~ok].any() and not ser[~ok].any()
    pk.close()


def test_provide_id_needs_signer_secret(gpu_pkg, ref):
    wl = workload.make_issuance_workload(2, 2, 1, seed=4)
    pk = _pk(gpu_pkg, wl.key, 6, with_secret=False)
    with pytest.raises(ValueError):
        gpu_pkg.PSSigner(pk)
    pk.close()


@pytest.mark.parametrize("n_attrs,n_hidden,with_id,lanes,w", [(5, 2, True, 60, 8), (10, 2, True, 36, 16), (4, 2, False, 24, 6),
                                                            (2, 2, True, 12, 8), (3, 3, False, 12, 5)])
def test_verify_id_matches_reference(gpu_pkg, ref, n_attrs, n_hidden, with_id, lanes, w):
    wl = workload.make_signon_workload(n_attrs, lanes, n_hidden, seed=3, with_id=with_id, tamper_every=3)
    ev = workload.expected_verify_id(wl)
    assert 0 < ev.sum() < lanes
    pk = _pk(gpu_pkg, wl.key, w)
    got = gpu_pkg.PSVerifier(pk).el_passo_verify_id(wl.proof, wl.proof_attrs, wl.ads, wl.service_pt, wl.y, wl.g, wl.h,
                                                    with_id=with_id)
    assert got.tolist() == ev.tolist(), wl.tampered.tolist()
    # second batch with the same service/authority points hits the cached per-batch tables
    got2 = gpu_pkg.PSVerifier(pk).el_passo_verify_id(wl.proof, wl.proof_attrs, wl.ads, wl.service_pt, wl.y, wl.g, wl.h,
                                                     with_id=with_id)
    assert got2.tolist() == ev.tolist()
    pk.close()


def test_verify_id_unnormalized_inputs(gpu_pkg, ref):
    """proof points with z != 1 (as they leave the prover, before any serialization round trip)."""
    wl = workload.make_signon_workload(5, 10, 2, seed=5, with_id=True)
    p = dict(wl.proof)
    for name in ("phi", "E1", "E2", "sig1", "sig2"):   # P = 2P - P leaves the group element, changes z
        d = ref.g1_op(ref.G_DBL, p[name])
        p[name] = ref.g1_op(ref.G_SUB, d, wl.proof[name])
        assert not np.array_equal(p[name], wl.proof[name])
    p["k"] = ref.g2_op(ref.G_SUB, ref.g2_op(ref.G_DBL, p["k"]), wl.proof["k"])
    wl.proof = p
    ev = workload.expected_verify_id(wl)
    assert ev.all()
    pk = _pk(gpu_pkg, wl.key, 8)
    got = gpu_pkg.PSVerifier(pk).el_passo_verify_id(p, wl.proof_attrs, wl.ads, wl.service_pt, wl.y, wl.g, wl.h)
    assert got.tolist() == ev.tolist()
    pk.close()


@bls_only
def test_elpasso_golden_fixtures_on_gpu(gpu_pkg):
    """committed reference outputs (tests/golden/elpasso.json): issuance bytes + sign-on verdicts."""
    keys = json.load(open(os.path.join(G, "keys.json")))["keys"]["5"]
    p = json.load(open(os.path.join(G, "elpasso.json")))
    arr = lambda h, w: np.frombuffer(bytes.fromhex(h), dtype=np.uint64).reshape(-1, w).copy()  # noqa: E731
    pk = gpu_pkg.PSPubKey(arr(keys["g"], 18), arr(keys["gg"], 36), arr(keys["XX"], 36), arr(keys["Y"], 18),
                          arr(keys["YY"], 36), X_secret=arr(keys["X"], 18), window_bits=8)
    q = p["provide_id"]
    N = len(q["verdict"])
    attrs = [[a.encode() for a in lane] for lane in q["attrs"]]
    v, s1, s2, ser = gpu_pkg.PSSigner(pk).el_passo_provide_id(arr(q["A"], 18), arr(q["c"], 4), arr(q["rs"], 4).reshape(N, -1, 4),
                                                              attrs, [a.encode() for a in q["ads"]], arr(q["u"], 4))
    assert v.tolist() == q["verdict"]
    assert ser.tobytes().hex() == q["ser"]
    for name in ("verify_id", "verify_id_without_id_retrieval"):
        q = p[name]
        N = len(q["verdict"])
        proof = {k: arr(q[k], w) for k, w in (("sig1", 18), ("sig2", 18), ("k", 36), ("phi", 18), ("E1", 18), ("E2", 18), ("c", 4))}
        proof["rs"] = arr(q["rs"], 4).reshape(N, -1, 4)
        attrs = [[a.encode() for a in lane] for lane in q["attrs"]]
        got = gpu_pkg.PSVerifier(pk).el_passo_verify_id(proof, attrs, [a.encode() for a in q["ads"]], arr(q["service_pt"], 18),
                                                        arr(q["y"], 18), arr(q["g"], 18), arr(q["h"], 18),
                                                        with_id=(name == "verify_id"))
        assert got.tolist() == q["verdict"], name
    pk.close()
