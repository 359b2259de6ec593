"""GPU parity of the prover-side batch entry points against the reference's own methods run under
seeded RandGens, with the scalars those methods drew handed to the GPU in draw order:
psb_request_id vs PSRequester::el_passo_request_id (src/ps-requester.cc:19-97)
psb_unblind    vs PSRequester::unblind_credential  (:99-113)
psb_prove_id   vs PSRequester::el_passo_prove_id   (:150-310) and el_passo_prove_id_without_id_retrieval (:312-432)
plus the round trips through the verifier-side batch entries (request -> provide_id -> unblind -> verify; prove -> verify_id)."""
import numpy as np
import pytest

from tests import workload
from tests.conftest import bls_only

pytestmark = pytest.mark.gpu


def _pk(gpu_pkg, key, w=8, with_secret=False):
    return gpu_pkg.PSPubKey(key.g, key.gg, key.XX, key.Y, key.YY, X_secret=key.X if with_secret else None, window_bits=w)


@pytest.mark.parametrize("n_attrs,n_hidden,lanes,w", [(5, 2, 70, 8), (20, 2, 33, 16), (1, 1, 8, 6), (2, 0, 8, 8), (3, 3, 9, 5)])
def test_request_id_and_unblind_match_reference(gpu_pkg, ref, n_attrs, n_hidden, lanes, w):
    pw = workload.make_prover_request_workload(n_attrs, lanes, n_hidden, seed=6)
    pk = _pk(gpu_pkg, pw.key, w, with_secret=True)
    rq = gpu_pkg.PSRequester(pk)
    A, c, rs = rq.el_passo_request_id(pw.attrs, pw.hidden, pw.ads, pw.rnd)
    assert np.array_equal(A, ref.g1_op(ref.G_NORM, pw.exp_A))
    assert np.array_equal(c, pw.exp_c)
    assert np.array_equal(rs, pw.exp_rs)
    s1, s2 = rq.unblind_credential(pw.blind_sig1, pw.blind_sig2, pw.rnd[:, 0])
    assert np.array_equal(s2, ref.g1_op(ref.G_NORM, pw.exp_unblind2))
    # round trip on the GPU: our request is accepted by the batched issuer, and the unblinded credential verifies
    ref.seed(99)
    u = ref.fr_rand(lanes)
    req_attrs = [[b"" if pw.hidden[i] else lane[i] for i in range(n_attrs)] for lane in pw.attrs]
    v, b1, b2, _ = gpu_pkg.PSSigner(pk).el_passo_provide_id(A, c, rs, req_attrs, pw.ads, u)
    assert v.all()
    if n_attrs != 1:   # a single-attribute request is signed as a bare commitment (sign_hybrid, ps-signer.cc:115-117)
        _, un2 = rq.unblind_credential(b1, b2, pw.rnd[:, 0])
        assert gpu_pkg.PSVerifier(pk).verify(b1, un2, pw.attrs).all()
    pk.close()


@pytest.mark.parametrize("n_attrs,n_hidden,with_id,lanes,w", [(5, 2, True, 40, 8), (10, 2, True, 33, 16), (4, 2, False, 24, 6),
                                                            (2, 0, True, 12, 8), (3, 3, False, 12, 5)])
def test_prove_id_matches_reference(gpu_pkg, ref, n_attrs, n_hidden, with_id, lanes, w):
    pw = workload.make_prover_signon_workload(n_attrs, lanes, n_hidden, seed=8, with_id=with_id)
    pk = _pk(gpu_pkg, pw.key, w)
    got = gpu_pkg.PSRequester(pk).el_passo_prove_id(pw.sig1, pw.sig2, pw.attrs, pw.hidden, pw.ads, pw.service_pt, pw.y, pw.g,
                                                    pw.h, rnd=pw.rnd, with_id=with_id)
    workload.assert_proof_equal(got, pw.exp, with_id)
    # the batched verifier accepts the batched prover's proofs
    proof_attrs = [[b"" if pw.hidden[i] else lane[i] for i in range(n_attrs)] for lane in pw.attrs]
    ok = gpu_pkg.PSVerifier(pk).el_passo_verify_id(got, proof_attrs, pw.ads, pw.service_pt, pw.y, pw.g, pw.h, with_id=with_id)
    if n_hidden >= (2 if with_id else 1):   # the statement is only well-formed when attributes 0 (and 1) are hidden
        assert ok.all()
    pk.close()


def test_prove_id_argument_checks(gpu_pkg, ref):
    pw = workload.make_prover_signon_workload(1, 2, 1, seed=8, with_id=False)
    pk = _pk(gpu_pkg, pw.key, 6)
    with pytest.raises(gpu_pkg.PsbError):   # id retrieval reads attributes[1]
        gpu_pkg.PSRequester(pk).el_passo_prove_id(pw.sig1, pw.sig2, pw.attrs, pw.hidden, pw.ads, pw.service_pt, pw.y, pw.g, pw.h,
                                                  rnd=np.zeros((2, 6, 4), np.uint64), with_id=True)
    with pytest.raises(ValueError):
        gpu_pkg.PSRequester(pk).el_passo_request_id(pw.attrs, np.zeros(3, np.uint8), pw.ads, pw.rnd)
    pk.close()


@bls_only
def test_prover_golden_fixtures_on_gpu(gpu_pkg):
    """committed reference outputs (tests/golden/prover.json): requests, unblinded credentials, proofs."""
    import json
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    keys = json.load(open(os.path.join(G, "keys.json")))["keys"]["5"]
    p = json.load(open(os.path.join(G, "prover.json")))
    arr = lambda h, w: np.frombuffer(bytes.fromhex(h), dtype=np.uint64).reshape(-1, w).copy()  # noqa: E731
    pk = gpu_pkg.PSPubKey(arr(keys["g"], 18), arr(keys["gg"], 36), arr(keys["XX"], 36), arr(keys["Y"], 18), arr(keys["YY"], 36),
                          window_bits=8)
    rq = gpu_pkg.PSRequester(pk)
    q = p["request_id"]
    N = len(q["ads"])
    hidden = np.array(q["hidden"], dtype=np.uint8)
    attrs = [[a.encode() for a in lane] for lane in q["attrs"]]
    rnd = arr(q["rnd"], 4).reshape(N, -1, 4)
    A, c, rs = rq.el_passo_request_id(attrs, hidden, [a.encode() for a in q["ads"]], rnd)
    assert A.tobytes().hex() == q["A"] and c.tobytes().hex() == q["c"] and rs.tobytes().hex() == q["rs"]
    _, un2 = rq.unblind_credential(arr(q["blind_sig1"], 18), arr(q["blind_sig2"], 18), rnd[:, 0])
    assert un2.tobytes().hex() == q["unblind_sig2"]
    for name, with_id in (("prove_id", True), ("prove_id_without_id_retrieval", False)):
        q = p[name]
        N = len(q["ads"])
        attrs = [[a.encode() for a in lane] for lane in q["attrs"]]
        got = rq.el_passo_prove_id(arr(q["in_sig1"], 18), arr(q["in_sig2"], 18), attrs, hidden, [a.encode() for a in q["ads"]],
                                   arr(q["service_pt"], 18), arr(q["y"], 18), arr(q["g"], 18), arr(q["h"], 18),
                                   rnd=arr(q["rnd"], 4).reshape(N, -1, 4), with_id=with_id)
        for f in ("sig1", "sig2", "k", "phi", "c", "rs") + (("E1", "E2") if with_id else ()):
            assert got[f].tobytes().hex() == q[f], (name, f)
    pk.close()
