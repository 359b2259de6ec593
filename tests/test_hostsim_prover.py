"""CPU: the prover-side lane functions of csrc/prover.cuh, compiled for the host by tests/hostsim,
against the reference's own PSRequester::el_passo_request_id / unblind_credential / el_passo_prove_id run under
seeded RandGens -- the scalars the reference draws are replayed from the same streams and handed to our lanes."""
import ctypes as C

import numpy as np
import pytest

from tests import workload
from tests.conftest import G1W, G2W, bls_only


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n_attrs,n_hidden", [(5, 2), (1, 1), (2, 0), (3, 3)])
def test_request_id_and_unblind_lanes(hostsim, ref, n_attrs, n_hidden):
    lanes = 4
    pw = workload.make_prover_request_workload(n_attrs, lanes, n_hidden, seed=6)
    blob, off = ref.pack_attrs(pw.attrs)
    ad_blob, ad_off = ref.pack_strings(pw.ads)
    A = np.zeros((lanes, G1W), dtype=np.uint64)
    c = np.zeros((lanes, 4), dtype=np.uint64)
    rs = np.zeros((lanes, n_hidden + 1, 4), dtype=np.uint64)
    hostsim.hostsim_request_id(C.c_int(n_attrs), C.c_int(4), _p(pw.key.g), _p(pw.key.Y), C.c_size_t(lanes), _p(pw.hidden),
                               _p(blob), _p(off), _p(ad_blob), _p(ad_off), _p(pw.rnd), _p(A), _p(c), _p(rs))
    assert np.array_equal(A, ref.g1_op(ref.G_NORM, pw.exp_A))
    assert np.array_equal(c, pw.exp_c)
    assert np.array_equal(rs, pw.exp_rs)
    out2 = np.zeros((lanes, G1W), dtype=np.uint64)
    hostsim.hostsim_unblind(C.c_size_t(lanes), _p(pw.blind_sig1), _p(pw.blind_sig2), _p(pw.rnd[:, 0].copy()), _p(out2))
    assert np.array_equal(out2, ref.g1_op(ref.G_NORM, pw.exp_unblind2))


@pytest.mark.parametrize("n_attrs,n_hidden,with_id", [(5, 2, True), (4, 2, False), (2, 0, True), (3, 3, False)])
def test_prove_id_lanes(hostsim, ref, n_attrs, n_hidden, with_id):
    lanes = 3
    pw = workload.make_prover_signon_workload(n_attrs, lanes, n_hidden, seed=8, with_id=with_id)
    blob, off = ref.pack_attrs(pw.attrs)
    ad_blob, ad_off = ref.pack_strings(pw.ads)
    per = n_hidden + (2 if with_id else 1)
    o = dict(sig1=np.zeros((lanes, G1W), np.uint64), sig2=np.zeros((lanes, G1W), np.uint64), k=np.zeros((lanes, G2W), np.uint64),
             phi=np.zeros((lanes, G1W), np.uint64), E1=np.zeros((lanes, G1W), np.uint64), E2=np.zeros((lanes, G1W), np.uint64),
             c=np.zeros((lanes, 4), np.uint64), rs=np.zeros((lanes, per, 4), np.uint64))
    hostsim.hostsim_prove_id(C.c_int(n_attrs), C.c_int(4), _p(pw.key.gg), _p(pw.key.XX), _p(pw.key.YY), C.c_size_t(lanes),
                             _p(pw.sig1), _p(pw.sig2), _p(pw.hidden), _p(blob), _p(off), _p(ad_blob), _p(ad_off),
                             _p(pw.service_pt), _p(pw.y), _p(pw.g), _p(pw.h), C.c_int(int(with_id)), _p(pw.rnd),
                             _p(o["sig1"]), _p(o["sig2"]), _p(o["k"]), _p(o["phi"]), _p(o["E1"]), _p(o["E2"]), _p(o["c"]),
                             _p(o["rs"]))
    workload.assert_proof_equal(o, pw.exp, with_id)


def _golden():
    import json
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    keys = json.load(open(os.path.join(G, "keys.json")))["keys"]["5"]
    p = json.load(open(os.path.join(G, "prover.json")))
    arr = lambda h, w: np.frombuffer(bytes.fromhex(h), dtype=np.uint64).reshape(-1, w).copy()  # noqa: E731
    return keys, p, arr


@bls_only
def test_prover_golden_fixtures_on_hostsim(hostsim):
    """committed reference outputs (tests/golden/prover.json), no live reference needed."""
    from oracle import ref as R
    keys, p, arr = _golden()
    q = p["request_id"]
    N = len(q["ads"])
    hidden = np.array(q["hidden"], dtype=np.uint8)
    h = int(hidden.sum())
    blob, off = R.pack_attrs([[a.encode() for a in lane] for lane in q["attrs"]])
    ad_blob, ad_off = R.pack_strings([a.encode() for a in q["ads"]])
    A = np.zeros((N, 18), np.uint64); c = np.zeros((N, 4), np.uint64); rs = np.zeros((N, h + 1, 4), np.uint64)
    rnd = arr(q["rnd"], 4).reshape(N, h + 2, 4)
    hostsim.hostsim_request_id(C.c_int(5), C.c_int(4), _p(arr(keys["g"], 18)), _p(arr(keys["Y"], 18)), C.c_size_t(N), _p(hidden),
                               _p(blob), _p(off), _p(ad_blob), _p(ad_off), _p(rnd), _p(A), _p(c), _p(rs))
    assert A.tobytes().hex() == q["A"] and c.tobytes().hex() == q["c"] and rs.tobytes().hex() == q["rs"]
    out2 = np.zeros((N, 18), np.uint64)
    hostsim.hostsim_unblind(C.c_size_t(N), _p(arr(q["blind_sig1"], 18)), _p(arr(q["blind_sig2"], 18)), _p(rnd[:, 0].copy()), _p(out2))
    assert out2.tobytes().hex() == q["unblind_sig2"]
    for name, with_id in (("prove_id", True), ("prove_id_without_id_retrieval", False)):
        q = p[name]
        N = len(q["ads"])
        per = h + (2 if with_id else 1)
        blob, off = R.pack_attrs([[a.encode() for a in lane] for lane in q["attrs"]])
        ad_blob, ad_off = R.pack_strings([a.encode() for a in q["ads"]])
        o = dict(sig1=np.zeros((N, 18), np.uint64), sig2=np.zeros((N, 18), np.uint64), k=np.zeros((N, 36), np.uint64),
                 phi=np.zeros((N, 18), np.uint64), E1=np.zeros((N, 18), np.uint64), E2=np.zeros((N, 18), np.uint64),
                 c=np.zeros((N, 4), np.uint64), rs=np.zeros((N, per, 4), np.uint64))
        hostsim.hostsim_prove_id(C.c_int(5), C.c_int(4), _p(arr(keys["gg"], 36)), _p(arr(keys["XX"], 36)), _p(arr(keys["YY"], 36)),
                                 C.c_size_t(N), _p(arr(q["in_sig1"], 18)), _p(arr(q["in_sig2"], 18)), _p(hidden), _p(blob), _p(off),
                                 _p(ad_blob), _p(ad_off), _p(arr(q["service_pt"], 18)), _p(arr(q["y"], 18)), _p(arr(q["g"], 18)),
                                 _p(arr(q["h"], 18)), C.c_int(int(with_id)), _p(arr(q["rnd"], 4)), _p(o["sig1"]), _p(o["sig2"]),
                                 _p(o["k"]), _p(o["phi"]), _p(o["E1"]), _p(o["E2"]), _p(o["c"]), _p(o["rs"]))
        for f in ("sig1", "sig2", "k", "phi", "c", "rs") + (("E1", "E2") if with_id else ()):
            assert o[f].tobytes().hex() == q[f], (name, f)
