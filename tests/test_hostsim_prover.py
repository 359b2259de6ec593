"""CPU: the prover-side lane functions of csrc/prover.cuh, compiled for the host by tests/hostsim,
against the reference's own PSRequester::el_passo_request_id / unblind_credential / el_passo_prove_id run under
seeded RandGens -- the scalars the reference draws are replayed from the same streams and handed to our lanes."""
import ctypes as C

import numpy as np
import pytest

from tests import workload


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n_attrs,n_hidden", [(5, 2), (1, 1), (2, 0), (3, 3)])
def test_request_id_and_unblind_lanes(hostsim, ref, n_attrs, n_hidden):
    lanes = 4
    pw = workload.make_prover_request_workload(n_attrs, lanes, n_hidden, seed=6)
    blob, off = ref.pack_attrs(pw.attrs)
    ad_blob, ad_off = ref.pack_strings(pw.ads)
    A = np.zeros((lanes, 18), dtype=np.uint64)
    c = np.zeros((lanes, 4), dtype=np.uint64)
    rs = np.zeros((lanes, n_hidden + 1, 4), dtype=np.uint64)
    hostsim.hostsim_request_id(C.c_int(n_attrs), C.c_int(4), _p(pw.key.g), _p(pw.key.Y), C.c_size_t(lanes), _p(pw.hidden),
                               _p(blob), _p(off), _p(ad_blob), _p(ad_off), _p(pw.rnd), _p(A), _p(c), _p(rs))
    assert np.array_equal(A, ref.g1_op(ref.G_NORM, pw.exp_A))
    assert np.array_equal(c, pw.exp_c)
    assert np.array_equal(rs, pw.exp_rs)
    out2 = np.zeros((lanes, 18), dtype=np.uint64)
    hostsim.hostsim_unblind(C.c_size_t(lanes), _p(pw.blind_sig1), _p(pw.blind_sig2), _p(pw.rnd[:, 0].copy()), _p(out2))
    assert np.array_equal(out2, ref.g1_op(ref.G_NORM, pw.exp_unblind2))


@pytest.mark.parametrize("n_attrs,n_hidden,with_id", [(5, 2, True), (4, 2, False), (2, 0, True), (3, 3, False)])
def test_prove_id_lanes(hostsim, ref, n_attrs, n_hidden, with_id):
    lanes = 3
    pw = workload.make_prover_signon_workload(n_attrs, lanes, n_hidden, seed=8, with_id=with_id)
    blob, off = ref.pack_attrs(pw.attrs)
    ad_blob, ad_off = ref.pack_strings(pw.ads)
    per = n_hidden + (2 if with_id else 1)
    o = dict(sig1=np.zeros((lanes, 18), np.uint64), sig2=np.zeros((lanes, 18), np.uint64), k=np.zeros((lanes, 36), np.uint64),
             phi=np.zeros((lanes, 18), np.uint64), E1=np.zeros((lanes, 18), np.uint64), E2=np.zeros((lanes, 18), np.uint64),
             c=np.zeros((lanes, 4), np.uint64), rs=np.zeros((lanes, per, 4), np.uint64))
    hostsim.hostsim_prove_id(C.c_int(n_attrs), C.c_int(4), _p(pw.key.gg), _p(pw.key.XX), _p(pw.key.YY), C.c_size_t(lanes),
                             _p(pw.sig1), _p(pw.sig2), _p(pw.hidden), _p(blob), _p(off), _p(ad_blob), _p(ad_off),
                             _p(pw.service_pt), _p(pw.y), _p(pw.g), _p(pw.h), C.c_int(int(with_id)), _p(pw.rnd),
                             _p(o["sig1"]), _p(o["sig2"]), _p(o["k"]), _p(o["phi"]), _p(o["E1"]), _p(o["E2"]), _p(o["c"]),
                             _p(o["rs"]))
    workload.assert_proof_equal(o, pw.exp, with_id)
