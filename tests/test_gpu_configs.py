"""GPU parity at the shapes BASELINE.json names (configs[2..4]) -- the attribute counts and hidden patterns of the
benchmark configs at thousands of lanes, not only the small batches of test_gpu_verify / test_gpu_elpasso:

  cfg5  PSVerifier::verify, 50 attributes (window tables of 50 bases, w = 16), 2^12 lanes, all four tamper kinds,
        every lane against the reference (src/ps-verifier.cc:13-35)
  cfg3  el_passo_verify_id, 10 attributes / 2 hidden, 2^14 proofs made by the reference's own prover, six tamper kinds;
        every tampered lane and every 4th lane against the reference (src/ps-verifier.cc:37-138)
  cfg4  el_passo_provide_id, 20 attributes / 2 hidden, 2^14 requests + randomize_credential of the issued credentials;
        verdicts and SERIALIZED credentials against the reference (src/ps-signer.cc:63-146, src/ps-requester.cc:139-148)

Expected values always come from the reference compiled here (oracle/_ref), never from the generator.
"""
import numpy as np
import pytest

from tests import workload
from tests.conftest import FPW

pytestmark = pytest.mark.gpu


def _sample(lanes, step, tampered):
    idx = np.union1d(np.arange(0, lanes, step), tampered)
    return idx.astype(np.int64)


def test_cfg5_verify_50_attributes(gpu_pkg, ref):
    lanes, n = 1 << 12, 50
    wl = workload.make_verify_workload(n_attrs=n, lanes=lanes, seed=31, tamper_every=64)
    assert len(wl.tampered) == lanes // 64            # sigma2 += g, attribute flipped, sigma1 = 0, swap: 16 lanes each
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=16)
    assert pk.table_bytes == n * 16 * 32768 * 32 * FPW   # 5.03 GB of affine G2 entries on BLS12-381
    got, gt = gpu_pkg.PSVerifier(pk).verify(wl.sig1, wl.sig2, (wl.blob, wl.off), want_gt=True)
    exp, exp_gt = workload.expected_verify(wl, want_gt=True)
    assert np.array_equal(got, exp)
    mask = np.ones(lanes, dtype=bool)
    mask[wl.tampered] = False
    assert exp[mask].all() and not exp[~mask].any()
    live = wl.sig1[:, 2 * FPW:].any(axis=1)           # the reference returns before the pairing when sigma1 == 0
    assert np.array_equal(gt[live], exp_gt[live])     # GT bytes of every lane, tampered ones included
    pk.close()


def test_cfg3_signon_10_attributes_2p14(gpu_pkg, ref):
    lanes = 1 << 14
    wl = workload.make_signon_workload(10, lanes, 2, seed=33, with_id=True, tamper_every=61)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=16)
    got = gpu_pkg.PSVerifier(pk).el_passo_verify_id(wl.proof, wl.proof_attrs, wl.ads, wl.service_pt, wl.y, wl.g, wl.h,
                                                    with_id=True)
    honest = np.ones(lanes, dtype=bool)
    honest[wl.tampered] = False
    assert got[honest].all()
    idx = _sample(lanes, 4, wl.tampered)
    sub = workload.SignonWorkload(wl.key, {k: np.ascontiguousarray(v[idx]) for k, v in wl.proof.items()},
                                  [wl.proof_attrs[i] for i in idx], [wl.ads[i] for i in idx], wl.service, wl.service_pt,
                                  wl.y, wl.g, wl.h, True, np.array([], dtype=np.int64))
    exp = workload.expected_verify_id(sub)
    assert np.array_equal(got[idx], exp)
    rejected = (~exp.astype(bool)).sum()
    # five of the six tamper kinds reject; sigma = (0, 0) passes in the reference (no zero check)
    assert rejected == len(wl.tampered) - len(wl.tampered[4::6])
    pk.close()


def test_cfg4_issuance_20_attributes_2p14_then_randomize(gpu_pkg, ref):
    lanes = 1 << 14
    wl = workload.make_issuance_workload(20, lanes, 2, seed=34, tamper_every=61)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, X_secret=wl.key.X, window_bits=16)
    v, s1, s2, ser = gpu_pkg.PSSigner(pk).el_passo_provide_id(wl.A, wl.c, wl.rs, wl.req_attrs, wl.ads, wl.u)
    honest = np.ones(lanes, dtype=bool)
    honest[wl.tampered] = False
    assert v[honest].all() and not v[~honest].any()
    idx = _sample(lanes, 4, wl.tampered)
    sub = workload.IssuanceWorkload(wl.key, np.ascontiguousarray(wl.A[idx]), np.ascontiguousarray(wl.c[idx]),
                                    np.ascontiguousarray(wl.rs[idx]), [wl.req_attrs[i] for i in idx],
                                    [wl.ads[i] for i in idx], np.ascontiguousarray(wl.u[idx]), np.array([], dtype=np.int64))
    ev, e1, e2, eser = workload.expected_provide_id(sub)
    assert np.array_equal(v[idx], ev)
    ok = ev.astype(bool)
    assert np.array_equal(ser[idx][ok], eser[ok])                 # serialized credentials, byte for byte
    assert not ser[idx][~ok].any()
    # randomize the issued credentials (config 4's second half): t host-supplied, serialized output against mcl's
    ref.seed(35)
    t = ref.fr_rand(lanes)
    o1, o2, rser = gpu_pkg.PSRequester.randomize_credential(s1, s2, t, want_serialized=True)
    j = idx[ok]
    r1, r2, exp_ser = ref.randomize(np.ascontiguousarray(s1[j]), np.ascontiguousarray(s2[j]), np.ascontiguousarray(t[j]))
    assert np.array_equal(rser[j], exp_ser)
    assert np.array_equal(o1[j], ref.g1_op(ref.G_NORM, r1)) and np.array_equal(o2[j], ref.g1_op(ref.G_NORM, r2))
    pk.close()


def test_pipelined_entry_points_across_chunk_boundaries(gpu_pkg, ref):
    """psb_provide_id / psb_randomize / psb_request_id / psb_unblind cut a batch of >= 2^16 lanes into four chunks that run as a
    two-stream copy / compute pipeline (csrc/psb_api.cu, pipe_chunk).  A reference-checked workload of 1 024 distinct lanes is
    tiled 80 times (81 920 lanes: chunks of 20 480 lanes, so every chunk boundary falls INSIDE a tile) with page-locked buffers and
    caller-owned outputs: tile 0 must equal the reference and every other tile must equal tile 0, byte for byte."""
    D, reps, n, nh = 1024, 80, 6, 2
    N = D * reps
    tile = lambda a: np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))  # noqa: E731
    same = lambda a: np.array_equal(a.reshape(reps, D, -1), np.broadcast_to(a[:D].reshape(1, D, -1), (reps, D, a[:D].size // D)))  # noqa: E731
    wl = workload.make_issuance_workload(n, D, nh, seed=71, tamper_every=37)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, X_secret=wl.key.X, window_bits=12)
    ev, e1, e2, eser = workload.expected_provide_id(wl)
    A, c, rs, u = (gpu_pkg.pinned_copy(tile(a)) for a in (wl.A, wl.c, wl.rs, wl.u))
    attrs, ads = gpu_pkg.pack_attrs(wl.req_attrs * reps), gpu_pkg.pack_strings(wl.ads * reps)
    G1W = wl.A.shape[1]
    out = (gpu_pkg.pinned_empty(N, np.uint8), gpu_pkg.pinned_empty((N, G1W), np.uint64), gpu_pkg.pinned_empty((N, G1W), np.uint64),
           gpu_pkg.pinned_empty((N, eser.shape[1]), np.uint8))
    v, s1, s2, ser = gpu_pkg.PSSigner(pk).el_passo_provide_id(A, c, rs, attrs, ads, u, out=out)
    ok = ev.astype(bool)
    assert np.array_equal(v[:D], ev) and np.array_equal(ser[:D][ok], eser[ok]) and 0 < ok.sum() < D
    assert same(v) and same(s1) and same(s2) and same(ser)
    # randomisation of the issued credentials
    ref.seed(72)
    t = ref.fr_rand(D)
    o1, o2, rser = gpu_pkg.PSRequester.randomize_credential(s1, s2, gpu_pkg.pinned_copy(tile(t)), want_serialized=True)
    _, _, exp_ser = ref.randomize(np.ascontiguousarray(s1[:D][ok]), np.ascontiguousarray(s2[:D][ok]), np.ascontiguousarray(t[ok]))
    assert np.array_equal(rser[:D][ok], exp_ser) and same(o1) and same(o2) and same(rser)
    pk.close()
    # prover side: request_id and unblind
    pw = workload.make_prover_request_workload(n, D, nh, seed=73)
    pk = gpu_pkg.PSPubKey(pw.key.g, pw.key.gg, pw.key.XX, pw.key.Y, pw.key.YY, window_bits=12)
    rq = gpu_pkg.PSRequester(pk)
    A2, c2, rs2 = rq.el_passo_request_id(gpu_pkg.pack_attrs(pw.attrs * reps), pw.hidden, gpu_pkg.pack_strings(pw.ads * reps),
                                         gpu_pkg.pinned_copy(tile(pw.rnd)))
    assert np.array_equal(A2[:D], ref.g1_op(ref.G_NORM, pw.exp_A)) and np.array_equal(c2[:D], pw.exp_c) and np.array_equal(rs2[:D], pw.exp_rs)
    assert same(A2) and same(c2) and same(rs2)
    _, un2 = rq.unblind_credential(tile(pw.blind_sig1), tile(pw.blind_sig2), tile(pw.rnd[:, 0].copy()))
    assert np.array_equal(un2[:D], ref.g1_op(ref.G_NORM, pw.exp_unblind2)) and same(un2)
    pk.close()
