"""GPU parity of the wire-format ingest: batched point decompression
(psb_g1_deserialize / psb_g2_deserialize vs mcl's G1::deserialize / G2::deserialize, ec.hpp:924-1057) and PS verification
straight from serialized credentials (psb_verify_ser vs PSCredential::fromBufferString + PSVerifier::verify)."""
import numpy as np
import pytest

from tests import workload
from tests.conftest import G1_SER

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("g2", [False, True])
def test_deserialize_matches_mcl(gpu_pkg, ref, g2):
    enc, want, okv = workload.deserialize_cases(g2, count=64)
    assert 0 < okv.sum() < len(okv)
    out, ok = (gpu_pkg.g2_deserialize if g2 else gpu_pkg.g1_deserialize)(enc)
    assert ok.tolist() == okv.tolist()
    good = okv.astype(bool)
    assert np.array_equal(out[good], want[good])
    assert not out[~good].any()
    # round trip at size: serialize(k * base) -> deserialize gives the normalised point back
    ref.seed(23)
    k = ref.fr_rand(2000)
    if g2:
        P = ref.g2_op(ref.G_NORM, ref.g2_mul(ref.hash_to_g2(b"edf"), k, ref.hw_threads()))
        back, ok2 = gpu_pkg.g2_deserialize(ref.g2_serialize(P))
    else:
        P = ref.g1_op(ref.G_NORM, ref.g1_mul(ref.hash_to_g1(b"abc"), k, ref.hw_threads()))
        back, ok2 = gpu_pkg.g1_deserialize(ref.g1_serialize(P))
    assert ok2.all() and np.array_equal(back, P)


def test_verify_serialized_credentials(gpu_pkg, ref):
    """credentials in the reference's own TLV encoding (PSCredential::toBufferString, 100 bytes), incl. tampered lanes,
    an undecodable lane (x not on the curve) and the infinity encoding: verdicts equal PSVerifier::verify on the points
    the reference's own parser produces."""
    lanes = 200
    wl = workload.make_verify_workload(n_attrs=5, lanes=lanes, seed=29, tamper_every=9)
    buf = ref.cred_encode(wl.sig1, wl.sig2)
    S, T = G1_SER, 2 + G1_SER                     # compressed G1 (48 / 32 bytes), one TLV-framed G1
    assert buf.shape == (lanes, 2 * T) and buf[0, 0] == 1 and buf[0, 1] == S and buf[0, T] == 1 and buf[0, T + 1] == S
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
    ver = gpu_pkg.PSVerifier(pk)
    s1, s2 = ref.cred_decode(buf)
    exp = ref.ps_verify(wl.key, s1, s2, wl.attrs)
    got, dec = ver.verify_serialized(buf, wl.attrs)
    assert dec.all() and np.array_equal(got, exp)
    assert exp.sum() == lanes - len(wl.tampered)
    # undecodable sigma2 on lane 0 (find an x with x^3 + 4 a non-residue), flipped y flag on lane 1
    bad = buf.copy()
    for v in range(1, 50):
        bad[0, T + 2:2 * T] = 0
        bad[0, T + 2] = v
        if not ref.g1_deserialize(bad[0:1, T + 2:2 * T].copy())[1]:
            break
    bad[1, T - 1] ^= 0x80
    got2, dec2 = ver.verify_serialized(bad, wl.attrs)
    assert dec2.tolist() == [0] + [1] * (lanes - 1)
    assert got2[0] == 0 and got2[1] == 0 and np.array_equal(got2[2:], exp[2:])
    # bare layout: sigma1 || sigma2 (96 / 64 bytes)
    raw = np.concatenate([buf[:, 2:T], buf[:, T + 2:2 * T]], axis=1)
    got3, _ = ver.verify_serialized(raw, wl.attrs, stride=2 * S, off1=0, off2=S)
    assert np.array_equal(got3, exp)
    pk.close()


# ---- IdProof / PSCredRequest straight from the wire (psb_verify_id_ser / psb_provide_id_ser) ----------------------------
def _wire_lanes(wire):
    blob, off = wire
    return [bytes(blob[int(off[j]):int(off[j + 1])]) for j in range(len(off) - 1)]


@pytest.mark.parametrize("n_attrs,with_id,b64", [(5, True, False), (10, True, True), (4, False, True), (3, False, False)])
def test_verify_id_from_wire(gpu_pkg, ref, n_attrs, with_id, b64):
    """sign-on proofs in the reference's own encoding (IdProof::toBufferString, optionally base64), parsed and decompressed
    on the GPU: verdicts equal IdProof::fromBufferString + el_passo_verify_id, tampered lanes included; malformed buffers
    (which the reference's parser answers with false, an exception or a crash -- tests/test_hostsim_wire.py) give 0."""
    lanes = 96
    wl = workload.make_signon_workload(n_attrs, lanes, 2, seed=61, with_id=with_id, tamper_every=5)
    wl.proof_attrs[7][n_attrs - 1] = b"y" * 280            # 3-byte length prefix
    exp = workload.expected_verify_id(wl)
    msgs = _wire_lanes(ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=with_id, base64=b64))
    # (the 280-byte attribute changes lane 7's K: take the expected verdicts from the reference on the WIRE form)
    rv, st = ref.verify_id_wire(wl.key, gpu_pkg.pack_strings(msgs), wl.ads, wl.service, wl.y, wl.g, wl.h, with_id=with_id,
                                base64=b64, nthreads=ref.hw_threads())
    assert not st.any() and rv.sum() > 0 and np.array_equal(np.delete(rv, 7), np.delete(exp, 7))
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
    ver = gpu_pkg.PSVerifier(pk)
    got, parsed = ver.el_passo_verify_id_wire(msgs, wl.ads, wl.service_pt, wl.y, wl.g, wl.h, with_id=with_id, base64=b64)
    assert parsed.all() and got.tolist() == rv.tolist()
    # malformed lanes inside an otherwise honest batch: the neighbours are unaffected
    bad = list(msgs)
    raw = _wire_lanes(ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=with_id, base64=False))
    import base64 as pyb64
    enc = (lambda b: pyb64.b64encode(b)) if b64 else (lambda b: b)
    bad[0] = enc(raw[0][:40])                               # short buffer
    bad[1] = enc(b"\x02" + raw[1][1:])                      # wrong type byte
    bad[2] = enc(raw[2][:1] + b"\xfe" + raw[2][2:])         # invalid length byte
    bad[3] = b""                                            # empty
    bad[5] = enc(raw[5][:2] + b"\xff" * (len(raw[5]) - 2))  # x >= p, garbage after
    got2, parsed2 = ver.el_passo_verify_id_wire(bad, wl.ads, wl.service_pt, wl.y, wl.g, wl.h, with_id=with_id, base64=b64)
    broken = [0, 1, 2, 3, 5]
    assert not parsed2[broken].any() and not got2[broken].any()
    keep = np.ones(lanes, dtype=bool)
    keep[broken] = False
    assert parsed2[keep].all() and np.array_equal(got2[keep], rv[keep])
    if with_id:   # a proof WITHOUT E1 / E2 is well formed but fails el_passo_verify_id (ps-verifier.cc:68-70) ...
        noe = _wire_lanes(ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=False, base64=b64))
        got3, parsed3 = ver.el_passo_verify_id_wire(noe, wl.ads, wl.service_pt, wl.y, wl.g, wl.h, with_id=True, base64=b64)
        assert not got3.any() and not parsed3.any()
    pk.close()


def test_verify_id_strict_rejects_zero_sigma(gpu_pkg, ref):
    """sigma1 = sigma2 = 0 with an honest NIZK passes the reference's el_passo_verify_id (no zero check);
    PSB_VID_REJECT_ZERO_SIGMA turns exactly those lanes into rejections"""
    lanes = 24
    wl = workload.make_signon_workload(5, lanes, 2, seed=62, with_id=True)
    wl.proof["sig1"][::4] = 0
    wl.proof["sig2"][::4] = 0
    exp = workload.expected_verify_id(wl)
    assert exp.all()                                        # the reference accepts the credential-less lanes
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
    ver = gpu_pkg.PSVerifier(pk)
    args = (wl.proof, wl.proof_attrs, wl.ads, wl.service_pt, wl.y, wl.g, wl.h)
    assert ver.el_passo_verify_id(*args).tolist() == exp.tolist()
    strict = ver.el_passo_verify_id(*args, strict=True)
    want = exp.copy()
    want[::4] = 0
    assert strict.tolist() == want.tolist()
    msgs = _wire_lanes(ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs))
    got, parsed = ver.el_passo_verify_id_wire(msgs, wl.ads, wl.service_pt, wl.y, wl.g, wl.h, strict=True)
    assert parsed.all() and got.tolist() == want.tolist()
    pk.close()


@pytest.mark.parametrize("n_attrs,b64", [(5, False), (20, True)])
def test_provide_id_from_wire(gpu_pkg, ref, n_attrs, b64):
    lanes = 64
    wl = workload.make_issuance_workload(n_attrs, lanes, 2, seed=63, tamper_every=6)
    msgs = _wire_lanes(ref.request_encode(wl.key, wl.A, wl.c, wl.rs, wl.req_attrs, base64=b64))
    ev, e1, e2, eser, st = ref.provide_id_wire(wl.key, gpu_pkg.pack_strings(msgs), wl.ads, wl.u, base64=b64, nthreads=ref.hw_threads())
    assert not st.any() and 0 < ev.sum() < lanes
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, X_secret=wl.key.X, window_bits=8)
    sg = gpu_pkg.PSSigner(pk)
    v, s1, s2, ser, parsed = sg.el_passo_provide_id_wire(msgs, wl.ads, wl.u, base64=b64)
    assert parsed.all() and v.tolist() == ev.tolist()
    ok = ev.astype(bool)
    assert np.array_equal(ser[ok], eser[ok]) and not ser[~ok].any()
    bad = list(msgs)
    bad[0] = msgs[0][:30]
    bad[1] = b""
    v2, _, _, ser2, parsed2 = sg.el_passo_provide_id_wire(bad, wl.ads, wl.u, base64=b64)
    assert not parsed2[:2].any() and not v2[:2].any() and not ser2[:2].any()
    assert parsed2[2:].all() and v2[2:].tolist() == ev[2:].tolist() and np.array_equal(ser2[2:][ok[2:]], eser[2:][ok[2:]])
    pk.close()


@pytest.mark.parametrize("n_attrs,na", [(5, 0), (5, 5), (5, 1), (5, 3), (20, 20)])
def test_sign_matches_reference(gpu_pkg, ref, n_attrs, na):
    """batched PSSigner::sign_commitment (na = 0) / sign_hybrid against the reference with the same u: serialized bytes"""
    lanes = 40
    key = ref.KeyMaterial(n_attrs, seed_=1)
    ref.seed(64)
    Cm = ref.g1_mul(key.g, ref.fr_rand(lanes), ref.hw_threads())
    Cm[3] = 0                                                # the zero commitment signs fine
    u = ref.fr_rand(lanes)
    attrs = [[b"" if (i + j) % 3 == 0 else b"v%d.%d" % (i, j) for i in range(na)] for j in range(lanes)] if na else None
    e1, e2, eser = ref.sign(key, Cm, attrs, u, nthreads=ref.hw_threads())
    pk = gpu_pkg.PSPubKey(key.g, key.gg, key.XX, key.Y, key.YY, X_secret=key.X, window_bits=8)
    s1, s2, ser = gpu_pkg.PSSigner(pk).sign_hybrid(Cm, attrs, u)
    assert np.array_equal(ser, eser)
    assert np.array_equal(s1, ref.g1_op(ref.G_NORM, e1)) and np.array_equal(s2, ref.g1_op(ref.G_NORM, e2))
    pk.close()


def test_key_outlives_context_as_dead_handle(gpu_pkg, ref):
    """psb_init / psb_shutdown release the device memory of live keys; a later call with such a key fails cleanly and
    psb_key_destroy stays valid (key and device lifetimes)"""
    wl = workload.make_verify_workload(n_attrs=2, lanes=4, seed=65)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=6)
    assert gpu_pkg.PSVerifier(pk).verify(wl.sig1, wl.sig2, wl.attrs).all()
    gpu_pkg.init([0])                                       # re-initialise: pk is now a dead handle
    with pytest.raises(gpu_pkg.PsbError):
        gpu_pkg.PSVerifier(pk).verify(wl.sig1, wl.sig2, wl.attrs)
    pk.close()                                              # no crash
    pk2 = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=6)
    assert gpu_pkg.PSVerifier(pk2).verify(wl.sig1, wl.sig2, wl.attrs).all()
    pk2.close()


@pytest.mark.parametrize("b64", [False, True])
def test_wire_encode_matches_reference_and_round_trips(gpu_pkg, ref, b64):
    """psb_wire_encode (the batched prover's output format) against IdProof::toBufferString / PSCredRequest::toBufferString /
    PSBuffer::toBase64 byte for byte, then back through psb_verify_id_ser: same verdicts as the reference on the same bytes."""
    n, lanes = 5, 300
    wl = workload.make_signon_workload(n, lanes, 2, seed=45, with_id=True, tamper_every=7)
    wl.proof_attrs[1][3] = b"y" * 300
    want = ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=True, base64=b64)
    blob, off = gpu_pkg.idproof_serialize(wl.proof, wl.proof_attrs, n, with_id=True, base64=b64)
    assert np.array_equal(off, want[1]) and np.array_equal(blob[:int(off[-1])], want[0][:int(off[-1])])
    nb, no = gpu_pkg.idproof_serialize(wl.proof, wl.proof_attrs, n, with_id=False, base64=b64)
    w2 = ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=False, base64=b64)
    assert np.array_equal(no, w2[1]) and np.array_equal(nb[:int(no[-1])], w2[0][:int(no[-1])])
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
    got, parsed = gpu_pkg.PSVerifier(pk).el_passo_verify_id_wire((blob, off), wl.ads, wl.service_pt, wl.y, wl.g, wl.h, base64=b64)
    exp = workload.expected_verify_id(wl)
    assert parsed.all() and np.array_equal(got, exp) and 0 < exp.sum() < lanes
    iw = workload.make_issuance_workload(n, 64, 2, seed=46)
    want = ref.request_encode(iw.key, iw.A, iw.c, iw.rs, iw.req_attrs, base64=b64)
    blob, off = gpu_pkg.request_serialize(iw.A, iw.c, iw.rs, iw.req_attrs, n, base64=b64)
    assert np.array_equal(off, want[1]) and np.array_equal(blob[:int(off[-1])], want[0][:int(off[-1])])
    pk.close()
