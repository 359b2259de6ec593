"""CPU: the wire-format lane functions of csrc/wire.cuh (base64, TLV walk, point decompression) and the signing lane
of csrc/protocol.cuh, compiled for the host by tests/hostsim, against the reference's own encoder / parser
(IdProof / PSCredRequest ::toBufferString, ::fromBufferString, PSBuffer::toBase64 / fromBase64, src/ps-encoding.cc) and
PSSigner::sign_hybrid / sign_commitment (src/ps-signer.cc:112-146).  The same lane functions run inside the GPU kernels
k_wire_base64 / k_wire_parse / k_wire_points / k_sign (tests/test_gpu_wire.py)."""
import base64 as pyb64
import ctypes as C

import numpy as np
import pytest

from tests import workload
from tests.conftest import FP_BYTES, G1W, G2W, GROUP_R


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


def _parse(hostsim, kind, n, msg: bytes, b64=False):
    g1 = np.zeros((5, G1W), dtype=np.uint64)
    g2 = np.zeros(G2W, dtype=np.uint64)
    c = np.zeros(4, dtype=np.uint64)
    rs = np.zeros((n + 2, 4), dtype=np.uint64)
    per, has_e = C.c_int(0), C.c_int(0)
    attr = np.zeros(len(msg) + 16, dtype=np.uint8)
    aoff = np.zeros(n + 1, dtype=np.uint64)
    buf = np.frombuffer(msg + b"\0" * 8, dtype=np.uint8).copy()
    ok = hostsim.hostsim_wire_parse(C.c_int(kind), C.c_int(n), _p(buf), C.c_size_t(len(msg)), C.c_int(int(b64)), _p(g1), _p(g2),
                                    _p(c), _p(rs), C.byref(per), _p(attr), _p(aoff), C.byref(has_e))
    attrs = [bytes(attr[int(aoff[i]):int(aoff[i + 1])]) for i in range(n)]
    return dict(ok=bool(ok), g1=g1, k=g2, c=c, rs=rs[:per.value], per=per.value, has_e=bool(has_e.value), attrs=attrs)


def _lanes(wire):
    blob, off = wire
    return [bytes(blob[int(off[j]):int(off[j + 1])]) for j in range(len(off) - 1)]


def _reference_outcome(ref, wl, msg: bytes):
    """IdProof::fromBufferString + el_passo_verify_id on ONE buffer in a forked child: the reference's parser is not memory
    safe on malformed input (lengths are trusted, failed deserializations ignored: SURVEY F9), so a lane may kill the
    process.  Returns accept / reject / throws / crash."""
    import multiprocessing as mp

    def child(q):
        import faulthandler, os
        faulthandler.disable()
        os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
        blob = np.frombuffer(msg + b"\0" * 8, dtype=np.uint8).copy()
        off = np.array([0, len(msg)], dtype=np.uint64)
        rv, st = ref.verify_id_wire(wl.key, (blob, off), wl.ads[:1], wl.service, wl.y, wl.g, wl.h, with_id=True)
        q.put("throws" if st[0] else ("accept" if rv[0] else "reject"))
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    pr = ctx.Process(target=child, args=(q,))
    pr.start()
    pr.join(60)
    if pr.is_alive():
        pr.kill()
        return "crash"
    return q.get() if pr.exitcode == 0 and not q.empty() else "crash"


def test_base64_decoder_rules(hostsim):
    """src/ps-encoding.cc:56-96: stop at '=' or at any byte outside the alphabet; a partial group of i sextets gives i - 1 bytes"""
    hostsim.hostsim_base64_decode.restype = C.c_size_t
    rng = np.random.default_rng(3)

    def dec(text: bytes):
        out = np.zeros(len(text) + 8, dtype=np.uint8)
        buf = np.frombuffer(text + b"\0" * 8, dtype=np.uint8).copy()
        k = hostsim.hostsim_base64_decode(_p(buf), C.c_size_t(len(text)), _p(out))
        return bytes(out[:k])
    for ln in list(range(0, 20)) + [100, 257]:
        raw = rng.bytes(ln)
        enc = pyb64.b64encode(raw)
        assert dec(enc) == raw
        assert dec(enc.rstrip(b"=")) == raw                 # unpadded text decodes the same
        assert dec(enc + b"!ignored") == raw or not enc.endswith(b"=") and dec(enc + b"!ignored") == raw
    assert dec(b"QUJD*QUJD") == b"ABC"                       # '*' ends the input
    assert dec(b"QUJDR") == b"ABC"                           # one dangling sextet: no byte
    assert dec(b"QUJDRE") == b"ABCD"                         # two sextets: one byte
    assert dec(b"") == b"" and dec(b"=QUJD") == b""


@pytest.mark.parametrize("with_id,b64", [(True, False), (True, True), (False, False), (False, True)])
def test_idproof_parse_matches_reference(hostsim, ref, with_id, b64):
    n, lanes = 5, 5
    wl = workload.make_signon_workload(n, lanes, 2, seed=41, with_id=with_id)
    wl.proof_attrs[1][3] = b"x" * 300                       # a 3-byte length prefix (253, hi, lo)
    wl.proof_attrs[2][4] = b"\xff\x00bin"
    wire = ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=with_id, base64=b64)
    msgs = _lanes(wire)
    if not b64:
        assert len(msgs[0]) == 3 * (2 + FP_BYTES) + (2 + 2 * FP_BYTES) + (2 + 32) + 2 + wl.proof["rs"].shape[1] * 33 + 2 + \
            sum(1 + len(a) for a in wl.proof_attrs[0]) + (2 * (2 + FP_BYTES) if with_id else 0)
    dec, per, has_e, status = ref.idproof_decode(wl.key, wire, n + 2, base64=b64)
    assert not status.any()
    for j, m in enumerate(msgs):
        got = _parse(hostsim, 0, n, m, b64)
        assert got["ok"] and got["has_e"] == with_id == bool(has_e[j]) and got["per"] == per[j] == wl.proof["rs"].shape[1]
        for slot, name in enumerate(("sig1", "sig2", "phi", "E1", "E2")):
            want = ref.g1_op(ref.G_NORM, dec[name][j:j + 1])[0] if (with_id or slot < 3) else np.zeros(G1W, np.uint64)
            assert np.array_equal(got["g1"][slot], want), name
        assert np.array_equal(got["k"], ref.g2_op(ref.G_NORM, dec["k"][j:j + 1])[0])
        assert np.array_equal(got["c"], dec["c"][j]) and np.array_equal(got["c"], wl.proof["c"][j])
        assert np.array_equal(got["rs"], dec["rs"][j][:per[j]])
        assert got["attrs"] == wl.proof_attrs[j]


def test_idproof_malformed_lanes_reject(hostsim, ref):
    """every way a buffer can be broken gives parsed = 0; the reference throws, or returns false, on the same buffers"""
    n = 5
    wl = workload.make_signon_workload(n, 1, 2, seed=42, with_id=True)
    msg = _lanes(ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=True))[0]
    T1, T2 = 2 + FP_BYTES, 2 + 2 * FP_BYTES
    at_c = 2 * T1 + T2 + T1                                   # offset of the Fr TLV
    at_rs = at_c + 34
    at_str = at_rs + 2 + 4 * 33
    end_attrs = at_str + 2 + sum(1 + len(a) for a in wl.proof_attrs[0])
    assert msg[at_c] == 3 and msg[at_rs] == 6 and msg[at_rs + 1] == 4 and msg[at_str] == 7 and msg[at_str + 1] == n
    assert len(msg) == end_attrs + 2 * T1
    assert _parse(hostsim, 0, n, msg)["ok"]
    # every proper prefix is malformed -- except the one that ends after the attribute list (a proof without E1 / E2)
    for cut in range(len(msg)):
        got = _parse(hostsim, 0, n, msg[:cut])
        assert got["ok"] == (cut == end_attrs), cut
        if cut == end_attrs:
            assert not got["has_e"]

    def mut(pos, val):
        b = bytearray(msg); b[pos] = val; return bytes(b)
    big_r = GROUP_R.to_bytes(32, "little")
    nonres = None
    for v in range(1, 60):                                    # an x with x^3 + b a non-residue: mcl's deserialize fails
        cand = bytes([v]) + b"\0" * (FP_BYTES - 1)
        if not ref.g1_deserialize(np.frombuffer(cand, dtype=np.uint8).reshape(1, -1))[1]:
            nonres = cand
            break
    bad = {
        "wrong type sig1": mut(0, 2), "wrong type k": mut(2 * T1, 1), "wrong type c": mut(at_c, 1), "wrong type rs": mut(at_rs, 7),
        "wrong type attributes": mut(at_str, 6), "wrong type E1": mut(end_attrs, 3),
        "length byte 254": mut(1, 254), "length byte 255": mut(2 * T1 + 1, 255), "short G1 payload": mut(1, FP_BYTES - 1),
        "short Fr payload": mut(at_c + 1, 31), "scalar >= r": msg[:at_c + 2] + big_r + msg[at_c + 34:],
        "response >= r": msg[:at_rs + 3] + big_r + msg[at_rs + 35:], "attribute count": mut(at_str + 1, n - 1),
        "too many responses": mut(at_rs + 1, n + 3), "attribute length past the end": mut(at_str + 2, 250),
        "sigma1 not on the curve": msg[:2] + nonres + msg[2 + FP_BYTES:],
        "phi x >= p": msg[:2 * T1 + T2 + 2] + b"\xff" * FP_BYTES + msg[2 * T1 + T2 + 2 + FP_BYTES:],
        "trailing garbage instead of E1": msg[:end_attrs] + b"\x01",
    }
    outcome = {name: _reference_outcome(ref, wl, m) for name, m in bad.items()}
    print(outcome)
    assert "accept" not in outcome.values(), outcome           # the reference never accepts one of these ...
    assert {"reject", "throws"} & set(outcome.values())        # ... it returns false, throws, or dies (memory-unsafe parser)
    for name, m in bad.items():
        assert not _parse(hostsim, 0, n, m)["ok"], name
    # a LONGER payload than the element is accepted like the reference accepts it (mcl reads the leading bytes)
    longer = msg[:1] + bytes([FP_BYTES + 3]) + msg[2:2 + FP_BYTES] + b"\x01\x02\x03" + msg[2 + FP_BYTES:]
    got = _parse(hostsim, 0, n, longer)
    assert got["ok"] and np.array_equal(got["g1"][0], ref.g1_op(ref.G_NORM, wl.proof["sig1"])[0])
    blob2 = np.frombuffer(longer + b"\0" * 8, dtype=np.uint8).copy()
    rv2, st2 = ref.verify_id_wire(wl.key, (blob2, np.array([0, len(longer)], dtype=np.uint64)), wl.ads, wl.service, wl.y, wl.g, wl.h)
    assert rv2[0] == 1 and st2[0] == 0


@pytest.mark.parametrize("b64", [False, True])
def test_request_parse_matches_reference(hostsim, ref, b64):
    n, lanes = 4, 4
    wl = workload.make_issuance_workload(n, lanes, 2, seed=43)
    msgs = _lanes(ref.request_encode(wl.key, wl.A, wl.c, wl.rs, wl.req_attrs, base64=b64))
    for j, m in enumerate(msgs):
        got = _parse(hostsim, 1, n, m, b64)
        assert got["ok"] and got["per"] == wl.rs.shape[1]
        assert np.array_equal(got["g1"][0], ref.g1_op(ref.G_NORM, wl.A[j:j + 1])[0])
        assert np.array_equal(got["c"], wl.c[j]) and np.array_equal(got["rs"], wl.rs[j])
        assert got["attrs"] == wl.req_attrs[j]
    raw = _lanes(ref.request_encode(wl.key, wl.A, wl.c, wl.rs, wl.req_attrs))[0]
    assert all(not _parse(hostsim, 1, n, raw[:cut])["ok"] for cut in range(len(raw)))
    assert _parse(hostsim, 1, n, raw + b"tail")["ok"]           # fromBufferString ignores bytes after the attribute list


@pytest.mark.parametrize("n_attrs,na", [(5, 0), (5, 5), (5, 1), (5, 3), (2, 2)])
def test_sign_lanes(hostsim, ref, n_attrs, na):
    """sign_commitment (na = 0) and sign_hybrid: a ONE-entry list is signed as a bare commitment (src/ps-signer.cc:114-116),
    "" entries are skipped, shorter lists than the key use the first bases"""
    lanes = 4
    key = ref.KeyMaterial(n_attrs, seed_=1)
    ref.seed(51)
    Cm = ref.g1_mul(key.g, ref.fr_rand(lanes))
    u = ref.fr_rand(lanes)
    attrs = None
    if na:
        attrs = [[b"" if (i + j) % 3 == 0 else b"v%d.%d" % (i, j) for i in range(na)] for j in range(lanes)]
    e1, e2, _ = ref.sign(key, Cm, attrs, u)
    blob, off = ref.pack_attrs(attrs) if na else (np.zeros(8, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    s1 = np.zeros((lanes, G1W), dtype=np.uint64)
    s2 = np.zeros((lanes, G1W), dtype=np.uint64)
    hostsim.hostsim_sign(C.c_int(n_attrs), C.c_int(na), C.c_int(4), _p(key.g), _p(key.X), _p(key.Y), C.c_size_t(lanes), _p(Cm),
                         _p(blob), _p(off), _p(u), _p(s1), _p(s2))
    assert np.array_equal(s1, ref.g1_op(ref.G_NORM, e1)) and np.array_equal(s2, ref.g1_op(ref.G_NORM, e2))


@pytest.mark.parametrize("with_id,b64", [(True, False), (True, True), (False, False), (False, True)])
def test_encoder_matches_reference(hostsim, ref, with_id, b64):
    """csrc/wire.cuh encode_message_lane / base64_encode_lane against IdProof::toBufferString, PSCredRequest::toBufferString and
    PSBuffer::toBase64 (src/ps-encoding.cc:14-54, :429-439, :452-468), byte for byte; the reference's proofs hold raw Jacobian
    points, so the normalisation inside the encoder is exercised too."""
    hostsim.hostsim_wire_encode.restype = C.c_size_t
    n, lanes = 5, 4
    wl = workload.make_signon_workload(n, lanes, 2, seed=43, with_id=with_id)
    wl.proof_attrs[1][3] = b"y" * 300                       # 253-prefixed length
    wl.proof_attrs[2][4] = b""
    want = _lanes(ref.idproof_encode(wl.key, wl.proof, wl.proof_attrs, with_e=with_id, base64=b64))
    per = wl.proof["rs"].shape[1]
    for j in range(lanes):
        g1 = np.stack([wl.proof[k][j] for k in ("sig1", "sig2", "phi", "E1", "E2")])
        blob, off = ref.pack_attrs([wl.proof_attrs[j]])
        out = np.zeros(4096, dtype=np.uint8)
        k = hostsim.hostsim_wire_encode(C.c_int(0), C.c_int(n), _p(g1), _p(wl.proof["k"][j]), C.c_int(int(with_id)), _p(wl.proof["c"][j]),
                                        _p(wl.proof["rs"][j]), C.c_int(per), _p(blob), _p(off), C.c_int(int(b64)), _p(out))
        assert bytes(out[:k]) == want[j]
    iw = workload.make_issuance_workload(n, lanes, 2, seed=44)
    iw.req_attrs[0][2] = b"z" * 260
    want = _lanes(ref.request_encode(iw.key, iw.A, iw.c, iw.rs, iw.req_attrs, base64=b64))
    for j in range(lanes):
        blob, off = ref.pack_attrs([iw.req_attrs[j]])
        out = np.zeros(4096, dtype=np.uint8)
        k = hostsim.hostsim_wire_encode(C.c_int(1), C.c_int(n), _p(iw.A[j]), None, C.c_int(0), _p(iw.c[j]), _p(iw.rs[j]),
                                        C.c_int(iw.rs.shape[1]), _p(blob), _p(off), C.c_int(int(b64)), _p(out))
        assert bytes(out[:k]) == want[j]
