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
