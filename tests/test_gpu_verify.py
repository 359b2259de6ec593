"""GPU parity of the batched PS verification (psb_verify through the C ABI) against the reference's
own PSVerifier::verify (oracle/_ref), incl. tampered lanes, edge cases, both attribute input forms,
and the fused-lane GT bytes."""
import numpy as np
import pytest

from tests import workload
from tests.conftest import FPW, bls_only

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_attrs,lanes,w", [(5, 130, 8), (5, 64, 16), (1, 33, 4), (3, 40, 11), (0, 8, 8), (2, 40, 22), (1, 33, 24)])
def test_verify_matches_reference(gpu_pkg, ref, n_attrs, lanes, w):
    wl = workload.make_verify_workload(n_attrs=n_attrs, lanes=lanes, seed=11, tamper_every=5)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=w)
    exp_v, exp_gt = workload.expected_verify(wl, want_gt=True)
    got_v, got_gt = gpu_pkg.PSVerifier(pk).verify(wl.sig1, wl.sig2, wl.attrs, want_gt=True)
    assert np.array_equal(got_v, exp_v)
    assert exp_v.sum() == lanes - len(wl.tampered)  # honest lanes accept, tampered lanes reject
    live = wl.sig1[:, 2 * FPW:].any(axis=1)  # reference returns before pairing when sig1 == 0
    assert np.array_equal(got_gt[live], exp_gt[live])
    # same lanes, scalars supplied by the host instead of attribute strings
    if n_attrs:
        m = ref.fr_set_hash_of_batch([a for lane in wl.attrs for a in lane])
        assert np.array_equal(gpu_pkg.PSVerifier(pk).verify(wl.sig1, wl.sig2, scalars=m), exp_v)
    pk.close()


def test_verify_unnormalized_inputs(gpu_pkg, ref):
    """credentials and key points with arbitrary Jacobian z (straight out of G1::mul) verify the same."""
    wl = workload.make_verify_workload(n_attrs=2, lanes=16, seed=3)
    ref.seed(77)
    t = ref.fr_rand(16)
    s1, s2, _ = ref.randomize(wl.sig1, wl.sig2, t)  # raw Jacobian outputs of mcl, z != 1
    assert (s1[:, 2 * FPW:] != wl.sig1[:, 2 * FPW:]).any()
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
    exp = ref.ps_verify(wl.key, s1, s2, wl.attrs)
    assert exp.all()
    assert np.array_equal(gpu_pkg.PSVerifier(pk).verify(s1, s2, wl.attrs), exp)


def test_verify_edge_lanes(gpu_pkg, ref):
    """sigma1 = 0 rejects; sigma2 = 0 with valid sigma1 rejects; empty attribute strings hash fine;
    N = 0 and N = 1 work."""
    wl = workload.make_verify_workload(n_attrs=2, lanes=6, seed=4)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
    v = gpu_pkg.PSVerifier(pk)
    wl.sig1[0] = 0
    wl.sig2[1] = 0
    wl.attrs[2][1] = b""
    wl.attrs[3][0] = b"x" * 200  # multi-block SHA-256
    exp = ref.ps_verify(wl.key, wl.sig1, wl.sig2, wl.attrs)
    assert np.array_equal(v.verify(wl.sig1, wl.sig2, wl.attrs), exp)
    assert list(exp) == [0, 0, 0, 0, 1, 1]
    assert v.verify(wl.sig1[:0], wl.sig2[:0], []).shape == (0,)
    assert np.array_equal(v.verify(wl.sig1[4:5], wl.sig2[4:5], wl.attrs[4:5]), exp[4:5])
    with pytest.raises(ValueError):
        v.verify(wl.sig1, wl.sig2, [a[:1] for a in wl.attrs])  # attribute size does not match


def test_verify_large_batch_properties(gpu_pkg, ref):
    """2^13 lanes: every honest lane accepts, every tampered lane rejects, and a sampled subset agrees
    with the reference lane by lane (full-size property check)."""
    lanes = 1 << 13
    wl = workload.make_verify_workload(n_attrs=5, lanes=lanes, seed=21, tamper_every=64)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY)
    got = gpu_pkg.PSVerifier(pk).verify(wl.sig1, wl.sig2, (wl.blob, wl.off))
    mask = np.ones(lanes, dtype=bool)
    mask[wl.tampered] = False
    assert got[mask].all() and not got[~mask].any()
    idx = np.concatenate([np.arange(0, lanes, 37), wl.tampered])
    sub = ref.ps_verify(wl.key, wl.sig1[idx], wl.sig2[idx], [wl.attrs[i] for i in idx], nthreads=ref.hw_threads())
    assert np.array_equal(got[idx], sub)


def test_attribute_hash_padding_boundaries(gpu_pkg, ref):
    """device SHA-256 -> Fr (Fr::setHashOf rule) at every padding boundary: attribute lengths around 55/56, 63/64/65,
    119/120 bytes, and lanes whose digest exceeds r after the 255-bit mask (second mask to 254 bits)."""
    lens = [0, 1, 54, 55, 56, 57, 63, 64, 65, 118, 119, 120, 121, 128, 191, 192, 300]
    key = ref.KeyMaterial(2, seed_=1)
    attrs = [[bytes([65 + (i % 26)]) * L, b"tail%d" % i] for i, L in enumerate(lens)]
    # hunt a few strings whose masked digest is >= r (probability ~ 1/10 each): exercises the 254-bit remask
    import hashlib
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    j = 0
    while sum(1 for a in attrs if a[1].startswith(b"big")) < 3:
        s = b"big%d" % j
        j += 1
        if int.from_bytes(hashlib.sha256(s).digest(), "little") & ((1 << 255) - 1) >= R:
            attrs.append([b"x", s])
    sig1, sig2 = workload.sign_lanes(key, attrs, seed=31)
    pk = gpu_pkg.PSPubKey(key.g, key.gg, key.XX, key.Y, key.YY, window_bits=8)
    exp = ref.ps_verify(key, sig1, sig2, attrs)
    assert exp.all()
    assert np.array_equal(gpu_pkg.PSVerifier(pk).verify(sig1, sig2, attrs), exp)
    # and the same signatures against attribute strings one byte longer must all fail
    bad = [[a[0] + b"!", a[1]] for a in attrs]
    assert not gpu_pkg.PSVerifier(pk).verify(sig1, sig2, bad).any()
    pk.close()


def test_randomize_verify_round_trip_large(gpu_pkg, ref):
    """size-independent property at 2^14 lanes: verify(randomize(sigma, t)) == verify(sigma) for every lane
    (honest and tampered), with per-lane random t; t = 0 sends a credential to (0, 0), which verify rejects."""
    lanes = 1 << 14
    wl = workload.make_verify_workload(n_attrs=3, lanes=lanes, seed=23, tamper_every=97)
    pk = gpu_pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=12)
    ver = gpu_pkg.PSVerifier(pk)
    base = ver.verify(wl.sig1, wl.sig2, (wl.blob, wl.off))
    ref.seed(99)
    t = ref.fr_rand(lanes)
    t[5] = 0
    r1, r2 = gpu_pkg.PSRequester.randomize_credential(wl.sig1, wl.sig2, t)
    again = ver.verify(r1, r2, (wl.blob, wl.off))
    exp = base.copy()
    exp[5] = 0
    assert np.array_equal(again, exp)
    assert not r1[5].any() and not r2[5].any()
    pk.close()


@bls_only
def test_full_size_2p20_properties(gpu_pkg, ref):
    """BASELINE.json configs[1] at its full size -- 2^20 signatures, 5 attributes -- through size-independent properties
    (the oracle needs 73 core-minutes for this batch): (i) the verdict bitmap equals the construction (honest lanes
    accept, the 1024 tampered lanes -- swapped, sigma1 = 0, attribute changed, sigma2 += g -- reject); (ii) idempotence: a second
    pass gives the same bitmap; (iii) verify(randomize(sigma, t)) == verify(sigma) lane by lane; (iv) a uniformly drawn
    sample of 512 lanes + every tampered lane of the first 2^16 agrees with the reference's PSVerifier::verify."""
    import bench
    lanes = 1 << 20
    key = bench.load_key(5)
    sig1, sig2, blob, off, expected = bench.make_batch(gpu_pkg, key, lanes, 0)
    pk = gpu_pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=16)
    ver = gpu_pkg.PSVerifier(pk)
    v1 = ver.verify(sig1, sig2, (blob, off))
    assert np.array_equal(v1, expected) and int(v1.sum()) == lanes - lanes // 1024
    assert np.array_equal(ver.verify(sig1, sig2, (blob, off)), v1)
    rng = np.random.default_rng(5)
    t = np.frombuffer(rng.bytes(32 * lanes), dtype=np.uint64).reshape(lanes, 4).copy()
    t[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    t[:, 0] |= np.uint64(1)
    r1, r2 = gpu_pkg.PSRequester.randomize_credential(sig1, sig2, t)
    assert np.array_equal(ver.verify(r1, r2, (blob, off)), v1)
    pick = np.unique(np.concatenate([rng.integers(0, lanes, 512), np.arange(1023, 1 << 16, 1024)]))
    km = ref.KeyMaterial(5, seed_=1)             # the key of tests/golden/keys.json (bench.load_key)
    assert np.array_equal(km.XX.reshape(-1), key["XX"].reshape(-1))
    want = ref.ps_verify(km, sig1[pick].copy(), sig2[pick].copy(), [bench.lane_strings(blob, off, j, 5) for j in pick], nthreads=ref.hw_threads())
    assert np.array_equal(v1[pick], want)
    pk.close()


@bls_only
def test_wave_boundaries(gpu_pkg, ref):
    """Launch shape (psb_api.cu, for_waves): whole waves of 512-thread blocks, then the ragged rest as one thinner block
    per SM with a lane offset.  Batches of one wave -1 / +0 / +1 lanes and of two waves + 33 (a rest of a single warp on
    33 SMs): the verdict bitmap equals the construction, the lanes either side of every launch boundary agree with the
    reference's PSVerifier::verify, and psb_pairing returns the same GT for a tiled input on both sides of the seam."""
    import torch
    import bench
    wave = torch.cuda.get_device_properties(0).multi_processor_count * 512
    key = bench.load_key(5)
    pk = gpu_pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=12)
    ver = gpu_pkg.PSVerifier(pk)
    km = ref.KeyMaterial(5, seed_=1)
    for lanes in (wave - 1, wave, wave + 1, 2 * wave + 33):
        sig1, sig2, blob, off, expected = bench.make_batch(gpu_pkg, key, lanes, 0, base=2048)
        got = ver.verify(sig1, sig2, (blob, off))
        assert np.array_equal(got, expected), lanes
        full = lanes // wave * wave
        pick = np.unique(np.clip(np.array([0, full - 2, full - 1, full, full + 1, lanes - 2, lanes - 1, 1023]), 0, lanes - 1))
        want = ref.ps_verify(km, sig1[pick].copy(), sig2[pick].copy(), [bench.lane_strings(blob, off, j, 5) for j in pick])
        assert np.array_equal(got[pick], want), lanes
    pk.close()
    ref.seed(7)
    k = ref.fr_rand(8)
    P, Q = ref.g1_mul(ref.hash_to_g1(b"abc"), k), ref.g2_mul(ref.hash_to_g2(b"edf"), k)
    lanes = wave + 40
    gt = gpu_pkg.pairing(np.tile(P, (lanes // 8 + 1, 1))[:lanes].copy(), np.tile(Q, (lanes // 8 + 1, 1))[:lanes].copy())
    exp = ref.pairing(P, Q)
    assert np.array_equal(gt[:8], exp) and np.array_equal(gt[wave - 8:wave], exp) and np.array_equal(gt[wave:wave + 8], exp)
    assert np.array_equal(gt[lanes - 8:], exp)


def test_batched_affine_sum_exceptional_pairs(gpu_pkg, ref):
    """The fixed-base sum of k_verify_msm adds table entries pairwise in AFFINE coordinates (AffBatch, csrc/curve.cuh); pairs the
    affine formula cannot take -- P = Q, P = -Q, an absent entry (zero digit) -- must come out as the same group element.
    Key with YY_1 = 2^252 YY_0 and YY_3 = -2^252 YY_2 and w = 7 (37 windows: the top window of one base pairs with window 0
    of the next), host-supplied scalars whose digits meet at those seams; lanes with small scalars (mostly absent entries)
    and ordinary random lanes share the warps.  Expected: K from the reference's G2 arithmetic, GT from its pairings,
    verdict 1 exactly on the lanes signed under the equivalent exponents."""
    from tests.conftest import GROUP_R
    rng = np.random.default_rng(21)
    km = ref.KeyMaterial(2, seed_=5)
    shift = (1 << 252) % GROUP_R
    sh = ref.fr_from_ints([shift])
    YY = np.stack([km.YY[0], ref.g2_mul(km.YY[0], sh)[0], km.YY[1], ref.g2_op(ref.G_NEG, ref.g2_mul(km.YY[1], sh))[0]])
    YY = ref.g2_op(ref.G_NORM, YY)
    Y = np.stack([km.Y[0], km.Y[0], km.Y[1], km.Y[1]])          # G1 side of the key: unused by verify
    y0, y1 = ref.fr_to_ints(km.y)
    x = ref.fr_to_ints(km.x.reshape(1, -1))[0]
    ya = [y0, shift * y0 % GROUP_R, y1, (GROUP_R - shift) * y1 % GROUP_R]
    lanes = 96
    low = lambda: int.from_bytes(rng.bytes(12), "little") << 7  # noqa: E731
    ms = []
    for j in range(lanes):
        kind = j % 4
        if kind == 0:      # P = Q at the 0/1 seam, P = -Q at the 2/3 seam
            ms.append([(2 << 252) + low(), 2 + low(), (2 << 252) + low(), 2 + low()])
        elif kind == 1:    # mostly absent entries
            ms.append([5, 0, 1 << 200, 77])
        else:              # ordinary scalars
            ms.append([int.from_bytes(rng.bytes(40), "little") % GROUP_R for _ in range(4)])
    m = ref.fr_from_ints([v for lane in ms for v in lane])
    ref.seed(9)
    h = ref.g1_mul(km.g, ref.fr_rand(lanes))
    e = [(x + sum(a * b for a, b in zip(lane, ya))) % GROUP_R for lane in ms]
    sig2 = ref.g1_mul(h, ref.fr_from_ints(e))
    bad = np.arange(lanes) % 5 == 4
    sig2[bad] = ref.g1_op(ref.G_ADD, sig2[bad], np.broadcast_to(km.g, (int(bad.sum()), km.g.size)).copy())
    sig1 = ref.g1_op(ref.G_NORM, h)
    sig2 = ref.g1_op(ref.G_NORM, sig2)
    K = np.broadcast_to(km.XX, (lanes, km.XX.size)).copy()
    mm = m.reshape(lanes, 4, -1)
    for i in range(4):
        K = ref.g2_op(ref.G_ADD, K, ref.g2_mul(np.broadcast_to(YY[i], (lanes, YY[i].size)).copy(), mm[:, i]))
    exp_gt = ref.pairing_ratio(sig1, K, sig2, np.broadcast_to(km.gg, (lanes, km.gg.size)).copy())
    for w in (7, 8):       # w = 8: 32 windows per base, no seam pairs -- the same lanes through ordinary pairs
        pk = gpu_pkg.PSPubKey(km.g, km.gg, km.XX, Y, YY, window_bits=w)
        got_v, got_gt = gpu_pkg.PSVerifier(pk).verify(sig1, sig2, scalars=m, want_gt=True)
        assert np.array_equal(got_gt, exp_gt), w
        assert np.array_equal(got_v.astype(bool), ~bad), w
        pk.close()
