"""CPU: the EL PASSO lane functions of csrc/protocol.cuh (issuance, sign-on verification), compiled for the
host by tests/hostsim, against the reference's own PSSigner::el_passo_provide_id and
PSVerifier::el_passo_verify_id -- the same code the GPU kernels wrap, minus the PTX carry chains."""
import ctypes as C

import numpy as np
import pytest

from tests import workload
from tests.conftest import G1W, G2W


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


def test_fixed_base_table_mul(hostsim, ref):
    ref.seed(21)
    k = ref.fr_rand(3)
    k[2] = ref.fr_from_ints([0])[0]
    g = ref.hash_to_g1(b"abc")
    for w in (4, 5):
        for i in range(3):
            out = np.zeros(G1W, dtype=np.uint64)
            hostsim.hostsim_fixed_mul_g1(C.c_int(w), _p(g), _p(k[i]), _p(out))
            assert np.array_equal(out, ref.g1_op(ref.G_NORM, ref.g1_mul(g, k[i:i + 1]))[0]), (w, i)


@pytest.mark.parametrize("n_attrs,n_hidden", [(5, 2), (1, 1), (1, 0), (3, 3)])
def test_provide_id_lanes(hostsim, ref, n_attrs, n_hidden):
    lanes = 6
    wl = workload.make_issuance_workload(n_attrs, lanes, n_hidden, seed=4, tamper_every=2)
    ev, e1, e2, _ = workload.expected_provide_id(wl)
    blob, off = ref.pack_attrs(wl.req_attrs)
    ad_blob, ad_off = ref.pack_strings(wl.ads)
    verdict = np.zeros(lanes, dtype=np.uint8)
    s1 = np.zeros((lanes, G1W), dtype=np.uint64)
    s2 = np.zeros((lanes, G1W), dtype=np.uint64)
    hostsim.hostsim_provide_id(C.c_int(n_attrs), C.c_int(4), _p(wl.key.g), _p(wl.key.X), _p(wl.key.Y), C.c_size_t(lanes),
                               _p(wl.A), _p(wl.c), _p(wl.rs), C.c_int(wl.rs.shape[1]), _p(blob), _p(off), _p(ad_blob),
                               _p(ad_off), _p(wl.u), _p(verdict), _p(s1), _p(s2))
    assert verdict.tolist() == ev.tolist()
    assert ev.sum() == lanes - len(wl.tampered)
    ok = ev.astype(bool)
    assert np.array_equal(s1[ok], ref.g1_op(ref.G_NORM, e1[ok]))
    assert np.array_equal(s2[ok], ref.g1_op(ref.G_NORM, e2[ok]))
    assert not s1[~ok].any() and not s2[~ok].any()


@pytest.mark.parametrize("n_attrs,n_hidden,with_id", [(5, 2, True), (4, 2, False), (2, 2, True)])
def test_verify_id_lanes(hostsim, ref, n_attrs, n_hidden, with_id):
    lanes = 7
    wl = workload.make_signon_workload(n_attrs, lanes, n_hidden, seed=3, with_id=with_id, tamper_every=1 if n_attrs == 5 else 3)
    ev = workload.expected_verify_id(wl)
    p = wl.proof
    blob, off = ref.pack_attrs(wl.proof_attrs)
    ad_blob, ad_off = ref.pack_strings(wl.ads)
    nizk = np.zeros(lanes, dtype=np.uint8)
    verdict = np.zeros(lanes, dtype=np.uint8)
    hostsim.hostsim_verify_id(C.c_int(n_attrs), C.c_int(4), _p(wl.key.gg), _p(wl.key.XX), _p(wl.key.YY), C.c_size_t(lanes),
                              _p(p["sig1"]), _p(p["sig2"]), _p(p["k"]), _p(p["phi"]), _p(p["E1"]), _p(p["E2"]), _p(p["c"]),
                              _p(p["rs"]), C.c_int(p["rs"].shape[1]), _p(blob), _p(off), _p(ad_blob), _p(ad_off),
                              _p(wl.service_pt), _p(wl.y), _p(wl.g), _p(wl.h), C.c_int(int(with_id)), _p(nizk), _p(verdict))
    assert verdict.tolist() == ev.tolist(), (nizk.tolist(), wl.tampered.tolist())
    if n_attrs != 5:
        assert ev.sum() == lanes - len(wl.tampered)


@pytest.mark.parametrize("g2", [False, True])
def test_point_decompression_lanes(hostsim, ref, g2):
    enc, want, okv = workload.deserialize_cases(g2)
    assert 0 < okv.sum() < len(okv)
    fn = hostsim.hostsim_g2_deserialize if g2 else hostsim.hostsim_g1_deserialize
    for j in range(len(okv)):
        out = np.zeros(G2W if g2 else G1W, dtype=np.uint64)
        r = fn(_p(enc[j]), _p(out))
        assert r == okv[j], j
        if okv[j]:
            assert np.array_equal(out, want[j]), j


@pytest.mark.parametrize("is_g2", [0, 1])
def test_fixed_base_sum_batched_affine(hostsim, ref, is_g2):
    """Fixed-base sums through the batched affine pair additions (AffBatch in csrc/curve.cuh, the path k_verify_msm runs)
    against the plain mixed-addition chain AND against the reference's G::mul / add: random scalars, absent entries (zero
    digits), a single small batch (below the inversion's break-even), and the two pairs the affine formula cannot take --
    P = Q and P = -Q -- built by giving base 1 = 2^252 base 0 (w = 7: 37 windows, so the top window of base 0 pairs with
    window 0 of base 1) and base 3 = -2^252 base 2."""
    from tests.conftest import GROUP_R
    W = G2W if is_g2 else G1W
    mul, op = (ref.g2_mul, ref.g2_op) if is_g2 else (ref.g1_mul, ref.g1_op)
    rng = np.random.default_rng(5 + is_g2)
    B0 = ref.hash_to_g2(b"base-0") if is_g2 else ref.hash_to_g1(b"base-0")
    B2 = ref.hash_to_g2(b"base-2") if is_g2 else ref.hash_to_g1(b"base-2")
    shift = ref.fr_from_ints([(1 << 252) % GROUP_R])
    B1 = mul(B0, shift)[0]
    B3 = op(ref.G_NEG, mul(B2, shift))[0]
    acc0 = mul(B2, ref.fr_from_ints([12345]))[0]

    def run(w, bases, ks):
        bases = op(ref.G_NORM, np.stack(bases))
        km = ref.fr_from_ints(ks)
        exp = acc0.reshape(1, -1)
        for b, k in zip(bases, km):
            exp = op(ref.G_ADD, exp, mul(b, k.reshape(1, -1)))
        exp = op(ref.G_NORM, exp)[0]
        for levels in (0, 1, 2):  # pair sums straight into the accumulator / paired up once more (AffPts) / table declared too large for the slot index: plain chain
            o1 = np.zeros(W, dtype=np.uint64)
            o2 = np.zeros(W, dtype=np.uint64)
            hostsim.hostsim_fixed_msm(C.c_int(is_g2), C.c_int(w), C.c_int(len(ks)), _p(bases), _p(km), _p(acc0), _p(o1), _p(o2),
                                      C.c_int(levels))
            assert np.array_equal(o1, exp), ("plain", w)
            assert np.array_equal(o2, exp), ("batched affine", w, levels)

    low = lambda: int.from_bytes(rng.bytes(12), "little")  # noqa: E731
    top = 2 << 252
    # P = Q at the seam of bases 0/1, P = -Q at the seam of bases 2/3 (sum = infinity), random scalars elsewhere
    run(7, [B0, B1, B2, B3], [top + (low() << 7), 2 + (low() << 7), top + (low() << 7), 2 + (low() << 7)])
    # zero digits: small scalars leave most windows empty; k = 0 leaves a whole base empty
    run(5, [B0, B2, B1], [5, 0, (1 << 200) + 77])
    # random full-size scalars, two flushes (3 x 64 slots = 96 pairs -> 2 batches of 48)
    run(4, [B0, B2, B3], [int.from_bytes(rng.bytes(40), "little") % GROUP_R for _ in range(3)])
    run(3, [B0, B2, B1, B3], [int.from_bytes(rng.bytes(40), "little") % GROUP_R for _ in range(4)])   # 344 slots: 3 + 3 flushes
    # one short batch (8 pairs: plain additions), and an even slot count in one flush (w = 5: 52 windows)
    run(16, [B0], [int.from_bytes(rng.bytes(40), "little") % GROUP_R])
    run(5, [B2], [int.from_bytes(rng.bytes(40), "little") % GROUP_R])


@pytest.mark.parametrize("is_g2", [0, 1])
def test_batched_affine_second_level(hostsim, ref, is_g2):
    """AffPts (the second level of the batched affine sums) on explicit points: pairs P + P, P + (-P), a point beside an empty
    slot, two empty slots, a flush that fills up in the middle (70 points: 64 + 6) and a short one (plain additions)."""
    from tests.conftest import GROUP_R
    W = G2W if is_g2 else G1W
    mul, op = (ref.g2_mul, ref.g2_op) if is_g2 else (ref.g1_mul, ref.g1_op)
    rng = np.random.default_rng(11 + is_g2)
    B = ref.hash_to_g2(b"l2") if is_g2 else ref.hash_to_g1(b"l2")
    acc0 = mul(B, ref.fr_from_ints([777]))[0]

    def rand_pts(n):
        return op(ref.G_NORM, mul(B, ref.fr_from_ints([int.from_bytes(rng.bytes(40), "little") % GROUP_R for _ in range(n)])))

    for total in (70, 24, 7):
        pts = rand_pts(total)
        present = np.ones(total, dtype=np.uint8)
        pts[1] = pts[0]                                   # P + P
        pts[3] = op(ref.G_NEG, pts[2:3])[0]               # P + (-P)
        present[5] = 0                                    # a point beside an empty slot
        if total > 8:
            present[6] = present[7] = 0                   # two empty slots
        out = np.zeros(W, dtype=np.uint64)
        hostsim.hostsim_aff_l2(C.c_int(is_g2), C.c_int(total), _p(pts), _p(present), _p(acc0), _p(out))
        exp = acc0.reshape(1, -1)
        for i in range(total):
            if present[i]:
                exp = op(ref.G_ADD, exp, pts[i:i + 1])
        assert np.array_equal(out, op(ref.G_NORM, exp)[0]), total
