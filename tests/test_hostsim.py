"""CPU: the engine's __host__ __device__ arithmetic (tower, curves, pairing, SHA-256), compiled for
the host by tests/hostsim, against the reference -- checks the LOGIC the GPU kernels share.  The
PTX carry chains themselves are only exercised by the -m gpu tests."""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import BN254, CURVE_Z, FIELD_P, FP_BYTES, GROUP_R, bls_only, rand_fp_raw


def hop(hs, op, a, b=None, c=None):
    s = (C.c_int * 4)()
    assert hs.hostsim_op_shape(op, s) == 0
    a32 = np.ascontiguousarray(a).view(np.uint32).reshape(-1, s[0])
    n = a32.shape[0]
    b32 = None if b is None or not s[1] else np.ascontiguousarray(b).view(np.uint32).reshape(n, s[1])
    c32 = None if c is None or not s[2] else np.ascontiguousarray(c).view(np.uint32).reshape(n, s[2])
    out = np.zeros((n, s[3]), dtype=np.uint32)
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert hs.hostsim_op(op, C.c_size_t(n), p(a32), p(b32), p(c32), p(out)) == 0
    return out.view(np.uint64)


@pytest.fixture(scope="module")
def rng():
    return np.random.default_rng(7)


def test_fields(hostsim, ref, rng):
    a, b = rand_fp_raw(ref, rng, 32), rand_fp_raw(ref, rng, 32)
    b[1] = 0
    for op in range(6):
        assert np.array_equal(hop(hostsim, op, a, b if op < 3 else None), ref.fp_op(op, a, b)), op
    a, b = rand_fp_raw(ref, rng, 16, 2), rand_fp_raw(ref, rng, 16, 2)
    for op in range(6):
        assert np.array_equal(hop(hostsim, 10 + op, a, b if op < 3 else None), ref.fp2_op(op, a, b)), op
    a, b = rand_fp_raw(ref, rng, 4, 6), rand_fp_raw(ref, rng, 4, 6)
    assert np.array_equal(hop(hostsim, 20, a, b), ref.fp6_op(ref.OP_MUL, a, b))
    assert np.array_equal(hop(hostsim, 21, a), ref.fp6_op(ref.OP_INV, a))
    a, b = rand_fp_raw(ref, rng, 3, 12), rand_fp_raw(ref, rng, 3, 12)
    assert np.array_equal(hop(hostsim, 30, a, b), ref.fp12_op(ref.OP_MUL, a, b))
    assert np.array_equal(hop(hostsim, 31, a), ref.fp12_op(ref.OP_SQR, a))
    assert np.array_equal(hop(hostsim, 32, a), ref.fp12_op(ref.OP_INV, a))
    for k in (1, 2, 3):
        assert np.array_equal(hop(hostsim, 32 + k, a), ref.fp12_frobenius(k, a))


def test_fp_inv_divsteps_edges(hostsim, ref, rng):
    """Fp::inv through the Bernstein-Yang divstep iteration (csrc/modinv.cuh) on raw limb patterns that stress the
    iteration count and the sign handling: 0, 1, p-1, powers of two (long runs of even g), p - 2^k, all-ones limbs, and
    random values -- against mcl's Fp::inv (ext-gcd), raw Montgomery limbs compared bit for bit."""
    P = FIELD_P
    bits = P.bit_length()
    raw = [0, 1, 2, 3, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, (1 << (bits - 1)) - 1, 1 << (bits - 1)]
    raw += [1 << k for k in range(1, bits - 1, 7)] + [P - (1 << k) for k in range(1, bits - 1, 11)]
    raw += [((1 << bits) - 1) % P, (0x5555555555555555 * (1 << 320 | 1 << 256 | 1 << 192 | 1 << 128 | 1 << 64 | 1)) % P]
    raw += [int.from_bytes(rng.bytes(FP_BYTES), "little") % P for _ in range(300)]
    a = np.frombuffer(b"".join(v.to_bytes(FP_BYTES, "little") for v in raw), dtype=np.uint64).reshape(len(raw), -1).copy()
    got = hop(hostsim, 5, a)
    assert not got[0].any()          # inv(0) = 0 (the reference's Vint ext-gcd divides by zero on 0: not asked)
    assert np.array_equal(got[1:], ref.fp_op(5, a[1:]))


def test_final_exp_compressed_pow_z_and_fallback(hostsim, ref, rng):
    """finalExp with pow_z on Karabina's compressed squarings (csrc/pairing.cuh): generic Fp12 inputs, and inputs whose
    cyclotomic image is 1 (the value 1 itself, elements of Fp2 and Fp6: a zero denominator in the decompression, which
    must take the Granger-Scott fallback) -- GT bytes against mcl's finalExp."""
    a = rand_fp_raw(ref, rng, 6, 12)
    one = ref.fp_from_ints([1]).reshape(-1)
    fp = a.shape[1] // 12
    a[0] = 0; a[0, :fp] = one                 # 1
    a[1, 2 * fp:] = 0                         # an element of Fp2
    a[2, 6 * fp:] = 0                         # an element of Fp6 (b = 0)
    assert np.array_equal(hop(hostsim, 61, a), ref.final_exp(a))
    assert np.array_equal(hop(hostsim, 61, a[:3]), np.repeat(a[0:1], 3, axis=0))


def test_groups_and_pairing(hostsim, ref):
    ref.seed(31)
    g, gg = ref.hash_to_g1(b"abc"), ref.hash_to_g2(b"edf")
    k1, k2, k3 = ref.fr_rand(4), ref.fr_rand(4), ref.fr_rand(4)
    P, P2, Q, Q2 = ref.g1_mul(g, k1), ref.g1_mul(g, k2), ref.g2_mul(gg, k1), ref.g2_mul(gg, k2)
    s1, s2 = ref.g1_serialize, ref.g2_serialize
    assert np.array_equal(s1(hop(hostsim, 40, P, P2)), s1(ref.g1_op(ref.G_ADD, P, P2)))
    assert np.array_equal(s1(hop(hostsim, 43, P, k3)), s1(ref.g1_mul(P, k3)))
    assert np.array_equal(hop(hostsim, 42, P), ref.g1_op(ref.G_NORM, P))
    assert np.array_equal(s2(hop(hostsim, 50, Q, Q2)), s2(ref.g2_op(ref.G_ADD, Q, Q2)))
    assert np.array_equal(s2(hop(hostsim, 53, Q, k3)), s2(ref.g2_mul(Q, k3)))
    assert np.array_equal(hop(hostsim, 60, P[:2], Q[:2]), ref.pairing(P[:2], Q[:2]))
    Q2n = ref.g2_op(ref.G_NORM, Q2)
    c = np.concatenate([P2[:2], Q2n[:2]], axis=1)
    assert np.array_equal(hop(hostsim, 62, P[:2], Q[:2], c), ref.pairing_ratio(P[:2], Q[:2], P2[:2], Q2n[:2]))


def test_sha256_to_scalar(hostsim, ref):
    """Fr::setHashOf (SHA-256, mask to the bit length of r, one bit less if still >= r) against mcl itself and -- on
    BLS12-381 -- against the Python restatement."""
    msgs = (b"", b"attr0", b"x" * 55, b"y" * 56, b"z" * 64, b"w" * 200)
    want = ref.fr_to_ints(ref.fr_set_hash_of_batch(list(msgs)))
    for msg, w in zip(msgs, want):
        k = np.zeros(8, dtype=np.uint32)
        hostsim.hostsim_set_hash_of(msg, C.c_size_t(len(msg)), k.ctypes.data_as(C.c_void_p))
        assert int.from_bytes(k.tobytes(), "little") == w
        if not BN254:
            from oracle import ps_oracle as O
            assert w == O.fr_set_hash_of(msg)


def test_glv_gls_scalar_edges(hostsim, ref):
    """GLV (G1) / GLS (G2) variable-base multiplication at the decomposition boundaries: 0, 1, lambda-1, lambda,
    lambda+1, multiples of |z|, r-1, and random scalars -- against mcl's G1::mul / G2::mul (normalised)."""
    Z, R = CURVE_Z, GROUP_R
    lam = Z * Z - 1          # BLS12-381's GLV eigenvalue; on BN254 just another boundary-sized scalar
    ks = [0, 1, 2, 15, 16, lam - 1, lam, lam + 1, 2 * lam, lam * lam, lam * (lam + 1), Z - 1, Z, Z + 1, Z * Z, Z ** 3, Z ** 3 - 1,
          (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 1 << 128, R - 1, R - 2]
    if BN254:                # BN254's eigenvalues: lambda = 36 z^4 - 1 (G1), mu = 6 z^2 (G2), and the sizes of the lattice vectors
        lb, mu = 36 * Z ** 4 - 1, 6 * Z * Z
        ks += [lb - 1, lb, lb + 1, 2 * lb, mu - 1, mu, mu + 1, mu * mu, mu ** 3, mu ** 3 + mu, 6 * Z * Z + 2 * Z, 2 * Z + 1, (1 << 127) - 1,
               1 << 127, (1 << 253) + 1, R // 2, R // 2 + 1, R // 3, 2 * R // 3]
    ks = [k % R for k in ks]
    rng = np.random.default_rng(5)
    ks += [int.from_bytes(rng.bytes(32), "little") % R for _ in range(9 if not BN254 else 40)]
    k = ref.fr_from_ints(ks)
    g, gg = ref.hash_to_g1(b"abc"), ref.hash_to_g2(b"edf")
    P = np.repeat(g.reshape(1, -1), len(ks), axis=0)
    Q = np.repeat(gg.reshape(1, -1), len(ks), axis=0)
    norm1 = lambda a: hop(hostsim, 42, a)  # noqa: E731  T_G1_NORM
    norm2 = lambda a: hop(hostsim, 52, a)  # noqa: E731  T_G2_NORM
    assert np.array_equal(norm1(hop(hostsim, 43, P, k)), ref.g1_op(ref.G_NORM, ref.g1_mul(g, k)))
    assert np.array_equal(norm2(hop(hostsim, 53, Q, k)), ref.g2_op(ref.G_NORM, ref.g2_mul(gg, k)))
    # un-normalised input points (z != 1)
    P2 = ref.g1_op(ref.G_DBL, P)
    Q2 = ref.g2_op(ref.G_DBL, Q)
    assert np.array_equal(norm1(hop(hostsim, 43, P2, k)), ref.g1_op(ref.G_NORM, ref.g1_mul(P2, k)))
    assert np.array_equal(norm2(hop(hostsim, 53, Q2, k)), ref.g2_op(ref.G_NORM, ref.g2_mul(Q2, k)))
    # the point at infinity (every entry of the affine table of multiples is masked out) and scalars made of the extreme signed
    # digits: 0x8 (digit 8), 0x9 (digit -7 with a carry), 0xf (carries rippling to the extra top window)
    zero1, zero2 = np.zeros_like(P[:3]), np.zeros_like(Q[:3])
    assert not hop(hostsim, 43, zero1, k[-3:]).any() and not hop(hostsim, 53, zero2, k[-3:]).any()
    # the table of affine multiples behind both multiplications (test ops 45 / 55: entry m of pt_affine_multiples8), z = 1 and z != 1
    ms = ref.fr_from_ints(list(range(1, 9)))
    for PP, QQ in ((P[:8], Q[:8]), (P2[:8], Q2[:8])):
        assert np.array_equal(hop(hostsim, 45, PP, ms), ref.g1_op(ref.G_NORM, ref.g1_mul(PP, ms)))
        assert np.array_equal(hop(hostsim, 55, QQ, ms), ref.g2_op(ref.G_NORM, ref.g2_mul(QQ, ms)))
    pat = [int(h * 64, 16) % R for h in "89f7"] + [int("f" * 31, 16), int("8" * 32, 16), int("9" * 48, 16) % R]
    kp = ref.fr_from_ints(pat)
    assert np.array_equal(norm1(hop(hostsim, 43, P[:len(pat)], kp)), ref.g1_op(ref.G_NORM, ref.g1_mul(g, kp)))
    assert np.array_equal(norm2(hop(hostsim, 53, Q[:len(pat)], kp)), ref.g2_op(ref.G_NORM, ref.g2_mul(gg, kp)))


@bls_only
def test_executed_mac_counts(hostsim, ref):
    """bench.py's `executed_mac32_per_lane`: the Fp-level calls one psb_verify lane makes in each phase, counted by the
    instrumented host build of the engine's own headers (PSB_COUNT_OPS in tests/hostsim), equal the constants bench.py
    reports and its mirror of the batched affine sums (msm_ops) at every level setting."""
    import bench
    from tests import workload
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)  # noqa: E731
    rng = np.random.default_rng(3)
    for n, w in ((5, 8), (5, 4), (2, 8)):
        wl = workload.make_verify_workload(n_attrs=n, lanes=1, seed=3)
        # scalars without a zero digit and without recoding carries: every w-bit digit in [1, 2^(w-1) - 1], top digit small (< r)
        ks = []
        for _ in range(n):
            digs = [int(rng.integers(1, 1 << (w - 1))) for _ in range(256 // w)]
            digs[-1] = 1 + digs[-1] % 2 if w == 4 else 1 + digs[-1] % 0x2f
            ks.append(sum(d << (w * j) for j, d in enumerate(digs)))
        km = ref.fr_from_ints(ks)
        for levels in (0, 1, 2):
            out = np.zeros(12, dtype=np.uint64)
            hostsim.hostsim_count_verify_ops(C.c_int(n), C.c_int(w), C.c_int(levels), p(wl.key.gg), p(wl.key.XX), p(wl.key.YY),
                                             p(km), p(wl.sig1[0]), p(wl.sig2[0]), p(out))
            got = out.reshape(3, 4)
            assert tuple(int(x) for x in got[0]) == bench.msm_ops(n, w, levels)[1], (n, w, levels)
            assert tuple(int(x) for x in got[1]) == bench.EXEC_OPS_MILLER
            assert tuple(int(x) for x in got[2]) == bench.EXEC_OPS_FINAL
    # the headline shape: fewer executed MACs with every level, and the FpMul-eq numerator follows
    f0, f1, f2 = (bench.msm_ops(5, 20, lv) for lv in (0, 1, 2))
    assert f0[0] == 5 * 13 * 29 and f0[0] > f1[0] > f2[0]
    assert bench.exec_mac32(f0[1]) > bench.exec_mac32(f1[1]) > bench.exec_mac32(f2[1])
