"""GPU parity, layer by layer: every arithmetic op of the engine (one GPU thread per element,
through the C ABI's psb_test_op) against the reference compiled in oracle/_ref (mcl), bit-exact on
raw Montgomery limbs / serialized points."""
import numpy as np
import pytest

from tests.conftest import BN254, CURVE_Z, FIELD_P, GROUP_R, rand_fp_raw

pytestmark = pytest.mark.gpu

N = 257  # ragged on purpose (not a multiple of the block size)


@pytest.fixture(scope="module")
def rng():
    return np.random.default_rng(20261017)


def edge_fp(ref, a, b):
    a[0] = 0
    b[1] = 0
    a[2] = ref.fp_from_ints([FIELD_P - 1])[0]
    b[2] = a[2]
    a[3] = ref.fp_from_ints([1])[0]
    return a, b


@pytest.mark.parametrize("op,name", [(0, "add"), (1, "sub"), (2, "mul"), (3, "sqr"), (4, "neg"), (5, "inv")])
def test_fp(gpu_pkg, ref, rng, op, name):
    a, b = edge_fp(ref, rand_fp_raw(ref, rng, N), rand_fp_raw(ref, rng, N))
    if op == 5:
        a[0] = a[5]  # mcl's inv(0) traps
    exp = ref.fp_op(op, a, b)
    got = gpu_pkg.test_op(op, a, b if op < 3 else None)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("op", range(6))
def test_fp2(gpu_pkg, ref, rng, op):
    a, b = rand_fp_raw(ref, rng, N, 2), rand_fp_raw(ref, rng, N, 2)
    a[0] = 0 if op != 5 else a[1]
    b[1] = 0
    assert np.array_equal(gpu_pkg.test_op(10 + op, a, b if op < 3 else None), ref.fp2_op(op, a, b))


def extreme_rows(ref, a, b):
    """worst-case magnitudes for the multiplier's unreduced operand sums (tower.cuh, engine A): every component p - 1;
    p - 1 against (p - 1, 0) pairs (the negated imaginary part becomes 2p); p - 1 against 0 and 1"""
    w = a.shape[1] // ref.fp_from_ints([1]).shape[1]
    top, one = ref.fp_from_ints([FIELD_P - 1])[0], ref.fp_from_ints([1])[0]
    fw = top.shape[0]
    a[0] = np.tile(top, w); b[0] = np.tile(top, w)
    a[1] = np.tile(top, w); b[1] = np.tile(np.concatenate([top, np.zeros(fw, dtype=top.dtype)]), w // 2)
    a[2] = np.tile(np.concatenate([np.zeros(fw, dtype=top.dtype), top]), w // 2); b[2] = np.tile(top, w)
    a[3] = np.tile(top, w); b[3] = np.tile(one, w)
    a[4] = np.tile(top, w); b[4] = 0
    return a, b


def test_fp6(gpu_pkg, ref, rng):
    a, b = extreme_rows(ref, rand_fp_raw(ref, rng, 65, 6), rand_fp_raw(ref, rng, 65, 6))
    assert np.array_equal(gpu_pkg.test_op(20, a, b), ref.fp6_op(ref.OP_MUL, a, b))
    assert np.array_equal(gpu_pkg.test_op(21, a), ref.fp6_op(ref.OP_INV, a))


def test_fp12(gpu_pkg, ref, rng):
    a, b = extreme_rows(ref, rand_fp_raw(ref, rng, 65, 12), rand_fp_raw(ref, rng, 65, 12))
    assert np.array_equal(gpu_pkg.test_op(30, a, b), ref.fp12_op(ref.OP_MUL, a, b))
    assert np.array_equal(gpu_pkg.test_op(31, a), ref.fp12_op(ref.OP_SQR, a))
    assert np.array_equal(gpu_pkg.test_op(32, a), ref.fp12_op(ref.OP_INV, a))
    for k in (1, 2, 3):
        assert np.array_equal(gpu_pkg.test_op(32 + k, a), ref.fp12_frobenius(k, a))


@pytest.fixture(scope="module")
def points(ref):
    ref.seed(99)
    g, gg = ref.hash_to_g1(b"abc"), ref.hash_to_g2(b"edf")
    n = 33
    k1, k2 = ref.fr_rand(n), ref.fr_rand(n)
    return dict(P=ref.g1_mul(g, k1), P2=ref.g1_mul(g, k2), Q=ref.g2_mul(gg, k1), Q2=ref.g2_mul(gg, k2),
                k=ref.fr_rand(n), n=n)


def test_g1(gpu_pkg, ref, points):
    P, P2, k = points["P"], points["P2"], points["k"]
    s = ref.g1_serialize
    assert np.array_equal(s(gpu_pkg.test_op(40, P, P2)), s(ref.g1_op(ref.G_ADD, P, P2)))
    assert np.array_equal(s(gpu_pkg.test_op(41, P)), s(ref.g1_op(ref.G_DBL, P)))
    assert np.array_equal(gpu_pkg.test_op(42, P), ref.g1_op(ref.G_NORM, P))  # raw limbs, z = 1
    assert np.array_equal(s(gpu_pkg.test_op(40, P, P)), s(ref.g1_op(ref.G_DBL, P)))  # P + P
    neg = ref.g1_op(ref.G_NEG, P)
    assert not gpu_pkg.test_op(40, P, neg).any()  # P + (-P) = canonical zero
    zero = np.zeros_like(P)
    assert np.array_equal(s(gpu_pkg.test_op(40, zero, P2)), s(P2))
    assert np.array_equal(s(gpu_pkg.test_op(43, P, k)), s(ref.g1_mul(P, k)))
    assert np.array_equal(s(gpu_pkg.test_op(44, P, ref.g1_op(ref.G_NORM, P2))), s(ref.g1_op(ref.G_ADD, P, P2)))
    # the table of affine multiples 1..8 behind G1::mul (one shared inversion), every entry, lanes with different entries per warp
    ms = ref.fr_from_ints([1 + (i % 8) for i in range(P.shape[0])])
    assert np.array_equal(gpu_pkg.test_op(45, P, ms), ref.g1_op(ref.G_NORM, ref.g1_mul(P, ms)))


def test_g2(gpu_pkg, ref, points):
    Q, Q2, k = points["Q"], points["Q2"], points["k"]
    s = ref.g2_serialize
    assert np.array_equal(s(gpu_pkg.test_op(50, Q, Q2)), s(ref.g2_op(ref.G_ADD, Q, Q2)))
    assert np.array_equal(s(gpu_pkg.test_op(51, Q)), s(ref.g2_op(ref.G_DBL, Q)))
    assert np.array_equal(gpu_pkg.test_op(52, Q), ref.g2_op(ref.G_NORM, Q))
    assert np.array_equal(s(gpu_pkg.test_op(50, Q, Q)), s(ref.g2_op(ref.G_DBL, Q)))
    assert np.array_equal(s(gpu_pkg.test_op(53, Q, k)), s(ref.g2_mul(Q, k)))
    assert np.array_equal(s(gpu_pkg.test_op(54, Q, ref.g2_op(ref.G_NORM, Q2))), s(ref.g2_op(ref.G_ADD, Q, Q2)))
    ms = ref.fr_from_ints([1 + (i % 8) for i in range(Q.shape[0])])
    assert np.array_equal(gpu_pkg.test_op(55, Q, ms), ref.g2_op(ref.G_NORM, ref.g2_mul(Q, ms)))


def test_scalar_edge_cases(gpu_pkg, ref, points):
    """scalars 0, 1, 2, r-1 (mcl's small-int fast path, ec.hpp:1140-1260, must give the same point)."""
    P = points["P"][:4]
    k = ref.fr_from_ints([0, 1, 2, GROUP_R - 1])
    assert np.array_equal(ref.g1_serialize(gpu_pkg.test_op(43, P, k)), ref.g1_serialize(ref.g1_mul(P, k)))


def test_pairing(gpu_pkg, ref, points):
    P, Q = points["P"][:9], points["Q"][:9]
    exp = ref.pairing(P, Q)
    assert np.array_equal(gpu_pkg.test_op(60, P, Q), exp)       # probe path
    assert np.array_equal(gpu_pkg.pairing(P, Q), exp)           # psb_pairing entry
    f = ref.miller_loop(P, Q)
    assert np.array_equal(gpu_pkg.test_op(61, f), ref.final_exp(f))  # final exponentiation on mcl's Miller value
    cyc = exp
    assert np.array_equal(gpu_pkg.test_op(36, cyc), ref.fp12_op(ref.OP_SQR, cyc))  # cyclotomic squaring


def test_pairing_zero_points(gpu_pkg, ref, points):
    """e(0, Q) = e(P, 0) = 1 (mcl: bn.hpp:1666-1669; bls12_test.cpp:288-296)."""
    P, Q = points["P"][:2].copy(), points["Q"][:2].copy()
    P[0] = 0
    Q[1] = 0
    assert np.array_equal(gpu_pkg.pairing(P, Q), ref.pairing(P, Q))


def test_pairing_ratio_fixed_lines(gpu_pkg, ref, points):
    P, Q, P2 = points["P"][:5], points["Q"][:5], points["P2"][:5]
    Q2 = ref.g2_op(ref.G_NORM, points["Q2"][:5])
    c = np.concatenate([P2, Q2], axis=1)
    assert np.array_equal(gpu_pkg.test_op(62, P, Q, c), ref.pairing_ratio(P, Q, P2, Q2))


def test_fr(gpu_pkg, ref):
    ref.seed(5)
    a, b = ref.fr_rand(64), ref.fr_rand(64)
    assert np.array_equal(gpu_pkg.test_op(71, a, b), ref.fr_op(ref.OP_MUL, a, b))
    assert np.array_equal(gpu_pkg.test_op(72, a, b), ref.fr_op(ref.OP_SUB, a, b))
    assert np.array_equal(gpu_pkg.test_op(73, a, b), ref.fr_op(ref.OP_ADD, a, b))
    ints = ref.fr_to_ints(a)
    got = gpu_pkg.test_op(70, a)
    assert [int.from_bytes(got[i].tobytes(), "little") for i in range(64)] == ints


def test_glv_gls_scalar_edges(gpu_pkg, ref):
    """GLV (G1) / GLS (G2) variable-base multiplication on the GPU at the decomposition boundaries."""
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
    ks += [int.from_bytes(rng.bytes(32), "little") % R for _ in range(105)]
    k = ref.fr_from_ints(ks)
    g, gg = ref.hash_to_g1(b"abc"), ref.hash_to_g2(b"edf")
    P = ref.g1_op(ref.G_DBL, np.repeat(g.reshape(1, -1), len(ks), axis=0))   # z != 1
    Q = ref.g2_op(ref.G_DBL, np.repeat(gg.reshape(1, -1), len(ks), axis=0))
    assert np.array_equal(ref.g1_serialize(gpu_pkg.test_op(43, P, k)), ref.g1_serialize(ref.g1_mul(P, k)))
    assert np.array_equal(ref.g2_serialize(gpu_pkg.test_op(53, Q, k)), ref.g2_serialize(ref.g2_mul(Q, k)))
    # lanes holding the point at infinity MIXED with ordinary lanes in the same warps (they build their table of multiples from a
    # stand-in point and must not disturb their neighbours), and scalars made of the extreme signed digits
    pat = [int(h * 64, 16) % GROUP_R for h in "89f7"] + [int("f" * 31, 16), int("8" * 32, 16), int("9" * 48, 16) % GROUP_R]
    kp = ref.fr_from_ints((pat * 10)[:64])
    P, Q = P[:1].repeat(64, axis=0).copy(), Q[:1].repeat(64, axis=0).copy()
    P[::3] = 0
    Q[1::4] = 0
    got1, got2 = gpu_pkg.test_op(42, gpu_pkg.test_op(43, P, kp)), gpu_pkg.test_op(52, gpu_pkg.test_op(53, Q, kp))
    assert np.array_equal(got1, ref.g1_op(ref.G_NORM, ref.g1_mul(P, kp))) and not got1[::3].any()
    assert np.array_equal(got2, ref.g2_op(ref.G_NORM, ref.g2_mul(Q, kp))) and not got2[1::4].any()


def test_fp_inv_divsteps_edges_gpu(gpu_pkg, ref):
    """Bernstein-Yang Fp inversion (csrc/modinv.cuh) on the device: raw limb patterns 1, p-1, powers of two, p - 2^k,
    random -- against mcl's Fp::inv; inv(0) = 0."""
    from tests.conftest import FIELD_P, FP_BYTES
    P, bits = FIELD_P, FIELD_P.bit_length()
    rng = np.random.default_rng(11)
    raw = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (1 << (bits - 1)) - 1, 1 << (bits - 1)]
    raw += [1 << k for k in range(1, bits - 1, 13)] + [P - (1 << k) for k in range(1, bits - 1, 17)]
    raw += [int.from_bytes(rng.bytes(FP_BYTES), "little") % P for _ in range(200)]
    a = np.frombuffer(b"".join(v.to_bytes(FP_BYTES, "little") for v in raw), dtype=np.uint64).reshape(len(raw), -1).copy()
    got = gpu_pkg.test_op(5, a)
    assert not got[0].any()
    assert np.array_equal(got[1:], ref.fp_op(5, a[1:]))


def test_final_exp_compressed_pow_z_mixed_warp(gpu_pkg, ref):
    """finalExp with pow_z on compressed squarings: lanes whose cyclotomic image is 1 (the value 1, elements of Fp2 /
    Fp6: zero denominators -> the whole warp takes the Granger-Scott fallback) interleaved with generic lanes in the
    same warps, plus warps with no such lane -- GT bytes against mcl's finalExp on every lane."""
    from tests.conftest import rand_fp_raw
    rng = np.random.default_rng(13)
    n = 96                                     # three warps; degenerate lanes only in the first
    a = rand_fp_raw(ref, rng, n, 12)
    fp = a.shape[1] // 12
    one = ref.fp_from_ints([1]).reshape(-1)
    a[3] = 0; a[3, :fp] = one
    a[7, 2 * fp:] = 0
    a[20, 6 * fp:] = 0
    want = ref.final_exp(a)
    assert np.array_equal(gpu_pkg.test_op(61, a), want)
    assert np.array_equal(want[3], a[3]) and np.array_equal(want[7], a[3]) and np.array_equal(want[20], a[3])
