"""The generated PTX row chains of the Montgomery multiplier (csrc/fp_cios*.cuh), INTERPRETED on Python integers.

tools/gen_cios.py describes every chain once and both emits the inline-PTX macro and executes it; here the rows are
composed exactly like csrc/cios.cuh composes the macros -- the fused multiplier (mul_rr), the wide product (mulpre_rr)
and the stand-alone Montgomery reduction (redc_rr) of the lazy-reduction Fp2 engine -- and checked against
big-integer arithmetic, so the limb bookkeeping (window shifts, stray limbs, carry hand-overs) is verified without a
GPU.  The device code itself is compared with mcl in tests/test_gpu_arith.py.
"""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_cios  # noqa: E402


def _operands(m, rng, bound):
    """edge patterns + random values below `bound`"""
    edge = [0, 1, 2, m.p - 1, m.p, m.p + 1, bound - 1, (1 << 32) - 1, 1 << 32, (1 << (32 * m.n - 1)) % bound,
            int("f" * (8 * m.n), 16) % bound, m.R - 1 if bound >= m.R else bound - 1]
    return edge + [rng.randrange(bound) for _ in range(40)]


@pytest.mark.parametrize("curve", ["bls12_381", "bn254"])
def test_generated_header_is_current(curve):
    c = gen_cios.CURVES[curve]
    path = os.path.join(ROOT, "ps-signature-and-el-passo_b200", "csrc", c["file"])
    assert open(path).read() == gen_cios.gen(c["p"], c["n"]), "run python tools/gen_cios.py"


@pytest.mark.parametrize("curve", ["bls12_381", "bn254"])
def test_fused_multiplier(curve):
    m = gen_cios.Model(curve)
    rng = random.Random(1)
    ri = pow(m.R, -1, m.p)
    # operands below 2p: the multiplicand side of the engines may be an unreduced sum
    vals = _operands(m, rng, 2 * m.p)
    for a in vals:
        for b in vals[::3]:
            t = m.mul(m.limbs(a), m.limbs(b))
            assert t < 2 * m.p and t % m.p == a * b * ri % m.p


@pytest.mark.parametrize("curve", ["bls12_381", "bn254"])
def test_wide_product(curve):
    m = gen_cios.Model(curve)
    rng = random.Random(2)
    vals = _operands(m, rng, m.R)           # any n-limb operands (sums up to 4p < R on the Karatsuba middle term)
    for a in vals:
        for b in vals[::3]:
            assert m.value(m.mulpre(m.limbs(a), m.limbs(b))) == a * b


@pytest.mark.parametrize("curve", ["bls12_381", "bn254"])
def test_standalone_reduction(curve):
    m = gen_cios.Model(curve)
    rng = random.Random(3)
    ri = pow(m.R, -1, m.p)
    # T < p R (the engine's contract: results then lie below 2p), plus the extremes of the 2n-limb range
    bound = m.p * m.R
    vals = [0, 1, m.p, m.R - 1, m.R, bound - 1, (m.p - 1) ** 2, 2 * (2 * m.p - 1) ** 2 % bound] + [rng.randrange(bound) for _ in range(300)]
    for t in vals:
        r = m.redc(m.limbs(t, 2 * m.n))
        assert r % m.p == t * ri % m.p
        assert r < 2 * m.p, (hex(t), hex(r))
    # any 2n-limb input still reduces correctly modulo p (result < T / R + p + 1)
    for t in [m.R * m.R - 1] + [rng.randrange(m.R * m.R) for _ in range(50)]:
        r = m.redc(m.limbs(t, 2 * m.n))
        assert r % m.p == t * ri % m.p and r <= t // m.R + m.p


@pytest.mark.parametrize("curve", ["bls12_381", "bn254"])
def test_karatsuba_fp2_bounds(curve):
    """the lazy-reduction Fp2 product of tower.cuh (engine A): three wide products, two reductions.
    re = T0 - T1 (+ pR on borrow), im = T2 - T0 - T1 with T2 = (xa + xb)(ya + yb); operands up to the engine's
    unreduced-sum bounds; both reductions must see T < pR so that ONE conditional subtraction canonicalises."""
    m = gen_cios.Model(curve)
    rng = random.Random(4)
    ri = pow(m.R, -1, m.p)
    xb_bound = 2 * m.p                       # multiplicand side: x = x1 + x2 unreduced
    yb_bound = 2 * m.p if curve == "bls12_381" else m.p   # multiplier side: unreduced on BLS12-381 only (tower.cuh PSB_LAZY_Y)
    assert 2 * xb_bound < m.R and 2 * yb_bound < m.R       # the Karatsuba operand sums fit n limbs
    assert 2 * xb_bound * yb_bound < m.p * m.R             # im < pR
    for _ in range(60):
        xa, xb = (rng.choice([0, xb_bound - 1, rng.randrange(xb_bound)]) for _ in range(2))
        ya, yb = (rng.choice([0, yb_bound - 1, rng.randrange(yb_bound)]) for _ in range(2))
        T0 = m.value(m.mulpre(m.limbs(xa), m.limbs(ya)))
        T1 = m.value(m.mulpre(m.limbs(xb), m.limbs(yb)))
        T2 = m.value(m.mulpre(m.limbs(xa + xb), m.limbs(ya + yb)))
        im = T2 - T0 - T1
        re = T0 - T1 + (m.p * m.R if T0 < T1 else 0)
        assert 0 <= im < m.p * m.R and 0 <= re < m.p * m.R
        r_re, r_im = m.redc(m.limbs(re, 2 * m.n)), m.redc(m.limbs(im, 2 * m.n))
        assert r_re < 2 * m.p and r_im < 2 * m.p
        assert r_re % m.p == (xa * ya - xb * yb) * ri % m.p
        assert r_im % m.p == (xa * yb + xb * ya) * ri % m.p


@pytest.mark.parametrize("curve", ["bls12_381", "bn254"])
def test_wide_square(curve):
    """the dedicated squaring (cios::sqrpre_rr): N (N - 1) / 2 cross products + N diagonal ones instead of N^2"""
    m = gen_cios.Model(curve)
    rng = random.Random(5)
    n_mac = sum(len([op for op in r.ops if op[0].endswith("lo.cc")]) for r in gen_cios.sqr_rows(m.n) + gen_cios.sqr_diag_rows(m.n))
    assert n_mac == m.n * (m.n + 1) // 2
    ri = pow(m.R, -1, m.p)
    for a in _operands(m, rng, m.R) + [rng.randrange(m.R) for _ in range(200)]:
        T = m.sqrpre(m.limbs(a))
        assert m.value(T) == a * a
    for a in _operands(m, rng, 2 * m.p):          # the Montgomery square: operands below 2p give T < 4 p^2 < pR
        r = m.redc(m.sqrpre(m.limbs(a)))
        assert r < 2 * m.p and r % m.p == a * a * ri % m.p


@pytest.mark.parametrize("curve", ["bls12_381", "bn254"])
def test_fused_dot2_with_unreduced_operands(curve):
    """engine A of csrc/tower.cuh: (re, im) = (xa ya + xb (2p - yb), xa yb + xb ya) through cios::dot2_rr.  The multiplicand
    side is always an unreduced sum (< 2p); on BLS12-381 (PSB_LAZY_Y) the multiplier side is too, so the accumulated pair of
    products reaches 8 p^2: the rows must not overflow their window and the result must stay below 2p, so that the single
    conditional subtraction canonicalises it ((8 p^2 + R p) / R < 1.82 p needs 8 p < R: true for 381 bits in 384, false for
    BN254, whose multiplier side therefore stays canonical)."""
    m = gen_cios.Model(curve)
    rng = random.Random(6)
    ri = pow(m.R, -1, m.p)
    lazy_y = curve == "bls12_381"
    assert (8 * m.p < m.R) == lazy_y
    xb, yb = 2 * m.p, (2 * m.p if lazy_y else m.p)
    xs = [xb - 1, xb - 2, m.p, 1, 0] + [rng.randrange(xb) for _ in range(12)]
    ys = [yb - 1, yb - 2, m.p - 1, 1, 0] + [rng.randrange(yb) for _ in range(12)]
    worst = 0
    for xa in xs:
        for xbv in xs[::2]:
            for ya in ys[::2]:
                for ybv in ys[::3]:
                    t = m.dot2(m.limbs(xa), m.limbs(ya), m.limbs(xbv), m.limbs(ybv))
                    assert t < 2 * m.p, (hex(xa), hex(ya), hex(xbv), hex(ybv))
                    assert t % m.p == (xa * ya + xbv * ybv) * ri % m.p
                    worst = max(worst, t)
    assert worst >= m.p          # the top of the range is exercised
