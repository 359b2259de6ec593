"""oracle/ps_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-Python (arbitrary precision integer) restatement of the reference's algorithm for the
north-star hot path: batched PS-signature verification / issuance / randomisation and EL PASSO
sign-on verification over BLS12-381, exactly as the reference computes them through mcl.

Parity status: PINNED.  tests/test_oracle.py checks this file against
  (i)  mcl's own known-answer vectors (third-parties/mcl/test/bls12_test.cpp:19-65 generator +
       e(g1,g2) 12-coefficient KAT; :398-436 finalExp KAT), committed under tests/golden/, and
  (ii) the reference itself compiled here (oracle/_ref/libpsref.so, see oracle/Makefile) on
       seeded random inputs, byte for byte (raw Montgomery limbs, serialized points, verdicts).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product path (libpsb.so + the host mirror) never does.

Every function cites the reference file:line it restates.  `mcl/` abbreviates
/root/reference/third-parties/mcl/.  Pure-Python loops: use for small cases (a pairing is ~0.1 s).
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence, Tuple

# ------------------------------------------------------------------------------------------------
# curve constants: mcl BLS12_381 (mcl/include/mcl/curve_type.h:91, bn.hpp:909-994 Param::init)
# ------------------------------------------------------------------------------------------------
Z = -0xD201000000010000  # BLS parameter z (negative)
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
assert P == (Z - 1) ** 2 * (Z ** 4 - Z ** 2 + 1) // 3 + Z
assert R == Z ** 4 - Z ** 2 + 1
B1 = 4  # E : y^2 = x^3 + 4
FP_BYTES, FR_BYTES = 48, 32
RP = 1 << 384  # Montgomery radix for Fp (6 x 64 bit limbs; SURVEY F4, mcl/src/fp.cpp:375-391)
RR = 1 << 256  # Montgomery radix for Fr

# ------------------------------------------------------------------------------------------------
# Fp2 = Fp[i]/(i^2+1)  (mcl/include/mcl/fp_tower.hpp:214-611)  elements are (a, b) = a + b i
# ------------------------------------------------------------------------------------------------
Fp2 = Tuple[int, int]
F2_ZERO: Fp2 = (0, 0)
F2_ONE: Fp2 = (1, 0)
XI: Fp2 = (1, 1)  # xi = 1 + i  (fp_tower.hpp:584-592 mul_xi)


def f2_add(x: Fp2, y: Fp2) -> Fp2:
    return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)


def f2_sub(x: Fp2, y: Fp2) -> Fp2:
    return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)


def f2_neg(x: Fp2) -> Fp2:
    return ((-x[0]) % P, (-x[1]) % P)


def f2_mul(x: Fp2, y: Fp2) -> Fp2:  # fp_tower.hpp:528-534 (mulC), result only
    a, b = x
    c, d = y
    return ((a * c - b * d) % P, (a * d + b * c) % P)


def f2_sqr(x: Fp2) -> Fp2:  # fp_tower.hpp:539-550
    a, b = x
    return ((a + b) * (a - b) % P, 2 * a * b % P)


def f2_mul_fp(x: Fp2, k: int) -> Fp2:
    return (x[0] * k % P, x[1] * k % P)


def f2_mul_xi(x: Fp2) -> Fp2:  # (a+bi)(1+i) = (a-b) + (a+b)i ; fp_tower.hpp:584-592
    return ((x[0] - x[1]) % P, (x[0] + x[1]) % P)


def f2_conj(x: Fp2) -> Fp2:  # Frobenius on Fp2 since p = 3 mod 4; fp_tower.hpp:367-379
    return (x[0], (-x[1]) % P)


def f2_inv(x: Fp2) -> Fp2:  # fp_tower.hpp:597-611: (a - bi)/(a^2+b^2)
    a, b = x
    t = pow((a * a + b * b) % P, -1, P)
    return (a * t % P, (-b) * t % P)


# ------------------------------------------------------------------------------------------------
# Fp6 = Fp2[v]/(v^3 - xi) (fp_tower.hpp:786-1060), Fp12 = Fp6[w]/(w^2 - v) (fp_tower.hpp:1066-1372)
# An Fp12 value is held as 6 Fp2 coefficients of w^0..w^5 (w^2 = v):  c[k] <-> w^k.
# mcl memory order is (a.a, a.b, a.c, b.a, b.b, b.c) = (w^0, w^2, w^4, w^1, w^3, w^5).
# ------------------------------------------------------------------------------------------------
Fp12 = Tuple[Fp2, Fp2, Fp2, Fp2, Fp2, Fp2]
F12_ONE: Fp12 = (F2_ONE, F2_ZERO, F2_ZERO, F2_ZERO, F2_ZERO, F2_ZERO)
_MCL_ORDER = (0, 2, 4, 1, 3, 5)  # mcl slot j holds the coefficient of w^_MCL_ORDER[j]


def f12_mul(x: Fp12, y: Fp12) -> Fp12:  # value of fp_tower.hpp:1131-1160 (schoolbook over w)
    acc = [[0, 0] for _ in range(11)]
    for i in range(6):
        a, b = x[i]
        if a == 0 and b == 0:
            continue
        for j in range(6):
            c, d = y[j]
            t = acc[i + j]
            t[0] += a * c - b * d
            t[1] += a * d + b * c
    out = []
    for k in range(6):
        lo = acc[k]
        if k + 6 < 11:
            hi = acc[k + 6]  # w^6 = xi
            lo = [lo[0] + hi[0] - hi[1], lo[1] + hi[0] + hi[1]]
        out.append((lo[0] % P, lo[1] % P))
    return tuple(out)  # type: ignore[return-value]


def f12_sqr(x: Fp12) -> Fp12:  # fp_tower.hpp:1166-1178
    return f12_mul(x, x)


def f12_conj(x: Fp12) -> Fp12:  # unitaryInv: (a + b w) -> (a - b w); fp_tower.hpp:1202-1206
    return (x[0], f2_neg(x[1]), x[2], f2_neg(x[3]), x[4], f2_neg(x[5]))


def _f2_pow(x: Fp2, e: int) -> Fp2:
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, x)
        x = f2_sqr(x)
        e >>= 1
    return r


# gamma[j][k] = xi^(k (p^j - 1)/6): Frobenius constants (fp_tower.hpp:412-438 builds g/g2/g3)
_GAMMA = {}
for _j in (1, 2, 3):
    _base = _f2_pow(XI, (P ** _j - 1) // 6)
    _row = [F2_ONE]
    for _k in range(1, 6):
        _row.append(f2_mul(_row[-1], _base))
    _GAMMA[_j] = _row


def f12_frobenius(x: Fp12, j: int = 1) -> Fp12:  # fp_tower.hpp:1225-1265 (Frobenius/2/3)
    out = []
    for k in range(6):
        c = x[k] if j % 2 == 0 else f2_conj(x[k])
        out.append(f2_mul(c, _GAMMA[j][k]))
    return tuple(out)  # type: ignore[return-value]


def _f6_from(x: Fp12, odd: int):
    return (x[odd], x[odd + 2], x[odd + 4])


def f6_mul(x, y):  # Fp6 over v: v^3 = xi ; fp_tower.hpp:978-1022
    acc = [[0, 0] for _ in range(5)]
    for i in range(3):
        a, b = x[i]
        for j in range(3):
            c, d = y[j]
            acc[i + j][0] += a * c - b * d
            acc[i + j][1] += a * d + b * c
    out = []
    for k in range(3):
        lo = acc[k]
        if k + 3 < 5:
            hi = acc[k + 3]
            lo = [lo[0] + hi[0] - hi[1], lo[1] + hi[0] + hi[1]]
        out.append((lo[0] % P, lo[1] % P))
    return tuple(out)


def f6_inv(x):  # fp_tower.hpp:917-948
    a, b, c = x
    t0 = f2_sub(f2_sqr(a), f2_mul_xi(f2_mul(b, c)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(c)), f2_mul(a, b))
    t2 = f2_sub(f2_sqr(b), f2_mul(a, c))
    d = f2_add(f2_mul(a, t0), f2_mul_xi(f2_add(f2_mul(c, t1), f2_mul(b, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


def f12_inv(x: Fp12) -> Fp12:  # fp_tower.hpp:1183-1198: (a - b w)/(a^2 - v b^2)
    a = _f6_from(x, 0)
    b = _f6_from(x, 1)
    a2 = f6_mul(a, a)
    b2 = f6_mul(b, b)
    vb2 = (f2_mul_xi(b2[2]), b2[0], b2[1])  # multiply by v
    d = tuple(f2_sub(a2[i], vb2[i]) for i in range(3))
    di = f6_inv(d)
    ra = f6_mul(a, di)
    rb = f6_mul(b, di)
    rb = tuple(f2_neg(t) for t in rb)
    return (ra[0], rb[0], ra[1], rb[1], ra[2], rb[2])


def f12_pow(x: Fp12, e: int) -> Fp12:
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_sqr(r)
        if bit == "1":
            r = f12_mul(r, x)
    return r


# ------------------------------------------------------------------------------------------------
# raw (mcl in-memory) encodings: little-endian limbs, Montgomery form
# ------------------------------------------------------------------------------------------------
def fp_to_raw(x: int) -> bytes:
    return (x * RP % P).to_bytes(48, "little")


def fp_from_raw(b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(RP, -1, P) % P


def fr_to_raw(x: int) -> bytes:
    return (x * RR % R).to_bytes(32, "little")


def fr_from_raw(b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(RR, -1, R) % R


def f2_to_raw(x: Fp2) -> bytes:
    return fp_to_raw(x[0]) + fp_to_raw(x[1])


def f2_from_raw(b: bytes) -> Fp2:
    return (fp_from_raw(b[:48]), fp_from_raw(b[48:96]))


def f12_to_raw(x: Fp12) -> bytes:  # GT layout: a.a, a.b, a.c, b.a, b.b, b.c
    return b"".join(f2_to_raw(x[k]) for k in _MCL_ORDER)


def f12_from_raw(b: bytes) -> Fp12:
    c = [None] * 6
    for j, k in enumerate(_MCL_ORDER):
        c[k] = f2_from_raw(b[96 * j:96 * (j + 1)])
    return tuple(c)  # type: ignore[return-value]


# ------------------------------------------------------------------------------------------------
# groups.  Points are affine tuples or None (= infinity); the group law's VALUE is what
# mcl/include/mcl/ec.hpp:138-284 (dblJacobi/addJacobi) computes -- coordinates are free because
# every compared output is normalized (ec.hpp:77-88) or serialized (ec.hpp:849-896).
# ------------------------------------------------------------------------------------------------
G1Pt = Optional[Tuple[int, int]]
G2Pt = Optional[Tuple[Fp2, Fp2]]
B2: Fp2 = f2_mul_fp(XI, 4)  # M-type twist E': y^2 = x^3 + 4 xi  (bn.hpp:946-961)


def g1_is_on_curve(pt: G1Pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B1) % P == 0


def g2_is_on_curve(pt: G2Pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), B2)) == F2_ZERO


def g1_neg(a: G1Pt) -> G1Pt:
    return None if a is None else (a[0], (-a[1]) % P)


def g1_add(a: G1Pt, b: G1Pt) -> G1Pt:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def g1_mul(a: G1Pt, k: int) -> G1Pt:  # value of ec.hpp:1124-1139 (any algorithm; result canonical)
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g1_add(acc, acc)
        if bit == "1":
            acc = g1_add(acc, a)
    return acc


def g2_neg(a: G2Pt) -> G2Pt:
    return None if a is None else (a[0], f2_neg(a[1]))


def g2_add(a: G2Pt, b: G2Pt) -> G2Pt:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        lam = f2_mul(f2_mul_fp(f2_sqr(x1), 3), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(a: G2Pt, k: int) -> G2Pt:  # value of bn.hpp:1039-1047 (mulArrayGLV2)
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, a)
    return acc


def g1_to_raw(pt: G1Pt) -> bytes:
    """mcl G1 object bytes of a NORMALIZED point (z = 1) or zero (x=y=z=0): 144 B."""
    if pt is None:
        return bytes(144)
    return fp_to_raw(pt[0]) + fp_to_raw(pt[1]) + fp_to_raw(1)


def g1_from_raw(b: bytes) -> G1Pt:
    """Jacobian (x, y, z) raw limbs -> affine (ec.hpp:77-88 normalizeJacobi)."""
    x, y, z = (fp_from_raw(b[48 * i:48 * (i + 1)]) for i in range(3))
    if z == 0:
        return None
    zi = pow(z, -1, P)
    return (x * zi * zi % P, y * zi * zi * zi % P)


def g2_to_raw(pt: G2Pt) -> bytes:
    if pt is None:
        return bytes(288)
    return f2_to_raw(pt[0]) + f2_to_raw(pt[1]) + f2_to_raw(F2_ONE)


def g2_from_raw(b: bytes) -> G2Pt:
    x, y, z = (f2_from_raw(b[96 * i:96 * (i + 1)]) for i in range(3))
    if z == F2_ZERO:
        return None
    zi = f2_inv(z)
    zi2 = f2_sqr(zi)
    return (f2_mul(x, zi2), f2_mul(y, f2_mul(zi2, zi)))


def g1_serialize(pt: G1Pt) -> bytes:
    """mcl compressed LE form (ec.hpp:849-896, non-ETH mode, isMSBserialize): x as 48 LE bytes,
    bit 7 of the last byte = y odd (normal form); infinity = 48 zero bytes."""
    if pt is None:
        return bytes(48)
    b = bytearray(pt[0].to_bytes(48, "little"))
    if pt[1] & 1:
        b[47] |= 0x80
    return bytes(b)


def g2_serialize(pt: G2Pt) -> bytes:
    """x.a || x.b as LE bytes; parity of y.a (fp_tower.hpp:312) in bit 7 of the last byte."""
    if pt is None:
        return bytes(96)
    b = bytearray(pt[0][0].to_bytes(48, "little") + pt[0][1].to_bytes(48, "little"))
    if pt[1][0] & 1:
        b[95] |= 0x80
    return bytes(b)


def hex_str(b: bytes) -> str:  # serializeToHexStr, mcl/include/mcl/operator.hpp:187-193
    return b.hex()


# ------------------------------------------------------------------------------------------------
# hashing to scalars
# ------------------------------------------------------------------------------------------------
def small_mask(buf: bytes, modulus: int, bit_size: int) -> int:
    """fp::copyAndMask(..., SmallMask) (mcl/src/fp.cpp:612-662): read little-endian, keep bit_size
    bits, and if the value is still >= modulus keep bit_size-1 bits.  NOT a modular reduction."""
    n_bytes = (bit_size + 7) // 8
    x = int.from_bytes(buf[:n_bytes], "little")
    x &= (1 << bit_size) - 1
    if x >= modulus:
        x &= (1 << (bit_size - 1)) - 1
    return x


def fr_set_hash_of(msg: bytes) -> int:
    """Fr::setHashOf (mcl/include/mcl/fp.hpp:430-435): SHA-256 (fp.cpp:552-556) + SmallMask."""
    return small_mask(hashlib.sha256(msg).digest(), R, 255)


def fr_from_csprng_bytes(buf: bytes) -> int:
    """Fr::setByCSPRNG (fp.hpp:408-414): 32 raw bytes + SmallMask."""
    return small_mask(buf, R, 255)


def det_rng_bytes(state: int, n: int) -> Tuple[bytes, int]:
    """the xorshift64* byte stream installed by oracle/ref_harness.cc."""
    out = bytearray()
    m = (1 << 64) - 1
    for _ in range(n):
        state ^= state >> 12
        state ^= (state << 25) & m
        state ^= state >> 27
        out.append(((state * 0x2545F4914F6CDD1D) & m) >> 56)
    return bytes(out), state


# ------------------------------------------------------------------------------------------------
# pairing: optimal ate, mcl/include/mcl/bn.hpp:1660-1715 (millerLoop, pairing) + :1494-1659
# ------------------------------------------------------------------------------------------------
def _line(lam: Fp2, T: Tuple[Fp2, Fp2], Pt: Tuple[int, int]) -> Fp12:
    """Line through T with slope lam on the twist, evaluated at P and scaled by w^3 (a factor in a
    proper subfield, erased by the final exponentiation):
        l = (lam*xT - yT) + (-lam*xP) w^2 + yP w^3
    Same sparsity as mcl's M-type line (slots 0,4,1 of bn.hpp:1436-1468 mul_041)."""
    xT, yT = T
    xP, yP = Pt
    c0 = f2_sub(f2_mul(lam, xT), yT)
    c2 = f2_neg(f2_mul_fp(lam, xP))
    c3 = (yP % P, 0)
    return (c0, F2_ZERO, c2, c3, F2_ZERO, F2_ZERO)


def miller_loop(Pt: G1Pt, Q: G2Pt) -> Fp12:
    """f_{|z|,Q}(P), conjugated because z < 0 (bn.hpp:1695-1697).  Q = 0 -> 1 (bn.hpp:1666-1669);
    P = 0 falls through with xP = yP = 0 (every line lands in Fp2 and dies in finalExp).
    The pre-final-exp VALUE differs from mcl's by subfield factors; only final_exp(miller_loop)
    is comparable."""
    if Q is None:
        return F12_ONE
    Pa = (0, 0) if Pt is None else Pt
    f = F12_ONE
    T = Q
    for bit in bin(-Z)[3:]:
        lam = f2_mul(f2_mul_fp(f2_sqr(T[0]), 3), f2_inv(f2_add(T[1], T[1])))
        f = f12_mul(f12_sqr(f), _line(lam, T, Pa))
        T = g2_add(T, T)
        if bit == "1":
            lam = f2_mul(f2_sub(Q[1], T[1]), f2_inv(f2_sub(Q[0], T[0])))
            f = f12_mul(f, _line(lam, T, Pa))
            T = g2_add(T, Q)
    return f12_conj(f)


def _pow_z(x: Fp12) -> Fp12:  # bn.hpp:1150-1176: x^|z| then unitaryInv because z < 0
    return f12_conj(f12_pow(x, -Z))


def final_exp(x: Fp12) -> Fp12:
    """bn.hpp:1643-1659: mapToCyclotomic (:1494-1502) then expHardPartBLS12 (:1508-1555), i.e.
    x^((p^6-1)(p^2+1) * 3(p^4-p^2+1)/r) -- note mcl's factor 3."""
    z = f12_mul(f12_frobenius(x, 2), x)  # x^(p^2+1)
    y = f12_mul(f12_inv(z), f12_conj(z))  # ^(p^6-1)
    x = y
    a0 = f12_conj(x)
    a1 = f12_sqr(a0)
    a2 = _pow_z(x)
    a3 = f12_sqr(a2)
    a1 = f12_mul(a1, a2)
    a7 = _pow_z(a1)
    a4 = _pow_z(a7)
    a5 = _pow_z(a4)
    a3 = f12_mul(a3, a5)
    a6 = _pow_z(a3)
    a1 = f12_conj(a1)
    a1 = f12_mul(a1, a6)
    a1 = f12_mul(a1, x)
    a3 = f12_mul(a3, a0)
    a3 = f12_frobenius(a3, 1)
    a1 = f12_mul(a1, a3)
    a4 = f12_mul(a4, a2)
    a4 = f12_frobenius(a4, 2)
    a1 = f12_mul(a1, a4)
    a7 = f12_mul(a7, x)
    y = f12_frobenius(a7, 3)
    return f12_mul(y, a1)


def pairing(Pt: G1Pt, Q: G2Pt) -> Fp12:  # bn.hpp:1711-1715
    return final_exp(miller_loop(Pt, Q))


def pairing_ratio(P1: G1Pt, Q1: G2Pt, P2: G1Pt, Q2: G2Pt) -> Fp12:
    """e(P1,Q1) * e(P2,Q2)^-1: the fused lane value, equal to lhs * unitaryInv(rhs) of
    src/ps-verifier.cc:31-34 (SURVEY 8d 'Parity check')."""
    return final_exp(f12_mul(miller_loop(P1, Q1), miller_loop(g1_neg(P2), Q2)))


# ------------------------------------------------------------------------------------------------
# protocol layer (reference src/*.cc)
# ------------------------------------------------------------------------------------------------
class PubKey:
    """PSPubKey (src/ps-encoding.h:111-133): g, gg, XX, Yi[], YYi[]."""

    def __init__(self, g: G1Pt, gg: G2Pt, XX: G2Pt, Yi: Sequence[G1Pt], YYi: Sequence[G2Pt]):
        self.g, self.gg, self.XX, self.Yi, self.YYi = g, gg, XX, list(Yi), list(YYi)


def keygen(g: G1Pt, gg: G2Pt, x: int, ys: Sequence[int]) -> Tuple[PubKey, G1Pt]:
    """PSSigner::key_gen (src/ps-signer.cc:29-55) with the exponents supplied by the caller."""
    pk = PubKey(g, gg, g2_mul(gg, x), [g1_mul(g, y) for y in ys], [g2_mul(gg, y) for y in ys])
    return pk, g1_mul(g, x)


def ps_verify_K(pk: PubKey, attrs: Sequence[bytes]) -> G2Pt:
    K = pk.XX
    for i, a in enumerate(attrs):  # src/ps-verifier.cc:24-29
        K = g2_add(K, g2_mul(pk.YYi[i], fr_set_hash_of(a)))
    return K


def ps_verify(pk: PubKey, sig1: G1Pt, sig2: G1Pt, attrs: Sequence[bytes]) -> bool:
    """PSVerifier::verify (src/ps-verifier.cc:13-35) == PSRequester::verify (ps-requester.cc:115-137)."""
    if sig1 is None:
        return False
    K = ps_verify_K(pk, attrs)
    return pairing(sig1, K) == pairing(sig2, pk.gg)


def ps_verify_gt(pk: PubKey, sig1: G1Pt, sig2: G1Pt, attrs: Sequence[bytes]) -> Fp12:
    return pairing_ratio(sig1, ps_verify_K(pk, attrs), sig2, pk.gg)


def randomize_credential(sig1: G1Pt, sig2: G1Pt, t: int) -> Tuple[G1Pt, G1Pt]:
    """PSRequester::randomize_credential (src/ps-requester.cc:139-148), t host-supplied."""
    return g1_mul(sig1, t), g1_mul(sig2, t)


def _challenge(points_hex: Sequence[str], ad: bytes) -> int:
    """c = Fr::setHashOf(Sha256(hex(points...) || ad)) -- double hash (SURVEY F5;
    src/ps-verifier.cc:111-122, src/ps-signer.cc:95-101)."""
    h = hashlib.sha256()
    for s in points_hex:
        h.update(s.encode())
    h.update(ad)
    return fr_set_hash_of(h.digest())


def nizk_verify_request(pk: PubKey, A: G1Pt, c: int, rs: Sequence[int], attrs: Sequence[bytes],
                        ad: bytes) -> bool:
    """PSSigner::el_passo_nizk_verify_request (src/ps-signer.cc:74-110)."""
    V = g1_add(g1_mul(A, c), g1_mul(pk.g, rs[0]))
    j = 1
    for i, a in enumerate(attrs):
        if a == b"":
            V = g1_add(V, g1_mul(pk.Yi[i], rs[j]))
            j += 1
    return _challenge([hex_str(g1_serialize(A)), hex_str(g1_serialize(V))], ad) == c


def sign_commitment(pk: PubKey, X: G1Pt, commitment: G1Pt, u: int) -> Tuple[G1Pt, G1Pt]:
    """PSSigner::sign_commitment (src/ps-signer.cc:132-146), u host-supplied."""
    return g1_mul(pk.g, u), g1_mul(g1_add(X, commitment), u)


def sign_hybrid(pk: PubKey, X: G1Pt, A: G1Pt, attrs: Sequence[bytes], u: int):
    """PSSigner::sign_hybrid (src/ps-signer.cc:112-130) incl. the size()==1 shortcut (F9)."""
    if len(attrs) == 1:
        return sign_commitment(pk, X, A, u)
    for i, a in enumerate(attrs):
        if a == b"":
            continue
        A = g1_add(A, g1_mul(pk.Yi[i], fr_set_hash_of(a)))
    return sign_commitment(pk, X, A, u)


def provide_id(pk: PubKey, X: G1Pt, A: G1Pt, c: int, rs: Sequence[int], attrs: Sequence[bytes],
               ad: bytes, u: int):
    """PSSigner::el_passo_provide_id (src/ps-signer.cc:63-72). Returns (ok, sig1, sig2)."""
    if not nizk_verify_request(pk, A, c, rs, attrs, ad):
        return False, None, None
    s1, s2 = sign_hybrid(pk, X, A, attrs, u)
    return True, s1, s2


def prepare_hybrid_verification(pk: PubKey, k: G2Pt, attrs: Sequence[bytes]) -> G2Pt:
    """src/ps-verifier.cc:214-229."""
    for i, a in enumerate(attrs):
        if a == b"":
            continue
        k = g2_add(k, g2_mul(pk.YYi[i], fr_set_hash_of(a)))
    return k


def verify_id(pk: PubKey, sig1: G1Pt, sig2: G1Pt, k: G2Pt, phi: G1Pt, E1: G1Pt, E2: G1Pt, c: int,
              rs: Sequence[int], attrs: Sequence[bytes], ad: bytes, service_pt: G1Pt,
              y: G1Pt, g: G1Pt, h: G1Pt, with_id: bool = True) -> bool:
    """PSVerifier::el_passo_verify_id (src/ps-verifier.cc:37-138) and
    el_passo_verify_id_without_id_retrieval (:140-212).  `service_pt` = hashAndMapToG1(service)
    (one value per batch, computed by the host: SURVEY a26)."""
    n_rs = len(rs)
    Vk = g2_mul(k, c)
    counter = 0
    for i, a in enumerate(attrs):
        if a == b"":
            Vk = g2_add(Vk, g2_mul(pk.YYi[i], rs[counter]))
            counter += 1
    Vk = g2_add(Vk, g2_mul(pk.gg, rs[n_rs - 2] if with_id else rs[n_rs - 1]))
    Vk = g2_add(Vk, g2_mul(pk.XX, (1 - c) % R))
    Vphi = g1_add(g1_mul(phi, c), g1_mul(service_pt, rs[0]))
    pts = [hex_str(g2_serialize(k)), hex_str(g1_serialize(phi))]
    if with_id:
        VE1 = g1_add(g1_mul(E1, c), g1_mul(g, rs[n_rs - 1]))
        VE2 = g1_add(g1_add(g1_mul(E2, c), g1_mul(y, rs[n_rs - 1])), g1_mul(h, rs[1]))
        pts += [hex_str(g1_serialize(E1)), hex_str(g1_serialize(E2))]
        pts += [hex_str(g2_serialize(Vk)), hex_str(g1_serialize(Vphi)),
                hex_str(g1_serialize(VE1)), hex_str(g1_serialize(VE2))]
    else:
        pts += [hex_str(g2_serialize(Vk)), hex_str(g1_serialize(Vphi))]
    if _challenge(pts, ad) != c:
        return False
    K = prepare_hybrid_verification(pk, k, attrs)
    return pairing(sig1, K) == pairing(sig2, pk.gg)
