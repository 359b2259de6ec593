#!/usr/bin/env python3
"""tools/gen_cios.py -- emits the even/odd row chains (inline PTX) used by csrc/cios.cuh, and can EXECUTE them.

One row of the operand-scanning Montgomery multiplier is two carry chains of N/2 wide MACs
(mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32.X); the modulus limbs of the
reduction rows are immediates.  N = limbs of the field (12 for BLS12-381, 8 for BN254).

Every chain is described ONCE as a list of abstract carry-flag operations; `emit` turns the list into an
inline-PTX macro and `Machine` interprets the same list on Python integers.  tests/test_cios_model.py composes the
interpreted rows exactly like csrc/cios.cuh composes the macros (fused multiplier, wide product, stand-alone
reduction) and checks them against big-integer arithmetic, so the limb bookkeeping is verified without a GPU.

Run:  python tools/gen_cios.py        (writes csrc/fp_cios.cuh and csrc/fp_cios_bn254.cuh)
"""
import os

CURVES = {
    "bls12_381": dict(
        p=0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
        n=12, file="fp_cios.cuh"),
    "bn254": dict(
        p=0x2523648240000001BA344D80000000086121000000000013A700000000000013,
        n=8, file="fp_cios_bn254.cuh"),
}
M32 = 0xFFFFFFFF


class Row:
    """one macro: parameter names (in macro order), which of them are written, and the op list.
    An op is (mnemonic, dst, a, b, c) for multiply-adds, (mnemonic, dst, a, b) for adds and plain multiplies; an operand is a
    parameter name or an int immediate."""

    def __init__(self, name, params, outs, ops, fresh=(), volatile=True):
        self.name, self.params, self.outs, self.ops, self.fresh, self.volatile = name, params, outs, ops, set(fresh), volatile

    def emit(self):
        order = [p for p in self.params if p in self.outs] + [p for p in self.params if p not in self.outs]
        num = {p: i for i, p in enumerate(order)}

        def opnd(x):
            return ("0x%08x" % x if x else "0") if isinstance(x, int) else "%%%d" % num[x]
        lines = ["%s.u32 %s" % (op[0], ", ".join(opnd(x) for x in op[1:])) for op in self.ops]
        body = ";\\n\\t".join(lines) + ";"
        cons_o = ", ".join('"%s"(%s)' % ("=&r" if p in self.fresh else "+r", p)
                           for p in order if p in self.outs)
        cons_i = ", ".join('"r"(%s)' % p for p in order if p not in self.outs)
        return "#define %s(%s) \\\n  asm%s(\"%s\" : %s : %s)\n" % (self.name, ", ".join(self.params),
                                                                  " volatile" if self.volatile else "", body, cons_o, cons_i)

    def emit_stmt(self):
        """the same chain as ONE asm statement whose operands are C expressions (array elements): for sequences that
        are emitted as a whole (the squaring) instead of being composed by hand in cios.cuh"""
        order = [p for p in self.params if p in self.outs] + [p for p in self.params if p not in self.outs]
        num = {p: i for i, p in enumerate(order)}

        def opnd(x):
            return ("0x%08x" % x if x else "0") if isinstance(x, int) else "%%%d" % num[x]
        body = ";\\n\\t".join("%s.u32 %s" % (op[0], ", ".join(opnd(x) for x in op[1:])) for op in self.ops) + ";"
        cons_o = ", ".join('"+r"(%s)' % p for p in order if p in self.outs)
        cons_i = ", ".join('"r"(%s)' % p for p in order if p not in self.outs)
        return 'asm volatile("%s" : %s : %s);' % (body, cons_o, cons_i)

    def run(self, env, cf_in=0):
        """interpret on a dict name -> u32 (ints are immediates); mutates env; returns the carry flag at the end"""
        cf = cf_in
        val = lambda x: x if isinstance(x, int) else env[x]  # noqa: E731
        for op in self.ops:
            m, d = op[0], op[1]
            if m in ("add.cc", "addc.cc", "addc"):
                s = val(op[2]) + val(op[3]) + (cf if m != "add.cc" else 0)
                env[d] = s & M32
                if m != "addc":
                    cf = s >> 32
            elif m in ("mul.lo", "mul.hi"):
                pr = val(op[2]) * val(op[3])
                env[d] = (pr & M32) if m == "mul.lo" else (pr >> 32)
            else:
                base, flags = m.split(".")[0], m.split(".")[1:]
                pr = val(op[2]) * val(op[3])
                part = (pr & M32) if flags[0] == "lo" else (pr >> 32)
                s = part + val(op[4]) + (cf if base == "madc" else 0)
                env[d] = s & M32
                if "cc" in flags:
                    cf = s >> 32
                else:
                    assert s >> 32 == 0, "carry out of a chain end in %s" % self.name
        return cf


def rows(p, n):
    """all macros of one curve, keyed by name"""
    h = n // 2
    n0 = (-pow(p, -1, 1 << 32)) & M32
    pl = [(p >> (32 * i)) & M32 for i in range(n)]
    allr = list(range(n))
    ev, od = list(range(0, n, 2)), list(range(1, n, 2))
    E = ["e%d" % i for i in allr]
    O = ["o%d" % i for i in allr]
    out = {}

    def add(r):
        out[r.name] = r

    def chain(dst, mult, mulb, addend, first_plain, last_open=False, swap=False):
        """pairs of wide MACs: (dst[2k], dst[2k+1]) = mult[k] * mulb + (addend[2k], addend[2k+1]) along one carry chain;
        first_plain: the chain starts here (mad, no carry in); last_open: the last high half closes the chain (no carry out)"""
        ops = []
        for k in range(len(mult)):
            lo = "mad.lo.cc" if (k == 0 and first_plain) else "madc.lo.cc"
            hi = "madc.hi" if (last_open and k == len(mult) - 1) else "madc.hi.cc"
            x, y = (mulb, mult[k]) if swap else (mult[k], mulb)     # immediates go second, as ptxas likes them
            ops.append((lo, dst[2 * k], x, y, addend[2 * k]))
            ops.append((hi, dst[2 * k + 1], x, y, addend[2 * k + 1]))
        return ops

    a_od, a_ev = ["a%d" % i for i in od], ["a%d" % i for i in ev]
    shifted = O[2:] + [0, 0]     # the odd array two limbs down, zeros entering at the top
    # --- ROW_ODD_RSHIFT: merge the stray limb e0 += o1, then O <- (O >> 2 limbs) + a_odd * b
    add(Row("PSB_ROW_ODD_RSHIFT", ["e0"] + O + a_od + ["b"], ["e0"] + O,
            [("add.cc", "e0", "e0", "o1")] + chain(O, a_od, "b", shifted, False, True)))
    # --- ROW_EVEN: E += a_even * b, carry out into o(n-1)
    add(Row("PSB_ROW_EVEN", E + [O[-1]] + a_ev + ["b"], E + [O[-1]],
            chain(E, a_ev, "b", E, True) + [("addc", O[-1], O[-1], 0)]))
    # --- ROW_ODD: O += a_odd * b in place
    add(Row("PSB_ROW_ODD", O + a_od + ["b"], O, chain(O, a_od, "b", O, True, True)))
    # --- RED_ODD / RED_EVEN: T += m p (modulus limbs as immediates)
    add(Row("PSB_RED_ODD", O + ["m"], O, chain(O, [pl[i] for i in od], "m", O, True, True, swap=True)))
    add(Row("PSB_RED_EVEN", E + [O[-1], "m"], E + [O[-1]],
            chain(E, [pl[i] for i in ev], "m", E, True, swap=True) + [("addc", O[-1], O[-1], 0)]))
    # --- stand-alone Montgomery reduction rows (csrc/cios.cuh redc_rr): the shift of the window is folded into the
    #     m p_odd chain of the NEXT row, whose m comes from the merged low limb:  e0 += o1;  m = e0 * (-1/p);
    #     O <- (O >> 2 limbs) + m p_odd
    add(Row("PSB_RED_ODD_RSHIFT", ["m", "e0"] + O, ["m", "e0"] + O,
            [("add.cc", "e0", "e0", "o1"), ("mul.lo", "m", "e0", n0)] + chain(O, [pl[i] for i in od], "m", shifted, False, True, swap=True),
            fresh=["m"]))
    # --- first rows: plain products
    for nm, arr, mult in (("PSB_ROW_FIRST_EVEN", E, a_ev), ("PSB_ROW_FIRST_ODD", O, a_od)):
        ops = []
        for k in range(h):
            ops += [("mul.lo", arr[2 * k], mult[k], "b"), ("mul.hi", arr[2 * k + 1], mult[k], "b")]
        add(Row(nm, arr + mult + ["b"], arr, ops, fresh=arr, volatile=False))
    ops = []
    for k in range(h):
        ops += [("mul.lo", O[2 * k], "m", pl[od[k]]), ("mul.hi", O[2 * k + 1], "m", pl[od[k]])]
    add(Row("PSB_RED_FIRST_ODD", O + ["m"], O, ops, fresh=O, volatile=False))
    return out


def sqr_rows(n):
    """cross products of a square, sum_{i<j} a_i a_j 2^(32 (i+j)), on two accumulator arrays of 2n limbs: EV holds the
    limb pairs at even positions, OD the pairs at odd positions (OD[k] sits at position k + 1), so that every 64-bit
    product lands on an aligned pair and a row is two carry chains of wide MACs -- the layout of the multiplier rows.
    Each chain ends by absorbing its carry in the next limb of its array (rows run in increasing i: that limb holds
    at most earlier absorbed carries, never a product)."""
    out = []
    for i in range(n - 1):
        for arr, js in (("OD", list(range(i + 1, n, 2))), ("EV", list(range(i + 2, n, 2)))):
            if not js:
                continue
            idx = (lambda pos: pos - 1) if arr == "OD" else (lambda pos: pos)
            ops, outs = [], []
            for k, j in enumerate(js):
                lo, hi = "%s[%d]" % (arr, idx(i + j)), "%s[%d]" % (arr, idx(i + j + 1))
                ops.append(("mad.lo.cc" if k == 0 else "madc.lo.cc", lo, "a[%d]" % i, "a[%d]" % j, lo))
                ops.append(("madc.hi.cc", hi, "a[%d]" % i, "a[%d]" % j, hi))
                outs += [lo, hi]
            top = "%s[%d]" % (arr, idx(i + js[-1] + 2))
            ops.append(("addc", top, top, 0))
            outs.append(top)
            out.append(Row("sqr_row_%d_%s" % (i, arr), outs + ["a[%d]" % i] + ["a[%d]" % j for j in js], outs, ops))
    return out


def sqr_diag_rows(n):
    """T += sum_i a_i^2 2^(64 i) along ONE carry chain over the 2n limbs, split into statements of <= 6 wide MACs"""
    out = []
    for c0 in range(0, n, 6):
        ops, outs, ins = [], [], []
        for i in range(c0, min(n, c0 + 6)):
            first, last = i == 0, i == n - 1
            lo, hi = "T[%d]" % (2 * i), "T[%d]" % (2 * i + 1)
            ops.append(("mad.lo.cc" if first else "madc.lo.cc", lo, "a[%d]" % i, "a[%d]" % i, lo))
            ops.append(("madc.hi" if last else "madc.hi.cc", hi, "a[%d]" % i, "a[%d]" % i, hi))
            outs += [lo, hi]
            ins.append("a[%d]" % i)
        out.append(Row("sqr_diag_%d" % c0, outs + ins, outs, ops))
    return out


def gen_sqr(n):
    lines = ["// cross products of a square on the EV / OD accumulator arrays (2N limbs each, zero on entry), see tools/gen_cios.py sqr_rows",
             "#define PSB_SQR_CROSS(a, EV, OD) do { \\"]
    lines += ["  %s \\" % r.emit_stmt() for r in sqr_rows(n)]
    lines += ["} while (0)", "// T (2N limbs) += sum a_i^2 2^(64 i)   (one carry chain across the statements)",
              "#define PSB_SQR_DIAG(a, T) do { \\"]
    lines += ["  %s \\" % r.emit_stmt() for r in sqr_diag_rows(n)]
    lines += ["} while (0)", ""]
    return "\n".join(lines)


def gen(p, n):
    return ("// GENERATED by tools/gen_cios.py -- even/odd row chains of the Montgomery multiplier (device only).\n#pragma once\n"
            + "".join(r.emit() for r in rows(p, n).values()) + gen_sqr(n))


# ---- the compositions of csrc/cios.cuh on the interpreted rows (used by tests/test_cios_model.py) ----------------
class Model:
    def __init__(self, curve):
        c = CURVES[curve]
        self.p, self.n = c["p"], c["n"]
        self.R = 1 << (32 * self.n)
        self.n0 = (-pow(self.p, -1, 1 << 32)) & M32
        self.rows = rows(self.p, self.n)

    def limbs(self, x, cnt=None):
        return [(x >> (32 * i)) & M32 for i in range(cnt or self.n)]

    @staticmethod
    def value(v):
        return sum(x << (32 * i) for i, x in enumerate(v))

    def _run(self, name, binding):
        """binding: macro parameter -> (list, index) reference or int value"""
        r = self.rows[name]
        env = {}
        for prm in r.params:
            b = binding[prm]
            env[prm] = b[0][b[1]] if isinstance(b, tuple) else b
        r.run(env)
        for prm in r.outs:
            b = binding[prm]
            b[0][b[1]] = env[prm]

    def _bind(self, E=None, O=None, a=None, **kw):
        n = self.n
        b = dict(kw)
        if E is not None:
            b.update({"e%d" % i: (E, i) for i in range(n)})
        if O is not None:
            b.update({"o%d" % i: (O, i) for i in range(n)})
        if a is not None:
            b.update({"a%d" % i: a[i] for i in range(n)})
        return b

    # cios.cuh: first / mac_shift / mac / reduce
    def first(self, E, O, a, b):
        self._run("PSB_ROW_FIRST_EVEN", self._bind(E=E, a=a, b=b))
        self._run("PSB_ROW_FIRST_ODD", self._bind(O=O, a=a, b=b))

    def mac_shift(self, E, O, a, b):
        bd = self._bind(O=O, a=a, b=b)
        bd["e0"] = (E, 0)
        self._run("PSB_ROW_ODD_RSHIFT", bd)
        self._run("PSB_ROW_EVEN", self._bind(E=E, a=a, b=b, **{"o%d" % (self.n - 1): (O, self.n - 1)}))

    def mac(self, E, O, a, b):
        self._run("PSB_ROW_ODD", self._bind(O=O, a=a, b=b))
        self._run("PSB_ROW_EVEN", self._bind(E=E, a=a, b=b, **{"o%d" % (self.n - 1): (O, self.n - 1)}))

    def reduce(self, E, O):
        m = (E[0] * self.n0) & M32
        self._run("PSB_RED_ODD", self._bind(O=O, m=m))
        self._run("PSB_RED_EVEN", self._bind(E=E, m=m, **{"o%d" % (self.n - 1): (O, self.n - 1)}))

    def finish_nored(self, E, O):
        """(E >> 32) + O as one integer (n limbs + possible carry reported)"""
        return self.value(E[1:]) + self.value(O)

    def mul(self, a, b):
        """cios::mul_rr -- returns the UNCANONICALISED sum (< 2p)"""
        n = self.n
        X, Y = [0] * n, [0] * n
        self.first(X, Y, a, b[0])
        self.reduce(X, Y)
        for i in range(1, n):
            if i & 1:
                self.mac_shift(Y, X, a, b[i]); self.reduce(Y, X)
            else:
                self.mac_shift(X, Y, a, b[i]); self.reduce(X, Y)
        return self.finish_nored(Y, X)

    def dot2(self, a, b, c, d):
        """cios::dot2_rr -- (a b + c d) / R, one reduction per row; returns the UNCANONICALISED sum"""
        n = self.n
        X, Y = [0] * n, [0] * n
        self.first(X, Y, a, b[0]); self.mac(X, Y, c, d[0]); self.reduce(X, Y)
        for i in range(1, n):
            if i & 1:
                self.mac_shift(Y, X, a, b[i]); self.mac(Y, X, c, d[i]); self.reduce(Y, X)
            else:
                self.mac_shift(X, Y, a, b[i]); self.mac(X, Y, c, d[i]); self.reduce(X, Y)
        return self.finish_nored(Y, X)

    def mulpre(self, a, b):
        """cios::mulpre_rr -- the 2n-limb product, low limbs peeled off row by row"""
        n = self.n
        X, Y = [0] * n, [0] * n
        T = []
        self.first(X, Y, a, b[0])
        T.append(X[0])
        for i in range(1, n):
            if i & 1:
                self.mac_shift(Y, X, a, b[i]); T.append(Y[0])
            else:
                self.mac_shift(X, Y, a, b[i]); T.append(X[0])
        hi = self.finish_nored(Y, X)
        assert hi < self.R
        return T + self.limbs(hi)

    def sqrpre(self, a):
        """cios::sqrpre_rr -- the 2n-limb square: cross products once on EV / OD, merged, doubled, plus the diagonal.
        NOTE the diagonal chain carries across statements (the interpreter threads the flag like the hardware does)."""
        n = self.n
        env = {"a[%d]" % i: a[i] for i in range(n)}
        env.update({"EV[%d]" % k: 0 for k in range(2 * n + 1)})
        env.update({"OD[%d]" % k: 0 for k in range(2 * n + 1)})
        for r in sqr_rows(n):
            r.run(env)
        assert env["EV[%d]" % (2 * n)] == 0 and env["OD[%d]" % (2 * n)] == 0 and env["OD[%d]" % (2 * n - 1)] == 0
        cross = self.value([env["EV[%d]" % k] for k in range(2 * n)]) + (self.value([env["OD[%d]" % k] for k in range(2 * n - 1)]) << 32)
        assert cross < 1 << (64 * n - 1)
        t = 2 * cross                                  # merge + doubling are plain carry chains in cios.cuh
        env.update({"T[%d]" % k: (t >> (32 * k)) & M32 for k in range(2 * n)})
        cf = 0
        for r in sqr_diag_rows(n):
            cf = r.run(env, cf_in=cf)
        return [env["T[%d]" % k] for k in range(2 * n)]

    def redc(self, T):
        """cios::redc_rr -- (T_lo + M p) / R + T_hi, uncanonicalised"""
        n = self.n
        X, Y = list(T[:n]), [0] * n          # X: even array = the low half; Y: odd array
        m = (X[0] * self.n0) & M32
        self._run("PSB_RED_FIRST_ODD", self._bind(O=Y, m=m))
        self._run("PSB_RED_EVEN", self._bind(E=X, m=m, **{"o%d" % (n - 1): (Y, n - 1)}))
        for i in range(1, n):
            E, O = (Y, X) if i & 1 else (X, Y)
            mm = [0]
            bd = self._bind(O=O)
            bd["e0"] = (E, 0)
            bd["m"] = (mm, 0)
            self._run("PSB_RED_ODD_RSHIFT", bd)
            self._run("PSB_RED_EVEN", self._bind(E=E, m=mm[0], **{"o%d" % (n - 1): (O, n - 1)}))
        E, O = (Y, X)                         # after row n-1 (odd index): even array = Y
        assert E[0] == 0
        return self.finish_nored(E, O) + self.value(T[n:])


if __name__ == "__main__":
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ps-signature-and-el-passo_b200", "csrc")
    for name, c in CURVES.items():
        path = os.path.join(base, c["file"])
        with open(path, "w") as f:
            f.write(gen(c["p"], c["n"]))
        print("wrote", os.path.normpath(path))
