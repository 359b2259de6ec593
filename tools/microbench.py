#!/usr/bin/env python3
"""tools/microbench.py -- device-timed micro-benchmarks (run under gpurun): the integer-multiply
issue rate of this B200 (the roofline denominator for every kernel here) and the
throughput of the engine's Fp / Fp2 / Fp12 multiplications at several occupancies.
Writes one JSON object to stdout (and to gpurun_out/microbench.json when that directory exists)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
pkg.init([0])
SM = 148
res = {"mad": [], "field": []}
for kind, name, per_iter in [(6, "imad32 (mul.lo+add)", 16), (4, "mad.lo.cc+madc.hi (MAC32)", 8), (5, "mad.wide.u32 with BOTH factors shared: ptxas folds it into 1 IMAD.WIDE + 8 64-bit adds per trip -- NOT a MAC rate", 8),
                             (7, "carry-chained IMAD.WIDE.U32.X rows (CIOS form, MAC32)", 36),
                             (8, "fma.rz.f64 (DFMA)", 8), (9, "mad.wide.u32 rows of 13 distinct limbs (radix-2^30 form, MAC)", 13),
                             (10, "mad.wide.u32, 16 accumulators, distinct multiplicands, shared multiplier (carry-free MAC32, not foldable)", 16),
                             (11, "mad.wide.u32, 16 accumulators, distinct multiplicands and multipliers (carry-free MAC32)", 16)]:
    for bps, thr in [(1, 256), (2, 256), (4, 256), (8, 256), (4, 128), (1, 128)]:
        iters = 20000
        ms = pkg.microbench(kind, SM * bps, thr, iters)
        ops = SM * bps * thr * iters * per_iter
        res["mad"].append({"probe": name, "blocks_per_sm": bps, "threads": thr, "ms": ms,
                           "Tops_per_s": ops / ms / 1e9, "ops_per_clk_per_sm_at_1965MHz": ops / (ms * 1e-3) / SM / 1.965e9})
for kind, name, iters in [(0, "fp_mul", 2000), (1, "fp_sqr", 2000), (2, "fp2_mul", 1000), (3, "fp12_mul", 50)]:
    for bps, thr in [(1, 128), (1, 256), (2, 128), (2, 256), (3, 128), (4, 128), (4, 256), (8, 128)]:
        ms = pkg.microbench(kind, SM * bps, thr, iters)
        ops = SM * bps * thr * iters
        res["field"].append({"op": name, "blocks_per_sm": bps, "threads": thr, "ms": ms, "Gops_per_s": ops / ms / 1e6})
out = json.dumps(res, indent=1)
print(out)
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
    open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w").write(out)
