#!/usr/bin/env python3
"""tools/time_pairing.py -- repeated end-to-end psb_pairing calls at one size (diagnostic, not a bench number)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
pkg = ge.load_package(); pkg.init([0])
key = bench.load_key(5)
rng = np.random.default_rng(77)
kk = np.frombuffer(rng.bytes(32 * NP), dtype=np.uint64).reshape(NP, 4).copy()
kk[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
Pp = pkg.g1_mul(key["g"], kk)
Qq = np.ascontiguousarray(np.tile(key["YY"], (NP // 5 + 1, 1))[:NP])
for i in range(5):
    t0 = time.perf_counter(); pkg.pairing(Pp, Qq); dt = time.perf_counter() - t0
    print("call", i, "%.1f ms" % (dt * 1e3), "%.0f pairings/s" % (NP / dt))
