#!/usr/bin/env python3
"""tools/prof_verify.py -- a short device-resident verify loop for ncu (never a bench number)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
wbits = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pkg = ge.load_package(); pkg.init([0])
key = bench.load_key(5)
pk = pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=wbits)
sig1, sig2, blob, off, expected = bench.make_batch(pkg, key, lanes, 0, base=min(lanes, 4096))
ver = pkg.PSVerifier(pk)
for _ in range(steps):
    v = ver.verify(sig1, sig2, (blob, off))
assert np.array_equal(v, expected)
print("ok", lanes, int(v.sum()))
