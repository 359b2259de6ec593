#!/usr/bin/env python3
"""tools/ab_sqr.py -- A/B of the dedicated Montgomery squaring (run under gpurun with PSB_LIB=<build>): the paths whose
time is square-root / Legendre chains of Fp squarings -- G1 decompression (psb_g1_deserialize), hashAndMapToG1, the
wire-format verification (psb_verify_ser) -- and the raw fp_mul / fp_sqr micro-benchmarks.  Wall clock around the
blocking host-buffer calls, best of 3.  Experiments only."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
pkg = ge.load_package(); pkg.init([0])
key = bench.load_key(5)
pk = pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=16)
sig1, sig2, blob, off, expected, _ = bench.make_batch(pkg, key, N, 0, base=min(N, 4096))
rng = np.random.default_rng(5)
t = np.frombuffer(rng.bytes(32 * N), dtype=np.uint64).reshape(N, 4).copy(); t[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
_, _, ser = pkg.PSRequester.randomize_credential(sig1, sig2, t, want_serialized=True)
live = sig1[:, 12:].any(axis=1)


def best(fn, reps=3):
    b = None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); dt = time.perf_counter() - t0
        b = dt if b is None else min(b, dt)
    return out, b


pts = np.ascontiguousarray(ser.reshape(-1, pkg.G1_SER))
(_, ok), dt = best(lambda: pkg.g1_deserialize(pts))
res = {"lib": os.path.basename(os.environ.get("PSB_LIB", "libpsb.so")), "lanes": N,
       "g1_deserialize_per_s": len(pts) / dt, "decoded_ok": int(ok.sum())}
msgs = pkg.pack_strings([b"svc%d" % i for i in range(N)])
(_, ok2), dt = best(lambda: pkg.hash_and_map_to_g1(msgs))
res["hash_to_g1_per_s"] = N / dt
ver = pkg.PSVerifier(pk)
cred = np.ascontiguousarray(np.concatenate([sig1, sig2], axis=1))   # unused; wire verify reads the serialized form of the ORIGINAL batch
_, _, ser0 = pkg.PSRequester.randomize_credential(sig1, sig2, np.tile(np.frombuffer(((1 << 256) % bench.R_ORDER).to_bytes(32, "little"), dtype=np.uint64), (N, 1)), want_serialized=True)
(v, dec), dt = best(lambda: ver.verify_serialized(ser0, (blob, off), stride=2 * pkg.G1_SER, off1=0, off2=pkg.G1_SER))
assert np.array_equal(v[live], expected[live]), "wire verdict mismatch"
res["verify_ser_per_s"] = N / dt
for kind, name, iters in [(0, "fp_mul", 2000), (1, "fp_sqr", 2000)]:
    ms = min(pkg.microbench(kind, 148 * 2, 256, iters) for _ in range(2))
    res[name + "_Gops"] = 148 * 2 * 256 * iters / ms / 1e6
print(json.dumps(res))
