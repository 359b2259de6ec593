import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
import __graft_entry__ as ge
from oracle import ref
pkg = ge.load_package(); pkg.init([0])
N = 1 << 18
msgs = [b"svc%d.example.org" % i for i in range(N)]
packed = pkg.pack_strings(msgs)
pkg.hash_and_map_to_g1(packed)
t0 = time.perf_counter(); out, ok = pkg.hash_and_map_to_g1(packed); dt = time.perf_counter() - t0
D = 2048
t0 = time.perf_counter(); want = np.stack([ref.hash_to_g1(m) for m in msgs[:D]]); cpu = time.perf_counter() - t0
assert ok.all() and np.array_equal(out[:D], ref.g1_op(ref.G_NORM, want))
print(json.dumps({"config": "hashAndMapToG1 (8f-4)", "lanes": N, "e2e_value": N / dt, "metric": "points_per_sec", "seconds": dt,
                  "cpu_reference": {"value": D / cpu, "cores": 1, "sample": f"{D} messages, mcl hashAndMapToG1"},
                  "parity": "normalised points identical to mcl on the sample"}))
