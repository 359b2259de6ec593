import sys, json
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package(); pkg.init([0])
geom = (148 * 8, 256)
for _ in range(3): pkg.microbench(7, *geom, 5000)
out = {}
for kind, name in ((7, "36 MAC32 (3 carry-chained rows)"), (12, "36 MAC32 + 24 DFMA"), (13, "36 MAC32 + 72 DFMA"), (14, "72 DFMA alone")):
    ms = min(pkg.microbench(kind, *geom, 5000) for _ in range(3))
    out[name] = ms
    print(name, "%.3f ms" % ms, "%.1f MAC32/clk/SM" % (148*8*256*5000*36/(ms*1e-3)/148/1.965e9) if kind != 14 else "%.1f DFMA/clk/SM" % (148*8*256*5000*72/(ms*1e-3)/148/1.965e9))
json.dump(out, open('/root/repo/gpurun_out/r2k_dfma_beside_imad.json', 'w'), indent=1)
