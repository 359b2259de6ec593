"""Host-side logic of the N > 1 path, on CPU: (a) the in-process lane split every batched call uses
(psb_shard_range: contiguous, balanced, exact cover -- SURVEY.md 8e), (b) bench.py's max-over-ranks timing and
whole-job throughput under a world_size-2 gloo process group (the GPU box uses the same code over NCCL)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_exactly(pkg):
    for N in (0, 1, 7, 8, 1000, 1 << 20, (1 << 24) + 5):
        for G in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for k in range(G):
                b, e = pkg.shard_range(N, G, k)
                assert b == prev and e >= b
                sizes.append(e - b)
                prev = e
            assert prev == N
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(pkg.PsbError):
        pkg.shard_range(10, 2, 2)


def test_bench_rank_aggregation_gloo(tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import json, os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        # rank r "measured" 100 + 10 r ms on the device arm and 200 - 5 r ms end to end
        ms, e2e = bench.max_over_ranks([100.0 + 10 * rank, 200.0 - 5 * rank], world)
        val = bench.job_throughput(1 << 20, world, 2, ms)
        if rank == 0:
            print(json.dumps({{"ms": ms, "e2e": e2e, "value": val, "world": world}}))
        dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["ms"] == 110.0 and d["e2e"] == 200.0
    assert abs(d["value"] - 2 * (1 << 20) * 2 / 0.110) < 1e-6


def test_bench_tile_packed_and_fr_rules():
    """host arithmetic of the workload generator: repeating a packed string list, Fr::setHashOf's mask rule on both curves."""
    import hashlib
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    strs = [b"a0:0", b"", b"xyz", b"\xff" * 300]
    off = np.zeros(len(strs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in strs])
    blob = np.frombuffer(b"".join(strs) + b"\0" * 8, dtype=np.uint8).copy()
    b3, o3 = bench.tile_packed(blob, off, 3)
    got = [bytes(b3[int(o3[i]):int(o3[i + 1])]) for i in range(len(o3) - 1)]
    assert got == strs * 3 and len(b3) == 3 * int(off[-1]) + 8
    assert bench.tile_packed(blob, off, 1)[1] is off
    for curve, bits in (("bls12_381", 255), ("bn254", 254)):
        r = bench.R_ORDERS[curve]
        assert r.bit_length() == bits
        old = bench.R_ORDER
        bench.R_ORDER = r
        try:
            for m in (b"a0:0", b"attr", b""):
                x = int.from_bytes(hashlib.sha256(m).digest(), "little") & ((1 << bits) - 1)
                assert bench.fr_hash(m) == (x if x < r else x & ((1 << (bits - 1)) - 1)) < r
            raw = np.frombuffer(((12345 << 256) % r).to_bytes(32, "little"), dtype=np.uint64)
            assert bench.fr_val(raw) == 12345 and np.array_equal(bench.fr_mont([12345])[0], raw)
        finally:
            bench.R_ORDER = old


def test_bench_config_ranks_side_group_gloo(tmp_path):
    """the configs' coordination (bench.Ranks over a gloo side group): MAX over ranks, and a rank that fails inside a config
    makes the OTHER rank's next collective time out instead of hanging -- both end with an error entry and exit 0."""
    script = tmp_path / "ranks.py"
    script.write_text(textwrap.dedent(f"""
        import datetime, json, os, sys, types
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        side = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=5))
        rk = bench.Ranks(rank, world, None, None, side)
        out, dt = rk.timed(lambda: 7, 2)
        red = rk.max_ms([10.0 + rank, 3.0 - rank])
        def good():
            rk.barrier()
            return {{"e2e_value": rk.max_ms([1.0 + rank])[0]}}
        def bad():
            if rank == 1:
                raise RuntimeError("boom")
            rk.barrier()
            return {{}}
        args = types.SimpleNamespace(configs="cfg3,cfg4,cfg5", steps=1, no_cpu_baseline=True, config_window_bits=8, config_lanes=8,
                                     window_bits50=8, lanes50=8)
        jobs = iter([good, bad, good])
        bench.cfg_signon = lambda *a, **k: next(jobs)()
        bench.cfg_issuance = lambda *a, **k: next(jobs)()
        bench.cfg_verify50 = lambda *a, **k: next(jobs)()
        res = bench.run_configs(None, args, rk, None)
        print(json.dumps({{"rank": rank, "red": red, "out": out, "res": {{k: sorted(v) for k, v in res.items()}},
                          "first": res["cfg3_signon"].get("e2e_value")}}), flush=True)
        os._exit(0)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29519", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 2
    for d in lines:
        assert d["red"] == [11.0, 3.0] and d["out"] == 7 and d["first"] == 2.0
        assert "error" in d["res"]["cfg4_issuance"] and "error" in d["res"]["cfg5_verify50"]
