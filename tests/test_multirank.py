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
