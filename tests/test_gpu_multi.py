"""In-process multi-GPU sharding (psb_init with several devices: one worker thread + stream per device, verdict
bytes written to disjoint slices -- SURVEY.md 8e).  Needs >= 2 GPUs; runs in a subprocess because the session's
engine is initialised on device 0 only."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_verify_sharded_over_all_devices(tmp_path):
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "multi.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        import __graft_entry__ as ge
        from tests import workload
        pkg = ge.load_package()
        pkg.init(list(range({ndev})))
        assert pkg.lib().psb_num_devices() == {ndev}
        wl = workload.make_verify_workload(n_attrs=5, lanes=301, seed=13, tamper_every=7)
        pk = pkg.PSPubKey(wl.key.g, wl.key.gg, wl.key.XX, wl.key.Y, wl.key.YY, window_bits=8)
        v, gt = pkg.PSVerifier(pk).verify(wl.sig1, wl.sig2, wl.attrs, want_gt=True)
        ev, egt = workload.expected_verify(wl, want_gt=True)
        assert np.array_equal(v, ev)
        live = wl.sig1[:, 2 * pkg.FP:].any(axis=1)
        assert np.array_equal(gt[live], egt[live])
        sw = workload.make_signon_workload(5, 41, 2, seed=3, with_id=True, tamper_every=4)
        got = pkg.PSVerifier(pk).el_passo_verify_id(sw.proof, sw.proof_attrs, sw.ads, sw.service_pt, sw.y, sw.g, sw.h)
        assert got.tolist() == workload.expected_verify_id(sw).tolist()
        print("multi ok", {ndev})
    """))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-3000:]
    assert "multi ok" in r.stdout
