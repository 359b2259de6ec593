"""BN254 instantiation: the curve `initPairing()` selects in the reference's shipped tests and WASM
build.  One curve per process -- mcl and the engine keep the curve in static state, and each curve is its own
library (libpsb_bn254.so / libhostsim_bn254.so / oracle/_ref/libpsref_bn254.so, all built from the same sources with
-DPSB_BUILD_BN254 resp. mcl's 256-bit configuration) -- so the curve-generic test modules are re-run in a child pytest
with PSB_CURVE=bn254.  Golden-fixture and Python-oracle tests are BLS12-381-only (`bls_only`) and skip themselves there.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU_MODULES = ["test_hostsim.py", "test_hostsim_protocol.py", "test_hostsim_prover.py", "test_hostsim_wire.py", "test_hash_to_curve.py", "test_abi.py"]
GPU_MODULES = ["test_gpu_arith.py", "test_gpu_verify.py", "test_gpu_protocol.py", "test_gpu_elpasso.py", "test_gpu_prover.py",
               "test_gpu_wire.py", "test_gpu_configs.py", "test_hash_to_curve.py"]


def _child(modules, marker, min_passed):
    if os.environ.get("PSB_CURVE", "").lower() == "bn254":
        pytest.skip("already inside the BN254 child run")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpsref_bn254.so")) and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libpsref_bn254.so unavailable (needs /root/reference to build)")
    env = dict(os.environ, PSB_CURVE="bn254")
    env.pop("PSB_LIB", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", marker, "-p", "no:cacheprovider"] +
                       [os.path.join(ROOT, "tests", m) for m in modules], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    passed = int(r.stdout.strip().splitlines()[-1].split(" passed")[0].split()[-1])
    assert passed >= min_passed, tail


def test_bn254_host_logic():
    """field / tower / curves / pairing / NIZK lanes / decompression / hashAndMapToG1 on the CPU hostsim vs mcl (BN254)."""
    _child(CPU_MODULES, "not gpu", 25)


@pytest.mark.gpu
def test_bn254_gpu_parity():
    """the same parity suite as BLS12-381, through libpsb_bn254.so's C ABI on the GPU, vs the BN254 build of the reference."""
    _child(GPU_MODULES, "gpu", 40)
