"""pytest configuration: `gpu` marker, import paths, shared fixtures.

-m "not gpu": oracle vs golden vectors / live reference, host logic (hostsim), C-ABI export check.
-m gpu      : parity of the CUDA path (through the C ABI) against the oracle.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# One curve per process (mcl and the engine keep the curve in static state).  The default run is BLS12-381;
# tests/test_bn254.py re-runs the curve-generic tests in a child pytest with PSB_CURVE=bn254.
BN254 = os.environ.get("PSB_CURVE", "bls12_381").lower() == "bn254"
bls_only = pytest.mark.skipif(BN254, reason="BLS12-381-specific (python oracle / GLV constants / golden fixtures)")
if BN254:
    _z = -0x4080000000000001
    FIELD_P = 36 * _z ** 4 + 36 * _z ** 3 + 24 * _z ** 2 + 6 * _z + 1
    GROUP_R = 36 * _z ** 4 + 36 * _z ** 3 + 18 * _z ** 2 + 6 * _z + 1
else:
    _z = -0xD201000000010000
    GROUP_R = _z ** 4 - _z ** 2 + 1
    FIELD_P = (_z - 1) ** 2 * GROUP_R // 3 + _z


CURVE_Z = abs(_z)                                    # |z| of the curve family parameter
FPW = 4 if BN254 else 6                              # u64 words: Fp, then G1 / G2 / GT Jacobian / tower objects
G1W, G2W, GTW = 3 * FPW, 6 * FPW, 12 * FPW
FP_BYTES = 8 * FPW                                   # serialized Fp = compressed G1
G1_SER, G2_SER = FP_BYTES, 2 * FP_BYTES


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def gpu_pkg(pkg):
    pkg.init([0])
    return pkg


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as r
    if not r.available() and not r.build():
        pytest.skip("oracle/_ref/libpsref.so unavailable (needs /root/reference to build)")
    r.lib()
    return r


@pytest.fixture(scope="session")
def hostsim():
    path = os.path.join(ROOT, "tests", "hostsim", "libhostsim_bn254.so" if BN254 else "libhostsim.so")
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build()
    return ctypes.CDLL(path)


def rand_fp_raw(ref, rng, n, k=1):
    """n x k random canonical Fp values as raw Montgomery limbs (n, FP * k) u64."""
    vals = [int.from_bytes(rng.bytes(FP_BYTES), "little") % FIELD_P for _ in range(n * k)]
    return ref.fp_from_ints(vals).reshape(n, k * ref.FP)
