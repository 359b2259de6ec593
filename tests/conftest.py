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
    path = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build()
    return ctypes.CDLL(path)


def rand_fp_raw(ref, rng, n, k=1):
    """n x k random canonical Fp values as raw Montgomery limbs (n, 6k) u64."""
    from oracle import ps_oracle as O
    vals = [int.from_bytes(rng.bytes(48), "little") % O.P for _ in range(n * k)]
    return ref.fp_from_ints(vals).reshape(n, k * 6)
