"""The C++ drop-in (ps-signature-and-el-passo_b200/host/ps_batch.hpp): psb::PSSigner / PSRequester / PSVerifier derive
from the reference's classes and add batched overloads over the C ABI.  tests/host/test_ps_batch.cc runs the
reference's own EL PASSO flow (test/ps-tests.cc:53-137) and compares every batched GPU result with the reference's
scalar method lane by lane; it is linked against the unmodified reference objects by oracle/Makefile (`hosttest`)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "test_ps_batch")


def _bin():
    if not os.path.exists(BIN):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/test_ps_batch not built (needs /root/reference)")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hosttest"], stdout=subprocess.DEVNULL)
    return BIN


def test_cpp_dropin_fails_loudly_without_gpu():
    """no CPU fallback behind the batched overloads: psb::init must throw when there is no device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([_bin(), "4"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_dropin_matches_reference_flow():
    r = subprocess.run([_bin(), "24"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
