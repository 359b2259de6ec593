"""The C++ drop-in (ps-signature-and-el-passo_b200/host/ps_batch.hpp): psb::PSSigner / PSRequester / PSVerifier derive
from the reference's classes and add batched overloads over the C ABI.  tests/host/test_ps_batch.cc runs the
reference's own EL PASSO flow (test/ps-tests.cc:53-137) and compares every batched GPU result with the reference's
scalar method lane by lane; it is linked against the unmodified reference objects by oracle/Makefile (`hosttest`)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "test_ps_batch")
BIN_BN254 = BIN + "_bn254"   # same source, mcl's 256-bit configuration + initPairing() = BN254, linked with libpsb_bn254.so


def _bin(path=BIN):
    if not os.path.exists(path):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/test_ps_batch not built (needs /root/reference)")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hosttest"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return path


def test_cpp_dropin_fails_loudly_without_gpu():
    """no CPU fallback behind the batched overloads: psb::init must throw when there is no device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for path in (BIN, BIN_BN254):
        r = subprocess.run([_bin(path), "4"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 3, r.stdout + r.stderr
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("path", [BIN, BIN_BN254], ids=["bls12_381", "bn254"])
def test_cpp_dropin_matches_reference_flow(path):
    r = subprocess.run([_bin(path), "24"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
