"""CPU: the C-ABI library loads and exports every symbol include/psb.h declares; the product has no
CPU path (calls fail loudly without a GPU) and never touches oracle/."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "psb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psb_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(pkg):
    L = pkg.lib()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/psb.h but not exported by libpsb.so"
    assert sorted(pkg.EXPORTS) == syms


def test_no_cpu_fallback(pkg):
    import ctypes as C
    L = pkg.lib()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    L.psb_last_error.restype = C.c_char_p
    assert L.psb_init(pkg.MCL_CURVE, None, 0) != 0  # no device -> error, never a silent CPU path
    assert b"no CUDA device" in L.psb_last_error()  # ... and the curve check passed: the library is built for ITS curve
    other = 0 if pkg.MCL_CURVE == 5 else 5
    assert L.psb_init(other, None, 0) == -4         # PSB_ERR_UNSUPPORTED: one curve per library
    assert b"is built for" in L.psb_last_error()
    assert L.psb_test_op(0, C.c_size_t(1), None, None, None, None) != 0


def test_product_does_not_reference_oracle():
    pkg_dir = os.path.join(ROOT, "ps-signature-and-el-passo_b200")
    for root, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle", "").replace("against the oracle", "") or \
                    "import" not in txt.split("oracle")[0][-40:], f
                assert "libpsref" not in txt and "hostsim" not in txt.replace('"hostsim"', "").replace("hostsim)", "") or f == "fp.cuh" or f == "testops.cuh", f
