"""ps-signature-and-el-passo_b200 -- B200-native batch engine for PS signatures / EL PASSO.

csrc/      hand-written sm_100a CUDA (field tower, curves, pairing, kernels) + the C ABI (include/psb.h)
host/      C++ mirror of the reference classes with batched overloads (PSVerifier, PSRequester, PSSigner)
engine.py  ctypes binding of libpsb.so + Python mirror of the same interface (tests / bench plumbing)

The directory name is not a Python identifier; import it through __graft_entry__.load_package().
"""
from .engine import *  # noqa: F401,F403
