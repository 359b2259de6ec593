"""build.py -- compiles the CUDA engine in-tree (sm_100a only): csrc/psb_api.cu -> libpsb.so (BLS12-381) and, with
-DPSB_BUILD_BN254, libpsb_bn254.so (BN254: same sources, 8-limb field, D-type twist, BN Miller loop / final
exponentiation).  One curve per library, like mcl's own bn256 / bn384 builds; the C ABI (include/psb.h) is the same.

nvcc cross-compiles without a GPU; the resulting .so travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpsb.so")
LIB_BN254 = os.path.join(HERE, "libpsb_bn254.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(HERE, "..", "include", "psb.h"))
    return out


def stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(s) > t for s in _sources())


def _compile(out: str, defs, verbose: bool):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PSB_EXTRA_FLAGS", "").split()
    cmd = [nvcc] + NVCC_FLAGS + defs + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, "psb_api.cu"), "-o", out]
    return subprocess.Popen(cmd, cwd=CSRC)


def build(force: bool = False, verbose: bool = False) -> str:
    """both curve libraries, compiled side by side"""
    jobs = []
    if force or stale(LIB):
        jobs.append(_compile(os.environ.get("PSB_LIB_OUT", LIB), [], verbose))
    if (force or stale(LIB_BN254)) and not os.environ.get("PSB_SKIP_BN254"):
        jobs.append(_compile(LIB_BN254, ["-DPSB_BUILD_BN254"], verbose))
    for j in jobs:
        if j.wait() != 0:
            raise subprocess.CalledProcessError(j.returncode, j.args)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
