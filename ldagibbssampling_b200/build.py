"""In-tree build of libb200lda.so (hand-written CUDA for sm_100a) with nvcc.

`python -m ldagibbssampling_b200.build` or `__graft_entry__.build()`. nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libb200lda.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the sampling spec rounds every fp32 op individually (no FMA contraction); the kernels also
    # use __f*_rn intrinsics on every decision path, this flag covers the rest of the TU
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200lda.so cannot be built (there is no CPU fallback)")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(INCLUDE, "b200lda.h")]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, out: str = None) -> str:
    """out: write the library somewhere else (tuning variants, see tools/build_variants.sh)."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("B200LDA_NVCC_EXTRA", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", INCLUDE, "-o", out or LIB_PATH, os.path.join(CSRC, "b200lda.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libb200lda.so")
    return out or LIB_PATH


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out))
