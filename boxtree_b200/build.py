"""Build ``libboxtree_b200.so`` in-tree with nvcc for sm_100a.

``python -m boxtree_b200.build`` (or :func:`build`) cross-compiles without a
GPU.  ``-fmad=false`` keeps every float expression un-contracted so results
match the CPU oracle's IEEE evaluation bit for bit ("Float
semantics").
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libboxtree_b200.so")
SOURCES = ["tree_build.cu", "traversal.cu", "distributed.cu"]
HEADERS = ["common.cuh", "scan.cuh", "radix_sort.cuh", os.path.join("..", "..", "include", "boxtree_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_rebuild() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and not needs_rebuild():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags,
           *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    flags = ["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else []
    build(force=True, verbose=True, extra_flags=flags)
