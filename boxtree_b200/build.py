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
SOURCES = ["tree_build.cu", "traversal.cu", "distributed.cu", "consumers.cu"]
HEADERS = ["common.cuh", "scan.cuh", "radix_sort.cuh", os.path.join("..", "..", "include", "boxtree_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC",
]
OBJDIR = os.path.join(HERE, "..", "build", "obj")


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
    # one object per translation unit, compiled concurrently; only stale objects are rebuilt
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS
                if os.path.exists(os.path.join(CSRC, h)))
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        src_path = os.path.join(CSRC, src)
        stale = force or extra_flags or not os.path.exists(obj) or \
            os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(src_path))
        if stale:
            cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", src_path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", *objs, "-o", LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    flags = ["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else []
    build(force="--force" in sys.argv or bool(flags), verbose=True, extra_flags=flags)
