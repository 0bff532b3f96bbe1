"""Build the CPU oracle's C kernels (test infrastructure, never the product).

``python -m oracle.build`` (or :func:`build`) compiles ``oracle_tree.c`` and
``oracle_trav.c`` twice -- once per coordinate dtype -- into
``oracle/_build/liboracle_{f32,f64}.so`` with strict IEEE flags
(``-ffp-contract=off``, no fast-math).

There is no ``oracle/_ref``: the reference is Python + run-time generated OpenCL
and cannot be compiled or imported in this image (no pyopencl / OpenCL ICD /
mako), see DESIGN.md.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
SOURCES = ["oracle_tree.c", "oracle_trav.c"]


def lib_path(tag: str) -> str:
    return os.path.join(BUILD_DIR, f"liboracle_{tag}.so")


def _needs_rebuild(out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(HERE, s)) > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> None:
    os.makedirs(BUILD_DIR, exist_ok=True)
    for tag, define in (("f32", "-DCOORD_F32"), ("f64", "-DCOORD_F64")):
        out = lib_path(tag)
        if not force and not _needs_rebuild(out):
            continue
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-std=gnu11",
               "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wno-unused-function", "-Wno-maybe-uninitialized",
               define, *[os.path.join(HERE, s) for s in SOURCES], "-o", out, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
