"""Run the reference's own code -- host drivers AND kernel sources, unmodified, from where they lie
under ``/root/reference`` -- on the CPU of this container, to pin ``oracle/`` against it.

The reference needs pyopencl (an OpenCL runtime), mako, pytools, arraycontext and cgen, none of
which is installed here, so its device code cannot run as shipped.  This package supplies
stand-ins for those third-party imports (``fakecl.py``, ``minimako.py``, ``cl_shim.h``) that
execute each kernel one work item at a time with g++-compiled code generated from the kernel text
the reference renders itself.  What is the reference's: every template, every kernel body, the
``FMMTraversalBuilder.__call__`` / ``TreeBuilder.__call__`` host logic.  What is restated here:
pyopencl's published semantics of ``ListOfListsBuilder``, ``GenericScanKernel``,
``ElementwiseKernel``, ``ReductionKernel`` and array operations (pyopencl is a dependency that is
absent from ``/root/reference``).

Test infrastructure only; usable only where ``/root/reference`` is mounted.  The vectors it
produces are committed under ``tests/golden/`` by ``tests/golden/make_refexec_golden.py``.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "boxtree", "traversal.py"))


@contextlib.contextmanager
def reference_modules():
    """Within the context, ``import boxtree.traversal`` etc. import the reference's files with the
    stand-in third-party modules; ``sys.modules`` is restored on exit (module objects handed out
    keep working)."""
    from . import fakecl
    fakes = fakecl.build_modules()
    pkg = types.ModuleType("boxtree")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "boxtree")]
    actx_stub = types.ModuleType("boxtree.array_context")
    actx_stub.dataclass_array_container = lambda cls: cls
    actx_stub.PyOpenCLArrayContext = fakecl.PyOpenCLArrayContext
    dist_pkg = types.ModuleType("boxtree.distributed")      # its __init__ needs mpi4py: skip it
    dist_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "boxtree", "distributed")]
    fakes["boxtree.distributed"] = dist_pkg

    def package_getattr(name):
        # `from boxtree import Tree, box_flags_enum, ...` of boxtree/__init__.py:26-52
        for modname in ("boxtree.tree", "boxtree.tree_build", "boxtree.traversal"):
            mod = importlib.import_module(modname)
            if hasattr(mod, name):
                return getattr(mod, name)
        raise AttributeError(name)
    pkg.__getattr__ = package_getattr
    fakes["boxtree"] = pkg
    fakes["boxtree.array_context"] = actx_stub
    touched = set(fakes) | {k for k in sys.modules if k == "boxtree" or k.startswith("boxtree.")}
    saved = {k: sys.modules.get(k) for k in touched}
    for k in list(sys.modules):
        if k == "boxtree" or k.startswith("boxtree."):
            del sys.modules[k]
    sys.modules.update(fakes)
    try:
        yield fakecl
    finally:
        for k in list(sys.modules):
            if k == "boxtree" or k.startswith("boxtree."):
                if k not in saved:
                    del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load(*names):
    """Import reference modules by dotted name; returns ``(fakecl, module, ...)``."""
    with reference_modules() as fakecl:
        mods = [importlib.import_module(n) for n in names]
        for m in mods:
            assert m.__file__.startswith(REFERENCE_ROOT), m.__file__
    return (fakecl, *mods)
