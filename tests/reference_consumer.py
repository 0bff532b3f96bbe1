"""Load the reference's OWN host-side consumer of the path -- ``boxtree/fmm.py`` (``drive_fmm``)
and ``boxtree/constant_one.py`` (``ConstantOneExpansionWrangler``), both pure numpy -- from
``/root/reference`` without importing the rest of the package (which needs pyopencl).

Only available where the reference checkout is mounted (this container); the GPU box has no
``/root/reference``.  Nothing from the reference is copied: the files are executed in place.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "boxtree", "constant_one.py"))


def load():
    """Returns ``(drive_fmm, ConstantOneExpansionWrangler, ConstantOneTreeIndependentDataForWrangler)``."""
    saved = {k: sys.modules.get(k) for k in ("boxtree", "boxtree.fmm", "boxtree.constant_one", "pytools")}
    try:
        # a bare namespace for `boxtree` so that boxtree/__init__.py (pyopencl) is not executed
        pkg = types.ModuleType("boxtree")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "boxtree")]
        sys.modules["boxtree"] = pkg
        try:
            import pytools  # noqa: F401
        except ImportError:
            stub = types.ModuleType("pytools")

            class ProcessLogger:                      # the one name boxtree/fmm.py imports
                def __init__(self, *a, **k):
                    pass

                def done(self, *a, **k):
                    pass
            stub.ProcessLogger = ProcessLogger
            sys.modules["pytools"] = stub
        sys.modules.pop("boxtree.fmm", None)
        sys.modules.pop("boxtree.constant_one", None)
        fmm = importlib.import_module("boxtree.fmm")
        c1 = importlib.import_module("boxtree.constant_one")
        assert fmm.__file__.startswith(REFERENCE_ROOT) and c1.__file__.startswith(REFERENCE_ROOT)
        return (fmm.drive_fmm, c1.ConstantOneExpansionWrangler,
                c1.ConstantOneTreeIndependentDataForWrangler)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_partition():
    """``get_box_ids_dfs_order`` and ``partition_work`` of the reference's
    ``boxtree/distributed/partition.py:38-121`` (pure Python host loops).  The module's other
    imports (mako, arraycontext, pyopencl: only used by the device code further down) are
    satisfied with empty stand-ins."""
    names = ("boxtree", "boxtree.distributed", "boxtree.distributed.partition", "mako",
             "mako.template", "arraycontext", "pyopencl", "pyopencl.elementwise",
             "pyopencl.tools", "pytools")
    saved = {k: sys.modules.get(k) for k in names}
    try:
        pkg = types.ModuleType("boxtree")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "boxtree")]
        sys.modules["boxtree"] = pkg
        dpkg = types.ModuleType("boxtree.distributed")
        dpkg.__path__ = [os.path.join(REFERENCE_ROOT, "boxtree", "distributed")]
        sys.modules["boxtree.distributed"] = dpkg

        def stub(name, **attrs):
            try:
                importlib.import_module(name)
            except ImportError:
                m = types.ModuleType(name)
                for k, v in attrs.items():
                    setattr(m, k, v)
                sys.modules[name] = m
        stub("mako")
        stub("mako.template", Template=object)
        stub("arraycontext", Array=object, ArrayContext=object, PyOpenCLArrayContext=object)
        stub("pyopencl")
        stub("pyopencl.elementwise", ElementwiseKernel=object)
        stub("pyopencl.tools", dtype_to_ctype=lambda dt: str(dt))
        stub("pytools", memoize_method=lambda f: f, ProcessLogger=object)
        sys.modules.pop("boxtree.distributed.partition", None)
        mod = importlib.import_module("boxtree.distributed.partition")
        assert mod.__file__.startswith(REFERENCE_ROOT)
        return mod.get_box_ids_dfs_order, mod.partition_work
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
