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
