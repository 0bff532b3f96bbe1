"""pytest plugin that lets the reference's OWN test files (``/root/reference/test/test_*.py``) run
against ``tests/refexec``'s CPU execution of the reference:

    PYTHONDONTWRITEBYTECODE=1 python -m pytest -p no:cacheprovider -p refexec.pytest_plugin \\
        --rootdir /tmp/refexec-root /root/reference/test/test_traversal.py

(see ``tests/refexec/run_reference_tests.py``).  It installs the stand-in third-party modules for
the whole session and provides the ``actx_factory`` fixture the reference's tests ask for.  A test
passing here says two things at once: the reference's code does what its authors assert, *as
executed by refexec* -- i.e. the stand-ins are a faithful executor.
"""
from __future__ import annotations

import sys

import pytest

from . import fakecl, reference_modules

_cm = reference_modules()
_fakecl = _cm.__enter__()           # for the whole pytest session

# what the reference's test modules import on top of the library's own imports
_arraycontext = sys.modules["arraycontext"]
_arraycontext.pytest_generate_tests_for_array_contexts = lambda factories: (lambda metafunc: None)
_boxtree_actx = sys.modules["boxtree.array_context"]
_boxtree_actx.PytestPyOpenCLArrayContextFactory = object
_boxtree_actx._acf = None


@pytest.fixture
def actx_factory():
    return lambda: fakecl.PyOpenCLArrayContext()


def pytest_configure(config):
    for marker in ("opencl", "mpi", "slowtest"):
        config.addinivalue_line("markers", f"{marker}: marker of the reference's test-suite")


# {{{ input generators that the reference writes in loopy (a code generator that is not installed)

def _surface_particles(actx, nparticles, dims, dtype, seed=15):
    """The point sets of ``boxtree/tools.py:120-184`` (a closed curve in 2-D, a torus in 3-D)
    evaluated with numpy instead of a loopy kernel: test INPUT, not code under test."""
    import numpy as np
    from pytools import obj_array
    if dims == 2:
        phi = 2 * np.pi / nparticles * np.arange(nparticles)
        pts = [0.5 * (3 * np.cos(phi) + 2 * np.sin(3 * phi)),
               0.5 * (1 * np.sin(phi) + 1.5 * np.sin(2 * phi))]
    elif dims == 3:
        n = int(nparticles ** 0.5)
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        phi, theta = 2 * np.pi / n * i, 2 * np.pi / n * j
        pts = [5 * np.cos(phi) * (3 + np.cos(theta)), 5 * np.sin(phi) * (3 + np.cos(theta)),
               5 * np.sin(theta)]
    else:
        raise NotImplementedError
    return obj_array.new_1d([actx.from_numpy(np.ascontiguousarray(p.ravel().astype(dtype)))
                             for p in pts])


def _uniform_particles(actx, nparticles, dims, dtype, seed=15):
    """``boxtree/tools.py:187-279``: a rotated regular grid."""
    import numpy as np
    from pytools import obj_array
    if dims == 2:
        n = int(nparticles ** 0.5)
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        xx, yy = 4 * i / (n - 1), 4 * j / (n - 1)
        s, c = np.sin(0.3), np.cos(0.3)
        pts = [c * xx + s * yy - 2, -s * xx + c * yy - 2]
    elif dims == 3:
        n = int(nparticles ** (1 / 3))
        i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        xx, yy, zz = i / (n - 1), j / (n - 1), k / (n - 1)
        s1, c1 = np.sin(0.3), np.cos(0.3)
        xxx, yyy, zzz = c1 * xx + s1 * yy, -s1 * xx + c1 * yy, zz
        s2, c2 = np.sin(0.7), np.cos(0.7)
        pts = [4 * (c2 * xxx + s2 * zzz) - 2, 4 * yyy - 2, 4 * (-s2 * xxx + c2 * zzz) - 2]
    else:
        raise NotImplementedError
    return obj_array.new_1d([actx.from_numpy(np.ascontiguousarray(p.ravel().astype(dtype)))
                             for p in pts])


import boxtree.tools as _tools  # noqa: E402  (the reference's module, imported with the stand-ins)

assert _tools.__file__.startswith("/root/reference/")
_tools.make_surface_particle_array = _surface_particles
_tools.make_uniform_particle_array = _uniform_particles

# }}}
