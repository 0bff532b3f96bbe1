"""ctypes access to the oracle's C kernels (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import numpy as np

from . import build as _build


@lru_cache(maxsize=None)
def lib_for(coord_dtype) -> C.CDLL:
    coord_dtype = np.dtype(coord_dtype)
    tag = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[coord_dtype]
    _build.build()
    return C.CDLL(_build.lib_path(tag))


def ptr(a):
    """void* of a C-contiguous numpy array (or NULL)."""
    if a is None:
        return C.c_void_p(0)
    assert a.flags.c_contiguous, "oracle kernels need contiguous arrays"
    return C.c_void_p(a.ctypes.data)


def ptr_array(arrays):
    """``T *const *``: an array of pointers to the given numpy arrays."""
    arr = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    return arr


def coord_arg(coord_dtype, v):
    return (C.c_float if np.dtype(coord_dtype) == np.float32 else C.c_double)(float(v))


i32 = C.c_int32
i64 = C.c_int64
cint = C.c_int
