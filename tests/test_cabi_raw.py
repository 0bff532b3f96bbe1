"""The C ABI driven WITHOUT the Python package: raw ``ctypes`` on ``libboxtree_b200.so`` with
device pointers, exactly what the binding stub of INTEGRATION.md does.  torch is used only to
own device memory (``data_ptr()``) and the stream."""
import ctypes as C
import os

import numpy as np
import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "boxtree_b200",
                   "libboxtree_b200.so")


class Particles(C.Structure):       # bt_particles, include/boxtree_b200.h
    _fields_ = [("sources", C.c_void_p * 3), ("targets", C.c_void_p * 3),
                ("source_radii", C.c_void_p), ("target_radii", C.c_void_p),
                ("nsources", C.c_int64), ("ntargets", C.c_int64)]


@pytest.mark.gpu
def test_bounding_box_keys_and_sort_through_raw_ctypes():
    import torch
    lib = C.CDLL(LIB)
    n, dim = 100_000, 3
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(dim, n))
    dev = [torch.from_numpy(np.ascontiguousarray(pts[a])).cuda() for a in range(dim)]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = Particles()
    for a in range(dim):
        p.sources[a] = dev[a].data_ptr()
    p.nsources, p.ntargets = n, 0
    # bt_bounding_box(dtype=BT_F64, dim, particles, out_minmax, stream)
    out = torch.empty(2 * dim, dtype=torch.float64, device="cuda")
    lib.bt_bounding_box.argtypes = [C.c_int, C.c_int, C.POINTER(Particles), C.c_void_p, C.c_void_p]
    assert lib.bt_bounding_box(1, dim, C.byref(p), C.c_void_p(out.data_ptr()), stream) == 0
    got = out.cpu().numpy()
    assert np.array_equal(got[0::2], pts.min(axis=1)) and np.array_equal(got[1::2], pts.max(axis=1))

    # bt_make_keys + bt_sort_particles: the ids come back as a permutation ordered by key
    lo = got[0::2].copy()
    ext = (got[1::2] - got[0::2]).max() * (1 + 1e-4)
    bmin = (C.c_double * 3)(*lo)
    bmax = (C.c_double * 3)(*(lo + ext))
    keys = [torch.empty(n, dtype=torch.int64, device="cuda") for _ in range(2)]
    ids = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
    lib.bt_make_keys.argtypes = [C.c_int, C.c_int, C.POINTER(Particles), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.bt_make_keys(1, dim, C.byref(p), bmin, bmax, 0, 0.0, 0,
                            C.c_void_p(keys[0].data_ptr()), None, None, stream) == 0
    in_alt = C.c_int(0)
    lib.bt_sort_particles.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
    unsorted = keys[0].clone()
    assert lib.bt_sort_particles(n, dim, 0, 0, C.c_void_p(keys[0].data_ptr()),
                                 C.c_void_p(keys[1].data_ptr()), C.c_void_p(ids[0].data_ptr()),
                                 C.c_void_p(ids[1].data_ptr()), C.byref(in_alt), stream) == 0
    torch.cuda.synchronize()
    k, i = keys[in_alt.value].cpu().numpy().view(np.uint64), ids[in_alt.value].cpu().numpy()
    assert np.all(k[1:] >= k[:-1])
    assert np.array_equal(np.sort(i), np.arange(n))
    assert np.array_equal(unsorted.cpu().numpy().view(np.uint64)[i], k)
    # stable: equal keys keep ascending particle ids
    same = k[1:] == k[:-1]
    assert np.all(i[1:][same] > i[:-1][same])
    assert lib.bt_launch_count() > 0
