"""Morton keys bit-identical: the per-level digits and the stop level that
``bt_make_keys`` packs into a particle's sort key are the Morton numbers the reference's
``scan_t_from_particle`` (``tree_build_kernels.py:308-470``, restated in
``oracle/oracle_tree.c::morton_nr_of_particle``) computes level by level:

    key = [digit(level 1) ... digit(level D)] << 6 | stop        (csrc/tree_build.cu)

digit(level L + 1) of a particle == the oracle's Morton number at particle level L for every
L below the particle's stop level, and the stop level is the first L where the oracle says
"stops here" (-1); 63 = never.  Checked through the C ABI for point particles and particles
with extent (both norms), fp64 and fp32, 1-3 D, and for every key depth the builder uses."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STOP_BITS, STOP_NEVER = 6, 63
NORM = {None: 0, "linf": 1, "l2": 2}


def oracle_digits(coords, radii, bmin, bmax, norm, sof, depth, dtype):
    from oracle._lib import coord_arg, lib_for, ptr, ptr_array
    lib = lib_for(dtype)
    d, n = len(coords), len(coords[0])
    out = np.empty((depth, n), np.int8)
    bmin, bmax = np.ascontiguousarray(bmin, dtype), np.ascontiguousarray(bmax, dtype)
    for lev in range(depth):
        lib.orc_particle_morton_nrs(C.c_int(d), C.c_int(NORM[norm]), C.c_int(lev), C.c_int64(n),
                                    ptr(bmin), ptr(bmax), ptr_array(coords), ptr(radii),
                                    coord_arg(dtype, sof), ptr(out[lev]))
    return out


@pytest.mark.parametrize("depth", [0, 7, 12])
@pytest.mark.parametrize("dims,dtype,norm", [(3, np.float64, None), (3, np.float64, "linf"),
                                            (3, np.float64, "l2"), (3, np.float32, "linf"),
                                            (2, np.float64, "linf"), (1, np.float64, "l2"),
                                            (2, np.float32, None)])
def test_key_digits_are_the_reference_morton_numbers(actx, dims, dtype, norm, depth):
    import torch

    from boxtree_b200 import _cabi
    from boxtree_b200._cabi import bt_particles, check, dptr, dtype_code
    lib = _cabi.load()
    n = 60_000
    rng = np.random.default_rng(dims * 100 + (0 if norm is None else len(norm)) + depth)
    coords = [np.ascontiguousarray(rng.normal(size=n).astype(dtype)) for _ in range(dims)]
    radii = None
    sof = 0.0
    if norm is not None:
        radii = np.ascontiguousarray((2.0 ** rng.uniform(-14, -1, n)).astype(dtype))
        radii[rng.random(n) < 0.2] = 0
        sof = 0.25
    lo = np.array([c.min() for c in coords], dtype)
    ext = dtype(max(float(c.max()) - float(c.min()) for c in coords) * (1 + 1e-4)
                + (float(radii.max()) * 2 if radii is not None else 0))
    bmin = (lo - (radii.max() if radii is not None else dtype(0))).astype(dtype)
    bmax = (bmin + ext).astype(dtype)

    D = int(lib.bt_max_key_level(dims)) if depth == 0 else min(depth, int(lib.bt_max_key_level(dims)))
    dev = [actx.from_numpy(c) for c in coords]
    drad = actx.from_numpy(radii) if radii is not None else None
    P = bt_particles()
    for a in range(dims):
        P.sources[a] = dptr(dev[a])
    P.source_radii = dptr(drad)
    P.nsources, P.ntargets = n, 0
    keys = actx.empty(n, np.int64)
    check(lib.bt_make_keys(dtype_code(np.dtype(dtype)), dims, C.byref(P), _cabi.darray(bmin),
                           _cabi.darray(bmax), NORM[norm], float(sof), D, dptr(keys), None, None,
                           actx.stream_handle), "bt_make_keys")
    torch.cuda.synchronize()
    k = keys.cpu().numpy().view(np.uint64)
    stop = (k & np.uint64(STOP_NEVER)).astype(np.int64)
    digits = k >> np.uint64(STOP_BITS)

    want = oracle_digits(coords, radii, bmin, bmax, norm, sof, D, dtype)      # [D, n]
    # the oracle's stop level: the first particle level whose Morton number is -1
    stops = want < 0
    want_stop = np.where(stops.any(axis=0), stops.argmax(axis=0), STOP_NEVER)
    assert np.array_equal(stop, want_stop)
    if norm is not None:
        assert np.count_nonzero(stop != STOP_NEVER) > n // 50     # (the case does exercise stops)
    for lev in range(D):        # digit of level lev + 1, for the particles that descend that far
        got = ((digits >> np.uint64((D - lev - 1) * dims)) & np.uint64((1 << dims) - 1)).astype(np.int64)
        live = want_stop > lev
        assert np.array_equal(got[live], want[lev][live].astype(np.int64)), lev
        # below its stop level a key is zero padded
        assert not got[~live].any(), lev
