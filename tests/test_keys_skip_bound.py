"""The bound behind ``make_keys_impl``'s ``points_never_stop`` (boxtree_b200/csrc/tree_build.cu):
for a particle of radius 0 the stick-out tests of ``make_keys_kernel`` (the reference's
``tree_build_kernels.py:388-428``) cannot fire at any level when

    stick_out_factor * min_extent / 2^(D+1)  >  64 * eps * max(|min|, |max|, extent)

so the kernel may skip its level loop for such particles.  Here the kernel's arithmetic is
replayed in numpy (IEEE, no contraction -- what ``-fmad=false`` gives) on adversarial points:
box corners and faces of every level, nudged by a few ulps, and random points, for boxes with
small and large offsets.  Where the condition holds no test may fire; where the offset is so
large that it does not hold, the replay shows why it is needed (tests do fire)."""
import numpy as np
import pytest

STOP_NEVER = 63


def condition(gmin, gext, sof, D, dtype):
    """Mirror of the host-side decision in ``make_keys_impl``."""
    eps = float(np.finfo(dtype).eps)
    gmax = gmin + gext
    M = max(np.max(np.abs(gmin)), np.max(np.abs(gmax)), np.max(gext))
    margin = float(dtype(sof)) * float(np.min(gext)) / float(1 << (D + 1))
    return bool(np.min(gext) > 0 and margin > 64.0 * eps * float(M))


def stop_levels(pos, radius, gmin, gext, sof, norm, D, dtype):
    """``make_keys_kernel``'s stop level for every particle (vectorised replay)."""
    T = dtype
    n, d = pos.shape
    half = T(1) / T(2)
    brf = T((1.0 + float(T(sof))) * float(half))
    scale = T(float(1 << D))
    q = (((pos - gmin) / gext) * scale).astype(np.uint64)
    stop = np.full(n, STOP_NEVER, np.int64)
    alive = np.ones(n, bool)
    for lev in range(D):
        size_factor = T(1) / T(float(1 << (1 + lev)))
        bits = (q >> np.uint64(D - 1 - lev)).astype(T)
        center = gmin + gext * (bits + half) * size_factor
        if norm == "linf":
            so = brf * gext * size_factor
            st = ((pos + radius[:, None] >= center + so) | (pos - radius[:, None] < center - so)).any(1)
        else:
            so = brf * gext[0] * size_factor
            acc = np.zeros(n, T)
            for a in range(d):
                acc = acc + (pos[:, a] - center[:, a]) * (pos[:, a] - center[:, a])
            dist = np.sqrt(acc) + radius
            st = dist * dist >= T(d) * so * so
        hit = alive & st
        stop[hit] = lev
        alive &= ~st
    return stop


def adversarial_points(gmin, gext, D, dtype, rng, d):
    pts = [rng.random((20000, d)).astype(dtype) * gext + gmin]
    for lev in (1, 2, 5, D - 3, D - 1, D):
        k = rng.integers(0, 1 << lev, (4000, d))
        base = (gmin + gext * (k.astype(dtype) / dtype(float(1 << lev)))).astype(dtype)
        for nudge in (-3, -1, 0, 1, 3):
            p = base.copy()
            for _ in range(abs(nudge)):
                p = np.nextafter(p, dtype(np.inf) if nudge > 0 else dtype(-np.inf))
            pts.append(p)
    p = np.concatenate(pts)
    lo, hi = gmin, gmin + gext
    inside = ((p >= lo) & (p < hi)).all(1)
    return np.ascontiguousarray(p[inside])


@pytest.mark.parametrize("norm", ["linf", "l2"])
@pytest.mark.parametrize("dtype,d,D", [(np.float64, 3, 19), (np.float64, 2, 28),
                                       (np.float32, 3, 19), (np.float64, 1, 31)])
@pytest.mark.parametrize("offset", [0.0, -3.5, 1e3, 1e9, 1e13])
@pytest.mark.parametrize("sof", [0.25, 0.01, 1e-6])
def test_points_never_stop_when_the_condition_holds(norm, dtype, d, D, offset, sof):
    rng = np.random.default_rng(hash((norm, d, D, offset, sof)) % 2 ** 32)
    ext = dtype(1.000100001 * 7.3)
    gmin = np.full(d, dtype(offset), dtype) + (rng.random(d) * 0.1).astype(dtype)
    gext = (gmin + ext) - gmin                      # as the kernel computes it: max - min
    if not np.all(gext > 0):
        pytest.skip("degenerate box at this offset")
    pos = adversarial_points(gmin, gext, D, dtype, rng, d)
    stop = stop_levels(pos, np.zeros(len(pos), dtype), gmin, gext, sof, norm, D, dtype)
    fired = int(np.count_nonzero(stop != STOP_NEVER))
    if condition(gmin, gext, sof, D, dtype):
        assert fired == 0, (fired, len(pos))
    # (no claim otherwise: the loop stays in the kernel)


def test_the_condition_is_not_vacuous():
    """It holds for the bench workload (fp64 unit box, stick-out 0.25) and not for fp32 or a far
    offset box -- where, indeed, radius-0 particles do trigger the tests."""
    one = np.ones(3)
    assert condition(np.zeros(3), one * 1.0001, 0.25, 19, np.float64)
    assert not condition(np.zeros(3, np.float32), (one * 1.0001).astype(np.float32), 0.25, 19,
                         np.float32)
    assert not condition(np.full(3, 1e13), one * 1.0001, 0.25, 19, np.float64)
    rng = np.random.default_rng(1)
    gmin = np.zeros(3, np.float32)
    gext = np.full(3, np.float32(1.0001))
    pos = adversarial_points(gmin, gext, 19, np.float32, rng, 3)
    stop = stop_levels(pos, np.zeros(len(pos), np.float32), gmin, gext, 0.0, "linf", 19,
                       np.float32)
    assert np.count_nonzero(stop != STOP_NEVER) > 0


def first_tested_level(radius, gmin, gext, sof, D, dtype):
    """Mirror of ``make_keys_kernel``'s ``lev0``: the loop level the walk starts at for a particle
    of the given radius (``make_keys_impl``: so_minext, skip_slack)."""
    eps = float(np.finfo(dtype).eps)
    gmax = gmin + gext
    M = float(max(np.max(np.abs(gmin)), np.max(np.abs(gmax)), np.max(gext)))
    min_ext = float(np.min(gext))
    if not (sof > 0 and min_ext > 0):
        return np.zeros(len(radius), np.int64)
    so_minext = float(dtype(sof)) * min_ext
    slack = 64.0 * eps * M
    with np.errstate(divide="ignore"):
        R = so_minext / (radius.astype(np.float64) + slack)
    lev0 = np.zeros(len(radius), np.int64)
    big = R > 8.0
    # ilogb(R) - 2, capped at D (ilogb(inf) is INT_MAX)
    e = np.where(np.isfinite(R[big]), np.floor(np.log2(np.where(np.isfinite(R[big]), R[big], 1.0))), 1e9)
    lev0[big] = np.minimum(e - 2, D).astype(np.int64)
    return lev0


@pytest.mark.parametrize("norm", ["linf", "l2"])
@pytest.mark.parametrize("dtype,d,D", [(np.float64, 3, 19), (np.float64, 2, 28),
                                       (np.float32, 3, 19), (np.float64, 1, 31)])
@pytest.mark.parametrize("offset", [0.0, -3.5, 1e3, 1e9])
@pytest.mark.parametrize("sof", [0.25, 0.01, 1.0])
def test_no_stop_test_fires_below_the_first_tested_level(norm, dtype, d, D, offset, sof):
    """The radius-aware skip of ``make_keys_kernel``: the full level loop never stops a particle
    at a level below ``lev0``, so starting the loop there gives the same stop level.  Radii sit
    on and around the per-level thresholds ``sof * extent / 2^(2+lev)`` (times 1/8 ... 8) and
    positions on and around box faces of every level."""
    rng = np.random.default_rng(hash((norm, d, D, offset, sof, "r")) % 2 ** 32)
    ext = dtype(1.000100001 * 7.3)
    gmin = np.full(d, dtype(offset), dtype) + (rng.random(d) * 0.1).astype(dtype)
    gext = (gmin + ext) - gmin
    if not np.all(gext > 0):
        pytest.skip("degenerate box at this offset")
    pos = adversarial_points(gmin, gext, D, dtype, rng, d)
    n = len(pos)
    lev = rng.integers(0, D, n)
    thr = float(dtype(sof)) * float(np.min(gext)) / np.exp2(2.0 + lev)
    radius = (thr * np.exp2(rng.integers(-3, 4, n)) * (1 + rng.integers(-2, 3, n) * 1e-15)).astype(dtype)
    radius[rng.random(n) < 0.05] = 0
    stop = stop_levels(pos, radius, gmin, gext, sof, norm, D, dtype)
    lev0 = first_tested_level(radius, gmin, gext, sof, D, dtype)
    assert np.all(stop >= lev0), int(np.count_nonzero(stop < lev0))
    # the skip is not vacuous: many particles start beyond level 0, and some stop right at lev0 + k
    assert np.count_nonzero(lev0 > 0) > n // 4
