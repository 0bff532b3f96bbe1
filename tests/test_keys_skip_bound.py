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
