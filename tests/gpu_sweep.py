"""Parity sweep (needs a GPU): CUDA path vs CPU oracle over many option combinations.

``python tests/gpu_sweep.py [--quick]`` prints one line per case; exit status 1 if any
case differs.  Used during development and by ``tests/test_gpu_parity.py``.
"""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.traversal import build_traversal as oracle_traversal  # noqa: E402
from oracle.tree_build import MaxLevelsExceeded as OracleMaxLevels  # noqa: E402
from oracle.tree_build import build_tree as oracle_tree  # noqa: E402
from tests.parity_util import (digest_mismatches, normal_particles, trav_digests,  # noqa: E402
                               trav_mismatches, tree_digests, tree_mismatches,
                               uniform_particles)


def make_cases(quick=False):
    cases = []
    dims_list = (2, 3) if quick else (1, 2, 3)
    for dims in dims_list:
        for dt in (np.float64, np.float32):
            base = dict(dims=dims, dtype=dt, n=3000 if quick else 20000)
            cases.append(dict(base, name="adaptive", tree={"max_particles_in_box": 30}))
            cases.append(dict(base, name="adaptive-5", tree={"max_particles_in_box": 5}))
            cases.append(dict(base, name="single-box", n=4, tree={"max_particles_in_box": 30}))
            cases.append(dict(base, name="two-level", n=50, tree={"max_particles_in_box": 30}))
            cases.append(dict(base, name="skip-prune", tree={"max_particles_in_box": 30,
                                                             "skip_prune": True}, trav=None))
            cases.append(dict(base, name="tiny-guess", tree={"max_particles_in_box": 30,
                                                             "nboxes_guess": 5}))
            cases.append(dict(base, name="non-adaptive", n=3000,
                              tree={"max_particles_in_box": 30, "kind": "non-adaptive"}))
            cases.append(dict(base, name="weights", weights=True,
                              tree={"max_leaf_refine_weight": 100}))
            cases.append(dict(base, name="lr", tree={"max_particles_in_box": 30,
                                                     "kind": "adaptive-level-restricted"}))
            cases.append(dict(base, name="lr-skip-prune",
                              tree={"max_particles_in_box": 30, "skip_prune": True,
                                    "kind": "adaptive-level-restricted"}, trav=None))
            cases.append(dict(base, name="src-tgt", ntargets=15000,
                              tree={"max_particles_in_box": 30}))
            cases.append(dict(base, name="uniform", uniform=True,
                              tree={"max_particles_in_box": 30}))
            # user-supplied (square, covering) bounding box, tree_build.py:477-508
            cases.append(dict(base, name="user-bbox", user_bbox=(-7.5, 8.5),
                              tree={"max_particles_in_box": 30}))
            # n-away 3: beyond the top-down colleague builder's reach in 3-D (walk-based fallback)
            cases.append(dict(base, name="nsep3", n=3000, tree={"max_particles_in_box": 30},
                              trav={"well_sep_is_n_away": 3}))
            for nsep in (1, 2):
                cases.append(dict(base, name=f"nsep{nsep}", tree={"max_particles_in_box": 30},
                                  trav={"well_sep_is_n_away": nsep}))
                for norm, crit in (("linf", "static_linf"), ("linf", "precise_linf"),
                                   ("l2", "static_l2"), ("l2", "precise_linf")):
                    cases.append(dict(
                        base, name=f"ext-{norm}-{crit}-n{nsep}", ntargets=15000, radii=True,
                        tree={"max_particles_in_box": 30, "stick_out_factor": 0.25,
                              "extent_norm": norm},
                        trav={"well_sep_is_n_away": nsep, "from_sep_smaller_crit": crit}))
                cases.append(dict(
                    base, name=f"ext-lr-n{nsep}", ntargets=15000, radii=True,
                    tree={"max_particles_in_box": 30, "stick_out_factor": 0.25,
                          "extent_norm": "linf", "kind": "adaptive-level-restricted"},
                    trav={"well_sep_is_n_away": nsep}))
            cases.append(dict(
                base, name="ext-minsrc", ntargets=15000, radii=True,
                tree={"max_particles_in_box": 30, "stick_out_factor": 0.25},
                trav={"_from_sep_smaller_min_nsources_cumul": 40}))
            # most targets are too fat to leave the top boxes: thousands of own particles per box
            cases.append(dict(
                base, name="ext-fat", ntargets=15000, radii=True, radii_scale=(2.0, -3),
                tree={"max_particles_in_box": 30, "stick_out_factor": 0.25, "extent_norm": "linf"}))
            cases.append(dict(base, name="coincident", n=12, coincident=True,
                              tree={"max_particles_in_box": 10}, expect_max_levels=True))
            # the reference's test_max_levels_error (test/test_tree.py:1103-1112): every point at
            # the origin, a bounding box without extent
            cases.append(dict(base, name="all-at-origin", n=11, at_origin=True,
                              tree={"max_particles_in_box": 10}, expect_max_levels=True))
            if dims > 1:
                # a tight cluster next to a sparse cloud: the tree is deeper than ONE 64-bit sort
                # key resolves (19 levels in 3-D, 28 in 2-D) -- two-word keys, up to the level 31
                # that the reference's `1U << (1 + level)` digits reach
                for kind in ("adaptive", "adaptive-level-restricted"):
                    cases.append(dict(base, name="deep-cluster" + ("-lr" if "restricted" in kind else ""),
                                      n=60, deep_cluster=22 if dims == 3 else 27,
                                      tree={"max_particles_in_box": 4, "kind": kind}))
                # test_same_tree_with_zero_weight_particles (test/test_tree.py:1050-1097): targets
                # with weight 0 and radii up to the domain size, stick-out factors 0 .. 1
                for sof in (0, 0.1, 0.3, 1):
                    cases.append(dict(base, name=f"zero-weight-sof{sof}", n=20, zero_weight=sof,
                                      tree={"max_leaf_refine_weight": 10,
                                            "stick_out_factor": sof}))
    return cases


def make_inputs(case):
    dims, dt, n = case["dims"], case["dtype"], case["n"]
    if case.get("at_origin"):
        return [np.zeros(n, dtype=dt) for _ in range(dims)], dict(case["tree"])
    if case.get("zero_weight") is not None:
        rng = np.random.default_rng(10)
        sources = rng.random((dims, n)) ** 2
        sources[:, 0] = -0.1
        sources[:, 1] = 1.1
        targets = rng.random((dims, 500))[:, :40].copy()
        radii = rng.random(500)[:40]
        weights = np.zeros(n + 40, np.int32)
        weights[:n] = 1
        kw = dict(case["tree"], targets=[t.astype(dt) for t in targets],
                  target_radii=radii.astype(dt), refine_weights=weights)
        return [s.astype(dt) for s in sources], kw
    if case.get("deep_cluster"):
        rng = np.random.default_rng(7)
        pts = rng.random((dims, n))
        pts[:, n // 2:] *= 2.0 ** -case["deep_cluster"]
        return [np.ascontiguousarray(pts[a]).astype(dt) for a in range(dims)], dict(case["tree"])
    if case.get("coincident"):
        # 11 coincident points (> max_particles_in_box) and one distinct point so the
        # bounding box is not degenerate: both implementations must give up
        src = [np.full(n, 0.25 + 0.1 * ax, dtype=dt) for ax in range(dims)]
        for ax in range(dims):
            src[ax][0] = 1.0
    elif case.get("uniform"):
        src = uniform_particles(n, dims, dt, seed=case.get("seed", 15))
    else:
        src = normal_particles(n, dims, dt, seed=case.get("seed", 15))
    kw = dict(case["tree"])
    if case.get("user_bbox"):
        lo, hi = case["user_bbox"]
        kw["bbox"] = np.array([[lo, hi]] * dims, dtype=dt)
    tgt = None
    if case.get("ntargets"):
        tgt = normal_particles(case["ntargets"], dims, dt, seed=case.get("seed", 15) + 3)
        kw["targets"] = tgt
        if case.get("radii"):
            rng = np.random.default_rng(case.get("seed", 15) - 2)
            scale, lo = case.get("radii_scale", (0.05, -10))
            kw["target_radii"] = (scale * 2 ** rng.uniform(lo, 0, case["ntargets"])).astype(dt)
    if case.get("weights"):
        kw["refine_weights"] = np.random.default_rng(10).integers(
            0, 10, n + (case.get("ntargets") or 0), dtype=np.int32)
    return src, kw


def run_case(case, actx, tb, travs, ref_digests=None):
    """CUDA path vs the oracle's arrays, and -- when *ref_digests* (the case's entry of
    ``tests/golden/refexec_digests.json``) is given -- vs the digests of the reference's own run."""
    from boxtree_b200 import FMMTraversalBuilder, MaxLevelsExceeded
    src, kw = make_inputs(case)
    t0 = time.time()
    try:
        ref_tree = oracle_tree(src, **kw)
        ref_err = None
    except OracleMaxLevels as e:
        ref_tree, ref_err = None, e

    dev_kw = {k: (v if k == "bbox" else              # the bounding box stays a host array
                  actx.from_numpy(v) if isinstance(v, np.ndarray) else
                  [actx.from_numpy(x) for x in v] if k == "targets" else v)
              for k, v in kw.items()}
    try:
        got_tree_dev, _ = tb(actx, [actx.from_numpy(s) for s in src], **dev_kw)
    except MaxLevelsExceeded:
        if case.get("expect_max_levels") or ref_err is not None:
            return []
        raise
    if ref_tree is None:
        return ["oracle raised MaxLevelsExceeded, CUDA path did not"]
    got_tree = actx.to_numpy(got_tree_dev)
    bad = ["tree." + b for b in tree_mismatches(ref_tree, got_tree)]
    if ref_digests is not None:
        bad += ["reference.tree." + b
                for b in digest_mismatches(ref_digests["tree"], tree_digests(got_tree))]
    if case.get("trav", {}) is not None and not bad:
        tkw = dict(case.get("trav") or {})
        ctor = {k: tkw.pop(k) for k in ("well_sep_is_n_away", "from_sep_smaller_crit")
                if k in tkw}
        ref_trav = oracle_traversal(ref_tree, **ctor, **tkw)
        key = tuple(sorted(ctor.items()))
        if key not in travs:
            travs[key] = FMMTraversalBuilder(actx, **ctor)
        got_trav_dev, _ = travs[key](actx, got_tree_dev, **tkw)
        case["_trav_stats"] = dict(travs[key].last_stats)
        got_trav = actx.to_numpy(got_trav_dev)
        bad += ["trav." + b for b in trav_mismatches(ref_trav, got_trav)]
        if ref_digests is not None:
            bad += ["reference.trav." + b
                    for b in digest_mismatches(ref_digests["trav"], trav_digests(got_trav))]
    case["_info"] = (f"nboxes={ref_tree.nboxes} nlevels={ref_tree.nlevels} "
                     f"{time.time() - t0:.2f}s")
    return bad


def main(quick=False):
    from boxtree_b200 import TorchArrayContext, TreeBuilder
    actx = TorchArrayContext()
    tb = TreeBuilder(actx)
    travs = {}
    nbad = 0
    cases = make_cases(quick)
    for case in cases:
        label = f"{case['dims']}d {np.dtype(case['dtype']).name:8s} {case['name']}"
        try:
            bad = run_case(case, actx, tb, travs)
        except Exception:  # noqa: BLE001
            bad = ["EXCEPTION: " + traceback.format_exc(limit=6).replace("\n", " | ")]
        if bad:
            nbad += 1
            print(f"FAIL {label}: {bad[:8]}", flush=True)
        else:
            print(f"ok   {label} {case.get('_info', '')}", flush=True)
    print(f"{len(cases) - nbad}/{len(cases)} cases match the oracle bit for bit")
    return nbad


if __name__ == "__main__":
    sys.exit(1 if main("--quick" in sys.argv) else 0)
