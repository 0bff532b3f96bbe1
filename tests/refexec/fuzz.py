"""Randomised pinning of the oracle against the REFERENCE's own run (CPU only; needs
``/root/reference``): the random case generator of ``tests/random_sweep.py`` (dimensions, dtype,
sizes, tree kind, targets, radii, norms, criteria, n-away 1-3, min-nsources), each case built by
the reference's ``TreeBuilder`` + ``FMMTraversalBuilder`` through ``tests/refexec`` and by the
oracle, all arrays compared bit for bit including dtypes.  Every fifth case runs the distributed
setup (``boxtree/distributed``: partition, masks, local trees, local traversals) on 1-6 ranks
instead and compares every per-rank output; every seventh runs the area queries of
``boxtree/area_query.py`` with random balls.

    python tests/refexec/fuzz.py [master_seed] [seconds] [nprocs] [max particles per case]
"""
from __future__ import annotations

import os
import sys
import time
import traceback
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def one(args):
    seed, deadline, nmax = args
    if time.time() > deadline:
        return None
    from oracle.traversal import build_traversal
    from oracle.tree_build import MaxLevelsExceeded, build_tree
    from refexec.run import reference_traversal, reference_tree
    from tests.gpu_sweep import make_inputs
    from tests.parity_util import trav_mismatches, tree_mismatches
    from tests.random_sweep import random_case
    case = random_case(np.random.default_rng(seed), nmax=nmax)
    desc = {k: (np.dtype(v).name if k == "dtype" else v) for k, v in case.items()}
    try:
        src, kw = make_inputs(case)
        tkw = dict(case["trav"])
        ctor = {k: tkw.pop(k) for k in ("well_sep_is_n_away", "from_sep_smaller_crit") if k in tkw}
        try:
            rtree = reference_tree(src, **kw)
        except Exception as e:  # noqa: BLE001
            if type(e).__name__ == "MaxLevelsExceeded":
                try:
                    build_tree(src, **kw)
                except MaxLevelsExceeded:
                    return desc, []
                return desc, ["reference raised MaxLevelsExceeded, oracle did not"]
            if case["dims"] == 1 and kw.get("kind") == "adaptive-level-restricted":
                return desc, []          # the reference's 1-D level-restriction kernel does not compile
            raise
        if rtree.nboxes >= 8 and case["seed"] % 5 == 0:
            # every fifth case: the distributed setup on 1-6 ranks instead
            from refexec.compare import distributed_mismatches
            nranks = 1 + case["seed"] % 6
            desc["nranks"] = nranks
            return desc, distributed_mismatches(src, kw, ctor, nranks)
        otree = build_tree(src, **kw)
        if case["seed"] % 7 == 0 and case["dims"] > 1:      # (in 1-D the reference's area-query
            # kernel does not compile: `ball_center.x` on a scalar coord_vec_t)
            # every seventh case: the area queries of boxtree/area_query.py with random balls,
            # some of them outside the bounding box
            from refexec.compare import area_query_mismatches
            rng = np.random.default_rng(case["seed"])
            nballs = int(rng.integers(1, 400))
            dt = np.dtype(case["dtype"])
            centers = [(rng.normal(size=nballs) * 1.3).astype(dt) for _ in range(case["dims"])]
            radii = (10 ** rng.uniform(-3, 0.3) * 2 ** rng.uniform(-8, 0, nballs)).astype(dt)
            desc["nballs"] = nballs
            return desc, area_query_mismatches(otree, centers, radii)
        bad = ["tree." + b for b in tree_mismatches(rtree, otree)]
        if not bad:
            bad = ["trav." + b for b in trav_mismatches(
                reference_traversal(rtree, **ctor, **tkw), build_traversal(otree, **ctor, **tkw))]
        return desc, bad
    except Exception:  # noqa: BLE001
        return desc, ["EXCEPTION: " + traceback.format_exc(limit=6).replace("\n", " | ")]


def main():
    master = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
    nprocs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    nmax = int(sys.argv[4]) if len(sys.argv) > 4 else 4000
    t0 = time.time()
    seeds = np.random.default_rng(master).integers(0, 2 ** 31, 100000)
    ncase = nbad = 0
    with ProcessPoolExecutor(nprocs) as pool:
        for res in pool.map(one, ((int(s), t0 + budget, nmax) for s in seeds), chunksize=4):
            if res is None:
                break
            desc, bad = res
            ncase += 1
            if bad:
                nbad += 1
                print(f"FAIL {desc}: {bad[:6]}", flush=True)
    print(f"refexec fuzz (master seed {master}): {ncase - nbad}/{ncase} random cases: oracle == "
          f"the reference's own run, bit for bit, in {time.time() - t0:.0f} s")
    return nbad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
