"""Generate ``refexec_digests.json`` and ``refexec_*.npz``: outputs of the REFERENCE ITSELF
(``TreeBuilder.__call__`` and ``FMMTraversalBuilder.__call__`` of ``/root/reference/boxtree``, host
code and kernel text unmodified, executed on the CPU through ``tests/refexec``) on the seeded inputs
of ``tests/gpu_sweep.py``.

    python tests/golden/make_refexec_golden.py [nprocs]

Needs ``/root/reference`` (this container).  The GPU box has neither the reference nor a way to run
it, so what travels is: a short sha256 per output field for every sweep case (both the quick and
the full case list), and three cases with all arrays.  ``tests/test_refexec.py`` checks the oracle
against them on the CPU, ``tests/test_gpu_parity.py`` the CUDA path on the B200.
"""
from __future__ import annotations

import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FULL_FIXTURES = {            # case key -> file stem
    "quick:2d-float64-adaptive-n3000": "refexec_2d_f64_adaptive",
    "quick:3d-float32-ext-lr-n1-n3000": "refexec_3d_f32_ext_lr",
    "quick:3d-float64-nsep2-n3000": "refexec_3d_f64_nsep2",
}


def case_key(case, quick):
    return (f"{'quick' if quick else 'full'}:{case['dims']}d-{np.dtype(case['dtype']).name}-"
            f"{case['name']}-n{case['n']}")


def split_trav_kwargs(case):
    tkw = dict(case.get("trav") or {})
    ctor = {k: tkw.pop(k) for k in ("well_sep_is_n_away", "from_sep_smaller_crit") if k in tkw}
    return ctor, tkw


def run_one(job):
    quick, index = job
    from refexec.run import reference_traversal, reference_tree
    from tests.gpu_sweep import make_cases, make_inputs
    from tests.parity_util import TRAV_FIELDS, trav_digests, tree_digests
    case = make_cases(quick)[index]
    key = case_key(case, quick)
    src, kw = make_inputs(case)
    try:
        tree = reference_tree(src, **kw)
    except Exception as e:  # noqa: BLE001
        return key, {"error": type(e).__name__}, None
    entry = {"tree": tree_digests(tree)}
    trav = None
    if case.get("trav", {}) is not None:
        ctor, tkw = split_trav_kwargs(case)
        trav = reference_traversal(tree, **ctor, **tkw)
        entry["trav"] = trav_digests(trav)
    arrays = None
    if key in FULL_FIXTURES:
        arrays = {}
        for name, v in vars(tree).items():
            if isinstance(v, np.ndarray):
                arrays["tree." + name] = v
            elif isinstance(v, (list, tuple)) and len(v) and isinstance(v[0], np.ndarray):
                arrays["tree." + name] = np.stack(v)
        for name in TRAV_FIELDS:
            v = getattr(trav, name)
            if v is not None:
                arrays["trav." + name] = v
        for lev, bl in enumerate(trav.from_sep_smaller_by_level):
            for f in ("starts", "lists", "nonempty_indices", "compressed_indices"):
                arrays[f"trav.from_sep_smaller_by_level.{lev}.{f}"] = getattr(bl, f)
            arrays[f"trav.target_boxes_sep_smaller_by_source_level.{lev}"] = \
                trav.target_boxes_sep_smaller_by_source_level[lev]
    return key, entry, arrays


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nprocs = int(args[0]) if args else 8
    from tests.gpu_sweep import make_cases
    jobs = [(quick, i) for quick in (True, False) for i in range(len(make_cases(quick)))]
    digests = {}
    if "--missing" in sys.argv:         # only the cases the committed file does not hold yet
        with open(os.path.join(HERE, "refexec_digests.json")) as f:
            digests = json.load(f)
        jobs = [(q, i) for q, i in jobs if case_key(make_cases(q)[i], q) not in digests]
    with ProcessPoolExecutor(nprocs) as pool:
        for n, (key, entry, arrays) in enumerate(pool.map(run_one, jobs)):
            digests[key] = entry
            if arrays is not None:
                np.savez_compressed(os.path.join(HERE, FULL_FIXTURES[key] + ".npz"), **arrays)
            print(f"[{n + 1}/{len(jobs)}] {key}", flush=True)
    with open(os.path.join(HERE, "refexec_digests.json"), "w") as f:
        json.dump(digests, f, indent=0, sort_keys=True)
    print(f"wrote {len(digests)} cases")


if __name__ == "__main__":
    main()
