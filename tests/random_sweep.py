"""Randomised parity sweep (needs a GPU): random shapes, sizes, options and seeds through the
same harness as tests/gpu_sweep.py, every array of Tree and FMMTraversalInfo compared bit for
bit with the oracle.   python tests/random_sweep.py [master_seed] [seconds]"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_sweep import run_case  # noqa: E402


def random_case(rng, nmax=6000):
    dims = int(rng.integers(1, 4))
    dt = [np.float64, np.float32][int(rng.integers(0, 2))]
    n = int(rng.integers(50, nmax))
    case = dict(dims=dims, dtype=dt, n=n, seed=int(rng.integers(100, 10 ** 6)), name="random")
    tree = {"max_particles_in_box": int(rng.integers(3, 40))}
    kind = ["adaptive", "adaptive-level-restricted", "non-adaptive"][int(rng.choice(3, p=[0.55, 0.4, 0.05]))]
    if kind == "non-adaptive":
        case["n"] = min(n, 1500)
        tree["max_particles_in_box"] = max(tree["max_particles_in_box"], 10)
    tree["kind"] = kind
    trav = {"well_sep_is_n_away": int(rng.choice([1, 1, 2, 3]))}
    if rng.random() < 0.2:
        case["uniform"] = True
    if rng.random() < 0.6:
        case["ntargets"] = int(rng.integers(50, nmax))
        if rng.random() < 0.7:
            case["radii"] = True
            case["radii_scale"] = (float(10 ** rng.uniform(-2.5, 0.3)), int(rng.integers(-12, -1)))
            norm = ["linf", "l2"][int(rng.integers(0, 2))]
            tree.update(stick_out_factor=float(rng.choice([0.0, 0.1, 0.25, 0.5])), extent_norm=norm)
            crits = ["precise_linf", "static_linf"] if norm == "linf" else ["precise_linf", "static_l2"]
            trav["from_sep_smaller_crit"] = crits[int(rng.integers(0, 2))]
            if rng.random() < 0.2:
                trav["_from_sep_smaller_min_nsources_cumul"] = int(rng.integers(1, 60))
    case["tree"], case["trav"] = tree, trav
    return case


def main():
    from boxtree_b200 import TorchArrayContext, TreeBuilder
    master = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    nmax = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
    rng = np.random.default_rng(master)
    actx = TorchArrayContext()
    tb, travs = TreeBuilder(actx), {}
    t0, ncase, nbad = time.time(), 0, 0
    while time.time() - t0 < budget:
        case = random_case(rng, nmax)
        ncase += 1
        try:
            bad = run_case(case, actx, tb, travs)
        except Exception:  # noqa: BLE001
            bad = ["EXCEPTION: " + traceback.format_exc(limit=8).replace("\n", " | ")]
        if bad:
            nbad += 1
            desc = {k: v for k, v in case.items() if not k.startswith("_")}
            desc["dtype"] = np.dtype(desc["dtype"]).name
            print(f"FAIL {desc}: {bad[:6]}", flush=True)
    print(f"random sweep (master seed {master}): {ncase - nbad}/{ncase} cases match the oracle bit for bit "
          f"in {time.time() - t0:.0f} s")
    return nbad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
