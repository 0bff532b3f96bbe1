"""BASELINE configurations at FULL size, run through the REFERENCE ITSELF on the CPU
(``tests/refexec``): sha256 per output array into ``full_size_digests.json``.

    python tests/golden/make_full_size_golden.py [config3|uniform] ...

Takes minutes and several GB per configuration (every kernel runs one work item at a time).  Needs
``/root/reference``.  ``tests/test_gpu_parity.py::test_full_size_matches_reference_run`` holds the
CUDA path to these digests on the B200 -- bit-level parity with the reference at the sizes the
bench is quoted on.
"""
from __future__ import annotations

import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from tests.golden.make_golden import digest, flatten                      # noqa: E402
from tests.parity_util import config3_inputs, uniform_particles          # noqa: E402
import numpy as np                                                        # noqa: E402

OUT = os.path.join(HERE, "full_size_digests.json")


def full_size_cases():
    """name -> (sources, tree kwargs, traversal kwargs); inputs are built lazily."""
    def config3():
        s, t, r = config3_inputs(5_000_000, 5_000_000)
        return s, dict(max_particles_in_box=30, targets=t, target_radii=r, stick_out_factor=0.25,
                       extent_norm="linf", kind="adaptive-level-restricted"), {}

    def uniform():
        return uniform_particles(10_000_000, 3, np.float64), dict(max_particles_in_box=30), {}

    def config2():
        return uniform_particles(1_000_000, 3, np.float64), dict(max_particles_in_box=30), {}

    def plummer():
        # BASELINE config 4's recipe (bench.py "config4": Plummer sphere, fp32) at 1e7 points
        from tests.parity_util import plummer_particles
        return plummer_particles(10_000_000, np.float32, seed=15), dict(max_particles_in_box=30), {}

    return {"config2_3d_1e6": config2, "config3_3d_1e7": config3, "uniform_3d_1e7_f64": uniform,
            "plummer_3d_1e7_f32": plummer}


def main():
    from refexec.run import reference_traversal, reference_tree
    wanted = sys.argv[1:] or list(full_size_cases())
    digests = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in wanted:
        src, tkw, vkw = full_size_cases()[name]()
        t0 = time.time()
        tree = reference_tree(src, **tkw)
        t1 = time.time()
        trav = reference_traversal(tree, **vkw)
        t2 = time.time()
        digests[name] = {k: digest(v) for k, v in flatten(tree, trav).items()}
        digests[name]["_nboxes"] = int(tree.nboxes)
        digests[name]["_nlevels"] = int(tree.nlevels)
        print(f"{name}: nboxes {tree.nboxes} nlevels {tree.nlevels}; reference tree build "
              f"{t1 - t0:.0f} s, traversal {t2 - t1:.0f} s (serial CPU)", flush=True)
        with open(OUT, "w") as f:
            json.dump(digests, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
