"""Tiny driver for ncu: one warm-up step, then one step inside cudaProfilerStart/Stop.

    ncu --profile-from-start off --set full -k regex:... python tests/ncu_driver.py uniform:10000000:f64
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder  # noqa: E402
from tests.perf_probe import make  # noqa: E402
import numpy as np  # noqa: E402


def main():
    actx = TorchArrayContext()
    tb, tg = TreeBuilder(actx), FMMTraversalBuilder(actx)
    src, kw = make(sys.argv[1])
    dsrc = [actx.from_numpy(s) for s in src]
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in kw.items()}
    for rep in range(2):
        if rep == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        tree, _ = tb(actx, dsrc, **dkw)
        trav, _ = tg(actx, tree)
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
