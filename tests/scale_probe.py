"""Scale probe (needs a GPU): config 4 (Plummer fp32) at large N; tree only beyond the int32
CSR range of list 2."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder  # noqa: E402
from tests.parity_util import plummer_particles  # noqa: E402

actx = TorchArrayContext()
tb, tg = TreeBuilder(actx), FMMTraversalBuilder(actx)
for n in [int(float(a)) for a in sys.argv[1:]]:
    t0 = time.time()
    src = plummer_particles(n, np.float32)
    dsrc = [actx.from_numpy(s) for s in src]
    print(f"n={n}: generated in {time.time() - t0:.1f}s", flush=True)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tree, _ = tb(actx, dsrc, max_particles_in_box=30)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        try:
            trav, _ = tg(actx, tree)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            msg = (f"trav {1e3 * (t2 - t1):.1f} ms list2={int(trav.from_sep_siblings_lists.shape[0])} "
                   f"-> {n / (t2 - t0) / 1e6:.1f} Mpts/s")
            del trav
        except OverflowError as e:
            msg = f"traversal: OverflowError ({str(e)[:80]}...)"
        print(f"  rep{rep}: tree {1e3 * (t1 - t0):.1f} ms ({n / (t1 - t0) / 1e6:.0f} Mpts/s) nboxes={tree.nboxes} "
              f"nlevels={tree.nlevels}; {msg}; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB",
              flush=True)
        del tree
        torch.cuda.empty_cache()
