"""Development probe (needs ONE GPU): cost of the sharded distributed setup for one emulated
rank of P, per profiling scope.   python tests/dist_probe.py config3:10000000 2 8"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder, _cabi  # noqa: E402
from boxtree_b200 import distributed as bd  # noqa: E402
from tests.perf_probe import make  # noqa: E402


class OneRankComm:
    """Rank *rank* of *size*; the all-gather of the multipole masks repeats the rank's own."""

    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def Get_rank(self):  # noqa: N802
        return self.rank

    def Get_size(self):  # noqa: N802
        return self.size

    def allgather_tensor(self, t):
        return torch.stack([t] * self.size)


def main():
    actx = TorchArrayContext()
    tb, tg = TreeBuilder(actx), FMMTraversalBuilder(actx)
    lib = _cabi.load()
    src, kw = make(sys.argv[1])
    dsrc = [actx.from_numpy(s) for s in src]
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in kw.items()}
    tree, _ = tb(actx, dsrc, **dkw)
    for size in [int(a) for a in sys.argv[2:]]:
        comm = OneRankComm(size // 2, size)
        for rep in range(3):
            torch.cuda.synchronize()
            if rep == 2:
                lib.bt_prof_reset()
                lib.bt_prof_enable(1)
            t0 = time.perf_counter()
            out = bd.sharded_setup(actx, tree, tg, comm)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep == 2:
                lib.bt_prof_enable(0)
            lt = out[0]
            print(f"P={size} rep{rep}: sharded_setup {1e3 * dt:.2f} ms  local src={int(lt.sources[0].shape[0])} "
                  f"tgt={int(lt.targets[0].shape[0])}", flush=True)
            del out
        rep_ = _cabi.profile_report()
        tot = sum(v[1] for k, v in rep_.items())
        for k, (c, ms) in sorted(rep_.items(), key=lambda kv: -kv[1][1])[:16]:
            print(f"    {k:28s} calls={c:4d} {ms:9.3f} ms")
        print(f"    sum of scopes (nested counted twice) {tot:.2f} ms")


if __name__ == "__main__":
    main()
