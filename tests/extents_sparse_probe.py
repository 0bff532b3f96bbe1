"""Development probe (needs a GPU): the own-particle pass of the box extents on a rank's SHARE of
a big global tree (distributed build, weak scaling) -- 8 lanes per box against one lane per box
(``bt_box_extents_phase`` flag 4).  The share is emulated: every box keeps a binomial(count, 1/R)
part of its own particles.

    python tests/extents_sparse_probe.py config3:40000000 8"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from boxtree_b200 import TorchArrayContext, TreeBuilder, _cabi  # noqa: E402
from boxtree_b200._cabi import check, dptr, dtype_code  # noqa: E402
from tests.perf_probe import make  # noqa: E402


def main():
    spec, R = sys.argv[1], int(sys.argv[2])
    actx = TorchArrayContext()
    lib = _cabi.load()
    src, kw = make(spec)
    dsrc = [actx.from_numpy(s) for s in src]
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in kw.items()}
    tree, _ = TreeBuilder(actx)(actx, dsrc, **dkw)
    nb, aligned, dims = int(tree.nboxes), int(tree.aligned_nboxes), int(tree.dimensions)
    nlevels = int(tree.nlevels)
    ls = tree.level_start_box_nrs.cpu().numpy()
    ls_host = (C.c_int32 * (nlevels + 1))(*[int(x) for x in ls[:nlevels + 1]])
    print(f"{spec}: nboxes {nb} nlevels {nlevels}")
    torch.manual_seed(0)
    for kind, counts, have_r in (("sources", tree.box_source_counts_nonchild, False),
                                 ("targets", tree.box_target_counts_nonchild,
                                  tree.targets_have_extent)):
        lcount = torch.binomial(counts.double(), torch.full_like(counts, 1.0 / R, dtype=torch.float64)
                                ).to(torch.int32)
        lstart = (torch.cumsum(lcount, 0) - lcount).to(torch.int32)
        n = int(lcount.sum())
        parts = [torch.rand(max(n, 1), dtype=torch.float64, device=actx.device) for _ in range(dims)]
        radii = torch.rand(max(n, 1), dtype=torch.float64, device=actx.device) * 1e-3 if have_r else None
        bmin = actx.empty((dims, aligned), np.float64)
        bmax = actx.empty((dims, aligned), np.float64)
        print(f"  {kind}: {n} local particles, {n / nb:.2f} per box, max {int(lcount.max())}")
        for phases, name in ((1, "8 lanes per box"), (1 | 4, "1 lane per box ")):
            def run():
                check(lib.bt_box_extents_phase(
                    dtype_code(np.dtype(np.float64)), dims, nb, aligned, nlevels, ls_host,
                    dptr(tree.box_child_ids), dptr(tree.box_centers), dptr(lstart), dptr(lcount),
                    _cabi.ptr_array(parts), dptr(radii), dptr(bmin), dptr(bmax), phases,
                    actx.stream_handle), "bt_box_extents_phase")
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ref = (bmin.clone(), bmax.clone()) if phases == 1 else ref
            same = phases == 1 or (torch.equal(ref[0][:, :nb], bmin[:, :nb])
                                   and torch.equal(ref[1][:, :nb], bmax[:, :nb]))
            print(f"    {name}: {e0.elapsed_time(e1) / 10:.3f} ms  same result: {same}")


if __name__ == "__main__":
    main()
