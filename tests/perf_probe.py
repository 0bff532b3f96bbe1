"""Development probe (needs a GPU): wall time + per-scope CUDA-event breakdown.

``python tests/perf_probe.py uniform:1000000:f64 config3:10000000 plummer:10000000:f32``
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder, _cabi  # noqa: E402
from tests.parity_util import config3_inputs, plummer_particles, uniform_particles  # noqa: E402


def make(spec):
    parts = spec.split(":")
    kind, n = parts[0], int(float(parts[1]))
    dt = np.float32 if (len(parts) > 2 and parts[2] == "f32") else np.float64
    kw = {"max_particles_in_box": 30}
    if kind == "uniform":
        src = uniform_particles(n, 3, dt)
    elif kind == "uniform2d":
        src = uniform_particles(n, 2, dt)
    elif kind == "plummer":
        src = plummer_particles(n, dt)
    elif kind == "config3":
        src, tgt, radii = config3_inputs(n // 2, n - n // 2, dt)
        kw.update(targets=tgt, target_radii=radii, stick_out_factor=0.25, extent_norm="linf",
                  kind="adaptive-level-restricted")
    else:
        raise ValueError(kind)
    return src, kw


def main():
    actx = TorchArrayContext()
    tb = TreeBuilder(actx)
    tg = FMMTraversalBuilder(actx)
    lib = _cabi.load()
    for spec in sys.argv[1:]:
        if spec.startswith("--"):
            continue
        src, kw = make(spec)
        n = len(src[0]) + (len(kw["targets"][0]) if "targets" in kw else 0)
        dsrc = [actx.from_numpy(s) for s in src]
        dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
                   [actx.from_numpy(x) for x in v] if k == "targets" else v)
               for k, v in kw.items()}
        for rep in range(3):
            torch.cuda.synchronize()
            prof = rep == 2
            if prof:
                lib.bt_prof_reset()
                lib.bt_prof_enable(1)
            t0 = time.perf_counter()
            tree, _ = tb(actx, dsrc, **dkw)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            trav, _ = tg(actx, tree)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if prof:
                lib.bt_prof_enable(0)
            print(f"{spec} rep{rep}: tree {1e3 * (t1 - t0):.2f} ms  trav {1e3 * (t2 - t1):.2f} ms  "
                  f"total {1e3 * (t2 - t0):.2f} ms  -> {n / (t2 - t0) / 1e6:.1f} Mpts/s  "
                  f"nboxes={tree.nboxes} nlevels={tree.nlevels} "
                  f"list2={int(trav.from_sep_siblings_lists.shape[0])}", flush=True)
        rep_ = _cabi.profile_report()
        # entry scopes only: nested ones (radix passes, the parts of the fused walk) are inside them
        tot = sum(ms for k, (c, ms) in rep_.items() if not k.startswith(("rs_", "l13_", "l13h_")))
        for k, (c, ms) in sorted(rep_.items(), key=lambda kv: -kv[1][1]):
            print(f"    {k:28s} calls={c:4d}  {ms:9.3f} ms  {100 * ms / max(tot, 1e-9):5.1f}%")
        print(f"    sum of entry scopes {tot:.3f} ms; stats {tb.last_stats}")
        if "--check" in sys.argv:
            from oracle.traversal import build_traversal
            from oracle.tree_build import build_tree
            from tests.parity_util import trav_mismatches, tree_mismatches
            t0 = time.perf_counter()
            rt = build_tree(src, **kw)
            t1 = time.perf_counter()
            rv = build_traversal(rt)
            t2 = time.perf_counter()
            print(f"    oracle: tree {t1 - t0:.2f}s trav {t2 - t1:.2f}s")
            bad = tree_mismatches(rt, actx.to_numpy(tree))
            bad += trav_mismatches(rv, actx.to_numpy(trav)) if not bad else []
            print("    PARITY:", "OK" if not bad else bad[:10])
        del tree, trav
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
