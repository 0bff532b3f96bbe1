"""Development probe (needs a GPU): where the GPU idles inside one warm step.

    python tests/gap_probe.py config3:10000000 [min_gap_us]

Runs the step under torch.profiler (CUPTI kernel records) and lists the idle gaps between
consecutive kernels / memcpys / memsets on the device, longest first, with the activity before
and after each gap."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder  # noqa: E402
from tests.perf_probe import make  # noqa: E402


def main():
    spec = sys.argv[1]
    min_gap = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
    actx = TorchArrayContext()
    tb, tg = TreeBuilder(actx), FMMTraversalBuilder(actx)
    src, kw = make(spec)
    dsrc = [actx.from_numpy(s) for s in src]
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in kw.items()}

    def step():
        tree, _ = tb(actx, dsrc, **dkw)
        trav, _ = tg(actx, tree)
        return tree, trav
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    print(f"{spec}: {len(evs)} device activities, span {(t1 - t0) / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms")
    gaps = []
    end = evs[0].time_range.end
    prev = evs[0]
    for e in evs[1:]:
        g = e.time_range.start - end
        if g > min_gap:
            gaps.append((g, prev.name[:60], e.name[:60], (end - t0) / 1e3))
        if e.time_range.end > end:
            end, prev = e.time_range.end, e
    print(f"idle in gaps > {min_gap} us: {sum(g[0] for g in gaps) / 1e3:.3f} ms in {len(gaps)} gaps")
    for g, a, b, at in sorted(gaps, reverse=True)[:40]:
        print(f"  {g:7.1f} us at {at:7.3f} ms   after {a}   before {b}")


if __name__ == "__main__":
    main()
