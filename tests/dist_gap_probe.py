"""Development probe (torchrun, one process per GPU): where rank 0's GPU idles inside one warm
step of the distributed build + setup, and how long the NCCL kernels run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/dist_gap_probe.py config3:10000000 [min_gap_us]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import numpy as np

    from bench import make_inputs
    from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder
    from boxtree_b200 import distributed as bd
    actx = TorchArrayContext(f"cuda:{local_rank}")
    comm = bd.TorchDistComm()
    tb, tg = TreeBuilder(actx), FMMTraversalBuilder(actx)
    recipe, n = sys.argv[1].split(":")
    min_gap = float(sys.argv[2]) if len(sys.argv) > 2 else 15.0
    src, kw = make_inputs(recipe, int(float(n)), "f32" if recipe == "plummer" else "f64",
                          seed_shift=rank)
    ssrc = [actx.from_numpy(x) for x in src]
    skw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in kw.items()}

    def step():
        dtree = bd.build_distributed_tree(actx, tb, comm, ssrc, defer_extents=True, **skw)
        return bd.distributed_tree_setup(actx, dtree, tg, comm)
    for _ in range(3):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
        nccl = [e for e in evs if "nccl" in e.name.lower()]
        print(f"{len(evs)} device activities, span {(t1 - t0) / 1e3:.3f} ms; "
              f"{len(nccl)} NCCL kernels, {sum(e.time_range.end - e.time_range.start for e in nccl) / 1e3:.3f} ms")
        for e in sorted(nccl, key=lambda e: -(e.time_range.end - e.time_range.start))[:12]:
            print(f"   nccl {(e.time_range.end - e.time_range.start):8.1f} us at "
                  f"{(e.time_range.start - t0) / 1e3:7.3f} ms  {e.name[:70]}")
        gaps = []
        end, prev = evs[0].time_range.end, evs[0]
        for e in evs[1:]:
            g = e.time_range.start - end
            if g > min_gap:
                gaps.append((g, prev.name[:50], e.name[:50], (end - t0) / 1e3))
            if e.time_range.end > end:
                end, prev = e.time_range.end, e
        print(f"idle in gaps > {min_gap} us: {sum(g[0] for g in gaps) / 1e3:.3f} ms in {len(gaps)} gaps")
        for g, a, b, at in sorted(gaps, reverse=True)[:45]:
            print(f"  {g:7.1f} us at {at:7.3f} ms   after {a}   before {b}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
