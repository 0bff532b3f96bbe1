"""Development probe (torchrun, one process per GPU): phases of the distributed step.

    BT_PHASE_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29515 tests/dist_build_probe.py config3:10000000 [weak]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    from bench import make_inputs
    from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder, _cabi, _timing
    from boxtree_b200 import distributed as bd
    actx = TorchArrayContext(f"cuda:{local_rank}")
    comm = bd.TorchDistComm()
    tb, tg = TreeBuilder(actx), FMMTraversalBuilder(actx)
    recipe, n = sys.argv[1].split(":")
    n = int(float(n))
    weak = len(sys.argv) > 2 and sys.argv[2] == "weak"
    dtype = "f32" if recipe == "plummer" else "f64"
    if weak:
        src, kw = make_inputs(recipe, n, dtype, seed_shift=rank)
    else:
        src, kw = make_inputs(recipe, n, dtype, seed_shift=0)

        def sl(a):
            m = len(a)
            return np.ascontiguousarray(a[rank * m // world:(rank + 1) * m // world])
        src = [sl(x) for x in src]
        kw = {k: (sl(v) if isinstance(v, np.ndarray) else [sl(x) for x in v] if k == "targets" else v)
              for k, v in kw.items()}
    ssrc = [actx.from_numpy(x) for x in src]
    skw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in kw.items()}
    lib = _cabi.load()
    nrep = int(os.environ.get("BT_PROBE_REPS", "5"))
    for rep in range(nrep):
        dist.barrier()
        torch.cuda.synchronize()
        _timing.report()
        if rep == nrep - 1 and rank == 0:
            lib.bt_prof_reset()
            lib.bt_prof_enable(1)
        t0 = time.perf_counter()
        dtree = bd.build_distributed_tree(actx, tb, comm, ssrc,
                                          defer_extents=os.environ.get("BT_DIST_DEFER", "1") != "0",
                                          **skw)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = bd.distributed_tree_setup(actx, dtree, tg, comm)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"rank {rank} rep{rep}: tree {1e3 * (t1 - t0):.2f} ms  setup {1e3 * (t2 - t1):.2f} ms  "
              f"total {1e3 * (t2 - t0):.2f} ms  nboxes={dtree.nboxes} "
              f"local src={int(out[2].shape[0])} tgt={int(out[3].shape[0])}", flush=True)
        if rep == nrep - 1:
            ph = _timing.report()
            if ph:
                print(f"rank {rank} phases: " + "  ".join(f"{k}={v:.2f}" for k, v in ph.items()),
                      flush=True)
            if rank == 0:
                lib.bt_prof_enable(0)
                for k, (c, ms) in sorted(_cabi.profile_report().items(), key=lambda kv: -kv[1][1])[:28]:
                    print(f"    {k:32s} calls={c:4d}  {ms:9.3f} ms")
        del out, dtree
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
