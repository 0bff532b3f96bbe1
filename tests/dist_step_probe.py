"""Development probe (torchrun, one process per GPU): wall time of the phases of one
distributed step (all-gather, replicated tree build, sharded setup).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29515 tests/dist_step_probe.py config3:10000000"""
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
    from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder
    from boxtree_b200 import distributed as bd
    from tests.perf_probe import make
    actx = TorchArrayContext(f"cuda:{local_rank}")
    comm = bd.TorchDistComm()
    tb, tg = TreeBuilder(actx), FMMTraversalBuilder(actx)
    src, kw = make(sys.argv[1])

    def sl(a):
        m = len(a)
        return actx.from_numpy(np.ascontiguousarray(a[rank * m // world:(rank + 1) * m // world]))

    ssrc = [sl(x) for x in src]
    skw = {k: (sl(v) if isinstance(v, np.ndarray) else [sl(x) for x in v] if k == "targets" else v)
           for k, v in kw.items()}
    for rep in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = bd.allgather_particles(actx, comm, ssrc)
        gk = dict(skw)
        if "targets" in skw:
            gk["targets"] = bd.allgather_particles(actx, comm, skw["targets"])
        if "target_radii" in skw:
            gk["target_radii"] = bd.allgather_particles(actx, comm, [skw["target_radii"]])[0]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tree, _ = tb(actx, g, **gk)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out = bd.sharded_setup(actx, tree, tg, comm)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"rank {rank} rep{rep}: allgather {1e3 * (t1 - t0):.2f} ms  tree {1e3 * (t2 - t1):.2f} ms  "
              f"sharded_setup {1e3 * (t3 - t2):.2f} ms  total {1e3 * (t3 - t0):.2f} ms", flush=True)
        del out, tree, g, gk
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
