"""Multi-GPU parity check of the DISTRIBUTED TREE BUILD (run under torchrun, one process per
GPU, NCCL over NVLink):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_build_nccl_check.py

Every rank contributes its slice of the particle set; the distributed build (all-reduced box
counts per level), the work partition, the all-to-all of particles and the local traversals
run over NCCL; each rank compares its global box arrays, local tree, local traversal and index
arrays bit for bit with the oracle's restatement of the reference's distributed flow for that
rank (and, for the committed rank counts, with the digests of the reference's own run)."""
import json
import os
import sys

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

    from boxtree_b200 import TorchArrayContext
    from boxtree_b200 import distributed as bd
    from tests.dist_build_util import check_rank, oracle_ranks, run_rank
    from tests.dist_cases import CASES
    from tests.parity_util import config3_inputs, uniform_particles

    actx = TorchArrayContext(f"cuda:{local_rank}")
    comm = bd.TorchDistComm()
    cases = {k: v() for k, v in CASES.items()}
    n = int(os.environ.get("BT_DIST_N", "400000"))
    s, t, r = config3_inputs(n // 2, n // 2)
    cases[f"config3-{n}"] = (s, dict(max_particles_in_box=30, targets=t, target_radii=r,
                                     stick_out_factor=0.25, extent_norm="linf",
                                     kind="adaptive-level-restricted"), {})
    cases[f"uniform-{n}"] = (uniform_particles(n, 3, np.float64), dict(max_particles_in_box=30), {})
    with open(os.path.join(os.path.dirname(__file__), "golden",
                           "refexec_distributed_digests.json")) as f:
        golden = json.load(f)
    all_ok = True
    for name, (src, tkw, vkw) in cases.items():
        rtree, want = oracle_ranks(src, tkw, vkw, world)
        out = run_rank(actx, comm, src, tkw, vkw)
        ref = golden.get(f"{name}:{world}")
        bad = check_rank(actx, rank, world, out, rtree, want[rank],
                         None if ref is None else ref[rank])
        ok = torch.tensor([0 if bad else 1], device=actx.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        print(f"{name}: rank {rank}/{world} nboxes={rtree.nboxes} "
              f"local src={int(out[3].shape[0])} tgt={int(out[4].shape[0])} "
              f"reference digests={'yes' if ref is not None else 'n/a'} "
              f"{'OK' if not bad else bad[:6]}", flush=True)
        all_ok = all_ok and int(ok.item()) == 1
    dist.barrier()
    if rank == 0:
        print("ALL OK" if all_ok else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
