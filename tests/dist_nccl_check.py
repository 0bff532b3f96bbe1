"""Multi-GPU check (run under torchrun, one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_nccl_check.py

Rank 0 builds the global tree; ``distributed_setup`` broadcasts it over NCCL, every rank
builds its local tree / local traversal, and each rank compares its result bit for bit with
the oracle's restatement of the reference's distributed setup for the same rank."""
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

    from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder
    from boxtree_b200 import distributed as bd
    from oracle import distributed as od
    from oracle.traversal import build_traversal
    from oracle.tree_build import build_tree
    from tests.parity_util import config3_inputs, trav_mismatches

    actx = TorchArrayContext(f"cuda:{local_rank}")
    comm = bd.TorchDistComm()
    n = int(os.environ.get("BT_DIST_N", "200000"))
    src, tgt, radii = config3_inputs(n // 2, n // 2)
    kw = dict(max_particles_in_box=30, stick_out_factor=0.25, extent_norm="linf",
              kind="adaptive-level-restricted")
    tree = None
    if rank == 0:
        tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(x) for x in src],
                                    targets=[actx.from_numpy(x) for x in tgt],
                                    target_radii=actx.from_numpy(radii), **kw)
    tg = FMMTraversalBuilder(actx)
    local_tree, local_trav, src_idx, tgt_idx, global_trav = bd.distributed_setup(
        actx, tree, tg, comm)

    # oracle for this rank
    rtree = build_tree(src, targets=tgt, target_radii=radii, **kw)
    rtrav = build_traversal(rtree)
    nb = rtree.nboxes
    cost = (1.0 + rtree.box_source_counts_nonchild[:nb]
            + rtree.box_target_counts_nonchild[:nb]).astype(np.float64)
    resp, _ = od.partition_work(cost, rtree, world)
    mp = np.stack([od.get_box_masks(rtrav, resp[r]).multipole_src_boxes for r in range(world)])
    wt, wsrc, wtgt = od.generate_local_tree(rtrav, resp[rank], mp)
    wtrav = od.generate_local_travs(wt)
    g = actx.to_numpy(local_tree)
    bad = []
    if not np.array_equal(np.asarray(g.responsible_boxes_list), resp[rank]):
        bad.append("responsible_boxes_list")
    for f in ("box_source_starts", "box_source_counts_cumul", "box_target_starts",
              "box_target_counts_nonchild", "box_flags"):
        if not np.array_equal(np.asarray(getattr(g, f)), np.asarray(getattr(wt, f))):
            bad.append(f)
    for f in ("box_to_user_rank_starts", "box_to_user_rank_lists"):
        if not np.array_equal(np.asarray(getattr(g, f)), wt.extra[f]):
            bad.append(f)
    if not np.array_equal(src_idx.cpu().numpy(), wsrc) or not np.array_equal(tgt_idx.cpu().numpy(), wtgt):
        bad.append("particle idx")
    bad += trav_mismatches(wtrav, actx.to_numpy(local_trav))
    ok = torch.tensor([0 if bad else 1], device=actx.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    print(f"rank {rank}/{world}: nresp={len(resp[rank])} local src={len(wsrc)} tgt={len(wtgt)} "
          f"{'OK' if not bad else bad[:6]}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
