"""Shared by the distributed-build parity tests (in-process ranks on one GPU,
``tests/test_gpu_dist_build.py``) and the NCCL check under torchrun
(``tests/dist_build_nccl_check.py``): slicing of the inputs over ranks, the per-rank run, and
the comparison of one rank's outputs with the single-GPU tree, the oracle's restatement of the
reference's distributed flow and the committed digests of the reference's own run."""
from __future__ import annotations

import numpy as np
import torch

TREE_BOX_FIELDS = (
    "level_start_box_nrs", "box_source_starts", "box_source_counts_nonchild",
    "box_source_counts_cumul", "box_target_starts", "box_target_counts_nonchild",
    "box_target_counts_cumul", "box_parent_ids", "box_child_ids", "box_centers", "box_levels",
    "box_flags", "box_source_bounding_box_min", "box_source_bounding_box_max",
    "box_target_bounding_box_min", "box_target_bounding_box_max")


def rank_slice(a, rank, size):
    n = int(a.shape[0])
    return a[rank * n // size:(rank + 1) * n // size]


def slice_inputs(src, tkw, rank, size):
    """Rank *rank*'s contiguous slice of every particle array (the global particle set is the
    concatenation over ranks)."""
    s = [np.ascontiguousarray(rank_slice(x, rank, size)) for x in src]
    kw = {}
    for k, v in tkw.items():
        if k == "targets":
            kw[k] = [np.ascontiguousarray(rank_slice(x, rank, size)) for x in v]
        elif isinstance(v, np.ndarray) and k in ("target_radii", "source_radii"):
            kw[k] = np.ascontiguousarray(rank_slice(v, rank, size))
        else:
            kw[k] = v
    return s, kw


def run_rank(actx, comm, src, tkw, vkw, defer=None):
    """The distributed build + setup of one rank (collective).  *defer*: the build returns with
    the all-reduce of the particle extents in flight, as ``bench.py`` runs it (default; set
    ``BT_DIST_DEFER=0`` or pass False for the build that completes them itself)."""
    import os
    if defer is None:
        defer = os.environ.get("BT_DIST_DEFER", "1") != "0"
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    from boxtree_b200 import distributed as bd
    rank, size = comm.Get_rank(), comm.Get_size()
    s, kw = slice_inputs(src, tkw, rank, size)
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in kw.items()}
    dtree = bd.build_distributed_tree(actx, TreeBuilder(actx), comm,
                                      [actx.from_numpy(x) for x in s], defer_extents=defer, **dkw)
    assert (dtree.pending is not None) == bool(defer)
    tg = FMMTraversalBuilder(actx, **vkw)
    lt, ltrav, sidx, tidx = bd.distributed_tree_setup(actx, dtree, tg, comm)
    return dtree, lt, ltrav, sidx, tidx


def check_rank(actx, rank, size, out, ref_tree, want, reference_digests=None):
    """*ref_tree*: numpy single-GPU (or oracle) tree of the concatenated input; *want*: the
    oracle's ``(resp, masks, local_tree, src_idx, tgt_idx, local_trav)`` of this rank.
    Returns a list of mismatching field names."""
    from tests.parity_util import (bits_equal, digest_mismatches, distributed_rank_digests,
                                   trav_mismatches)
    dtree, lt, ltrav, sidx, tidx = out
    bad = []
    d = actx.to_numpy(dtree)
    nb = ref_tree.nboxes
    if d.nboxes != nb:
        return [f"nboxes {d.nboxes} != {nb}"]
    for f in TREE_BOX_FIELDS:
        a, b = np.asarray(getattr(d, f)), np.asarray(getattr(ref_tree, f))
        if a.dtype != b.dtype or a.shape != b.shape or not bits_equal(a, b):
            bad.append("dtree." + f)
    for f in ("root_extent", "stick_out_factor", "extent_norm"):
        if getattr(d, f) != getattr(ref_tree, f):
            bad.append("dtree." + f)
    wresp, wmasks, wt, wsrc, wtgt, wtrav = want
    g = actx.to_numpy(lt)
    if not np.array_equal(np.asarray(g.responsible_boxes_list), wresp):
        bad.append("responsible_boxes_list")
    if not np.array_equal(sidx.cpu().numpy(), wsrc):
        bad.append("src_idx")
    if not np.array_equal(tidx.cpu().numpy(), wtgt):
        bad.append("tgt_idx")
    for f in ("box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
              "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul",
              "box_flags", "box_parent_ids", "box_levels", "box_child_ids"):
        a, b = np.asarray(getattr(g, f)), np.asarray(getattr(wt, f))
        if a.dtype != b.dtype or not np.array_equal(a, b):
            bad.append("local_tree." + f)
    for ax in range(ref_tree.dimensions):
        if not bits_equal(g.sources[ax], wt.sources[ax]):
            bad.append(f"local_tree.sources[{ax}]")
        if not bits_equal(g.targets[ax], wt.targets[ax]):
            bad.append(f"local_tree.targets[{ax}]")
    if ref_tree.targets_have_extent and not bits_equal(g.target_radii, wt.target_radii):
        bad.append("local_tree.target_radii")
    for f in ("box_to_user_rank_starts", "box_to_user_rank_lists", "responsible_boxes_mask",
              "ancestor_mask"):
        if not np.array_equal(np.asarray(getattr(g, f)), wt.extra[f]):
            bad.append("local_tree." + f)
    gtrav = actx.to_numpy(ltrav)
    tb = [b for b in trav_mismatches(wtrav, gtrav) if "same_level_non_well_sep" not in b]
    bad += tb
    if reference_digests is not None and not bad:
        # colleague rows are only built where they are read: digest the rest
        got = distributed_rank_digests(
            wresp, {f: np.asarray(getattr(wmasks, f)) for f in (
                "responsible_boxes", "ancestor_boxes", "point_src_boxes", "multipole_src_boxes")},
            {**{f: getattr(g, f) for f in (
                "box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
                "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul",
                "box_flags", "box_parent_ids", "box_levels", "box_child_ids",
                "box_to_user_rank_starts", "box_to_user_rank_lists", "responsible_boxes_mask",
                "ancestor_mask", "sources", "targets")},
             "target_radii": g.target_radii if ref_tree.targets_have_extent else None},
            sidx.cpu().numpy(), tidx.cpu().numpy(), gtrav, nb)
        skip = ("mask.", "trav.same_level_non_well_sep")
        bad += ["reference digest: " + k for k in digest_mismatches(reference_digests, got)
                if not k.startswith(skip)]
    return bad


def oracle_ranks(src, tkw, vkw, size):
    """The oracle's restatement of the reference's distributed flow for every rank."""
    from oracle import distributed as od
    from oracle.traversal import build_traversal
    from oracle.tree_build import build_tree
    from tests.dist_cases import box_cost
    rtree = build_tree(src, **tkw)
    rtrav = build_traversal(rtree, **vkw)
    resp, _ = od.partition_work(box_cost(rtree), rtree, size)
    masks = [od.get_box_masks(rtrav, resp[r]) for r in range(size)]
    mp = np.stack([m.multipole_src_boxes for m in masks])
    want = []
    for r in range(size):
        wt, wsrc, wtgt = od.generate_local_tree(rtrav, resp[r], mp)
        want.append((resp[r], masks[r], wt, wsrc, wtgt, od.generate_local_travs(wt, **vkw)))
    return rtree, want


def run_threads(size, fn):
    """Run ``fn(comm)`` on *size* in-process ranks (one thread each); returns the results."""
    import threading

    from boxtree_b200.distributed import ThreadComm, ThreadGroup
    group = ThreadGroup(size)
    results, errors = [None] * size, [None] * size

    def work(r):
        try:
            results[r] = fn(ThreadComm(group, r))
        except BaseException as e:  # noqa: BLE001
            errors[r] = e
            group.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results
