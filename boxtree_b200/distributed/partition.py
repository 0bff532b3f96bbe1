"""Work partition and box masks: ``boxtree/distributed/partition.py`` on B200."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from .. import _cabi
from .._cabi import check, dptr


def get_box_ids_dfs_order(actx, tree):
    """``partition.py:38-57``: box ids in depth-first order, HIGHEST Morton child first (the
    reference pops a stack).  Computed on the device; returns a device int32 tensor."""
    lib = _cabi.load()
    nboxes = int(tree.nboxes)
    with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
        child_ids = _dev(actx, tree.box_child_ids)
        ls = _dev(actx, tree.level_start_box_nrs).to(torch.int32)
        size = actx.empty(max(nboxes, 1), np.int32)
        rank = actx.empty(max(nboxes, 1), np.int32)
        order = actx.empty(max(nboxes, 1), np.int32)
        check(lib.bt_dist_dfs_order(int(tree.dimensions), nboxes, int(child_ids.shape[-1]),
                                    int(tree.nlevels), dptr(ls), dptr(child_ids), dptr(size),
                                    dptr(rank), dptr(order), actx.stream_handle),
              "bt_dist_dfs_order")
    return order[:nboxes]


def _dev(actx, a):
    if isinstance(a, np.ndarray):
        a = actx.from_numpy(np.ascontiguousarray(a))
    return a.contiguous()


def partition_segments(cost_in_dfs_order, mpi_size, total_workload=None):
    """The root-rank loop of ``partition.py:81-116`` on the costs already arranged in DFS
    order: returns the ``(mpi_size, 2)`` int32 array of ``[start, end)`` DFS positions.

    The running sum is a sequential float accumulation exactly like the reference's
    ``workload_count += cost``; the thresholds use the reference's expression."""
    cost = np.asarray(cost_in_dfs_order)
    nboxes = len(cost)
    if total_workload is None:
        # callers with non-integer costs pass np.sum over the costs in BOX order, the
        # reference's summation order (partition.py:95; pairwise sums depend on the order)
        total_workload = np.sum(cost)
    segments = np.empty((mpi_size, 2), dtype=np.int32)
    cum = np.cumsum(cost)                       # sequential, like the += loop
    monotone = bool(np.all(cost >= 0))
    start = 0
    for segment_idx in range(mpi_size - 1):
        thr = (segment_idx + 1) * total_workload / mpi_size
        if monotone:
            i = int(np.searchsorted(cum, thr, side="right"))
        else:
            hit = np.nonzero(cum[start:] > thr)[0]
            i = start + int(hit[0]) if len(hit) else nboxes - 1
        i = min(max(i, start), nboxes - 1)
        segments[segment_idx] = [start, i + 1]
        start = i + 1
    segments[mpi_size - 1] = [start, nboxes]
    return segments


def partition_segments_device(actx, cost_per_box, dfs_order, mpi_size):
    """:func:`partition_segments` without moving the per-box arrays to the host, for costs that
    are non-negative integers in float64 with an exactly representable total (the default
    ``1 + particle counts``; any summation order then gives the reference's running sums bit
    for bit).  Returns ``None`` when the costs do not qualify (the caller takes the host path,
    which accumulates sequentially like the reference)."""
    cost = cost_per_box if isinstance(cost_per_box, torch.Tensor) else \
        actx.from_numpy(np.ascontiguousarray(np.asarray(cost_per_box, np.float64)))
    if cost.dtype != torch.float64:
        return None
    c = cost[dfs_order.long()]
    cum = torch.cumsum(c, 0)
    nboxes = int(c.shape[0])
    total = cum[-1]
    ok = ((c == torch.round(c)) & (c >= 0)).all() & (total < 2.0 ** 52)
    # thresholds with the reference's expression ((k + 1) * total / size, IEEE float64 on the
    # device as on the host) and ONE readback: [qualifies, cut positions...]
    ks = torch.arange(1, mpi_size, dtype=torch.float64, device=cum.device)
    hits = torch.searchsorted(cum, ks * total / mpi_size, right=True)
    packed = actx.read_back(torch.cat([ok.to(torch.int64).view(1), hits.to(torch.int64)]))
    if not packed[0]:
        return None
    hits = packed[1:]
    segments = np.empty((mpi_size, 2), dtype=np.int32)
    start = 0
    for k in range(mpi_size - 1):
        i = min(max(int(hits[k]), start), nboxes - 1)
        segments[k] = [start, i + 1]
        start = i + 1
    segments[mpi_size - 1] = [start, nboxes]
    return segments


def partition_segments_default_cost(actx, tree, dfs_order, mpi_size):
    """:func:`partition_segments` for the default cost ``1 + own sources + own targets`` of a
    box, on the device (``bt_dist_partition_cuts``: one scan of the costs gathered in depth-first
    order, one binary search per cut, ONE readback).  Integer costs: the running sums are exact,
    so the cuts are the reference's (``partition.py:81-116``) bit for bit."""
    lib = _cabi.load()
    nboxes = int(tree.nboxes)
    cuts = actx.empty(mpi_size, np.int64)
    check(lib.bt_dist_partition_cuts(nboxes, mpi_size, dptr(dfs_order),
                                     dptr(tree.box_source_counts_nonchild),
                                     dptr(tree.box_target_counts_nonchild), dptr(cuts),
                                     actx.stream_handle), "bt_dist_partition_cuts")
    hits = actx.read_back(cuts)[:mpi_size - 1]
    segments = np.empty((mpi_size, 2), dtype=np.int32)
    start = 0
    for k in range(mpi_size - 1):
        i = min(max(int(hits[k]), start), nboxes - 1)
        segments[k] = [start, i + 1]
        start = i + 1
    segments[mpi_size - 1] = [start, nboxes]
    return segments


def partition_work(actx, cost_per_box, traversal, comm):
    """``partition.py:60-121``.  *cost_per_box* (numpy) is only significant on the root rank.
    Returns the numpy array of boxes the calling rank is responsible for."""
    tree = traversal.tree
    mpi_rank, mpi_size = comm.Get_rank(), comm.Get_size()
    if mpi_size > tree.nboxes:
        raise RuntimeError("Fail to partition work because the number of boxes is "
                           "less than the number of processes.")
    dfs_order = get_box_ids_dfs_order(actx, tree).cpu().numpy()
    segments = None
    if mpi_rank == 0:
        cost = np.asarray(actx.to_numpy(cost_per_box) if isinstance(cost_per_box, torch.Tensor)
                          else cost_per_box)
        segments = partition_segments(cost[dfs_order], mpi_size, total_workload=np.sum(cost))
    mine = comm.scatter_rows(segments, root=0)
    return dfs_order[int(mine[0]):int(mine[1])]


@dataclass(frozen=True)
class BoxMasks:
    """``partition.py:300-327``: int8 device masks of length ``tree.nboxes``."""
    responsible_boxes: Any
    ancestor_boxes: Any
    point_src_boxes: Any
    multipole_src_boxes: Any


def _responsible_and_ancestors(actx, lib, tree, responsible_boxes_list):
    nb = int(tree.nboxes)
    sh = actx.stream_handle
    resp_list = _dev(actx, np.asarray(responsible_boxes_list, np.int32)
                     if not isinstance(responsible_boxes_list, torch.Tensor)
                     else responsible_boxes_list.to(torch.int32))
    responsible = actx.zeros(nb, np.int8)
    check(lib.bt_dist_mask_from_list(int(resp_list.shape[0]), dptr(resp_list),
                                     dptr(responsible), sh), "bt_dist_mask_from_list")
    ancestors = actx.zeros(nb, np.int8)
    check(lib.bt_dist_ancestor_mask(nb, dptr(responsible), dptr(_dev(actx, tree.box_parent_ids)),
                                    dptr(ancestors), sh), "bt_dist_ancestor_mask")
    return responsible, ancestors


def _masks_from_traversal(actx, lib, traversal, responsible, ancestors, into=None,
                          all_rows=False) -> BoxMasks:
    """Point-source and multipole-source masks from the rows of *traversal* that belong to
    responsible boxes / their ancestors (``partition.py:197-297``).  *into*: masks of an earlier
    call on other rows of the same global traversal, which this call adds to.  *all_rows*: the
    caller guarantees that every non-empty row of every list qualifies (a traversal that was
    built for exactly those rows): entries are marked without looking up their row."""
    tree = traversal.tree
    nb = int(tree.nboxes)
    sh = actx.stream_handle

    def add(box_list, mask_a, mask_b, starts, lists, out):
        if all_rows:
            check(lib.bt_dist_mark_list_boxes(int(lists.shape[0]), dptr(_dev(actx, lists)),
                                              dptr(out), sh), "bt_dist_mark_list_boxes")
            return
        check(lib.bt_dist_add_list_boxes(int(box_list.shape[0]), dptr(_dev(actx, box_list)),
                                         dptr(mask_a), dptr(mask_b), dptr(_dev(actx, starts)),
                                         dptr(_dev(actx, lists)), dptr(out), sh),
              "bt_dist_add_list_boxes")

    src = responsible.clone() if into is None else into.point_src_boxes
    add(traversal.target_boxes, responsible, None, traversal.neighbor_source_boxes_starts,
        traversal.neighbor_source_boxes_lists, src)
    add(traversal.target_or_target_parent_boxes, responsible, ancestors,
        traversal.from_sep_bigger_starts, traversal.from_sep_bigger_lists, src)
    if tree.targets_have_extent:
        if traversal.from_sep_close_smaller_starts is not None:
            add(traversal.target_boxes, responsible, None,
                traversal.from_sep_close_smaller_starts,
                traversal.from_sep_close_smaller_lists, src)
        if traversal.from_sep_close_bigger_starts is not None:
            add(traversal.target_boxes, responsible, ancestors,
                traversal.from_sep_close_bigger_starts,
                traversal.from_sep_close_bigger_lists, src)
    mpole = actx.zeros(nb, np.int8) if into is None else into.multipole_src_boxes
    add(traversal.target_or_target_parent_boxes, responsible, ancestors,
        traversal.from_sep_siblings_starts, traversal.from_sep_siblings_lists, mpole)
    for ilevel in range(int(tree.nlevels)):
        bl = traversal.from_sep_smaller_by_level[ilevel]
        add(traversal.target_boxes_sep_smaller_by_source_level[ilevel], responsible, None,
            bl.starts, bl.lists, mpole)
    return BoxMasks(responsible, ancestors, src, mpole)


def get_box_masks(actx, traversal, responsible_boxes_list) -> BoxMasks:
    """``partition.py:330-357``."""
    lib = _cabi.load()
    with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
        responsible, ancestors = _responsible_and_ancestors(actx, lib, traversal.tree,
                                                            responsible_boxes_list)
        return _masks_from_traversal(actx, lib, traversal, responsible, ancestors)


def get_box_masks_sharded(actx, tree, responsible_boxes_list, traversal_builder):
    """The same masks WITHOUT the global traversal (which every rank of the reference
    builds in full, ``distributed/__init__.py:201``): only the rows ``get_box_masks`` reads
    -- those of the rank's responsible boxes and their ancestors -- are built, by running
    the traversal builder on a copy of the tree whose target flags are cleared elsewhere.
    Row contents depend on geometry and source flags only, so the masks are identical.

    :returns: ``(BoxMasks, partial_traversal, need_mask)``"""
    import dataclasses
    lib = _cabi.load()
    nb = int(tree.nboxes)
    with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
        responsible, ancestors = _responsible_and_ancestors(actx, lib, tree, responsible_boxes_list)
        flags = actx.empty(nb, np.uint8)
        need = actx.empty(nb, np.int8)
        check(lib.bt_dist_restrict_target_flags(nb, dptr(_dev(actx, tree.box_flags)), dptr(responsible),
                                                dptr(ancestors), dptr(flags), dptr(need),
                                                actx.stream_handle), "bt_dist_restrict_target_flags")
        partial_tree = dataclasses.replace(tree, box_flags=flags)
        # lists 1 / 3 are only read on responsible rows (partition.py:214-218, 287-295)
        partial_trav, _ = traversal_builder(actx, partial_tree, _colleague_row_mask=need,
                                            _list13_row_mask=responsible, _keep_shared=True)
        masks = _masks_from_traversal(actx, lib, partial_trav, responsible, ancestors)
    return masks, partial_trav, need
