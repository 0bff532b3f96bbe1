"""CPU oracle for the distributed setup rows.

TEST INFRASTRUCTURE ONLY.  PARITY PINNED: the reference's own ``partition_work``,
``get_box_masks``, ``generate_local_tree`` and ``generate_local_travs`` run on the CPU through
``tests/refexec`` (one thread per rank) and every per-rank output equals this restatement's
(``tests/test_refexec.py``, ``tests/golden/refexec_distributed_digests.json``).
numpy restatement of
* ``boxtree/distributed/partition.py:38-121`` (``get_box_ids_dfs_order``,
  ``partition_work`` without the MPI Scatter: all segments are returned),
* ``boxtree/distributed/partition.py:174-357`` (``get_box_masks``),
* ``boxtree/distributed/local_tree.py:198-284, 316-495`` (``generate_local_tree``; the
  Gather/bcast of the multipole masks is replaced by passing all ranks' masks),
* ``boxtree/distributed/local_traversal.py:34-62`` (``generate_local_travs``).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Any

import numpy as np

from .traversal import build_traversal, merge_close_lists

IS_TARGET, HAS_TGT_CHILD = 2, 8


def get_box_ids_dfs_order(tree):
    """partition.py:38-57: explicit stack, children pushed in Morton order and popped
    last-first."""
    nb = tree.nboxes
    dfs_order = np.empty(nb, dtype=np.int32)
    idx = 0
    stack = [0]
    child_ids = tree.box_child_ids
    while stack:
        box_id = stack.pop()
        dfs_order[idx] = box_id
        idx += 1
        for i in range(child_ids.shape[0]):
            c = child_ids[i][box_id]
            if c > 0:
                stack.append(int(c))
    return dfs_order


def partition_work(cost_per_box, tree, mpi_size):
    """partition.py:60-121 -> list of per-rank responsible box arrays."""
    if mpi_size > tree.nboxes:
        raise RuntimeError("Fail to partition work because the number of boxes is "
                           "less than the number of processes.")
    dfs_order = get_box_ids_dfs_order(tree)
    total_workload = np.sum(cost_per_box)
    segments = np.empty((mpi_size, 2), dtype=np.int32)
    segment_idx = 0
    start = 0
    workload_count = 0
    for box_idx_dfs_order in range(tree.nboxes):
        if segment_idx + 1 == mpi_size:
            segments[segment_idx, :] = [start, tree.nboxes]
            break
        box_idx = dfs_order[box_idx_dfs_order]
        workload_count += cost_per_box[box_idx]
        if (workload_count > (segment_idx + 1) * total_workload / mpi_size
                or box_idx_dfs_order == tree.nboxes - 1):
            segments[segment_idx, :] = [start, box_idx_dfs_order + 1]
            start = box_idx_dfs_order + 1
            segment_idx += 1
    return [dfs_order[s:e] for s, e in segments], segments


@dataclass
class BoxMasks:
    responsible_boxes: np.ndarray
    ancestor_boxes: np.ndarray
    point_src_boxes: np.ndarray
    multipole_src_boxes: np.ndarray


def _add_interaction_list_boxes(box_list, mask, starts, lists, out_mask):
    """partition.py:135-162"""
    starts = np.asarray(starts, np.int64)
    rows = np.nonzero(mask[box_list] != 0)[0]
    for i in rows:
        out_mask[lists[starts[i]:starts[i + 1]]] = 1


def get_box_masks(trav, responsible_boxes_list) -> BoxMasks:
    tree = trav.tree
    nb = tree.nboxes
    responsible = np.zeros(nb, np.int8)
    responsible[responsible_boxes_list] = 1
    # partition.py:174-194
    ancestors = np.zeros(nb, np.int8)
    last = responsible.copy()
    parents = tree.box_parent_ids[:nb]
    while last.any():
        new = np.zeros(nb, np.int8)
        cur = np.nonzero(last)[0]
        cur = cur[cur != 0]
        new[parents[cur]] = 1
        new = new & (~ancestors)
        ancestors = ancestors | new
        last = new
    # partition.py:197-252
    src = responsible.copy()
    _add_interaction_list_boxes(trav.target_boxes, responsible,
                                trav.neighbor_source_boxes_starts,
                                trav.neighbor_source_boxes_lists, src)
    _add_interaction_list_boxes(trav.target_or_target_parent_boxes, responsible | ancestors,
                                trav.from_sep_bigger_starts, trav.from_sep_bigger_lists, src)
    if tree.targets_have_extent:
        if trav.from_sep_close_smaller_starts is not None:
            _add_interaction_list_boxes(trav.target_boxes, responsible,
                                        trav.from_sep_close_smaller_starts,
                                        trav.from_sep_close_smaller_lists, src)
        if trav.from_sep_close_bigger_starts is not None:
            _add_interaction_list_boxes(trav.target_boxes, responsible | ancestors,
                                        trav.from_sep_close_bigger_starts,
                                        trav.from_sep_close_bigger_lists, src)
    # partition.py:255-297
    mpole = np.zeros(nb, np.int8)
    _add_interaction_list_boxes(trav.target_or_target_parent_boxes, responsible | ancestors,
                                trav.from_sep_siblings_starts, trav.from_sep_siblings_lists, mpole)
    for ilevel in range(tree.nlevels):
        _add_interaction_list_boxes(trav.target_boxes_sep_smaller_by_source_level[ilevel],
                                    responsible, trav.from_sep_smaller_by_level[ilevel].starts,
                                    trav.from_sep_smaller_by_level[ilevel].lists, mpole)
    return BoxMasks(responsible, ancestors, src, mpole)


def _local_particles_and_lists(box_mask, particles, radii, starts, counts_nonchild, counts_cumul):
    """local_tree.py:198-284"""
    n = len(particles[0])
    nb = len(box_mask)
    starts = np.asarray(starts[:nb], np.int64)
    particle_mask = np.zeros(n, np.int32)
    for b in np.nonzero(box_mask)[0]:
        particle_mask[starts[b]:starts[b] + counts_nonchild[b]] = 1
    g2l = np.zeros(n + 1, np.int32)
    g2l[1:] = np.cumsum(particle_mask)
    sel = particle_mask.astype(bool)
    local_particles = [p[sel] for p in particles]
    local_radii = radii[sel] if radii is not None else None
    local_starts = g2l[starts]
    local_nonchild = np.where(box_mask, counts_nonchild[:nb], 0).astype(np.int32)
    ends = starts + counts_cumul[:nb]
    local_cumul = (g2l[ends] - g2l[starts]).astype(np.int32)
    return (local_particles, local_radii, local_starts.astype(np.int32), local_nonchild,
            local_cumul, np.arange(n)[sel])


def mask_compressor(mask2d):
    """tools.py:647-740 (2-D case): CSR of the true columns of every row."""
    counts = mask2d.astype(bool).sum(axis=1)
    starts = np.zeros(mask2d.shape[0] + 1, np.int32)
    starts[1:] = np.cumsum(counts)
    lists = np.nonzero(mask2d)[1].astype(np.int32)
    return starts, lists


def generate_local_tree(global_trav, responsible_boxes_list, multipole_masks_all_ranks):
    """local_tree.py:316-495.  *multipole_masks_all_ranks*: [nranks, nboxes] int8, what the
    reference Gathers on the root rank."""
    gt = global_trav.tree
    nb = gt.nboxes
    masks = get_box_masks(global_trav, responsible_boxes_list)
    src = _local_particles_and_lists(masks.point_src_boxes, gt.sources,
                                     gt.source_radii if gt.sources_have_extent else None,
                                     gt.box_source_starts, gt.box_source_counts_nonchild,
                                     gt.box_source_counts_cumul)
    tgt = _local_particles_and_lists(masks.responsible_boxes, gt.targets,
                                     gt.target_radii if gt.targets_have_extent else None,
                                     gt.box_target_starts, gt.box_target_counts_nonchild,
                                     gt.box_target_counts_cumul)
    b2u_starts, b2u_lists = mask_compressor(np.ascontiguousarray(multipole_masks_all_ranks.T))
    # local_tree.py:163-185 (modify_target_flags)
    flags = gt.box_flags[:nb].copy()
    flags &= np.uint8(~IS_TARGET & 0xff)
    flags &= np.uint8(~HAS_TGT_CHILD & 0xff)
    flags[tgt[3] != 0] |= IS_TARGET
    flags[tgt[3] < tgt[4]] |= HAS_TGT_CHILD
    local_tree = replace(
        gt, sources=src[0], targets=tgt[0],
        source_radii=src[1] if gt.sources_have_extent else None,
        target_radii=tgt[1] if gt.targets_have_extent else None,
        box_source_starts=src[2], box_source_counts_nonchild=src[3],
        box_source_counts_cumul=src[4], box_target_starts=tgt[2],
        box_target_counts_nonchild=tgt[3], box_target_counts_cumul=tgt[4],
        box_flags=flags, user_source_ids=None, sorted_target_ids=None,
        box_parent_ids=gt.box_parent_ids[:nb], box_levels=gt.box_levels[:nb])
    local_tree.extra = dict(
        responsible_boxes_list=np.asarray(responsible_boxes_list),
        responsible_boxes_mask=masks.responsible_boxes, ancestor_mask=masks.ancestor_boxes,
        point_src_boxes=masks.point_src_boxes, multipole_src_boxes=masks.multipole_src_boxes,
        box_to_user_rank_starts=b2u_starts, box_to_user_rank_lists=b2u_lists)
    return local_tree, src[5], tgt[5]


def generate_local_travs(local_tree, well_sep_is_n_away=1, from_sep_smaller_crit=None,
                         merge_close=False):
    """local_traversal.py:34-62"""
    trav = build_traversal(local_tree, well_sep_is_n_away=well_sep_is_n_away,
                           from_sep_smaller_crit=from_sep_smaller_crit,
                           source_boxes_mask=local_tree.extra["responsible_boxes_mask"],
                           source_parent_boxes_mask=local_tree.extra["ancestor_mask"])
    if merge_close and local_tree.targets_have_extent:
        trav = merge_close_lists(trav)
    return trav
