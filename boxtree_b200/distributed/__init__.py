"""Distributed tree/traversal setup (``boxtree/distributed/``), B200-native.

One process per GPU; the plumbing is ``torch.distributed`` (NCCL over NVLink 5 /
NVSwitch), the per-rank work is CUDA kernels of ``csrc/distributed.cu``.  Mirrors the
reference's setup functions: ``partition_work``, ``get_box_masks``,
``generate_local_tree``, ``generate_local_travs``
(``boxtree/distributed/__init__.py:156-266`` without the FMM wrangler)."""
from __future__ import annotations

import numpy as np
import torch

from .comm import SingleProcessComm, ThreadComm, ThreadGroup, TorchDistComm
from .exchange import _local_ranges, exchange_particle_kinds, exchange_particles, preorder
from .local_traversal import generate_local_travs
from .local_tree import LocalTree, assemble_local_tree, box_to_user_rank, generate_local_tree
from .partition import (BoxMasks, get_box_ids_dfs_order, get_box_masks, get_box_masks_sharded,
                        partition_segments, partition_segments_default_cost,
                        partition_segments_device, partition_work)
from .tree_build import DistributedTree, build_distributed_tree

__all__ = [
    "SingleProcessComm", "TorchDistComm", "LocalTree", "BoxMasks", "get_box_ids_dfs_order",
    "partition_segments", "partition_work", "get_box_masks", "generate_local_tree",
    "box_to_user_rank", "generate_local_travs", "broadcast_tree", "distributed_setup",
    "get_box_masks_sharded", "allgather_particles", "sharded_setup", "ThreadComm", "ThreadGroup",
    "DistributedTree", "build_distributed_tree", "exchange_particles", "distributed_tree_setup",
]


def broadcast_tree(actx, tree, comm, root=0):
    """Ship the global tree from the root to every rank (``distributed/__init__.py:185-199``
    does this with a pickled ``comm.bcast``); arrays travel as raw buffers."""
    import dataclasses

    from ..array_context import make_obj_array
    from ..tree import Tree
    if comm.Get_size() == 1:
        return tree
    names = [f.name for f in dataclasses.fields(Tree)]
    is_root = comm.Get_rank() == root

    def kind_of(v):
        if isinstance(v, torch.Tensor):
            return ("tensor", None)
        if isinstance(v, np.ndarray) and v.dtype == object:
            return ("objarray", len(v))
        if isinstance(v, tuple) and all(isinstance(x, np.ndarray) for x in v):
            return ("nptuple", len(v))
        return ("object", v)

    manifest = [[(name, *kind_of(getattr(tree, name))) for name in names]] if is_root else [None]
    comm.dist.broadcast_object_list(manifest, src=root, group=comm.group)
    out = {}
    for name, kind, extra in manifest[0]:
        v = getattr(tree, name) if is_root else None
        if kind == "tensor":
            out[name] = actx.from_numpy(comm.bcast_array(actx.to_numpy(v) if is_root else None, root))
        elif kind == "objarray":
            out[name] = make_obj_array([
                actx.from_numpy(comm.bcast_array(actx.to_numpy(v[i]) if is_root else None, root))
                for i in range(extra)])
        elif kind == "nptuple":
            out[name] = tuple(comm.bcast_array(v[i] if is_root else None, root)
                              for i in range(extra))
        else:
            out[name] = extra
    if out["sources_are_targets"]:
        # keep the reference's aliasing (tree_build.py:1469-1474, 1572, 1739-1741)
        for a, b in (("targets", "sources"), ("box_target_starts", "box_source_starts"),
                     ("box_target_counts_nonchild", "box_source_counts_nonchild"),
                     ("box_target_counts_cumul", "box_source_counts_cumul"),
                     ("box_target_bounding_box_min", "box_source_bounding_box_min"),
                     ("box_target_bounding_box_max", "box_source_bounding_box_max")):
            out[a] = out[b]
    return Tree(**out)


def distributed_setup(actx, global_tree, traversal_builder, comm, cost_per_box=None,
                      merge_close_lists=False, level_orders=None, calibration_params=None):
    """The tree/traversal part of ``make_distributed_wrangler``
    (``distributed/__init__.py:156-266``): broadcast the root's global tree, build the
    global traversal on every rank, partition the boxes by cost in DFS order, build the
    rank's local tree and local traversal.

    *cost_per_box* (root only): with *level_orders* (the wrangler's expansion order per level)
    it is the reference's choice, ``FMMCostModel().cost_per_box`` on the global traversal with
    unit calibration parameters unless *calibration_params* is given
    (``distributed/__init__.py:208-230``); otherwise ``1 + own source count + own target count``.
    Returns ``(local_tree, local_trav, src_idx, tgt_idx, global_trav)``."""
    tree = broadcast_tree(actx, global_tree, comm)
    global_trav, _ = traversal_builder(actx, tree)
    if cost_per_box is None and level_orders is not None and comm.Get_rank() == 0:
        from ..cost import FMMCostModel
        if calibration_params is None:
            calibration_params = FMMCostModel.get_unit_calibration_params()
        cost_per_box = FMMCostModel().cost_per_box(
            actx, global_trav, level_orders, dict(calibration_params)).cpu().numpy()
    if cost_per_box is None and comm.Get_rank() == 0:
        cost_per_box = (1.0 + tree.box_source_counts_nonchild.double()
                        + tree.box_target_counts_nonchild.double()).cpu().numpy()
    responsible = partition_work(actx, cost_per_box, global_trav, comm)
    local_tree, src_idx, tgt_idx = generate_local_tree(actx, global_trav, responsible, comm)
    local_trav = generate_local_travs(actx, local_tree, traversal_builder, merge_close_lists)
    return local_tree, local_trav, src_idx, tgt_idx, global_trav


def allgather_particles(actx, comm, arrays):
    """Every rank contributes its slice of each 1-D array; all ranks end up with the
    concatenation in rank order (NCCL all-gather over NVLink).  Slices may differ in length."""
    if comm.Get_size() == 1:
        return list(arrays)
    size = comm.Get_size()
    n_local = torch.tensor([int(arrays[0].shape[0])], dtype=torch.int64, device=actx.device)
    counts = comm.allgather_tensor(n_local).view(-1).cpu().tolist()
    nmax = max(counts)
    out = []
    for a in arrays:
        pad = a if int(a.shape[0]) == nmax else torch.cat(
            [a, torch.zeros(nmax - int(a.shape[0]), dtype=a.dtype, device=a.device)])
        g = comm.allgather_tensor(pad)                       # [size, nmax]
        out.append(g.reshape(-1) if all(c == nmax for c in counts)
                   else torch.cat([g[r, :counts[r]] for r in range(size)]))
    return out


def sharded_setup(actx, tree, traversal_builder, comm, cost_per_box=None,
                  merge_close_lists=False):
    """Scalable variant of :func:`distributed_setup` for a global tree that is already
    replicated on every rank: no global traversal is built.  Every rank partitions the boxes
    itself (same deterministic host arithmetic on every rank), builds only the traversal rows
    of its responsible boxes / their ancestors to derive the box masks, then its local tree
    and local traversal.  Masks, local tree and local traversal are identical to those of
    :func:`distributed_setup`; the traversal work per rank is ~2/nranks of the global one.

    Returns ``(local_tree, local_trav, src_idx, tgt_idx)``."""
    from types import SimpleNamespace
    rank, size = comm.Get_rank(), comm.Get_size()
    if cost_per_box is None:
        cost_per_box = (1.0 + tree.box_source_counts_nonchild.double()
                        + tree.box_target_counts_nonchild.double())
    dfs_order = get_box_ids_dfs_order(actx, tree)
    with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
        segs = partition_segments_device(actx, cost_per_box, dfs_order, size)
    if segs is None:      # general float costs: the reference's sequential accumulation, on the host
        cost_host = cost_per_box.cpu().numpy() if isinstance(cost_per_box, torch.Tensor) \
            else np.asarray(cost_per_box)
        segs = partition_segments(cost_host[dfs_order.cpu().numpy()], size)
    seg = segs[rank]
    responsible = dfs_order[int(seg[0]):int(seg[1])]
    masks, _partial, need = get_box_masks_sharded(actx, tree, responsible, traversal_builder)
    local_tree, src_idx, tgt_idx = generate_local_tree(
        actx, SimpleNamespace(tree=tree), responsible, comm, box_masks=masks)
    # same box geometry, source flags and colleague rows as the partial traversal: its depth-first
    # ranks, colleagues and list-2 masks are reused
    shared = getattr(traversal_builder, "last_shared", None)
    local_trav, _ = traversal_builder(
        actx, local_tree, source_boxes_mask=local_tree.responsible_boxes_mask,
        source_parent_boxes_mask=local_tree.ancestor_mask, _colleague_row_mask=need, _shared=shared)
    traversal_builder.last_shared = None
    if merge_close_lists and local_tree.targets_have_extent:
        local_trav = local_trav.merge_close_lists(actx)
    return local_tree, local_trav, src_idx, tgt_idx


def distributed_tree_setup(actx, dtree, traversal_builder, comm, cost_per_box=None,
                           merge_close_lists=False, traversal_pieces=None):
    """The tree/traversal part of ``make_distributed_wrangler``
    (``distributed/__init__.py:156-266``) for a :class:`DistributedTree`: no rank holds all
    particles and nothing is broadcast.  Collective over *comm*.

    Every rank partitions the boxes itself (``partition.py:60-121``: DFS-order cost prefix, the
    same deterministic arithmetic on the replicated box arrays) and builds its LOCAL traversal
    (``local_traversal.py:37-60``) straight away -- the local target flags follow from the
    responsible boxes and the global per-box target counts, no particle is needed.  Its rows
    are rows of the global traversal, so the box masks (``partition.py:330-357``) are read off
    it, plus a small second traversal over the few rows that the global traversal has on
    responsible boxes / their ancestors and the local one lacks (boxes that are target or
    target-parent boxes only through other ranks' particles).  The masks are all-gathered, and
    the sources of the rank's point-source boxes and the targets of its responsible boxes arrive
    from their owners in one all-to-all each (:func:`exchange_particles`, replacing
    ``local_tree.py:408-470``).  Local tree, local traversal, masks and index arrays are
    identical to the reference flow's on the concatenated particle set.

    Returns ``(local_tree, local_trav, src_idx, tgt_idx)``; *local_trav* is a list of row
    pieces when the rank's lists do not fit the int32 CSR range (or *traversal_pieces* > 1)."""
    import dataclasses

    from .._cabi import check, dptr, load
    from .._timing import mark
    from .partition import _masks_from_traversal, _responsible_and_ancestors
    lib = load()
    rank, size = comm.Get_rank(), comm.Get_size()
    nb = int(dtree.nboxes)
    sh = actx.stream_handle
    with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
        mark("start")
        dfs_order = get_box_ids_dfs_order(actx, dtree)
        segs = None
        if cost_per_box is None:
            # the default cost, 1 + own sources + own targets: cut positions from one scan
            segs = partition_segments_default_cost(actx, dtree, dfs_order, size)
        if segs is None:
            if cost_per_box is None:
                cost_per_box = (1.0 + dtree.box_source_counts_nonchild.double()
                                + dtree.box_target_counts_nonchild.double())
            segs = partition_segments_device(actx, cost_per_box, dfs_order, size)
        if segs is None:
            cost_host = cost_per_box.cpu().numpy() if isinstance(cost_per_box, torch.Tensor) \
                else np.asarray(cost_per_box)
            segs = partition_segments(cost_host[dfs_order.cpu().numpy()], size,
                                      total_workload=np.sum(cost_host))
        seg = segs[rank]
        responsible = dfs_order[int(seg[0]):int(seg[1])]
        resp_mask, anc_mask = _responsible_and_ancestors(actx, lib, dtree, responsible)
        need = resp_mask | anc_mask
        mark("ds:partition")

        # {{{ local target ranges and flags: targets of the responsible boxes (local_tree.py:163-185)

        pre = preorder(actx, dtree)
        tgt_ranges = _local_ranges(actx, lib, nb, resp_mask, dtree.box_target_counts_nonchild, pre)
        local_flags = dtree.box_flags.clone()
        check(lib.bt_dist_modify_target_flags(nb, dptr(tgt_ranges[1]), dptr(tgt_ranges[2]),
                                              dptr(local_flags), sh), "bt_dist_modify_target_flags")
        corner_flags = actx.empty(nb, np.uint8)
        any_corner = actx.empty(1, np.int32)
        check(lib.bt_dist_corner_flags(nb, dptr(dtree.box_flags), dptr(local_flags), dptr(resp_mask),
                                       dptr(anc_mask), dptr(corner_flags), dptr(any_corner), sh),
              "bt_dist_corner_flags")

        # }}}

        # {{{ local traversal, then the masks from its rows + the corner rows

        # a deferred build (``build_distributed_tree(defer_extents=True)``): the all-reduce of the
        # particle extents has been running beside the partition; it is waited for where the
        # traversal first reads the extents (after the colleague pass), at the latest below
        pending = getattr(dtree, "pending", None)

        def finish_pending():
            if pending is not None:
                pending.finish()

        def local_piece(row_mask):
            # rows of the boxes of row_mask only (None: all rows of the local traversal)
            fl = local_flags
            if row_mask is not None:
                fl = actx.empty(nb, np.uint8)
                scratch = actx.empty(nb, np.int8)
                check(lib.bt_dist_restrict_target_flags(
                    nb, dptr(local_flags), dptr(row_mask), dptr(row_mask), dptr(fl), dptr(scratch),
                    sh), "bt_dist_restrict_target_flags")
            prev = traversal_builder.last_shared        # colleagues etc. of an earlier piece
            return traversal_builder(
                actx, dataclasses.replace(dtree, box_flags=fl),
                source_boxes_mask=resp_mask if row_mask is None else (resp_mask & row_mask),
                source_parent_boxes_mask=anc_mask if row_mask is None else (anc_mask & row_mask),
                _colleague_row_mask=need, _keep_shared=prev is None, _shared=prev,
                _before_extents=finish_pending, _preorder=(pre[2], pre[0]))[0]

        # one piece, or -- when a list of the rank's rows exceeds the int32 CSR range
        # (``traversal_pieces`` > 1, or on OverflowError) -- row pieces as in
        # FMMTraversalBuilder.build_in_chunks: piece k holds the rows of the k-th part of the
        # responsible boxes (depth-first order), piece 0 also those of the ancestors
        npieces = max(int(traversal_pieces or 1), 1)
        traversal_builder.last_shared = None
        while True:
            try:
                if npieces == 1:
                    local_travs = [local_piece(None)]
                else:
                    local_travs = []
                    nresp = int(responsible.shape[0])
                    for k in range(npieces):
                        part = responsible[k * nresp // npieces:(k + 1) * nresp // npieces]
                        pm = actx.zeros(nb, np.int8)
                        check(lib.bt_dist_mask_from_list(int(part.shape[0]), dptr(part.contiguous()),
                                                         dptr(pm), sh), "bt_dist_mask_from_list")
                        if k == 0:
                            pm = pm | (anc_mask & ~resp_mask)
                        local_travs.append(local_piece(pm))
                break
            except OverflowError:
                if traversal_pieces:
                    raise
                npieces *= 2
                traversal_builder.last_shared = None
                torch.cuda.empty_cache()
        shared = traversal_builder.last_shared
        traversal_builder.last_shared = None
        finish_pending()
        mark("ds:local traversal")
        # every row of the local traversal is a row the masks read: its target boxes are
        # responsible boxes, its target-or-target-parent boxes responsible boxes or ancestors;
        # the corner traversal has no list 1 / 3 rows and only rows of such boxes otherwise
        masks = None
        for piece in local_travs:
            masks = _masks_from_traversal(actx, lib, piece, resp_mask, anc_mask, into=masks,
                                          all_rows=True)
        # rows of lists 2 / 4 / 4-close that the global traversal has on responsible boxes and
        # their ancestors but the local one lacks: marked without building the lists
        import ctypes as C

        from .._cabi import bt_list_args, bt_tree_view, dtype_code
        tv = bt_tree_view()
        tv.dim, tv.nboxes, tv.aligned_nboxes = int(dtree.dimensions), nb, int(dtree.aligned_nboxes)
        tv.nlevels, tv.root_extent = int(dtree.nlevels), float(dtree.root_extent)
        tv.box_centers, tv.box_levels = dptr(dtree.box_centers), dptr(dtree.box_levels)
        tv.box_child_ids, tv.box_flags = dptr(dtree.box_child_ids), dptr(corner_flags)
        tv.box_parent_ids = dptr(dtree.box_parent_ids)
        tv.well_sep_is_n_away = int(traversal_builder.well_sep_is_n_away)
        tv.box_child_ids_t = dptr(shared["child_t"])
        la = bt_list_args()
        la.coll_starts, la.coll_lists = dptr(shared["coll"][0]), dptr(shared["coll"][1])
        la.stick_out_factor = float(dtree.stick_out_factor)
        la.with_extent = int(bool(dtree.sources_have_extent or dtree.targets_have_extent))
        check(lib.bt_trav_mark_rows(dtype_code(dtree.coord_dtype), C.byref(tv), C.byref(la),
                                    dptr(masks.point_src_boxes), dptr(masks.multipole_src_boxes),
                                    sh), "bt_trav_mark_rows")
        del shared
        mark("ds:masks")

        # }}}

        # every rank's masks as one bit field per box: who needs which box's sources (1) /
        # targets (2) / multipoles (4)
        mine = (masks.point_src_boxes | (masks.responsible_boxes << 1)
                | (masks.multipole_src_boxes << 2))
        allm = comm.allgather_tensor(mine)                                   # [size, nboxes]
        mark("ds:allgather masks")
        src, tgt = exchange_particle_kinds(
            actx, comm, dtree, allm,
            [(1, masks.point_src_boxes, "source", None),
             (2, masks.responsible_boxes, "target", tgt_ranges)], pre)
        mark("ds:exchange")
        local_tree = assemble_local_tree(actx, dtree, src, tgt, masks, allm, responsible, bitsel=4)
        local_travs = [dataclasses.replace(t, tree=local_tree) for t in local_travs]
        if merge_close_lists and local_tree.targets_have_extent:
            local_travs = [t.merge_close_lists(actx) for t in local_travs]
        mark("ds:local tree")
    local_trav = local_travs[0] if len(local_travs) == 1 else local_travs
    return local_tree, local_trav, src[5], tgt[5]
