"""``FMMTraversalBuilder`` / ``FMMTraversalInfo``: drop-in for ``boxtree.traversal``.

Same constructor, call signature, error behaviour and output record as the
reference (``boxtree/traversal.py:1353-1705`` FMMTraversalInfo, ``:1721-2345``
builder); lists are built by sm_100a kernels (``csrc/traversal.cu``) that
follow the reference's count -> scan -> write protocol with rows in the append
order of the reference's tree walks, so every CSR array is identical.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, replace
from typing import Any

import numpy as np
import torch

from . import _cabi
from ._cabi import (HCTL_NHEAVY, HCTL_NWALK, HCTL_OVERFLOW, HCTL_SIZE, bt_heavy_ws, bt_list3_args,
                    bt_list_args, bt_tree_view, check, dptr)
from .array_context import TorchArrayContext, make_obj_array
from .tree import Tree, TreeOfBoxes

CRIT_CODE = {"static_linf": 0, "precise_linf": 1, "static_l2": 2}

#: child visits one thread may spend on one row of list 1 / list 3 before the row is handed
#: to the grid-wide "heavy row" path (csrc/traversal.cu); BT_WALK_BUDGET overrides (tests).
#: Measured on config 3 (B200, profiles/r02_budget_sweep.txt): the breadth-first expansion of the
#: heavy path handles mid-size rows (a few hundred entries) faster than the row walk -- budget
#: 2048 / 1024 / 512 / 256 / 128: 15.5 / 14.3 / 13.2 / 13.1 / 14.1 ms per step (below 256 the
#: per-row overhead of the maps takes over)
DEFAULT_WALK_BUDGET = 256

#: entries per row the count pass of the fused list-1+3 walk may stage for the fill pass
#: (rows with more are walked a second time); BT_STAGE_STRIDE overrides, 0 disables.
#: A row within the budget makes at most budget / 2^d + 1 expansions of 2^d children, plus its
#: <= 27 roots and <= 24 near-field boxes from above: 320 holds every such row in 3-D
DEFAULT_STAGE_STRIDE = 320
_STAGE_MAX_BYTES = 8 << 30

# bits of bt_set_walk_mode (include/boxtree_b200.h) that select a different host sequence
WALK_MODE_COLL_TOPDOWN = 64
WALK_MODE_FUSED13 = 128
WALK_MODE_HEAVY_SORT = 512


class _HeavyWorkspace:
    """Device buffers of the heavy-row path for one list (lists 1 and 3)."""

    def __init__(self, actx, nrows, nboxes, dfs_rank, budget, row_mask=None):
        self.actx = actx
        self.row_mask = row_mask
        self.nrows = nrows
        self.row_heavy = actx.empty(max(nrows, 1), np.uint8)
        self.heavy_rows = actx.empty(max(nrows, 1), np.int32)
        self.hctl = actx.zeros(HCTL_SIZE, np.int32)
        self.heavy_total = actx.zeros(1, np.int64)
        self.dfs_rank = dfs_rank
        self.budget = budget
        self.frontier = None
        self.ekeys = self.evals = None
        self.ecap = 0
        self.stage = self.stage_count = None
        self.stage_cap = 0
        self.dfs_order = None
        self.keys_only = False
        # heavy rows by position map (fused walk)
        self.subtree_size = self.hrow_base = self.hplan = None
        self.seg_stride = 0
        self.hseg_rank = self.hseg_prefix = self.hseg_kind = self.hseg_n = None
        self.hmap = self.chunk_cnt = self.hctx = None
        self.hmap_cap = 0
        self._alloc_frontier(max(nrows, 8 * nboxes, 1 << 16))

    def enable_map(self, subtree_size, seg_stride):
        """Heavy rows of the fused walk by position map instead of a sort."""
        self.subtree_size = subtree_size
        self.seg_stride = int(seg_stride)
        self.hrow_base = self.actx.empty(self.nrows + 1, np.int64)
        self.hplan = self.actx.zeros(2, np.int64)

    def alloc_map(self, nheavy, map_bytes, nslots):
        n = max(nheavy, 1) * self.seg_stride
        self.hseg_rank = self.actx.empty(n, np.int32)
        self.hseg_prefix = self.actx.empty(n, np.int32)
        self.hseg_kind = self.actx.empty(n, np.uint8)
        self.hseg_n = self.actx.empty(max(nheavy, 1), np.int32)
        self.hctx = self.actx.empty(max(nheavy, 1) * 16, np.float64)     # 128 bytes per row
        self.hmap_cap = int(map_bytes)
        self.hmap = self.actx.empty(max(self.hmap_cap, 1), np.uint8)
        self.chunk_cnt = self.actx.empty(max(self.hmap_cap // 1024, 1) * nslots, np.int32)

    def enable_staging(self, stride):
        """Fused list-1+3 walk: room for *stride* staged entries per row (count pass)."""
        if stride > 0 and self.nrows > 0:
            self.stage_cap = int(stride)
            self.stage = self.actx.empty(self.nrows * self.stage_cap, np.int32)
            self.stage_count = self.actx.empty(self.nrows, np.int32)

    def _alloc_frontier(self, cap):
        self.frontier_cap = cap
        self.frontier = [self.actx.empty(cap, np.int64), self.actx.empty(cap, np.int64)]

    def grow_frontier(self):
        self._alloc_frontier(4 * self.frontier_cap)

    def alloc_entries(self, n):
        self.ecap = n
        if n > 0:
            self.ekeys = [self.actx.empty(n, np.int64), self.actx.empty(n, np.int64)]
            if not self.keys_only:
                self.evals = [self.actx.empty(n, np.int32), self.actx.empty(n, np.int32)]

    def struct(self) -> bt_heavy_ws:
        w = bt_heavy_ws()
        w.walk_budget = self.budget
        w.row_heavy = dptr(self.row_heavy)
        w.heavy_rows = dptr(self.heavy_rows)
        w.hctl = dptr(self.hctl)
        w.heavy_total = dptr(self.heavy_total)
        w.frontier[0] = dptr(self.frontier[0])
        w.frontier[1] = dptr(self.frontier[1])
        w.frontier_cap = self.frontier_cap
        w.dfs_rank = dptr(self.dfs_rank)
        if self.ekeys is not None:
            w.ekeys[0], w.ekeys[1] = dptr(self.ekeys[0]), dptr(self.ekeys[1])
        if self.evals is not None:
            w.evals[0], w.evals[1] = dptr(self.evals[0]), dptr(self.evals[1])
        w.dfs_order = dptr(self.dfs_order)
        w.subtree_size = dptr(self.subtree_size)
        w.hrow_base = dptr(self.hrow_base)
        w.hplan = dptr(self.hplan)
        w.seg_stride = self.seg_stride
        w.hseg_rank = dptr(self.hseg_rank)
        w.hseg_prefix = dptr(self.hseg_prefix)
        w.hseg_kind = dptr(self.hseg_kind)
        w.hseg_n = dptr(self.hseg_n)
        w.hmap = dptr(self.hmap)
        w.hmap_cap = self.hmap_cap
        w.chunk_cnt = dptr(self.chunk_cnt)
        w.hctx = dptr(self.hctx)
        w.ecap = self.ecap
        w.row_mask = dptr(self.row_mask)
        w.stage = dptr(self.stage)
        w.stage_cap = self.stage_cap
        w.stage_count = dptr(self.stage_count)
        return w


_INT32_MAX = 2**31 - 1


@dataclass(frozen=True)
class BuiltList:
    """Mirror of ``pyopencl.algorithm.BuiltList`` as used by the reference
    (``boxtree/array_context.py:222-238``, ``traversal.py:1523-1538``)."""
    count: int | None
    starts: Any
    lists: Any
    num_nonempty_lists: int | None = None
    nonempty_indices: Any = None
    compressed_indices: Any = None


@dataclass(frozen=True)
class FMMTraversalInfo:
    """Interaction lists for an FMM (``boxtree/traversal.py:1353-1705``)."""
    tree: Any
    well_sep_is_n_away: int

    source_boxes: Any
    target_boxes: Any
    level_start_source_box_nrs: Any
    level_start_target_box_nrs: Any
    source_parent_boxes: Any
    level_start_source_parent_box_nrs: Any
    target_or_target_parent_boxes: Any
    level_start_target_or_target_parent_box_nrs: Any

    same_level_non_well_sep_boxes_starts: Any
    same_level_non_well_sep_boxes_lists: Any

    neighbor_source_boxes_starts: Any
    neighbor_source_boxes_lists: Any

    from_sep_siblings_starts: Any
    from_sep_siblings_lists: Any

    from_sep_smaller_by_level: Any
    target_boxes_sep_smaller_by_source_level: Any
    from_sep_close_smaller_starts: Any
    from_sep_close_smaller_lists: Any

    from_sep_bigger_starts: Any
    from_sep_bigger_lists: Any
    from_sep_close_bigger_starts: Any
    from_sep_close_bigger_lists: Any

    @property
    def nboxes(self):
        return self.tree.nboxes

    @property
    def nlevels(self):
        return self.tree.nlevels

    @property
    def ntarget_boxes(self):
        return int(self.target_boxes.shape[0])

    @property
    def ntarget_or_target_parent_boxes(self):
        return int(self.target_or_target_parent_boxes.shape[0])

    def merge_close_lists(self, actx: TorchArrayContext, debug: bool = False):
        """``traversal.py:1650-1693``: fold both "close" lists into list 1."""
        starts, lists = _merge_lists(
            actx, _cabi.load(), None, self.ntarget_boxes,
            [self.neighbor_source_boxes_starts, self.from_sep_close_smaller_starts,
             self.from_sep_close_bigger_starts],
            [self.neighbor_source_boxes_lists, self.from_sep_close_smaller_lists,
             self.from_sep_close_bigger_lists])
        return replace(self, neighbor_source_boxes_starts=starts,
                       neighbor_source_boxes_lists=lists,
                       from_sep_close_smaller_starts=None, from_sep_close_smaller_lists=None,
                       from_sep_close_bigger_starts=None, from_sep_close_bigger_lists=None)

    def get_box_list(self, what, index):
        starts = getattr(self, f"{what}_starts")
        lists = getattr(self, f"{what}_lists")
        start, stop = (int(x) for x in starts[index:index + 2])
        return lists[start:stop]


def _read_i64(actx, dev: torch.Tensor) -> np.ndarray:
    return _read_finish(*_read_start(actx, dev))


def _read_start(actx, dev: torch.Tensor):
    """Enqueue the device -> pinned-host copy of *dev* and an event behind it.  Work launched
    between this and :func:`_read_finish` keeps the GPU busy while the host waits for the
    numbers, digests them and prepares the launches that depend on them."""
    host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
    host.copy_(dev, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(actx.stream)
    return host, ev


def _read_finish(host, ev) -> np.ndarray:
    ev.synchronize()
    return host.numpy().copy()


def _check_int32(total: int, what: str) -> None:
    if total > _INT32_MAX:
        raise OverflowError(
            f"{what} has {total} entries, which exceeds the int32 CSR index range "
            "(the reference's ListOfListsBuilder uses int32 starts as well)")


def _merge_lists(actx, lib, output_to_input_box, noutput, starts, lists):
    """_ListMerger, ``traversal.py:1298-1344``."""
    sh = actx.stream_handle
    with torch.cuda.stream(actx.stream):
        nl = len(starts)
        sp = _cabi.ptr_array(starts)
        lp = _cabi.ptr_array(lists)
        new_starts = actx.empty(noutput + 1, np.int32)
        totals = actx.zeros(2, np.int64)
        check(lib.bt_trav_merge_lists(0, noutput, dptr(output_to_input_box), nl, sp, lp,
                                      dptr(new_starts), None, dptr(totals), sh),
              "bt_trav_merge_lists")
        total = int(_read_i64(actx, totals)[0])
        _check_int32(total, "merged list")
        new_lists = actx.empty(total, np.int32)
        check(lib.bt_trav_merge_lists(1, noutput, dptr(output_to_input_box), nl, sp, lp,
                                      dptr(new_starts), dptr(new_lists), dptr(totals), sh),
              "bt_trav_merge_lists")
    return new_starts, new_lists


class FMMTraversalBuilder:
    """Mirrors ``boxtree.traversal.FMMTraversalBuilder`` (``traversal.py:1721``)."""

    def __init__(self, array_context: TorchArrayContext, *, well_sep_is_n_away: int = 1,
                 from_sep_smaller_crit: str | None = None) -> None:
        assert isinstance(array_context, TorchArrayContext)
        self._setup_actx = array_context
        self.well_sep_is_n_away = well_sep_is_n_away
        self.from_sep_smaller_crit = from_sep_smaller_crit
        self._lib = _cabi.load()
        self.last_stats: dict = {}
        self.last_shared: dict | None = None

    def _resolve_crit(self, extent_norm, sources_have_extent, targets_have_extent) -> str:
        # traversal.py:1776-1805
        crit = self.from_sep_smaller_crit
        if crit is None:
            crit = "precise_linf"
        if extent_norm == "linf":
            pass
        elif extent_norm == "l2":
            if crit == "static_linf":
                raise ValueError("the static l^inf from-sep-smaller criterion "
                                 "cannot be used with the l^2 extent norm")
        elif extent_norm is None:
            assert not (sources_have_extent or targets_have_extent)
        else:
            raise ValueError(f"unexpected value of 'extent_norm': {extent_norm}")
        if crit not in CRIT_CODE:
            raise ValueError(f"unexpected value of 'from_sep_smaller_crit': {crit}")
        return crit

    def build_in_chunks(self, actx: TorchArrayContext, tree: Tree, nchunks: int | None = None,
                        cost_per_box=None, **kwargs):
        """Large-list mode: the traversal of *tree* cut by ROWS into *nchunks* pieces, one
        :class:`FMMTraversalInfo` per piece.

        The reference's ``ListOfListsBuilder`` (``traversal.py:1854, 1950``) indexes every list
        with int32 ``starts``, so a tree whose ``from_sep_siblings`` (or any other) list has more
        than 2**31 - 1 entries cannot be traversed at all (BASELINE config 4: 1e8 Plummer points,
        2.2e9 list-2 entries).  Here the boxes are cut into *nchunks* contiguous segments of the
        depth-first order of ``boxtree/distributed/partition.py:38-121`` (equal cost; default cost
        ``1 + own sources + own targets``); piece *k* is the traversal of the tree whose target
        flags are kept only on the boxes of segment *k*: its rows are exactly the rows of the
        global traversal that belong to those boxes, every row of the global traversal is in
        exactly one piece, and every piece has int32 CSR arrays; the box lists of a piece
        (``source_boxes``, ``target_boxes``, ...) are the segment's boxes of each kind.  With
        ``nchunks=None`` the number of pieces is doubled until every list fits.

        :returns: a list of ``(boxes, trav)``: the boxes of the segment (device int32, DFS
            order) and the traversal restricted to their rows.  ``trav.same_level_non_well_sep_boxes``
            is the full list in every piece (built once, shared)."""
        import dataclasses

        from .distributed.partition import (_responsible_and_ancestors, get_box_ids_dfs_order,
                                            partition_segments, partition_segments_device)
        lib = self._lib
        nb = int(tree.nboxes)
        if cost_per_box is None:
            cost_per_box = (1.0 + tree.box_source_counts_nonchild.double()
                            + tree.box_target_counts_nonchild.double())
        dfs_order = get_box_ids_dfs_order(actx, tree)
        tries = [nchunks] if nchunks else [1, 2, 4, 8, 16, 32, 64]
        for n in tries:
            try:
                if n == 1:
                    trav, _ = self(actx, tree, **kwargs)
                    return [(dfs_order, trav)]
                with torch.cuda.stream(actx.stream):
                    segs = partition_segments_device(actx, cost_per_box, dfs_order, n)
                if segs is None:
                    cost_host = cost_per_box.cpu().numpy() if isinstance(
                        cost_per_box, torch.Tensor) else np.asarray(cost_per_box)
                    segs = partition_segments(cost_host[dfs_order.cpu().numpy()], n,
                                              total_workload=np.sum(cost_host))
                out = []
                self.last_shared = None
                for k in range(n):
                    boxes = dfs_order[int(segs[k][0]):int(segs[k][1])]
                    with torch.cuda.stream(actx.stream):
                        mine, anc = _responsible_and_ancestors(actx, lib, tree, boxes)
                        flags = actx.empty(nb, np.uint8)
                        need = actx.empty(nb, np.int8)
                        check(lib.bt_dist_restrict_target_flags(
                            nb, dptr(tree.box_flags), dptr(mine), dptr(actx.zeros(nb, np.int8)),
                            dptr(flags), dptr(need), actx.stream_handle),
                            "bt_dist_restrict_target_flags")
                    # box geometry, source flags and depth-first ranks are the same for every
                    # piece: colleagues and list-2 masks of ALL boxes are built once
                    prev = self.last_shared
                    piece, _ = self(actx, dataclasses.replace(tree, box_flags=flags),
                                    source_boxes_mask=mine, source_parent_boxes_mask=mine,
                                    _keep_shared=prev is None, _shared=prev, **kwargs)
                    out.append((boxes, dataclasses.replace(piece, tree=tree)))
                    del piece, flags, need
                self.last_shared = None
                return out
            except OverflowError:
                self.last_shared = None
                if nchunks:
                    raise
                torch.cuda.empty_cache()
        raise OverflowError("a single row piece still exceeds the int32 CSR index range")

    def __call__(self, actx: TorchArrayContext, tree: Tree | TreeOfBoxes, wait_for=None,
                 debug: bool = False, _from_sep_smaller_min_nsources_cumul: int | None = None,
                 source_boxes_mask=None, source_parent_boxes_mask=None,
                 _colleague_row_mask=None, _list13_row_mask=None, _keep_shared=False,
                 _shared=None, _before_extents=None, _preorder=None):
        """See ``boxtree/traversal.py:1969-1990``.

        :arg _preorder: (internal) ``(subtree_size, dfs_rank)`` of the tree's boxes when the caller
            has them already (``bt_trav_dfs_rank``).
        :arg _before_extents: (internal, deferred distributed build) called once, after the
            colleagues and the count passes of lists 2 and 4 have been launched and before the
            first kernel that reads ``box_target_bounding_box_{min,max}`` (trees whose targets
            have extent; otherwise the traversal never reads them and the hook is not called).

        :arg _colleague_row_mask: (internal, used by the sharded distributed setup) int8
            ``[nboxes]``; same-level non-well-separated boxes are only computed for boxes
            with a non-zero entry -- the caller guarantees that no other row is read.
        :arg _list13_row_mask: (internal) likewise for the rows of lists 1 and 3 (and list 3
            close): rows of target boxes with a zero entry are left empty.
        :arg _keep_shared: (internal) keep the flag-independent intermediates of this call
            (depth-first ranks, transposed child table, colleagues, list-2 counts and masks) in
            ``self.last_shared``.
        :arg _shared: (internal) such a dict from a call on the SAME box geometry, source flags
            and ``_colleague_row_mask`` (the sharded setup's second traversal): reused as is.

        :returns: ``(trav, event)``; *event* is a :class:`torch.cuda.Event`.
        """
        assert isinstance(actx, TorchArrayContext)
        lib = self._lib

        min_nsrc = _from_sep_smaller_min_nsources_cumul
        if min_nsrc is None:
            min_nsrc = 0

        if not tree._is_pruned:
            raise ValueError("tree must be pruned for traversal generation")
        if tree.sources_have_extent:
            raise NotImplementedError(
                "trees with source extent are not supported for traversal generation")

        crit = self._resolve_crit(tree.extent_norm, tree.sources_have_extent,
                                  tree.targets_have_extent)

        # a TreeOfBoxes (host numpy) is moved to the device first (traversal.py:1971, 2010)
        if isinstance(tree.box_flags, np.ndarray):
            dev_tree = actx.from_numpy(tree)
        else:
            dev_tree = tree

        nlevels = int(tree.nlevels)
        nboxes = int(tree.nboxes)
        dimensions = int(tree.dimensions)
        sources_are_targets = getattr(tree, "sources_are_targets", True)
        coord_dtype = np.dtype(tree.coord_dtype)
        dcode = _cabi.dtype_code(coord_dtype)
        stream = actx.stream
        sh = actx.stream_handle
        with_extent = bool(tree.sources_have_extent or tree.targets_have_extent)

        def dev(a, dt=None):
            if isinstance(a, np.ndarray):
                a = actx.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt)))
            return a.contiguous()

        for dep in (wait_for or ()):          # traversal.py:1969-1990
            if isinstance(dep, torch.cuda.Stream):
                stream.wait_stream(dep)
            else:
                stream.wait_event(dep)

        with torch.cuda.stream(stream), torch.cuda.device(actx.device):
            box_flags = dev(dev_tree.box_flags)
            box_levels = dev(dev_tree.box_levels)
            box_parent_ids = dev(dev_tree.box_parent_ids)
            box_centers = dev(dev_tree.box_centers)
            box_child_ids = dev(dev_tree.box_child_ids)
            # a TreeOfBoxes made by boxtree/tree_of_boxes.py carries int32 levels, -1 as the
            # parent of the root (its kernels then index starts[-1]: here the root is its own
            # parent, as in a Tree), IS_LEAF_BOX flags and no level starts
            if box_levels.dtype != torch.uint8:
                box_levels = box_levels.to(torch.uint8)
            if box_flags.dtype != torch.uint8:
                box_flags = box_flags.to(torch.uint8)
            if box_parent_ids.dtype != torch.int32:
                box_parent_ids = box_parent_ids.to(torch.int32)
            if box_child_ids.dtype != torch.int32:
                box_child_ids = box_child_ids.to(torch.int32)
            if isinstance(tree, TreeOfBoxes):
                box_parent_ids = torch.where(
                    box_parent_ids < 0,
                    torch.arange(nboxes, dtype=torch.int32, device=box_parent_ids.device),
                    box_parent_ids)
            if dev_tree.level_start_box_nrs is None:
                level_start_box_nrs = torch.searchsorted(
                    box_levels.to(torch.int32).contiguous(),
                    torch.arange(nlevels + 1, dtype=torch.int32, device=box_levels.device)
                ).to(torch.int32)
            else:
                level_start_box_nrs = dev(dev_tree.level_start_box_nrs, np.int32)
                if level_start_box_nrs.dtype != torch.int32:
                    level_start_box_nrs = level_start_box_nrs.to(torch.int32)

            tv = bt_tree_view()
            tv.dim = dimensions
            tv.nboxes = nboxes
            tv.aligned_nboxes = int(box_child_ids.shape[-1])
            tv.nlevels = nlevels
            tv.root_extent = float(tree.root_extent)
            tv.box_centers = dptr(box_centers)
            tv.box_levels = dptr(box_levels)
            tv.box_child_ids = dptr(box_child_ids)
            tv.box_flags = dptr(box_flags)
            tv.box_parent_ids = dptr(box_parent_ids)
            tv.well_sep_is_n_away = int(self.well_sep_is_n_away)
            if int(box_centers.shape[-1]) != tv.aligned_nboxes:
                raise ValueError("box_centers and box_child_ids must share their padded length")
            # scratch copy of the child table with the children of a box side by side
            if _shared is not None:
                child_t = _shared["child_t"]
            else:
                child_t = actx.empty(max(tv.aligned_nboxes, 1) * 2 ** dimensions, np.int32)
                check(lib.bt_trav_transpose_children(dimensions, tv.aligned_nboxes,
                                                     dptr(box_child_ids), dptr(child_t), sh),
                      "bt_trav_transpose_children")
            tv.box_child_ids_t = dptr(child_t)

            # {{{ b1/b2: box lists and their level starts (traversal.py:2054-2124)

            sbm = None if source_boxes_mask is None else dev(source_boxes_mask, np.int8)
            spbm = None if source_parent_boxes_mask is None else \
                dev(source_parent_boxes_mask, np.int8)
            nwhich = 3 if sources_are_targets else 4
            raw = [actx.empty(max(nboxes, 1), np.int32) for _ in range(nwhich)]
            counts_dev = actx.zeros(4, np.int32)
            masks = [spbm, sbm, None, None]
            for which in range(nwhich):
                check(lib.bt_trav_box_list(which, nboxes, dptr(box_flags), dptr(masks[which]),
                                           dptr(raw[which]), dptr(counts_dev[which:which + 1]),
                                           sh), "bt_trav_box_list")
            # (their sizes are read back together with the colleague total below: one
            # synchronisation for both)

            # }}}

            totals = actx.zeros(8, np.int64)

            crm = None if _colleague_row_mask is None else dev(_colleague_row_mask, np.int8)

            def list_args(row_boxes, coll=None, row_mask=None):
                a = bt_list_args()
                a.row_mask = dptr(row_mask)
                a.row_boxes = dptr(row_boxes)
                a.coll_starts = dptr(coll[0]) if coll else None
                a.coll_lists = dptr(coll[1]) if coll else None
                a.stick_out_factor = float(tree.stick_out_factor)
                a.with_extent = int(with_extent)
                return a

            # pre-order rank of every box: orders colleagues and the entries of heavy rows
            budget = int(os.environ.get("BT_WALK_BUDGET", DEFAULT_WALK_BUDGET))
            if _shared is not None:
                subtree_size, dfs_rank = _shared["subtree_size"], _shared["dfs_rank"]
            elif _preorder is not None:
                subtree_size, dfs_rank = _preorder
            else:
                subtree_size = actx.empty(max(nboxes, 1), np.int32)
                dfs_rank = actx.empty(max(nboxes, 1), np.int32)
                check(lib.bt_trav_dfs_rank(dimensions, nboxes, tv.aligned_nboxes, nlevels,
                                           dptr(level_start_box_nrs), dptr(box_child_ids),
                                           dptr(subtree_size), dptr(dfs_rank), sh),
                      "bt_trav_dfs_rank")

            # {{{ b3: same-level non-well-separated boxes (traversal.py:2135-2141)

            walk_mode = int(lib.bt_get_walk_mode())
            topdown = bool(walk_mode & WALK_MODE_COLL_TOPDOWN)
            # the top-down builder keeps (2n+1)^d parent-level boxes per warp in shared memory
            # (n <= 2 in 3-D); wider neighbourhoods use the reference's per-row walks
            nparents_max = (2 * int(self.well_sep_is_n_away) + 1) ** dimensions
            if nparents_max > 128 or (nparents_max * 2 ** dimensions + 31) // 32 + 1 > 40:
                topdown = False
            coll_starts = actx.empty(nboxes + 1, np.int32)
            l2_count_by_box = xflags = None
            reuse_coll = _shared is not None and _shared.get("topdown") == topdown
            if reuse_coll:
                coll_starts, coll_lists = _shared["coll"]
                l2_count_by_box, xflags = _shared["l2_count_by_box"], _shared["xflags"]
                l2_masks, mask_words = _shared["l2_masks"], _shared["mask_words"]
            elif topdown:
                # level by level from the parent's colleagues; list-2 counts come for free
                stride = (2 * int(self.well_sep_is_n_away) + 1) ** dimensions - 1
                staging = actx.empty(max(nboxes, 1) * stride, np.int32)
                l2_count_by_box = actx.empty(max(nboxes, 1), np.int32)
                xflags = actx.empty(max(nboxes, 1), np.uint8)
                mask_words = ((stride + 1) * 2 ** dimensions + 31) // 32 + 1
                l2_masks = actx.empty(max(nboxes, 1) * mask_words, np.int32)

                def colleagues(phase, lists):
                    check(lib.bt_trav_colleagues(
                        dcode, phase, C.byref(tv), dptr(level_start_box_nrs), dptr(dfs_rank),
                        dptr(crm), stride, dptr(staging), dptr(coll_starts), dptr(lists),
                        dptr(l2_count_by_box), dptr(xflags), dptr(l2_masks), mask_words,
                        dptr(totals), sh), "bt_trav_colleagues")
            else:
                a_coll = list_args(None, row_mask=crm)

                def colleagues(phase, lists):
                    check(lib.bt_trav_build_list(dcode, 0, phase, C.byref(tv), C.byref(a_coll),
                                                 nboxes, dptr(coll_starts), dptr(lists), None,
                                                 None, dptr(totals), sh), "colleagues")
            if not reuse_coll:
                colleagues(0, None)
                both = _read_i64(actx, torch.cat([totals[:1], counts_dev.to(torch.int64)]))
                total, counts = int(both[0]), both[1:]
                _check_int32(total, "same_level_non_well_sep_boxes")
                coll_lists = actx.empty(total, np.int32)
                colleagues(1, coll_lists)
                if topdown:
                    del staging
            else:
                counts = counts_dev.cpu().numpy()
            coll = (coll_starts, coll_lists)

            # {{{ b1/b2 continued: the box lists' sizes and level starts

            source_parent_boxes = raw[0][:int(counts[0])]
            source_boxes = raw[1][:int(counts[1])]
            target_or_target_parent_boxes = raw[2][:int(counts[2])]
            target_boxes = source_boxes if sources_are_targets else raw[3][:int(counts[3])]
            ntb = int(target_boxes.shape[0])
            ntp = int(target_or_target_parent_boxes.shape[0])

            def level_starts(box_list):
                out = actx.empty(nlevels + 1, np.int32)
                check(lib.bt_trav_level_starts(nlevels, dptr(level_start_box_nrs),
                                               dptr(box_list), int(box_list.shape[0]),
                                               dptr(out), sh), "bt_trav_level_starts")
                return out

            lss = level_starts(source_boxes)
            lssp = level_starts(source_parent_boxes)
            lst = lss if sources_are_targets else level_starts(target_boxes)
            lstp = level_starts(target_or_target_parent_boxes)

            # }}}

            if _keep_shared:
                self.last_shared = {
                    "child_t": child_t, "subtree_size": subtree_size, "dfs_rank": dfs_rank,
                    "topdown": topdown, "coll": coll, "l2_count_by_box": l2_count_by_box,
                    "xflags": xflags, "l2_masks": l2_masks if topdown else None,
                    "mask_words": mask_words if topdown else 0}

            # }}}

            # {{{ count phases of lists 1-4, then ONE readback

            # lists 1 and 3 from one fused walk (needs the top-down colleagues' xflags)
            fused13 = topdown and bool(walk_mode & WALK_MODE_FUSED13)
            rm13 = None if _list13_row_mask is None else dev(_list13_row_mask, np.int8)
            ws1 = None if fused13 else _HeavyWorkspace(actx, ntb, nboxes, dfs_rank, budget, rm13)
            ws3 = _HeavyWorkspace(actx, ntb, nboxes, dfs_rank, budget, rm13)
            if fused13:
                # the heavy rows sort (slot, row, rank) keys alone; the box comes back from its rank
                dfs_order = actx.empty(max(nboxes, 1), np.int32)
                check(lib.bt_reverse_index(nboxes, dptr(dfs_rank), dptr(dfs_order), sh),
                      "bt_reverse_index")
                ws3.dfs_order = dfs_order
                ws3.keys_only = True
                nroots_max = (2 * int(self.well_sep_is_n_away) + 1) ** dimensions
                # positions inside one row's map are int32: rows can span nroots_max subtrees
                if not (walk_mode & WALK_MODE_HEAVY_SORT) and nroots_max * nboxes < 2 ** 31 - 2048:
                    ws3.enable_map(subtree_size, nroots_max + 136)
            if fused13 and nboxes < (1 << 27) and nlevels + 2 <= 31:
                stage_stride = int(os.environ.get("BT_STAGE_STRIDE", DEFAULT_STAGE_STRIDE))
                while stage_stride > 32 and ntb * stage_stride * 4 > _STAGE_MAX_BYTES:
                    stage_stride //= 2
                if ntb * stage_stride * 4 <= _STAGE_MAX_BYTES:
                    ws3.enable_staging(stage_stride)

            l1_starts = actx.empty(ntb + 1, np.int32)

            def count_list1():
                check(lib.bt_trav_list1(dcode, 0, C.byref(tv), dptr(target_boxes), ntb,
                                        dptr(l1_starts), None, dptr(totals[0:]),
                                        C.byref(ws1.struct()), 0, sh), "list 1 count")
            if not fused13:
                count_list1()

            l2_starts = actx.empty(ntp + 1, np.int32)
            a2 = list_args(target_or_target_parent_boxes, coll)
            if topdown:
                check(lib.bt_trav_list2_starts(ntp, dptr(target_or_target_parent_boxes),
                                               dptr(l2_count_by_box), dptr(l2_starts),
                                               dptr(totals[1:]), sh), "list 2 starts")
            else:
                check(lib.bt_trav_build_list(dcode, 2, 0, C.byref(tv), C.byref(a2), ntp,
                                             dptr(l2_starts), None, None, None, dptr(totals[1:]),
                                             sh), "list 2 count")

            l4_starts = actx.empty(ntp + 1, np.int32)
            l4c_starts_raw = actx.empty(ntp + 1, np.int32) if with_extent else None
            a4 = list_args(target_or_target_parent_boxes, coll)

            def count_list4():
                check(lib.bt_trav_build_list(dcode, 4, 0, C.byref(tv), C.byref(a4), ntp,
                                             dptr(l4_starts), None, dptr(l4c_starts_raw), None,
                                             dptr(totals[2:]), sh), "list 4 count")
            # With the heavy-row maps the host reads the map sizes in the middle of the list-1+3
            # count (count_list3): the list-4 count is launched right behind that copy, so the GPU
            # works on it while the host allocates the maps.  (Not with deferred extents: there
            # it belongs to the work that hides the reduction, before the first use of extents.)
            l4_in_shadow = (fused13 and ws3.hrow_base is not None
                            and not (_before_extents is not None and tree.targets_have_extent))
            if not l4_in_shadow:
                count_list4()

            a3 = bt_list3_args()
            a3.target_boxes = dptr(target_boxes)
            a3.coll_starts = dptr(coll_starts)
            a3.coll_lists = dptr(coll_lists)
            a3.stick_out_factor = float(tree.stick_out_factor)
            a3.targets_have_extent = int(tree.targets_have_extent)
            a3.sources_have_extent = int(tree.sources_have_extent)
            a3.crit = CRIT_CODE[crit]
            keep_alive = []
            if _before_extents is not None and tree.targets_have_extent:
                _before_extents()
            if tree.targets_have_extent:
                bbmin = dev(dev_tree.box_target_bounding_box_min)
                bbmax = dev(dev_tree.box_target_bounding_box_max)
                bsc = dev(dev_tree.box_source_counts_cumul)
                keep_alive += [bbmin, bbmax, bsc]
                a3.box_target_bounding_box_min = dptr(bbmin)
                a3.box_target_bounding_box_max = dptr(bbmax)
                a3.box_source_counts_cumul = dptr(bsc)
            a3.min_nsources_cumul = int(min_nsrc)
            rowlen = ntb + 1
            nslots = nlevels + (2 if fused13 else 1)     # source levels, close list[, list 1]
            G = actx.empty(nslots * rowlen + 1, np.int32)
            Cc = actx.empty(nslots * rowlen + 1, np.int32)
            summary = actx.zeros(2 * (nslots + 1) + 1, np.int64)

            early = {}

            def count_list3(in_shadow=None):
                if fused13:
                    check(lib.bt_trav_list13(dcode, 0, C.byref(tv), C.byref(a3), dptr(xflags), ntb,
                                             dptr(G), dptr(Cc), None, dptr(summary),
                                             C.byref(ws3.struct()), 0, 0, 0, sh), "list 1+3 count")
                    if ws3.hrow_base is not None:
                        # heavy rows by position map: size the maps, expand the rows once
                        # (the list-2 total rides along: its fill can then be launched early)
                        pending = _read_start(actx, torch.cat([
                            ws3.hctl[:1].to(torch.int64), ws3.hplan[:1], totals[1:2]]))
                        if in_shadow is not None:
                            in_shadow()
                        plan = _read_finish(*pending)
                        early["l2_total"] = int(plan[2])
                        ws3.alloc_map(int(plan[0]), int(plan[1]), nslots)
                        check(lib.bt_trav_list13(dcode, 2, C.byref(tv), C.byref(a3), dptr(xflags),
                                                 ntb, dptr(G), dptr(Cc), None, dptr(summary),
                                                 C.byref(ws3.struct()), 0, int(plan[0]), 0, sh),
                              "list 1+3 heavy expansion")
                else:
                    check(lib.bt_trav_list3(dcode, 0, C.byref(tv), C.byref(a3), ntb, dptr(G),
                                            dptr(Cc), None, dptr(summary), C.byref(ws3.struct()),
                                            0, sh), "list 3 count")
            count_list3(count_list4 if l4_in_shadow else None)

            zero1 = actx.zeros(1, np.int64)
            zero2 = actx.zeros(2, np.int64)

            def fill_list2(total):
                lists = actx.empty(total, np.int32)
                if topdown:
                    check(lib.bt_trav_list2_fill_masked(
                        dimensions, ntp, dptr(target_or_target_parent_boxes), dptr(box_parent_ids),
                        dptr(coll_starts), dptr(coll_lists), dptr(child_t), dptr(l2_masks),
                        mask_words, dptr(l2_starts), dptr(lists), sh), "list 2 fill")
                else:
                    check(lib.bt_trav_build_list(dcode, 2, 1, C.byref(tv), C.byref(a2), ntp,
                                                 dptr(l2_starts), dptr(lists), None, None, None,
                                                 sh), "list 2 fill")
                return lists

            l2_lists = None

            def read_counts():
                nonlocal l2_lists
                pending = _read_start(actx, torch.cat([
                    totals, summary, zero1 if fused13 else ws1.heavy_total, ws3.heavy_total,
                    zero2 if fused13 else ws1.hctl[:2].to(torch.int64),
                    ws3.hctl[:4].to(torch.int64)]))
                # the list-2 fill needs none of these numbers: launched behind the copy, it keeps
                # the GPU busy while the host digests the totals and sets up the other fills
                if l2_lists is None and "l2_total" in early:
                    _check_int32(early["l2_total"], "from_sep_siblings")
                    l2_lists = fill_list2(early["l2_total"])
                both = _read_finish(*pending)
                ns = summary.shape[0]
                return both[:8], both[8:8 + ns], both[8 + ns:]

            tot, summ, heavy = read_counts()
            # frontier overflow of the heavy-row BFS: enlarge and count again (rare)
            while heavy[2 + HCTL_OVERFLOW] or heavy[4 + HCTL_OVERFLOW]:
                if heavy[2 + HCTL_OVERFLOW]:
                    ws1.grow_frontier()
                    count_list1()
                if heavy[4 + HCTL_OVERFLOW]:
                    ws3.grow_frontier()
                    count_list3()
                tot, summ, heavy = read_counts()
            heavy1_total, heavy3_total = int(heavy[0]), int(heavy[1])
            self.last_stats = {"heavy_rows_list1": int(heavy[2 + HCTL_NHEAVY]),
                               "heavy_rows_list3": int(heavy[4 + HCTL_NHEAVY]),
                               "heavy_entries_list1": heavy1_total,
                               "heavy_entries_list3": heavy3_total,
                               "rewalked_rows_list13": int(heavy[4 + HCTL_NWALK]) if fused13 else 0,
                               "fused13": fused13}
            g0 = summ[:nslots + 1]               # G[l][0], l = 0..nslots-1, then grand total
            c0 = summ[nslots + 1:2 * (nslots + 1)]
            if fused13:
                # the flattened scan is int32: the 64-bit grand total tells whether it wrapped
                _check_int32(int(summ[2 * (nslots + 1)]), "from_sep_smaller + neighbor_source_boxes")
                l1_total = int(g0[nlevels + 2] - g0[nlevels + 1])
            else:
                l1_total = int(tot[0])
            _check_int32(l1_total, "neighbor_source_boxes")
            _check_int32(int(tot[1]), "from_sep_siblings")
            _check_int32(int(tot[2]), "from_sep_bigger")
            _check_int32(int(g0[nslots]), "from_sep_smaller")

            # }}}

            # {{{ fill phases

            if not fused13:
                l1_lists = actx.empty(l1_total, np.int32)
                ws1.alloc_entries(heavy1_total)
                check(lib.bt_trav_list1(dcode, 1, C.byref(tv), dptr(target_boxes), ntb,
                                        dptr(l1_starts), dptr(l1_lists), None,
                                        C.byref(ws1.struct()), heavy1_total, sh), "list 1 fill")
            del ws1
            if l2_lists is None:
                l2_lists = fill_list2(int(tot[1]))
            assert int(l2_lists.shape[0]) == int(tot[1])
            if topdown and not _keep_shared:
                del l2_masks
            l4_lists = actx.empty(int(tot[2]), np.int32)
            l4c_lists_raw = actx.empty(int(tot[3]), np.int32) if with_extent else None
            check(lib.bt_trav_build_list(dcode, 4, 1, C.byref(tv), C.byref(a4), ntp,
                                         dptr(l4_starts), dptr(l4_lists), dptr(l4c_starts_raw),
                                         dptr(l4c_lists_raw), None, sh), "list 4 fill")

            l3_all = actx.empty(int(g0[nslots]), np.int32)
            if ws3.hrow_base is None:          # sort-based heavy rows need key buffers
                ws3.alloc_entries(heavy3_total)
            if fused13:
                check(lib.bt_trav_list13(dcode, 1, C.byref(tv), C.byref(a3), dptr(xflags), ntb,
                                         dptr(G), dptr(Cc), dptr(l3_all), dptr(summary),
                                         C.byref(ws3.struct()), heavy3_total,
                                         int(heavy[4 + HCTL_NHEAVY]), int(heavy[4 + HCTL_NWALK]),
                                         sh), "list 1+3 fill")
                l1_lists = l3_all[int(g0[nlevels + 1]):int(g0[nlevels + 2])]
            else:
                check(lib.bt_trav_list3(dcode, 1, C.byref(tv), C.byref(a3), ntb, dptr(G), dptr(Cc),
                                        dptr(l3_all), dptr(summary), C.byref(ws3.struct()),
                                        heavy3_total, sh), "list 3 fill")
            del ws3
            nne_total = int(c0[nlevels])         # non-empty rows over all source levels
            cstarts = actx.empty(nne_total + nlevels, np.int32)
            nonempty_all = actx.empty(max(nne_total, 1), np.int32)
            tb_nonempty_all = actx.empty(max(nne_total, 1), np.int32)
            comp_idx = actx.empty((nlevels, rowlen), np.int32)
            close3_starts = actx.empty(rowlen, np.int32) if with_extent else None
            check(lib.bt_trav_list3_compress(nlevels, ntb, dptr(G), dptr(Cc), dptr(target_boxes),
                                             dptr(cstarts), dptr(nonempty_all),
                                             dptr(tb_nonempty_all), dptr(comp_idx),
                                             dptr(close3_starts),
                                             dptr(l1_starts) if fused13 else None, sh),
                  "list 3 compress")

            from_sep_smaller_by_level = []
            target_boxes_sep_smaller_by_source_level = []
            for lev in range(nlevels):
                nne = int(c0[lev + 1] - c0[lev])
                noff = int(c0[lev])
                soff = noff + lev
                cnt = int(g0[lev + 1] - g0[lev])
                from_sep_smaller_by_level.append(BuiltList(
                    count=cnt,
                    starts=cstarts[soff:soff + nne + 1],
                    lists=l3_all[int(g0[lev]):int(g0[lev + 1])],
                    num_nonempty_lists=nne,
                    nonempty_indices=nonempty_all[noff:noff + nne],
                    compressed_indices=comp_idx[lev]))
                target_boxes_sep_smaller_by_source_level.append(tb_nonempty_all[noff:noff + nne])
            if with_extent:
                close3_lists = l3_all[int(g0[nlevels]):int(g0[nlevels + 1])]
            else:
                close3_starts = close3_lists = None

            # list 4 close: re-index from target_or_target_parent_boxes to target_boxes
            # (traversal.py:2255-2287, 1293-1304)
            if with_extent:
                rev = actx.zeros(max(nboxes, 1), np.int32)
                check(lib.bt_reverse_index(ntp, dptr(target_or_target_parent_boxes), dptr(rev),
                                           sh), "bt_reverse_index")
                out_to_in = actx.empty(max(ntb, 1), np.int32)
                check(lib.bt_gather_i32(ntb, dptr(rev), dptr(target_boxes), dptr(out_to_in), sh),
                      "bt_gather_i32")
                close4_starts, close4_lists = _merge_lists(
                    actx, lib, out_to_in, ntb, [l4c_starts_raw], [l4c_lists_raw])
            else:
                close4_starts = close4_lists = None

            # }}}

            evt = torch.cuda.Event()
            evt.record(stream)

        info = FMMTraversalInfo(
            tree=tree, well_sep_is_n_away=self.well_sep_is_n_away,
            source_boxes=source_boxes, target_boxes=target_boxes,
            level_start_source_box_nrs=lss, level_start_target_box_nrs=lst,
            source_parent_boxes=source_parent_boxes,
            level_start_source_parent_box_nrs=lssp,
            target_or_target_parent_boxes=target_or_target_parent_boxes,
            level_start_target_or_target_parent_box_nrs=lstp,
            same_level_non_well_sep_boxes_starts=coll_starts,
            same_level_non_well_sep_boxes_lists=coll_lists,
            neighbor_source_boxes_starts=l1_starts, neighbor_source_boxes_lists=l1_lists,
            from_sep_siblings_starts=l2_starts, from_sep_siblings_lists=l2_lists,
            from_sep_smaller_by_level=make_obj_array(from_sep_smaller_by_level),
            target_boxes_sep_smaller_by_source_level=make_obj_array(
                target_boxes_sep_smaller_by_source_level),
            from_sep_close_smaller_starts=close3_starts,
            from_sep_close_smaller_lists=close3_lists,
            from_sep_bigger_starts=l4_starts, from_sep_bigger_lists=l4_lists,
            from_sep_close_bigger_starts=close4_starts,
            from_sep_close_bigger_lists=close4_lists)
        return actx.freeze(info), evt
