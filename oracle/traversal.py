"""CPU oracle: restatement of ``boxtree.traversal.FMMTraversalBuilder.__call__``.

TEST INFRASTRUCTURE ONLY -- never imported by ``boxtree_b200``.  PARITY PINNED
against the reference's own ``FMMTraversalBuilder`` executed on the CPU
(``tests/refexec``, ``tests/test_refexec.py``; see ``oracle_trav.c``).  Follows the host driver at
``/root/reference/boxtree/traversal.py:1969-2345`` and the list merger at
``:1222-1344`` on numpy arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from ._lib import cint, i64, lib_for, ptr

CRIT_CODE = {"static_linf": 0, "precise_linf": 1, "static_l2": 2}


@dataclass
class OracleBuiltList:
    """Mirror of ``pyopencl.algorithm.BuiltList``."""
    count: int
    starts: np.ndarray
    lists: np.ndarray
    num_nonempty_lists: Any = None
    nonempty_indices: Any = None
    compressed_indices: Any = None


@dataclass
class OracleTraversal:
    """Mirror of ``FMMTraversalInfo`` (traversal.py:1595-1630)."""
    tree: Any
    well_sep_is_n_away: int
    source_boxes: np.ndarray
    target_boxes: np.ndarray
    level_start_source_box_nrs: np.ndarray
    level_start_target_box_nrs: np.ndarray
    source_parent_boxes: np.ndarray
    level_start_source_parent_box_nrs: np.ndarray
    target_or_target_parent_boxes: np.ndarray
    level_start_target_or_target_parent_box_nrs: np.ndarray
    same_level_non_well_sep_boxes_starts: np.ndarray
    same_level_non_well_sep_boxes_lists: np.ndarray
    neighbor_source_boxes_starts: np.ndarray
    neighbor_source_boxes_lists: np.ndarray
    from_sep_siblings_starts: np.ndarray
    from_sep_siblings_lists: np.ndarray
    from_sep_smaller_by_level: list
    target_boxes_sep_smaller_by_source_level: list
    from_sep_close_smaller_starts: Any
    from_sep_close_smaller_lists: Any
    from_sep_bigger_starts: np.ndarray
    from_sep_bigger_lists: np.ndarray
    from_sep_close_bigger_starts: Any
    from_sep_close_bigger_lists: Any
    extra: dict = field(default_factory=dict)

    @property
    def nboxes(self):
        return self.tree.nboxes

    @property
    def nlevels(self):
        return self.tree.nlevels

    @property
    def ntarget_boxes(self):
        return len(self.target_boxes)

    @property
    def ntarget_or_target_parent_boxes(self):
        return len(self.target_or_target_parent_boxes)


def _coord_c(coord_dtype):
    return C.c_float if np.dtype(coord_dtype) == np.float32 else C.c_double


def _make_structs(coord_dtype):
    ct = _coord_c(coord_dtype)
    vp = C.c_void_p

    class TreeView(C.Structure):
        _fields_ = [("d", C.c_int), ("aligned_nboxes", C.c_int64), ("root_extent", ct),
                    ("box_centers", vp), ("box_levels", vp), ("box_child_ids", vp),
                    ("box_flags", vp), ("box_parent_ids", vp),
                    ("well_sep_is_n_away", C.c_int)]

    class L1(C.Structure):
        _fields_ = [("target_boxes", vp)]

    class L2(C.Structure):
        _fields_ = [("tp_boxes", vp), ("coll_starts", vp), ("coll_lists", vp)]

    class L3(C.Structure):
        _fields_ = [("stick_out_factor", ct), ("target_boxes", vp), ("coll_starts", vp),
                    ("coll_lists", vp), ("targets_have_extent", C.c_int),
                    ("sources_have_extent", C.c_int), ("crit", C.c_int),
                    ("bb_min", vp), ("bb_max", vp), ("box_source_counts_cumul", vp),
                    ("min_nsources_cumul", C.c_int32), ("source_level", C.c_int)]

    class L4(C.Structure):
        _fields_ = [("stick_out_factor", ct), ("tp_boxes", vp), ("coll_starts", vp),
                    ("coll_lists", vp), ("with_extent", C.c_int)]

    class TravArgs(C.Structure):
        _fields_ = [("tree", TreeView), ("l1", L1), ("l2", L2), ("l3", L3), ("l4", L4)]

    return TravArgs


def _p(a):
    return None if a is None else a.ctypes.data


class _ListBuilder:
    """count -> exclusive scan -> write, like pyopencl's ListOfListsBuilder."""

    def __init__(self, lib, args):
        self.lib = lib
        self.args = args

    def __call__(self, kind, nrows, want=(True, False), eliminate_empty=False):
        lib, A = self.lib, self.args
        counts = [np.zeros(nrows, np.int32) if w else None for w in want]
        lib.orc_build_lists(cint(kind), C.byref(A), i64(nrows), cint(0),
                            ptr(counts[0]), ptr(counts[1]), None, None, None, None)
        starts, lists = [None, None], [None, None]
        for k in range(2):
            if counts[k] is None:
                continue
            st = np.zeros(nrows + 1, np.int64)
            np.cumsum(counts[k], out=st[1:])
            assert st[-1] < 2**31, "list too long for int32 starts (reference limit)"
            starts[k] = st.astype(np.int32)
            lists[k] = np.zeros(int(st[-1]), np.int32)
        lib.orc_build_lists(cint(kind), C.byref(A), i64(nrows), cint(1), None, None,
                            ptr(starts[0]), ptr(starts[1]), ptr(lists[0]), ptr(lists[1]))
        out = []
        for k in range(2):
            if counts[k] is None:
                out.append(None)
                continue
            bl = OracleBuiltList(count=int(starts[k][-1]), starts=starts[k], lists=lists[k])
            if eliminate_empty and k == 0:
                nonempty = counts[k] != 0
                bl.nonempty_indices = np.nonzero(nonempty)[0].astype(np.int32)
                bl.num_nonempty_lists = int(len(bl.nonempty_indices))
                ci = np.zeros(nrows + 1, np.int32)
                np.cumsum(nonempty, out=ci[1:])
                bl.compressed_indices = ci
                bl.starts = np.concatenate(
                    [starts[k][:-1][nonempty], starts[k][-1:]]).astype(np.int32)
            out.append(bl)
        return out


def build_traversal(tree, well_sep_is_n_away=1, from_sep_smaller_crit=None,
                    _from_sep_smaller_min_nsources_cumul=None,
                    source_boxes_mask=None, source_parent_boxes_mask=None,
                    debug=False) -> OracleTraversal:
    min_nsrc = _from_sep_smaller_min_nsources_cumul
    if min_nsrc is None:
        min_nsrc = 0
    if not tree._is_pruned:
        raise ValueError("tree must be pruned for traversal generation")
    if tree.sources_have_extent:
        raise NotImplementedError("trees with source extent are not supported for "
                                  "traversal generation")

    # crit processing -- traversal.py:1776-1805
    crit = from_sep_smaller_crit
    if crit is None:
        crit = "precise_linf"
    if tree.extent_norm == "l2" and crit == "static_linf":
        raise ValueError("the static l^inf from-sep-smaller criterion "
                         "cannot be used with the l^2 extent norm")
    if tree.extent_norm not in ("linf", "l2", None):
        raise ValueError(f"unexpected value of 'extent_norm': {tree.extent_norm}")
    if crit not in CRIT_CODE:
        raise ValueError(f"unexpected value of 'from_sep_smaller_crit': {crit}")

    nlevels = tree.nlevels
    sources_are_targets = getattr(tree, "sources_are_targets", True)
    coord_dtype = np.dtype(tree.coord_dtype)
    lib = lib_for(coord_dtype)
    nboxes = tree.nboxes
    box_flags = np.ascontiguousarray(tree.box_flags)
    box_levels = np.ascontiguousarray(tree.box_levels)
    box_parent_ids = np.ascontiguousarray(tree.box_parent_ids)
    box_centers = np.ascontiguousarray(tree.box_centers)
    box_child_ids = np.ascontiguousarray(tree.box_child_ids)
    level_start_box_nrs = np.ascontiguousarray(tree.level_start_box_nrs)

    # {{{ b1 -- traversal.py:2054-2067
    cnt = (C.c_int64 * 4)()
    sbm = None if source_boxes_mask is None else np.ascontiguousarray(source_boxes_mask, np.int8)
    spbm = None if source_parent_boxes_mask is None else \
        np.ascontiguousarray(source_parent_boxes_mask, np.int8)
    lib.orc_sources_parents_and_targets(i64(nboxes), ptr(box_flags), cint(sources_are_targets),
                                        ptr(sbm), ptr(spbm), None, None, None, None, cnt)
    source_parent_boxes = np.zeros(cnt[0], np.int32)
    source_boxes = np.zeros(cnt[1], np.int32)
    tp_boxes = np.zeros(cnt[2], np.int32)
    target_boxes_sep = np.zeros(cnt[3], np.int32)
    lib.orc_sources_parents_and_targets(i64(nboxes), ptr(box_flags), cint(sources_are_targets),
                                        ptr(sbm), ptr(spbm), ptr(source_parent_boxes),
                                        ptr(source_boxes), ptr(tp_boxes),
                                        ptr(target_boxes_sep), cnt)
    target_boxes = source_boxes if sources_are_targets else target_boxes_sep
    # }}}

    # {{{ b2 -- traversal.py:2073-2124
    def extract_level_start_box_nrs(box_list):
        result = np.full(nlevels + 1, len(box_list), np.int32)
        lib.orc_extract_level_start_box_nrs(i64(len(box_list)), ptr(level_start_box_nrs),
                                            ptr(box_levels), ptr(box_list), ptr(result))
        prev_start = len(box_list)
        for ilev in range(nlevels - 1, -1, -1):
            result[ilev] = prev_start = min(result[ilev], prev_start)
        return result

    lss = extract_level_start_box_nrs(source_boxes)
    lssp = extract_level_start_box_nrs(source_parent_boxes)
    lst = extract_level_start_box_nrs(target_boxes)
    lstp = extract_level_start_box_nrs(tp_boxes)
    # }}}

    TravArgs = _make_structs(coord_dtype)
    A = TravArgs()
    ct = _coord_c(coord_dtype)
    A.tree.d = tree.dimensions
    A.tree.aligned_nboxes = tree.aligned_nboxes
    A.tree.root_extent = float(tree.root_extent)
    A.tree.box_centers = _p(box_centers)
    A.tree.box_levels = _p(box_levels)
    A.tree.box_child_ids = _p(box_child_ids)
    A.tree.box_flags = _p(box_flags)
    A.tree.box_parent_ids = _p(box_parent_ids)
    A.tree.well_sep_is_n_away = well_sep_is_n_away
    builder = _ListBuilder(lib, A)

    # b3 colleagues -- traversal.py:2135-2141
    coll, _ = builder(0, nboxes)

    # b4 list 1 -- traversal.py:2149-2156
    A.l1.target_boxes = _p(target_boxes)
    list1, _ = builder(1, len(target_boxes))

    # b5 list 2 -- traversal.py:2164-2173
    A.l2.tp_boxes = _p(tp_boxes)
    A.l2.coll_starts = _p(coll.starts)
    A.l2.coll_lists = _p(coll.lists)
    list2, _ = builder(2, len(tp_boxes))

    with_extent = tree.sources_have_extent or tree.targets_have_extent

    # b6 list 3 -- traversal.py:2183-2231
    A.l3.stick_out_factor = float(ct(float(tree.stick_out_factor)).value)
    A.l3.target_boxes = _p(target_boxes)
    A.l3.coll_starts = _p(coll.starts)
    A.l3.coll_lists = _p(coll.lists)
    A.l3.targets_have_extent = int(tree.targets_have_extent)
    A.l3.sources_have_extent = int(tree.sources_have_extent)
    A.l3.crit = CRIT_CODE[crit]
    keep = []
    if tree.targets_have_extent:
        bbmin = np.ascontiguousarray(tree.box_target_bounding_box_min)
        bbmax = np.ascontiguousarray(tree.box_target_bounding_box_max)
        bsc = np.ascontiguousarray(tree.box_source_counts_cumul)
        keep += [bbmin, bbmax, bsc]
        A.l3.bb_min, A.l3.bb_max, A.l3.box_source_counts_cumul = _p(bbmin), _p(bbmax), _p(bsc)
    A.l3.min_nsources_cumul = int(min_nsrc)
    from_sep_smaller_by_level = []
    target_boxes_sep_smaller_by_source_level = []
    for ilevel in range(nlevels):
        A.l3.source_level = ilevel
        res, _ = builder(3, len(target_boxes), want=(True, False), eliminate_empty=True)
        target_boxes_sep_smaller_by_source_level.append(target_boxes[res.nonempty_indices])
        from_sep_smaller_by_level.append(res)
    if with_extent:
        A.l3.source_level = -1
        _, close3 = builder(3, len(target_boxes), want=(False, True))
        close3_starts, close3_lists = close3.starts, close3.lists
    else:
        close3_starts = close3_lists = None

    # b7 list 4 -- traversal.py:2242-2290
    A.l4.stick_out_factor = float(ct(float(tree.stick_out_factor)).value)
    A.l4.tp_boxes = _p(tp_boxes)
    A.l4.coll_starts = _p(coll.starts)
    A.l4.coll_lists = _p(coll.lists)
    A.l4.with_extent = int(with_extent)
    list4, close4_raw = builder(4, len(tp_boxes), want=(True, with_extent))
    if with_extent:
        # _ListMerger, TARGET_OR_TARGET_PARENT_BOXES -> TARGET_BOXES, traversal.py:1293-1344
        rev = np.zeros(nboxes, np.int32)            # tools.reverse_index_array
        rev[tp_boxes] = np.arange(len(tp_boxes), dtype=np.int32)
        out_to_in = np.ascontiguousarray(rev[target_boxes])
        close4_starts, close4_lists = merge_lists(lib, out_to_in, [close4_raw.starts],
                                                  [close4_raw.lists])
    else:
        close4_starts = close4_lists = None

    return OracleTraversal(
        tree=tree, well_sep_is_n_away=well_sep_is_n_away,
        source_boxes=source_boxes, target_boxes=target_boxes,
        level_start_source_box_nrs=lss, level_start_target_box_nrs=lst,
        source_parent_boxes=source_parent_boxes,
        level_start_source_parent_box_nrs=lssp,
        target_or_target_parent_boxes=tp_boxes,
        level_start_target_or_target_parent_box_nrs=lstp,
        same_level_non_well_sep_boxes_starts=coll.starts,
        same_level_non_well_sep_boxes_lists=coll.lists,
        neighbor_source_boxes_starts=list1.starts, neighbor_source_boxes_lists=list1.lists,
        from_sep_siblings_starts=list2.starts, from_sep_siblings_lists=list2.lists,
        from_sep_smaller_by_level=from_sep_smaller_by_level,
        target_boxes_sep_smaller_by_source_level=target_boxes_sep_smaller_by_source_level,
        from_sep_close_smaller_starts=close3_starts, from_sep_close_smaller_lists=close3_lists,
        from_sep_bigger_starts=list4.starts, from_sep_bigger_lists=list4.lists,
        from_sep_close_bigger_starts=close4_starts, from_sep_close_bigger_lists=close4_lists)


def merge_lists(lib, output_to_input_box, starts, lists):
    """traversal.py:1310-1344: count kernel, cumsum, write kernel."""
    nout = len(output_to_input_box)
    nl = len(starts)
    starts = [np.ascontiguousarray(s) for s in starts]
    lists = [np.ascontiguousarray(s) for s in lists]
    sp = (C.c_void_p * nl)(*[s.ctypes.data for s in starts])
    lp = (C.c_void_p * nl)(*[s.ctypes.data for s in lists])
    new_counts = np.zeros(nout + 1, np.int32)
    lib.orc_merge_lists_count(i64(nout), ptr(output_to_input_box), cint(nl), sp, ptr(new_counts))
    new_starts = np.cumsum(new_counts).astype(np.int32)
    new_lists = np.full(int(new_starts[-1]), 999999999, np.int32)
    lib.orc_merge_lists_write(i64(nout), ptr(output_to_input_box), cint(nl), sp, lp,
                              ptr(new_starts), ptr(new_lists))
    return new_starts, new_lists


def merge_close_lists(trav: OracleTraversal) -> OracleTraversal:
    """FMMTraversalInfo.merge_close_lists, traversal.py:1650-1693."""
    from dataclasses import replace
    lib = lib_for(trav.tree.coord_dtype)
    out_to_in = np.arange(trav.ntarget_boxes, dtype=np.int32)
    st, li = merge_lists(
        lib, out_to_in,
        [trav.neighbor_source_boxes_starts, trav.from_sep_close_smaller_starts,
         trav.from_sep_close_bigger_starts],
        [trav.neighbor_source_boxes_lists, trav.from_sep_close_smaller_lists,
         trav.from_sep_close_bigger_lists])
    return replace(trav, neighbor_source_boxes_starts=st, neighbor_source_boxes_lists=li,
                   from_sep_close_smaller_starts=None, from_sep_close_smaller_lists=None,
                   from_sep_close_bigger_starts=None, from_sep_close_bigger_lists=None)


def find_peer_lists(tree):
    """``PeerListFinder.__call__`` (``/root/reference/boxtree/area_query.py:1148-1186``): CSR
    ``(peer_list_starts, peer_lists)`` over all boxes."""
    coord_dtype = np.dtype(tree.coord_dtype)
    lib = lib_for(coord_dtype)
    TravArgs = _make_structs(coord_dtype)
    A = TravArgs()
    arrays = [np.ascontiguousarray(a) for a in (tree.box_centers, tree.box_levels, tree.box_child_ids,
                                                tree.box_flags, tree.box_parent_ids)]
    A.tree.d = tree.dimensions
    A.tree.aligned_nboxes = tree.aligned_nboxes
    A.tree.root_extent = float(tree.root_extent)
    (A.tree.box_centers, A.tree.box_levels, A.tree.box_child_ids, A.tree.box_flags,
     A.tree.box_parent_ids) = (_p(a) for a in arrays)
    A.tree.well_sep_is_n_away = 1
    peers, _ = _ListBuilder(lib, A)(5, tree.nboxes)
    return peers.starts, peers.lists


def area_query(tree, ball_centers, ball_radii, peer_lists=None):
    """``AreaQueryBuilder.__call__`` (``/root/reference/boxtree/area_query.py:757-807``):
    ``(leaves_near_ball_starts, leaves_near_ball_lists)``."""
    coord_dtype = np.dtype(tree.coord_dtype)
    lib = lib_for(coord_dtype)
    TravArgs = _make_structs(coord_dtype)
    A = TravArgs()
    arrays = [np.ascontiguousarray(a) for a in (tree.box_centers, tree.box_levels, tree.box_child_ids,
                                                tree.box_flags, tree.box_parent_ids)]
    A.tree.d = tree.dimensions
    A.tree.aligned_nboxes = tree.aligned_nboxes
    A.tree.root_extent = float(tree.root_extent)
    (A.tree.box_centers, A.tree.box_levels, A.tree.box_child_ids, A.tree.box_flags,
     A.tree.box_parent_ids) = (_p(a) for a in arrays)
    A.tree.well_sep_is_n_away = 1
    if peer_lists is None:
        peer_lists = find_peer_lists(tree)
    pst, pli = (np.ascontiguousarray(a, np.int32) for a in peer_lists)
    centers = [np.ascontiguousarray(c, coord_dtype) for c in ball_centers]
    radii = np.ascontiguousarray(ball_radii, coord_dtype)
    nballs = len(radii)
    cptr = (C.c_void_p * 3)(*[c.ctypes.data for c in centers] + [None] * (3 - len(centers)))
    bbox_min = np.ascontiguousarray(tree.bounding_box[0], coord_dtype)
    counts = np.zeros(nballs, np.int32)
    lib.orc_area_query(C.byref(A.tree), ptr(pst), ptr(pli), i64(nballs), cptr, ptr(radii),
                       ptr(bbox_min), cint(0), ptr(counts), None, None)
    starts = np.zeros(nballs + 1, np.int32)
    np.cumsum(counts, out=starts[1:])
    lists = np.zeros(int(starts[-1]), np.int32)
    lib.orc_area_query(C.byref(A.tree), ptr(pst), ptr(pli), i64(nballs), cptr, ptr(radii),
                       ptr(bbox_min), cint(1), None, ptr(starts), ptr(lists))
    return starts, lists


def leaves_to_balls(tree, ball_centers, ball_radii):
    """``LeavesToBallsLookupBuilder.__call__`` (``area_query.py:844-905``): the area query's pairs
    sorted (stably) by leaf box."""
    starts, lists = area_query(tree, ball_centers, ball_radii)
    nballs = len(starts) - 1
    ball_of_pair = np.repeat(np.arange(nballs, dtype=np.int32), np.diff(starts))
    order = np.argsort(lists, kind="stable")
    box_starts = np.zeros(tree.nboxes + 1, np.int32)
    np.cumsum(np.bincount(lists, minlength=tree.nboxes), out=box_starts[1:])
    return box_starts, ball_of_pair[order]


def space_invader_query(tree, ball_centers, ball_radii):
    """``SpaceInvaderQueryBuilder.__call__`` (``area_query.py:613-650, 990-1048``): per leaf the
    largest centre distance to an overlapping ball, maximum taken in float32."""
    coord_dtype = np.dtype(tree.coord_dtype)
    starts, lists = area_query(tree, ball_centers, ball_radii)
    out = np.zeros(tree.nboxes, np.float32)
    for i in range(len(starts) - 1):
        for leaf in lists[starts[i]:starts[i + 1]]:
            max_dist = coord_dtype.type(0)
            for a in range(tree.dimensions):
                max_dist = max(max_dist, abs(coord_dtype.type(ball_centers[a][i])
                                             - coord_dtype.type(tree.box_centers[a, leaf])))
            out[leaf] = max(out[leaf], np.float32(max_dist))
    return out.astype(coord_dtype)
