"""``PeerListFinder`` / ``PeerListLookup`` (``boxtree/area_query.py:92-113, 1057-1188``): for every
box the adjacent boxes that are at least as large and have no child with both properties.  The
reference's level-restriction test (``test/test_tree.py:929-974``) is written in terms of these
lists.  Same count -> scan -> write protocol and row order as the reference's walk
(``csrc/traversal.cu``, ``gen_peers``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from . import _cabi
from ._cabi import bt_list_args, bt_tree_view, check, dptr
from .array_context import TorchArrayContext


@dataclass(frozen=True)
class PeerListLookup:
    """``area_query.py:92-113``: ``peer_lists[peer_list_starts[b]:peer_list_starts[b+1]]``."""
    tree: Any
    peer_list_starts: Any
    peer_lists: Any


class PeerListFinder:
    def __init__(self, array_context: TorchArrayContext) -> None:
        assert isinstance(array_context, TorchArrayContext)
        self._setup_actx = array_context
        self._lib = _cabi.load()

    def __call__(self, actx, tree, wait_for=None):
        """:returns: ``(PeerListLookup, event)`` (``area_query.py:1148-1186``)."""
        assert isinstance(actx, TorchArrayContext)
        lib = self._lib
        nboxes = int(tree.nboxes)
        sh = actx.stream_handle
        with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
            tv = bt_tree_view()
            tv.dim = int(tree.dimensions)
            tv.nboxes = nboxes
            tv.aligned_nboxes = int(tree.box_child_ids.shape[-1])
            tv.nlevels = int(tree.nlevels)
            tv.root_extent = float(tree.root_extent)
            keep = [t.contiguous() for t in (tree.box_centers, tree.box_levels, tree.box_child_ids,
                                            tree.box_flags, tree.box_parent_ids)]
            tv.box_centers, tv.box_levels, tv.box_child_ids, tv.box_flags, tv.box_parent_ids = \
                (dptr(t) for t in keep)
            tv.well_sep_is_n_away = 1
            dcode = _cabi.dtype_code(np.dtype(tree.coord_dtype))
            args = bt_list_args()
            starts = actx.empty(nboxes + 1, np.int32)
            totals = actx.zeros(2, np.int64)
            check(lib.bt_trav_build_list(dcode, 5, 0, C.byref(tv), C.byref(args), nboxes,
                                         dptr(starts), None, None, None, dptr(totals), sh),
                  "peer lists count")
            total = int(totals[0].item())
            lists = actx.empty(total, np.int32)
            check(lib.bt_trav_build_list(dcode, 5, 1, C.byref(tv), C.byref(args), nboxes,
                                         dptr(starts), dptr(lists), None, None, dptr(totals), sh),
                  "peer lists fill")
            evt = torch.cuda.Event()
            evt.record(actx.stream)
        return actx.freeze(PeerListLookup(tree=tree, peer_list_starts=starts, peer_lists=lists)), evt


@dataclass(frozen=True)
class AreaQueryResult:
    """``area_query.py:115-140``: the leaf boxes near every ball, CSR over the balls."""
    tree: Any
    leaves_near_ball_starts: Any
    leaves_near_ball_lists: Any


def _tree_view(tree):
    tv = bt_tree_view()
    tv.dim = int(tree.dimensions)
    tv.nboxes = int(tree.nboxes)
    tv.aligned_nboxes = int(tree.box_child_ids.shape[-1])
    tv.nlevels = int(tree.nlevels)
    tv.root_extent = float(tree.root_extent)
    keep = [t.contiguous() for t in (tree.box_centers, tree.box_levels, tree.box_child_ids,
                                    tree.box_flags, tree.box_parent_ids)]
    tv.box_centers, tv.box_levels, tv.box_child_ids, tv.box_flags, tv.box_parent_ids = \
        (dptr(t) for t in keep)
    tv.well_sep_is_n_away = 1
    return tv, keep


class AreaQueryBuilder:
    r"""Look-up table from :math:`l^\infty` balls to the leaf boxes that intersect them
    (``area_query.py:657-807``)."""

    def __init__(self, array_context: TorchArrayContext) -> None:
        assert isinstance(array_context, TorchArrayContext)
        self._setup_actx = array_context
        self.peer_list_finder = PeerListFinder(array_context)
        self._lib = _cabi.load()

    def __call__(self, actx, tree, ball_centers, ball_radii, peer_lists=None, wait_for=None):
        assert isinstance(actx, TorchArrayContext)
        coord_dtype = np.dtype(tree.coord_dtype)
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[coord_dtype]
        if any(bc.dtype != tdt for bc in ball_centers):
            raise TypeError("ball_centers dtype must match tree.coord_dtype")
        if ball_radii.dtype != tdt:
            raise TypeError("ball_radii dtype must match tree.coord_dtype")
        if peer_lists is None:
            peer_lists, _ = self.peer_list_finder(actx, tree, wait_for=wait_for)
        if int(peer_lists.peer_list_starts.shape[0]) != int(tree.nboxes) + 1:
            raise ValueError("size of peer lists must match with number of boxes")
        lib = self._lib
        sh = actx.stream_handle
        nballs = int(ball_radii.shape[0])
        with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
            tv, keep = _tree_view(tree)
            centers = [bc.contiguous() for bc in ball_centers]
            radii = ball_radii.contiguous()
            dcode = _cabi.dtype_code(coord_dtype)
            starts = actx.empty(nballs + 1, np.int32)
            totals = actx.zeros(2, np.int64)
            cptr = _cabi.ptr_array(centers)
            bbox_min = _cabi.darray(np.asarray(tree.bounding_box[0], np.float64))

            def run(phase, lists):
                check(lib.bt_area_query(dcode, phase, C.byref(tv), dptr(peer_lists.peer_list_starts),
                                        dptr(peer_lists.peer_lists), nballs, cptr, dptr(radii),
                                        bbox_min, dptr(starts), dptr(lists), dptr(totals), sh),
                      "bt_area_query")
            run(0, None)
            lists = actx.empty(int(totals[0].item()), np.int32)
            run(1, lists)
            evt = torch.cuda.Event()
            evt.record(actx.stream)
        return actx.freeze(AreaQueryResult(tree=tree, leaves_near_ball_starts=starts,
                                           leaves_near_ball_lists=lists)), evt


@dataclass(frozen=True)
class LeavesToBallsLookup:
    """``area_query.py:143-163``: the balls overlapping every leaf box, CSR over ALL boxes."""
    tree: Any
    balls_near_box_starts: Any
    balls_near_box_lists: Any


class LeavesToBallsLookupBuilder:
    """``area_query.py:810-905``: the area query turned around (expand the starts, stable
    key-value sort by leaf box)."""

    def __init__(self, array_context: TorchArrayContext) -> None:
        self._setup_actx = array_context
        self.area_query_builder = AreaQueryBuilder(array_context)

    def __call__(self, actx, tree, ball_centers, ball_radii, peer_lists=None, wait_for=None):
        aq, _ = self.area_query_builder(actx, tree, ball_centers, ball_radii, peer_lists, wait_for)
        with torch.cuda.stream(actx.stream):
            nboxes = int(tree.nboxes)
            starts = aq.leaves_near_ball_starts.long()
            nballs = int(starts.shape[0]) - 1
            npairs = int(aq.leaves_near_ball_lists.shape[0])
            dev = starts.device
            # STARTS_EXPANDER_TEMPLATE: [0 2 5 6] -> [0 0 1 1 1 2]
            ball_of_pair = torch.repeat_interleave(torch.arange(nballs, device=dev),
                                                   starts[1:] - starts[:-1], output_size=npairs)
            # KeyValueSorter: stable by key, so the balls of a box stay in ascending order
            keys = aq.leaves_near_ball_lists.long()
            order = torch.argsort(keys, stable=True)
            lists = ball_of_pair[order].to(torch.int32)
            box_starts = torch.zeros(nboxes + 1, dtype=torch.int32, device=dev)
            box_starts[1:] = torch.cumsum(torch.bincount(keys, minlength=nboxes), 0).to(torch.int32)
            evt = torch.cuda.Event()
            evt.record(actx.stream)
        return actx.freeze(LeavesToBallsLookup(tree=tree, balls_near_box_starts=box_starts,
                                               balls_near_box_lists=lists)), evt


class SpaceInvaderQueryBuilder:
    r"""``area_query.py:908-1048``: per leaf box the *outer space invader distance*
    :math:`\max_{b^* \cap b \ne \emptyset} d_\infty(\mathrm{center}(b), \mathrm{center}(b^*))`
    (0 for other boxes).  Like the reference the maximum is taken in float32 and cast back to the
    coordinate dtype (``:613-650, 1036-1044``)."""

    def __init__(self, array_context: TorchArrayContext) -> None:
        self._setup_actx = array_context
        self.area_query_builder = AreaQueryBuilder(array_context)

    def __call__(self, actx, tree, ball_centers, ball_radii, peer_lists=None, wait_for=None):
        aq, _ = self.area_query_builder(actx, tree, ball_centers, ball_radii, peer_lists, wait_for)
        with torch.cuda.stream(actx.stream):
            nboxes = int(tree.nboxes)
            starts = aq.leaves_near_ball_starts.long()
            nballs = int(starts.shape[0]) - 1
            npairs = int(aq.leaves_near_ball_lists.shape[0])
            dev = starts.device
            ball = torch.repeat_interleave(torch.arange(nballs, device=dev), starts[1:] - starts[:-1],
                                           output_size=npairs)
            leaf = aq.leaves_near_ball_lists.long()
            max_dist = torch.zeros(npairs, dtype=ball_radii.dtype, device=dev)
            for a, bc in enumerate(ball_centers):
                max_dist = torch.maximum(max_dist, (bc[ball] - tree.box_centers[a][leaf]).abs())
            out = torch.zeros(nboxes, dtype=torch.float32, device=dev)
            out.scatter_reduce_(0, leaf, max_dist.to(torch.float32), reduce="amax", include_self=True)
            out = out.to(ball_radii.dtype)
            evt = torch.cuda.Event()
            evt.record(actx.stream)
        return actx.freeze(out), evt
