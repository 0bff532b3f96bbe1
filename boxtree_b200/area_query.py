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
