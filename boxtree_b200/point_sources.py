"""``link_point_sources`` / ``TreeWithLinkedPointSources`` (``boxtree/tree.py:720-955`` with the
kernels of ``boxtree/tree_build_kernels.py:1869-1950``): point sources attached to the
(extent-having) sources of a tree are reordered so that every box's point sources are
contiguous.  Consumer-side utility built from torch scans and gathers.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any

import numpy as np
import torch

from .array_context import TorchArrayContext, make_obj_array
from .tree import Tree


@dataclass(frozen=True)
class TreeWithLinkedPointSources(Tree):
    """``tree.py:720-770``: a :class:`Tree` plus the linked point sources in tree order."""
    npoint_sources: int = 0
    point_source_starts: Any = None
    point_source_counts: Any = None
    point_sources: Any = None
    user_point_source_ids: Any = None
    box_point_source_starts: Any = None
    box_point_source_counts_nonchild: Any = None
    box_point_source_counts_cumul: Any = None


def link_point_sources(actx, tree, point_source_starts, point_sources, *, debug=False):
    """See ``boxtree/tree.py:773-955``.  *point_source_starts* ``[nsources + 1]`` is indexed in
    user source order; *point_sources* is an object array of coordinate arrays.  Every source
    is expected to own at least one point source (with empty ranges the reference's
    ``multi_put`` writes two values to one slot)."""
    assert isinstance(actx, TorchArrayContext)
    if not tree.sources_have_extent:
        raise ValueError("only allowed on trees whose sources have extent")
    from . import _cabi
    from ._cabi import check, dptr
    lib = _cabi.load()
    sh = actx.stream_handle
    with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
        nboxes, nsources = int(tree.nboxes), int(tree.nsources)
        pss = point_source_starts
        pss = (actx.from_numpy(pss) if isinstance(pss, np.ndarray) else pss).to(torch.int32).contiguous()
        tree_order_starts = actx.empty(nsources + 1, np.int32)
        cnt = actx.empty(max(nsources, 1), np.int32)
        total = actx.zeros(1, np.int32)
        ps_start = actx.empty(nboxes, np.int32)
        ps_nonchild = actx.empty(nboxes, np.int32)
        ps_cumul = actx.empty(nboxes, np.int32)

        def call(phase, ids):
            # POINT_SOURCE_LINKING_SOURCE_SCAN_TPL (:1872-1897), then multi_put + segmented scan
            # (:1899-1911, tree.py:838-890) and POINT_SOURCE_LINKING_BOX_POINT_SOURCES (:1913-1950)
            check(lib.bt_link_point_sources(
                phase, nboxes, nsources, dptr(pss), dptr(tree.user_source_ids),
                dptr(tree_order_starts), dptr(cnt), dptr(total), dptr(ids),
                dptr(tree.box_source_starts), dptr(tree.box_source_counts_nonchild),
                dptr(tree.box_source_counts_cumul), dptr(ps_start), dptr(ps_nonchild),
                dptr(ps_cumul), sh), "bt_link_point_sources")

        call(0, None)
        npoint_sources = int(total.item()) if nsources else 0
        user_point_source_ids = actx.empty(npoint_sources, np.int32)
        call(1, user_point_source_ids)
        pts = [actx.from_numpy(p) if isinstance(p, np.ndarray) else p for p in point_sources]
        dcode = _cabi.dtype_code(tree.coord_dtype)
        tops = []
        for p in pts:
            o = actx.empty(npoint_sources, tree.coord_dtype)
            check(lib.bt_gather_coords(dcode, npoint_sources, dptr(p.contiguous()),
                                       dptr(user_point_source_ids), dptr(o), sh), "bt_gather_coords")
            tops.append(o)
        extra = dict(
            npoint_sources=npoint_sources,
            point_source_starts=tree_order_starts[:nsources],
            point_source_counts=cnt[:nsources],
            point_sources=make_obj_array(tops),
            user_point_source_ids=user_point_source_ids,
            box_point_source_starts=ps_start,
            box_point_source_counts_nonchild=ps_nonchild,
            box_point_source_counts_cumul=ps_cumul)
    base = {f.name: getattr(tree, f.name) for f in fields(Tree)}
    return actx.freeze(TreeWithLinkedPointSources(**base, **extra))
