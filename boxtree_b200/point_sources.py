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
    with torch.cuda.stream(actx.stream):
        dev = tree.box_flags.device
        nboxes, nsources = int(tree.nboxes), int(tree.nsources)
        pss = point_source_starts
        pss = (actx.from_numpy(pss) if isinstance(pss, np.ndarray) else pss).long()
        usi = tree.user_source_ids.long()
        # POINT_SOURCE_LINKING_SOURCE_SCAN_TPL (:1872-1897)
        cnt = pss[usi + 1] - pss[usi]
        incl = torch.cumsum(cnt, 0)
        tree_order_starts = incl - cnt
        npoint_sources = int(incl[-1].item()) if nsources else 0
        # multi_put + segmented scan (:1899-1911, tree.py:838-890): within the segment of tree-order
        # source i the ids count up from that source's first point source in user order
        seg_first = pss[usi] - tree_order_starts
        user_point_source_ids = (torch.repeat_interleave(seg_first, cnt, output_size=npoint_sources)
                                 + torch.arange(npoint_sources, device=dev)).to(torch.int32)
        pts = [actx.from_numpy(p) if isinstance(p, np.ndarray) else p for p in point_sources]
        tree_order_point_sources = make_obj_array([p[user_point_source_ids.long()] for p in pts])
        # POINT_SOURCE_LINKING_BOX_POINT_SOURCES (:1913-1950)
        s_start = tree.box_source_starts[:nboxes].long()
        tos = torch.cat([tree_order_starts, incl[-1:] if nsources else torch.zeros(1, dtype=torch.long, device=dev)])
        ps_start = tos[s_start.clamp(max=nsources)]

        def counts(s_count):
            s_count = s_count[:nboxes].long()
            last = (s_start + s_count - 1).clamp(min=0, max=max(nsources - 1, 0))
            beyond = tree_order_starts[last] + cnt[last]
            return torch.where(s_count == 0, torch.zeros_like(beyond), beyond - ps_start).to(torch.int32)

        extra = dict(
            npoint_sources=npoint_sources,
            point_source_starts=tree_order_starts.to(torch.int32),
            point_source_counts=cnt.to(torch.int32),
            point_sources=tree_order_point_sources,
            user_point_source_ids=user_point_source_ids,
            box_point_source_starts=ps_start.to(torch.int32),
            box_point_source_counts_nonchild=counts(tree.box_source_counts_nonchild),
            box_point_source_counts_cumul=counts(tree.box_source_counts_cumul))
    base = {f.name: getattr(tree, f.name) for f in fields(Tree)}
    return actx.freeze(TreeWithLinkedPointSources(**base, **extra))
