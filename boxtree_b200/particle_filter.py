"""``ParticleListFilter`` (``boxtree/tree.py:1040-1239``): subsets of the per-box target lists
selected by a flag per target (user order).  A consumer-side utility of the tree (SURVEY
§8(f) N2), built from torch primitives -- scans, gathers and a stable sort -- on the
array context's stream; same class names, fields and dtypes as the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from .array_context import TorchArrayContext, make_obj_array


@dataclass(frozen=True)
class FilteredTargetListsInUserOrder:
    """``tree.py:957-996``: CSR ``target_starts [nboxes+1]`` / ``target_lists`` with the
    filtered targets of every box (own targets only) as *user* target ids."""
    nfiltered_targets: int
    target_starts: Any
    target_lists: Any


@dataclass(frozen=True)
class FilteredTargetListsInTreeOrder:
    """``tree.py:999-1037``: a renumbering of the targets that counts only the filtered ones,
    with per-box ``box_target_starts`` / ``box_target_counts_nonchild`` into it."""
    nfiltered_targets: int
    box_target_starts: Any
    box_target_counts_nonchild: Any
    targets: Any
    unfiltered_from_filtered_target_indices: Any


class ParticleListFilter:
    def __init__(self, array_context: TorchArrayContext):
        assert isinstance(array_context, TorchArrayContext)
        self._setup_actx = array_context

    def filter_target_lists_in_user_order(self, actx, tree, flags) -> FilteredTargetListsInUserOrder:
        """``tree.py:1097-1130``; *flags*: int8 ``[ntargets]`` in user target order."""
        assert isinstance(actx, TorchArrayContext)
        with torch.cuda.stream(actx.stream):
            nboxes, ntargets = int(tree.nboxes), int(tree.ntargets)
            dev = tree.box_flags.device
            flags = actx.from_numpy(flags) if isinstance(flags, np.ndarray) else flags
            sti = tree.sorted_target_ids.long()
            # user_target_ids[tree position] = user id  (tree.py:1108-1111)
            user_target_ids = torch.empty(ntargets, dtype=torch.int32, device=dev)
            user_target_ids[sti] = torch.arange(ntargets, dtype=torch.int32, device=dev)
            starts = tree.box_target_starts[:nboxes].long()
            counts = tree.box_target_counts_nonchild[:nboxes].long()
            # owning box of every tree-order target position: the own-target ranges of the boxes
            # tile [0, ntargets)
            has = torch.nonzero(counts > 0).flatten()
            order = torch.argsort(starts[has], stable=True)
            owner = torch.repeat_interleave(has[order], counts[has[order]])
            keep = flags[user_target_ids.long()] != 0
            kept_pos = torch.nonzero(keep).flatten()
            kept_owner = owner[kept_pos]
            # rows in box order, entries in the order of the generate() loop (ascending position)
            by_box = torch.argsort(kept_owner, stable=True)
            lists = user_target_ids[kept_pos[by_box]]
            per_box = torch.bincount(kept_owner, minlength=nboxes)
            target_starts = torch.zeros(nboxes + 1, dtype=torch.int32, device=dev)
            target_starts[1:] = torch.cumsum(per_box, 0).to(torch.int32)
            n = int(lists.shape[0])
        return actx.freeze(FilteredTargetListsInUserOrder(
            nfiltered_targets=n, target_starts=target_starts, target_lists=lists))

    def filter_target_lists_in_tree_order(self, actx, tree, flags) -> FilteredTargetListsInTreeOrder:
        """``tree.py:1160-1239`` with ``TREE_ORDER_TARGET_FILTER_SCAN_TPL`` /
        ``TREE_ORDER_TARGET_FILTER_INDEX_TPL`` (``tree_build_kernels.py:1954-2021``)."""
        assert isinstance(actx, TorchArrayContext)
        with torch.cuda.stream(actx.stream):
            nboxes, ntargets = int(tree.nboxes), int(tree.ntargets)
            dev = tree.box_flags.device
            flags = actx.from_numpy(flags) if isinstance(flags, np.ndarray) else flags
            tree_order_flags = torch.zeros(ntargets, dtype=torch.int8, device=dev)
            tree_order_flags[tree.sorted_target_ids.long()] = flags.to(torch.int8)
            f = (tree_order_flags != 0).to(torch.int32)
            incl = torch.cumsum(f, 0, dtype=torch.int32)
            filtered_from_unfiltered = incl - f                     # prev_item
            unfiltered_from_filtered = torch.nonzero(f).flatten().to(torch.int32)
            nfiltered = int(unfiltered_from_filtered.shape[0])
            targets = make_obj_array([t[unfiltered_from_filtered.long()] for t in tree.targets])
            # one extra entry for ranges that end at ntargets (the reference branches, :2004-2013)
            ffu = torch.cat([filtered_from_unfiltered,
                             torch.full((1,), nfiltered, dtype=torch.int32, device=dev)])
            starts = tree.box_target_starts[:nboxes].long()
            counts = tree.box_target_counts_nonchild[:nboxes].long()
            fstart = ffu[starts.clamp(max=ntargets)]
            fend = ffu[(starts + counts).clamp(max=ntargets)]
            fcount = torch.where(counts > 0, fend - fstart, torch.zeros_like(fstart))
        return actx.freeze(FilteredTargetListsInTreeOrder(
            nfiltered_targets=nfiltered, box_target_starts=fstart.to(torch.int32),
            box_target_counts_nonchild=fcount.to(torch.int32), targets=targets,
            unfiltered_from_filtered_target_indices=unfiltered_from_filtered))
