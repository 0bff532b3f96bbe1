"""``ParticleListFilter`` (``boxtree/tree.py:1040-1239``): subsets of the per-box target lists
selected by a flag per target (user order).  A consumer-side utility of the tree (SURVEY
§8(f) N2): the kernels ``bt_filter_targets_user_order`` / ``bt_filter_targets_tree_order`` of
``csrc/consumers.cu`` (ballot-ordered compaction per box; one look-back scan) behind the C
ABI, on the array context's stream; same class names, fields and dtypes as the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from . import _cabi
from ._cabi import check, dptr
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
        lib = _cabi.load()
        sh = actx.stream_handle
        with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
            nboxes, ntargets = int(tree.nboxes), int(tree.ntargets)
            flags = (actx.from_numpy(flags) if isinstance(flags, np.ndarray) else flags) \
                .to(torch.int8).contiguous()
            # user_target_ids[tree position] = user id  (tree.py:1108-1111)
            user_target_ids = actx.empty(max(ntargets, 1), np.int32)
            check(lib.bt_reverse_index(ntargets, dptr(tree.sorted_target_ids), dptr(user_target_ids),
                                       sh), "bt_reverse_index")
            target_starts = actx.empty(nboxes + 1, np.int32)
            total = actx.zeros(1, np.int64)
            args = (nboxes, dptr(tree.box_target_starts), dptr(tree.box_target_counts_nonchild),
                    dptr(user_target_ids), dptr(flags), dptr(target_starts))
            check(lib.bt_filter_targets_user_order(0, *args, None, dptr(total), sh),
                  "bt_filter_targets_user_order")
            n = int(total.item())
            lists = actx.empty(n, np.int32)
            check(lib.bt_filter_targets_user_order(1, *args, dptr(lists), dptr(total), sh),
                  "bt_filter_targets_user_order")
        return actx.freeze(FilteredTargetListsInUserOrder(
            nfiltered_targets=n, target_starts=target_starts, target_lists=lists))

    def filter_target_lists_in_tree_order(self, actx, tree, flags) -> FilteredTargetListsInTreeOrder:
        """``tree.py:1160-1239`` with ``TREE_ORDER_TARGET_FILTER_SCAN_TPL`` /
        ``TREE_ORDER_TARGET_FILTER_INDEX_TPL`` (``tree_build_kernels.py:1954-2021``)."""
        assert isinstance(actx, TorchArrayContext)
        lib = _cabi.load()
        sh = actx.stream_handle
        with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
            nboxes, ntargets = int(tree.nboxes), int(tree.ntargets)
            flags = (actx.from_numpy(flags) if isinstance(flags, np.ndarray) else flags) \
                .to(torch.int8).contiguous()
            user_target_ids = actx.empty(max(ntargets, 1), np.int32)
            check(lib.bt_reverse_index(ntargets, dptr(tree.sorted_target_ids), dptr(user_target_ids),
                                       sh), "bt_reverse_index")
            ffu = actx.empty(ntargets + 1, np.int32)
            uff = actx.empty(max(ntargets, 1), np.int32)
            nf = actx.zeros(1, np.int32)
            fstart = actx.empty(nboxes, np.int32)
            fcount = actx.empty(nboxes, np.int32)
            check(lib.bt_filter_targets_tree_order(
                nboxes, ntargets, dptr(tree.box_target_starts),
                dptr(tree.box_target_counts_nonchild), dptr(user_target_ids), dptr(flags), dptr(ffu),
                dptr(uff), dptr(nf), dptr(fstart), dptr(fcount), sh), "bt_filter_targets_tree_order")
            nfiltered = int(nf.item())
            unfiltered_from_filtered = uff[:nfiltered]
            dcode = _cabi.dtype_code(tree.coord_dtype)
            targets = []
            for t in tree.targets:
                o = actx.empty(nfiltered, tree.coord_dtype)
                check(lib.bt_gather_coords(dcode, nfiltered, dptr(t), dptr(unfiltered_from_filtered),
                                           dptr(o), sh), "bt_gather_coords")
                targets.append(o)
        return actx.freeze(FilteredTargetListsInTreeOrder(
            nfiltered_targets=nfiltered, box_target_starts=fstart,
            box_target_counts_nonchild=fcount, targets=make_obj_array(targets),
            unfiltered_from_filtered_target_indices=unfiltered_from_filtered))
