"""Constant-one FMM on the device: the reference's strongest end-to-end check of a tree +
traversal pair, at full problem size.

``drive_fmm`` (``boxtree/fmm.py:342-532``) specialised to
``ConstantOneExpansionWrangler`` (``boxtree/constant_one.py:50-237``): the kernel is
identically 1, a multipole/local "expansion" is one number, and every translation is an
addition, so each interaction list becomes a CSR segmented sum.  A complete, non-overlapping
set of lists delivers ``sum(weights)`` to every target (``test/test_fmm.py:284-285``).
Integer arithmetic (int64) throughout, hence exact at any size.  Built from torch
primitives (gather, cumsum, index_add_): it is a consumer of the traversal, not part of
the hot path.
"""
from __future__ import annotations

import torch


def _row_sums(starts, lists, values):
    """``out[i] = sum(values[lists[starts[i]:starts[i+1]]])``."""
    v = values[lists.long()]
    cs = torch.cat([torch.zeros(1, dtype=values.dtype, device=values.device), torch.cumsum(v, 0)])
    st = starts.long()
    return cs[st[1:]] - cs[st[:-1]]


def constant_one_fmm(tree, trav, src_weights_user):
    """Potentials in user target order (int64 tensor) for integer *src_weights_user*
    (indexed like the particles handed to ``TreeBuilder``)."""
    dev = tree.box_flags.device
    nboxes = int(tree.nboxes)
    ntargets = int(tree.ntargets)
    w = torch.as_tensor(src_weights_user, device=dev).to(torch.int64)
    w = w[tree.user_source_ids.long()]                                     # reorder_sources
    W = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(w, 0)])
    s0 = tree.box_source_starts[:nboxes].long()
    own_src = W[s0 + tree.box_source_counts_nonchild[:nboxes].long()] - W[s0]
    t0 = tree.box_target_starts[:nboxes].long()
    tcnt = tree.box_target_counts_nonchild[:nboxes].long()
    child_ids = tree.box_child_ids[:, :nboxes].long()
    parents = tree.box_parent_ids[:nboxes].long()

    def to_targets(boxes, vals):
        """Per-target array with vals[i] on the own targets of boxes[i] (ranges are disjoint)."""
        b = boxes.long()
        diff = torch.zeros(ntargets + 1, dtype=torch.int64, device=dev)
        diff.index_add_(0, t0[b], vals)
        diff.index_add_(0, t0[b] + tcnt[b], -vals)
        return torch.cumsum(diff, 0)[:ntargets]

    # form_multipoles, coarsen_multipoles (fmm.py:391-406, constant_one.py:104-152)
    mpoles = torch.zeros(nboxes, dtype=torch.int64, device=dev)
    sb = trav.source_boxes.long()
    mpoles[sb] += own_src[sb]
    lssp = trav.level_start_source_parent_box_nrs.cpu().tolist()
    spb = trav.source_parent_boxes.long()
    for source_level in range(int(tree.nlevels) - 1, 2, -1):
        boxes = spb[lssp[source_level - 1]:lssp[source_level]]
        ch = child_ids[:, boxes]
        mpoles[boxes] += torch.where(ch != 0, mpoles[ch], torch.zeros_like(ch)).sum(0)

    tb = trav.target_boxes
    pot = to_targets(tb, _row_sums(trav.neighbor_source_boxes_starts,
                                   trav.neighbor_source_boxes_lists, own_src))      # list 1
    local = torch.zeros(nboxes, dtype=torch.int64, device=dev)
    tp = trav.target_or_target_parent_boxes.long()
    local[tp] += _row_sums(trav.from_sep_siblings_starts, trav.from_sep_siblings_lists, mpoles)
    for lev, ssn in enumerate(trav.from_sep_smaller_by_level):                       # list 3
        pot += to_targets(trav.target_boxes_sep_smaller_by_source_level[lev],
                          _row_sums(ssn.starts, ssn.lists, mpoles))
    if trav.from_sep_close_smaller_starts is not None:
        pot += to_targets(tb, _row_sums(trav.from_sep_close_smaller_starts,
                                        trav.from_sep_close_smaller_lists, own_src))
    local[tp] += _row_sums(trav.from_sep_bigger_starts, trav.from_sep_bigger_lists, own_src)
    if trav.from_sep_close_bigger_starts is not None:
        pot += to_targets(tb, _row_sums(trav.from_sep_close_bigger_starts,
                                        trav.from_sep_close_bigger_lists, own_src))
    lstp = trav.level_start_target_or_target_parent_box_nrs.cpu().tolist()           # downward
    for target_lev in range(1, int(tree.nlevels)):
        boxes = tp[lstp[target_lev]:lstp[target_lev + 1]]
        local[boxes] += local[parents[boxes]]
    pot += to_targets(tb, local[tb.long()])
    return pot[tree.sorted_target_ids.long()]                                        # reorder_potentials
