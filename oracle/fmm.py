"""Constant-one FMM check (test infrastructure only).

Restates ``drive_fmm`` (``/root/reference/boxtree/fmm.py:342-532``) specialised
to ``ConstantOneExpansionWrangler`` (``boxtree/constant_one.py:50-237``): the
Green's function is identically 1, so a correct tree + traversal delivers
``sum(weights)`` to every target (``test/test_fmm.py:284-285``).  Works on any
object exposing the ``Tree`` / ``FMMTraversalInfo`` attribute names with numpy
arrays (the oracle's or a host copy of the CUDA result).
"""
from __future__ import annotations

import numpy as np


def _csr_row_sums(starts, lists, values):
    cs = np.concatenate([[0.0], np.cumsum(values[lists], dtype=np.float64)])
    starts = np.asarray(starts, np.int64)
    return cs[starts[1:]] - cs[starts[:-1]]


def constant_one_fmm(tree, trav, src_weights_user):
    nboxes = tree.nboxes
    w = np.asarray(src_weights_user, np.float64)[tree.user_source_ids]   # reorder_sources
    W = np.concatenate([[0.0], np.cumsum(w)])
    s0 = np.asarray(tree.box_source_starts[:nboxes], np.int64)
    own_src = W[s0 + tree.box_source_counts_nonchild[:nboxes]] - W[s0]    # _get_source_slice
    t0 = np.asarray(tree.box_target_starts[:nboxes], np.int64)
    tcnt = np.asarray(tree.box_target_counts_nonchild[:nboxes], np.int64)

    def scatter_to_targets(pot, boxes, vals, assign=False):
        for ibox, v in zip(boxes, vals):
            sl = slice(t0[ibox], t0[ibox] + tcnt[ibox])
            if assign:
                pot[sl] = v
            else:
                pot[sl] += v

    # form_multipoles / coarsen_multipoles
    mpoles = np.zeros(nboxes)
    mpoles[trav.source_boxes] += own_src[trav.source_boxes]
    lssp = trav.level_start_source_parent_box_nrs
    for source_level in range(tree.nlevels - 1, 2, -1):
        target_level = source_level - 1
        start, stop = lssp[target_level:target_level + 2]
        for ibox in trav.source_parent_boxes[start:stop]:
            for child in tree.box_child_ids[:, ibox]:
                if child:
                    mpoles[ibox] += mpoles[child]

    def eval_direct(starts, lists):
        pot = np.zeros(tree.ntargets)
        scatter_to_targets(pot, trav.target_boxes, _csr_row_sums(starts, lists, own_src),
                           assign=True)
        return pot

    pot = eval_direct(trav.neighbor_source_boxes_starts, trav.neighbor_source_boxes_lists)

    local = np.zeros(nboxes)
    tp = trav.target_or_target_parent_boxes
    local[tp] += _csr_row_sums(trav.from_sep_siblings_starts, trav.from_sep_siblings_lists,
                               mpoles)

    for lev, ssn in enumerate(trav.from_sep_smaller_by_level):
        tb = trav.target_boxes_sep_smaller_by_source_level[lev]
        scatter_to_targets(pot, tb, _csr_row_sums(ssn.starts, ssn.lists, mpoles))

    if trav.from_sep_close_smaller_starts is not None:
        pot = pot + eval_direct(trav.from_sep_close_smaller_starts,
                                trav.from_sep_close_smaller_lists)

    local[tp] += _csr_row_sums(trav.from_sep_bigger_starts, trav.from_sep_bigger_lists,
                               own_src)
    if trav.from_sep_close_bigger_starts is not None:
        pot = pot + eval_direct(trav.from_sep_close_bigger_starts,
                                trav.from_sep_close_bigger_lists)

    lstp = trav.level_start_target_or_target_parent_box_nrs
    for target_lev in range(1, tree.nlevels):
        start, stop = lstp[target_lev:target_lev + 2]
        boxes = tp[start:stop]
        local[boxes] += local[tree.box_parent_ids[boxes]]

    scatter_to_targets(pot, trav.target_boxes, local[trav.target_boxes])
    return pot[tree.sorted_target_ids]                                   # reorder_potentials
