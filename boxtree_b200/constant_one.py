"""Constant-one FMM on the device: the reference's strongest end-to-end check of a tree +
traversal pair, at full problem size.

``drive_fmm`` (``boxtree/fmm.py:342-532``) specialised to
``ConstantOneExpansionWrangler`` (``boxtree/constant_one.py:50-237``): the kernel is
identically 1, a multipole/local "expansion" is one number, and every translation is an
addition, so each interaction list becomes a CSR row sum.  A complete, non-overlapping set of
lists delivers ``sum(weights)`` to every target (``test/test_fmm.py:284-285``).  Integer
arithmetic (int64) throughout, hence exact at any size.  Every stage is a kernel of
``csrc/consumers.cu`` behind the C ABI (``bt_csr_row_sums``, ``bt_range_sums_i64``,
``bt_add_to_ranges_i64``, ``bt_fmm_upward_i64``, ``bt_fmm_downward_i64``, ``bt_gather_i64``);
torch only owns the buffers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi
from ._cabi import check, dptr


def constant_one_fmm(tree, trav, src_weights_user, actx=None):
    """Potentials in user target order (int64 tensor) for integer *src_weights_user*
    (indexed like the particles handed to ``TreeBuilder``)."""
    lib = _cabi.load()
    dev = tree.box_flags.device
    stream = actx.stream if actx is not None else torch.cuda.current_stream(dev)
    sh = stream.cuda_stream
    nboxes, ntargets, nsources = int(tree.nboxes), int(tree.ntargets), int(tree.nsources)
    dims, aligned = int(tree.dimensions), int(tree.aligned_nboxes)

    def i64(n):
        return torch.zeros(max(int(n), 1), dtype=torch.int64, device=dev)

    with torch.cuda.stream(stream), torch.cuda.device(dev):
        w_user = torch.as_tensor(src_weights_user, device=dev).to(torch.int64).contiguous()
        w = i64(nsources)                                                       # reorder_sources
        check(lib.bt_gather_i64(nsources, dptr(w_user), dptr(tree.user_source_ids), dptr(w), sh),
              "bt_gather_i64")
        own_src = i64(nboxes)
        check(lib.bt_range_sums_i64(nboxes, dptr(tree.box_source_starts),
                                    dptr(tree.box_source_counts_nonchild), dptr(w), dptr(own_src),
                                    sh), "bt_range_sums_i64")

        def row_sums(starts, lists, values, out, out_index=None, accumulate=0):
            nrows = int(starts.shape[0]) - 1
            check(lib.bt_csr_row_sums(0, nrows, dptr(starts), dptr(lists), dptr(values),
                                      dptr(out_index), dptr(out), accumulate, 1.0, sh),
                  "bt_csr_row_sums")

        def to_targets(boxes, vals, pot):
            check(lib.bt_add_to_ranges_i64(int(boxes.shape[0]), dptr(boxes), dptr(vals),
                                           dptr(tree.box_target_starts),
                                           dptr(tree.box_target_counts_nonchild), dptr(pot), sh),
                  "bt_add_to_ranges_i64")

        # form_multipoles, coarsen_multipoles (fmm.py:391-406, constant_one.py:104-152)
        mpoles = i64(nboxes)
        sb = trav.source_boxes
        mpoles[sb.long()] = own_src[sb.long()]
        lssp = trav.level_start_source_parent_box_nrs.cpu().tolist()
        spb = trav.source_parent_boxes
        for source_level in range(int(tree.nlevels) - 1, 2, -1):
            lo, hi = lssp[source_level - 1], lssp[source_level]
            check(lib.bt_fmm_upward_i64(dims, hi - lo, dptr(spb[lo:hi]), dptr(tree.box_child_ids),
                                        aligned, dptr(mpoles), sh), "bt_fmm_upward_i64")

        tb = trav.target_boxes
        ntb = int(tb.shape[0])
        pot = i64(ntargets)
        tmp = i64(ntb)
        row_sums(trav.neighbor_source_boxes_starts, trav.neighbor_source_boxes_lists, own_src, tmp)
        to_targets(tb, tmp, pot)                                                     # list 1
        local = i64(nboxes)
        tp = trav.target_or_target_parent_boxes
        row_sums(trav.from_sep_siblings_starts, trav.from_sep_siblings_lists, mpoles, local,
                 out_index=tp, accumulate=1)                                         # list 2
        for lev, ssn in enumerate(trav.from_sep_smaller_by_level):                   # list 3
            boxes = trav.target_boxes_sep_smaller_by_source_level[lev]
            if int(boxes.shape[0]) == 0:
                continue
            t3 = i64(boxes.shape[0])
            row_sums(ssn.starts, ssn.lists, mpoles, t3)
            to_targets(boxes, t3, pot)
        if trav.from_sep_close_smaller_starts is not None:
            row_sums(trav.from_sep_close_smaller_starts, trav.from_sep_close_smaller_lists, own_src,
                     tmp)
            to_targets(tb, tmp, pot)
        row_sums(trav.from_sep_bigger_starts, trav.from_sep_bigger_lists, own_src, local,
                 out_index=tp, accumulate=1)                                         # list 4
        if trav.from_sep_close_bigger_starts is not None:
            row_sums(trav.from_sep_close_bigger_starts, trav.from_sep_close_bigger_lists, own_src,
                     tmp)
            to_targets(tb, tmp, pot)
        lstp = trav.level_start_target_or_target_parent_box_nrs.cpu().tolist()       # downward
        for target_lev in range(1, int(tree.nlevels)):
            lo, hi = lstp[target_lev], lstp[target_lev + 1]
            check(lib.bt_fmm_downward_i64(hi - lo, dptr(tp[lo:hi]), dptr(tree.box_parent_ids),
                                          dptr(local), sh), "bt_fmm_downward_i64")
        check(lib.bt_gather_i64(ntb, dptr(local), dptr(tb), dptr(tmp), sh), "bt_gather_i64")
        to_targets(tb, tmp, pot)
        out = i64(ntargets)                                                   # reorder_potentials
        check(lib.bt_gather_i64(ntargets, dptr(pot), dptr(tree.sorted_target_ids), dptr(out), sh),
              "bt_gather_i64")
    return out[:ntargets]
