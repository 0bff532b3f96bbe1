"""CPU restatement of ``ParticleListFilter`` (test infrastructure only): the loops of
``/root/reference/boxtree/tree.py:1062-1081`` (``generate``) and
``boxtree/tree_build_kernels.py:1954-2021`` written out in numpy/Python."""
from __future__ import annotations

import numpy as np


def filter_target_lists_in_user_order(tree, flags):
    nboxes, ntargets = tree.nboxes, tree.ntargets
    user_target_ids = np.zeros(ntargets, np.int32)
    user_target_ids[tree.sorted_target_ids] = np.arange(ntargets, dtype=np.int32)   # :1108-1111
    starts = np.zeros(nboxes + 1, np.int32)
    lists = []
    for i in range(nboxes):                                                          # :1062-1081
        s, c = int(tree.box_target_starts[i]), int(tree.box_target_counts_nonchild[i])
        ids = user_target_ids[s:s + c]
        ids = ids[flags[ids] != 0]
        lists.append(ids)
        starts[i + 1] = starts[i] + len(ids)
    lists = np.concatenate(lists) if lists else np.zeros(0, np.int32)
    return len(lists), starts, lists.astype(np.int32)


def filter_target_lists_in_tree_order(tree, flags):
    nboxes, ntargets = tree.nboxes, tree.ntargets
    tree_order_flags = np.zeros(ntargets, np.int8)
    tree_order_flags[tree.sorted_target_ids] = flags                                 # :1173-1174
    f = (tree_order_flags != 0).astype(np.int32)
    item = np.cumsum(f, dtype=np.int32)
    prev_item = item - f
    filtered_from_unfiltered = prev_item                                             # kernels :1968
    unfiltered_from_filtered = np.nonzero(f)[0].astype(np.int32)                     # :1969-1970
    nfiltered = int(item[-1]) if ntargets else 0
    bstart = np.zeros(nboxes, np.int32)
    bcount = np.zeros(nboxes, np.int32)
    for i in range(nboxes):                                                          # :1990-2019
        us, uc = int(tree.box_target_starts[i]), int(tree.box_target_counts_nonchild[i])
        fs = filtered_from_unfiltered[us] if us < ntargets else nfiltered
        bstart[i] = fs
        if uc > 0:
            upl = us + uc
            fpl = filtered_from_unfiltered[upl] if upl < ntargets else nfiltered
            bcount[i] = fpl - fs
    targets = [np.asarray(t)[unfiltered_from_filtered] for t in tree.targets]
    return nfiltered, bstart, bcount, targets, unfiltered_from_filtered


def link_point_sources(tree, point_source_starts, point_sources):
    """``/root/reference/boxtree/tree.py:773-955`` with the kernels of
    ``boxtree/tree_build_kernels.py:1872-1950`` as plain loops (every source owns >= 1 point)."""
    nsources, nboxes = tree.nsources, tree.nboxes
    pss = np.asarray(point_source_starts, np.int64)
    usi = tree.user_source_ids
    tos = np.zeros(nsources, np.int32)
    toc = np.zeros(nsources, np.int32)
    run = 0
    for i in range(nsources):                                    # SOURCE_SCAN_TPL
        c = int(pss[usi[i] + 1] - pss[usi[i]])
        tos[i], toc[i] = run, c
        run += c
    npoint = run
    ids = np.ones(npoint, np.int32)                              # tree.py:838-890
    bnd = np.zeros(npoint, np.int8)
    for i in range(nsources):
        ids[tos[i]] = pss[usi[i]]
        bnd[tos[i]] = 1
    for j in range(1, npoint):                                   # segmented inclusive scan
        if not bnd[j]:
            ids[j] = ids[j - 1] + ids[j]
    pts = [np.asarray(p)[ids] for p in point_sources]
    bstart = np.zeros(nboxes, np.int32)
    bnon = np.zeros(nboxes, np.int32)
    bcum = np.zeros(nboxes, np.int32)
    for ibox in range(nboxes):                                   # BOX_POINT_SOURCES
        s_start = int(tree.box_source_starts[ibox])
        ps_start = int(tos[s_start]) if s_start < nsources else npoint
        bstart[ibox] = ps_start
        for out, cnt in ((bnon, tree.box_source_counts_nonchild), (bcum, tree.box_source_counts_cumul)):
            s_count = int(cnt[ibox])
            if s_count:
                last = s_start + s_count - 1
                out[ibox] = tos[last] + toc[last] - ps_start
    return dict(npoint_sources=npoint, point_source_starts=tos, point_source_counts=toc,
                point_sources=pts, user_point_source_ids=ids, box_point_source_starts=bstart,
                box_point_source_counts_nonchild=bnon, box_point_source_counts_cumul=bcum)
