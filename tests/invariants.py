"""Structural checks restated from the reference's own tests (numpy, vectorised where cheap)
plus an independent brute-force list checker on integer box coordinates.

* ``check_tree``       -- ``test/test_tree.py:86-226`` (run_build_test)
* ``check_traversal``  -- ``test/test_traversal.py:58-267`` (test_tree_connectivity)
* ``brute_force_lists``-- independent O(nboxes^2) definition of colleagues / lists 1-4 for
  point-particle trees with well_sep_is_n_away = 1 (guards against the oracle and the CUDA
  code sharing a misreading of the reference).
"""
from __future__ import annotations

import numpy as np

IS_SOURCE, IS_TARGET, HAS_SRC_CHILD, HAS_TGT_CHILD = 1, 2, 4, 8


def check_tree(tree, unsorted_sources, max_particles_in_box=None, refine_weights=None,
               max_leaf_refine_weight=None, unsorted_targets=None):
    dtype = np.dtype(tree.coord_dtype)
    tol = 1e-4 if dtype == np.float32 else 1e-12
    sorted_particles = np.array(list(tree.sources))
    unsorted = np.array(list(unsorted_sources))
    assert np.all(sorted_particles == unsorted[:, tree.user_source_ids])
    if unsorted_targets is not None:
        tg = np.array(list(tree.targets))
        ut = np.array(list(unsorted_targets))
        assert np.all(tg[:, tree.sorted_target_ids] == ut)
    nb = tree.nboxes
    scaled_tol = tol * tree.root_extent
    levels = tree.box_levels[:nb].astype(np.int64)
    box_size = tree.root_extent / (2.0 ** levels)
    centers = tree.box_centers[:, :nb]
    lo = centers - 0.5 * box_size
    hi = lo + box_size
    flags = tree.box_flags[:nb]
    occupied = (flags & (IS_SOURCE | IS_TARGET)) != 0
    bmin, bmax = tree.bounding_box
    assert np.all(lo[:, occupied] >= bmin[:, None] - scaled_tol)
    assert np.all(hi[:, occupied] <= bmax[:, None] + scaled_tol)
    if not (tree.sources_have_extent or tree.targets_have_extent):
        for mn, mx in ((tree.box_source_bounding_box_min, tree.box_source_bounding_box_max),
                       (tree.box_target_bounding_box_min, tree.box_target_bounding_box_max)):
            mn, mx = mn[:, :nb], mx[:, :nb]
            assert np.all((lo - scaled_tol <= mn)[:, occupied])
            assert np.all((mn - scaled_tol <= centers)[:, occupied])
            assert np.all((mx - scaled_tol <= hi)[:, occupied])
            assert np.all((centers - scaled_tol <= mx)[:, occupied])
    children = tree.box_child_ids[:, :nb]
    kid_sum = np.where(children != 0, tree.box_source_counts_cumul[children], 0).sum(axis=0)
    ok = tree.box_source_counts_nonchild[:nb] + kid_sum == tree.box_source_counts_cumul[:nb]
    assert np.all(ok[occupied])
    # every particle of a box lies inside the box
    starts = tree.box_source_starts[:nb].astype(np.int64)
    counts = tree.box_source_counts_cumul[:nb].astype(np.int64)
    if not tree.sources_have_extent:
        for ibox in np.nonzero(occupied)[0]:
            p = sorted_particles[:, starts[ibox]:starts[ibox] + counts[ibox]]
            assert np.all((p < hi[:, ibox, None] + scaled_tol) & (lo[:, ibox, None] - scaled_tol <= p)), ibox
    leaf = occupied & ((flags & (HAS_SRC_CHILD | HAS_TGT_CHILD)) == 0)
    if max_particles_in_box is not None and unsorted_targets is None:
        assert np.all(counts[leaf] <= max_particles_in_box)
    elif refine_weights is not None and unsorted_targets is None:
        w = np.concatenate([[0], np.cumsum(refine_weights[tree.user_source_ids], dtype=np.int64)])
        assert np.all((w[starts + counts] - w[starts])[leaf] <= max_leaf_refine_weight)


def check_traversal(tree, trav, sources_are_targets):
    nb = tree.nboxes
    levels = tree.box_levels[:nb].astype(np.int64)
    parents = tree.box_parent_ids[:nb]
    children = tree.box_child_ids[:, :nb].T
    centers = tree.box_centers[:, :nb].T
    ids = np.arange(1, nb)
    assert np.all(levels[parents[ids]] + 1 == levels[ids])
    assert np.all((children[parents[ids]] == ids[:, None]).any(axis=1))

    def rows(starts, lists):
        starts = np.asarray(starts, np.int64)
        return np.repeat(np.arange(len(starts) - 1), np.diff(starts)), np.asarray(lists)

    # list 1 consists of leaves and contains the box itself (points, sources == targets)
    r, l1 = rows(trav.neighbor_source_boxes_starts, trav.neighbor_source_boxes_lists)
    if not tree.targets_have_extent:
        assert np.all(children[l1] == 0)
    if sources_are_targets:
        has_self = np.zeros(len(trav.target_boxes), bool)
        has_self[r[l1 == trav.target_boxes[r]]] = True
        assert np.all(has_self)
    # list 2: same level and farther than 2.5 box radii
    r, l2 = rows(trav.from_sep_siblings_starts, trav.from_sep_siblings_lists)
    tgt = trav.target_or_target_parent_boxes[r]
    assert np.all(levels[l2] == levels[tgt])
    mindist = 2.5 * 0.5 * 2.0 ** -levels[tgt] * tree.root_extent
    assert np.all(np.linalg.norm(centers[l2] - centers[tgt], axis=1) > mindist)
    # lists 3 / 4 level relations
    pairs3 = set()
    for lev, ssn in enumerate(trav.from_sep_smaller_by_level):
        tb = trav.target_boxes_sep_smaller_by_source_level[lev]
        r, l3 = rows(ssn.starts, ssn.lists)
        assert np.all(levels[tb[r]] < levels[l3])
        assert np.all(levels[l3] == lev)
        pairs3.update(zip(tb[r].tolist(), l3.tolist()))
    r, l4 = rows(trav.from_sep_bigger_starts, trav.from_sep_bigger_lists)
    tgt4 = trav.target_or_target_parent_boxes[r]
    assert np.all(levels[tgt4] > levels[l4])
    if sources_are_targets and not tree.targets_have_extent:
        # list 3 and list 4 are duals (test_traversal.py:141-218)
        assert np.all(trav.target_or_target_parent_boxes == np.arange(nb))
        pairs4 = set(zip(l4.tolist(), tgt4.tolist()))
        assert pairs3 == pairs4
    for name, ref_array in [("level_start_source_box_nrs", trav.source_boxes),
                            ("level_start_source_parent_box_nrs", trav.source_parent_boxes),
                            ("level_start_target_box_nrs", trav.target_boxes),
                            ("level_start_target_or_target_parent_box_nrs",
                             trav.target_or_target_parent_boxes)]:
        ls = getattr(trav, name)
        for lev in range(tree.nlevels):
            assert np.all(levels[ref_array[ls[lev]:ls[lev + 1]]] == lev), name


def integer_box_coords(tree):
    """(level, integer coordinates at the finest level) from the child links only."""
    nb = tree.nboxes
    d = tree.dimensions
    L = tree.nlevels - 1
    lev = np.zeros(nb, np.int64)
    lo = np.zeros((nb, d), np.int64)
    order = np.argsort(tree.box_levels[:nb], kind="stable")
    for b in order:
        for m in range(2 ** d):
            c = tree.box_child_ids[m, b]
            if c:
                lev[c] = lev[b] + 1
                half = 1 << (L - lev[c])
                bits = np.array([(m >> (d - 1 - a)) & 1 for a in range(d)])
                lo[c] = lo[b] + bits * half
    size = 1 << (L - lev)
    return lev, lo, size


def brute_force_lists(tree):
    """Set-valued colleagues and lists 1-4 (points, sources are targets, n_away = 1)."""
    nb = tree.nboxes
    lev, lo, size = integer_box_coords(tree)
    hi = lo + size[:, None]
    # closed boxes intersect in every axis
    adj = np.all((lo[:, None, :] <= hi[None, :, :]) & (lo[None, :, :] <= hi[:, None, :]), axis=2)
    parent = tree.box_parent_ids[:nb]
    leaf = np.all(tree.box_child_ids[:, :nb] == 0, axis=0)
    same = lev[:, None] == lev[None, :]
    coll = adj & same & ~np.eye(nb, dtype=bool)
    list1 = adj & leaf[None, :] & leaf[:, None]
    list2 = same & ~adj & coll[parent][:, parent]
    list2[0, :] = False
    # ancestors matrix anc[b, a] = a is a (strict) ancestor of b
    anc = np.zeros((nb, nb), bool)
    for b in np.argsort(lev, kind="stable")[1:]:
        anc[b] = anc[parent[b]]
        anc[b, parent[b]] = True
    # list 3: w below a colleague of b, parent(w) adjacent to b, w not adjacent
    desc_of_coll = (coll.astype(np.int32) @ anc.T.astype(np.int32)) > 0     # [b, w]
    list3 = leaf[:, None] & desc_of_coll & adj[:, parent] & ~adj
    # list 4: leaf c, colleague of an ancestor of b, not adjacent to b, adjacent to parent(b)
    coll_of_anc = (anc.astype(np.int32) @ coll.astype(np.int32)) > 0        # [b, c]
    list4 = coll_of_anc & leaf[None, :] & ~adj & adj[parent]
    list4[0, :] = False
    return {"coll": coll, "list1": list1, "list2": list2, "list3": list3, "list4": list4,
            "leaf": leaf}


def check_against_brute_force(tree, trav):
    bf = brute_force_lists(tree)
    nb = tree.nboxes

    def to_matrix(row_boxes, starts, lists):
        m = np.zeros((nb, nb), bool)
        starts = np.asarray(starts, np.int64)
        r = np.repeat(np.asarray(row_boxes), np.diff(starts))
        m[r, np.asarray(lists)] = True
        assert len(set(zip(r.tolist(), np.asarray(lists).tolist()))) == len(lists), "duplicates"
        return m

    allb = np.arange(nb)
    assert np.array_equal(to_matrix(allb, trav.same_level_non_well_sep_boxes_starts,
                                    trav.same_level_non_well_sep_boxes_lists), bf["coll"])
    assert np.array_equal(to_matrix(trav.target_boxes, trav.neighbor_source_boxes_starts,
                                    trav.neighbor_source_boxes_lists), bf["list1"])
    assert np.array_equal(to_matrix(trav.target_or_target_parent_boxes,
                                    trav.from_sep_siblings_starts,
                                    trav.from_sep_siblings_lists), bf["list2"])
    m3 = np.zeros((nb, nb), bool)
    for lev, ssn in enumerate(trav.from_sep_smaller_by_level):
        m3 |= to_matrix(trav.target_boxes_sep_smaller_by_source_level[lev], ssn.starts, ssn.lists)
    assert np.array_equal(m3, bf["list3"])
    assert np.array_equal(to_matrix(trav.target_or_target_parent_boxes,
                                    trav.from_sep_bigger_starts, trav.from_sep_bigger_lists),
                          bf["list4"])
