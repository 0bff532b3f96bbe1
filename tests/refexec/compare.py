"""Comparisons between what the reference itself produced (``tests/refexec/run.py``) and the
oracle's restatement, shared by ``tests/test_refexec.py`` and ``tests/refexec/fuzz.py``."""
from __future__ import annotations

import numpy as np


def oracle_distributed_digests(src, tkw, vkw, nranks):
    """Per-rank digests of the oracle's distributed setup (``oracle/distributed.py``) for the cost
    vector of ``tests.dist_cases.box_cost``."""
    from oracle import distributed as od
    from oracle.traversal import build_traversal
    from oracle.tree_build import build_tree
    from tests.dist_cases import box_cost
    from tests.parity_util import DIST_LOCAL_TREE_FIELDS, DIST_MASK_FIELDS, distributed_rank_digests
    tree = build_tree(src, **tkw)
    trav = build_traversal(tree, **vkw)
    resp, _ = od.partition_work(box_cost(tree), tree, nranks)
    masks = [od.get_box_masks(trav, resp[r]) for r in range(nranks)]
    mp = np.stack([m.multipole_src_boxes for m in masks])
    out = []
    for r in range(nranks):
        lt, src_idx, tgt_idx = od.generate_local_tree(trav, resp[r], mp)
        fields = {f: (lt.extra[f] if f in lt.extra else getattr(lt, f))
                  for f in DIST_LOCAL_TREE_FIELDS}
        fields.update(sources=lt.sources, targets=lt.targets,
                      target_radii=lt.target_radii if tree.targets_have_extent else None)
        out.append(distributed_rank_digests(
            resp[r], {f: getattr(masks[r], f) for f in DIST_MASK_FIELDS}, fields, src_idx, tgt_idx,
            od.generate_local_travs(lt, **vkw), tree.nboxes))
    return out


def reference_distributed_digests(src, tkw, vkw, nranks):
    """The same digests from the reference's own distributed setup (one thread per rank)."""
    from refexec.run import reference_distributed_setup
    from tests.dist_cases import box_cost
    from tests.parity_util import distributed_rank_digests
    tree, _, ranks = reference_distributed_setup(src, tkw, vkw, nranks, box_cost)
    return [distributed_rank_digests(r["responsible_boxes_list"], r["masks"], r["local_tree"],
                                     r["src_idx"], r["tgt_idx"], r["local_trav"], tree.nboxes)
            for r in ranks]


def distributed_mismatches(src, tkw, vkw, nranks):
    from tests.parity_util import digest_mismatches
    want = reference_distributed_digests(src, tkw, vkw, nranks)
    got = oracle_distributed_digests(src, tkw, vkw, nranks)
    bad = []
    for r in range(nranks):
        bad += [f"rank{r}.{k}" for k in digest_mismatches(want[r], got[r])]
    return bad


def area_query_mismatches(tree, ball_centers, ball_radii):
    """``boxtree/area_query.py`` (peer lists, area query, leaves-to-balls, space invaders) executed
    by the reference vs ``oracle/traversal.py``."""
    from oracle import traversal as ot
    from refexec.run import reference_area_queries
    ref = reference_area_queries(tree, ball_centers, ball_radii)
    got = dict(zip(("peer_list_starts", "peer_lists"), ot.find_peer_lists(tree)))
    got.update(zip(("leaves_near_ball_starts", "leaves_near_ball_lists"),
                   ot.area_query(tree, ball_centers, ball_radii)))
    got.update(zip(("balls_near_box_starts", "balls_near_box_lists"),
                   ot.leaves_to_balls(tree, ball_centers, ball_radii)))
    got["outer_space_invader_dists"] = ot.space_invader_query(tree, ball_centers, ball_radii)
    bad = []
    for k, v in ref.items():
        g = np.ascontiguousarray(got[k])
        v = np.ascontiguousarray(v)
        if v.dtype != g.dtype or v.shape != g.shape or not np.array_equal(
                v.view(np.uint8), g.view(np.uint8)):
            bad.append("area_query." + k)
    return bad
