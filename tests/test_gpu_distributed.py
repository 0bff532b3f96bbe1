"""GPU parity of the distributed setup rows (partition, masks, local tree, local traversal)
against the oracle, emulating the ranks one after the other in a single process."""
import numpy as np
import pytest
import torch

from oracle import distributed as od
from oracle.traversal import build_traversal
from oracle.tree_build import build_tree
from tests.parity_util import (bits_equal, config3_inputs, normal_particles, trav_mismatches)

pytestmark = pytest.mark.gpu


class FakeComm:
    """One emulated rank of *size*; the root's segments are computed by every emulated rank."""

    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def Get_rank(self):  # noqa: N802
        return 0            # every emulated rank holds the costs, so each acts as the root

    def Get_size(self):  # noqa: N802
        return self.size

    def scatter_rows(self, rows, root=0):
        return np.asarray(rows[self.rank])


from tests.dist_cases import CASES, box_cost  # noqa: E402


@pytest.mark.parametrize("nranks", [1, 3, 4])
@pytest.mark.parametrize("name", sorted(CASES))
def test_distributed_rows_match_oracle(actx, name, nranks):
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    from boxtree_b200 import distributed as bd
    src, tkw, vkw = CASES[name]()
    rtree = build_tree(src, **tkw)
    rtrav = build_traversal(rtree, **vkw)
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in tkw.items()}
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], **dkw)
    tg = FMMTraversalBuilder(actx, **vkw)
    trav, _ = tg(actx, tree)

    nb = rtree.nboxes
    cost = box_cost(rtree)
    # what the reference's own distributed setup produced for this case (tests/refexec)
    import json
    import os
    from tests.parity_util import digest_mismatches, distributed_rank_digests
    with open(os.path.join(os.path.dirname(__file__), "golden",
                           "refexec_distributed_digests.json")) as f:
        reference = json.load(f)[f"{name}:{nranks}"]
    want_resp, _ = od.partition_work(cost, rtree, nranks)
    assert np.array_equal(bd.get_box_ids_dfs_order(actx, tree).cpu().numpy(),
                          od.get_box_ids_dfs_order(rtree))

    got_resp = [bd.partition_work(actx, cost, trav, FakeComm(r, nranks)) for r in range(nranks)]
    for r in range(nranks):
        assert np.array_equal(got_resp[r], want_resp[r])
    want_masks = [od.get_box_masks(rtrav, want_resp[r]) for r in range(nranks)]
    got_masks = [bd.get_box_masks(actx, trav, got_resp[r]) for r in range(nranks)]
    for r in range(nranks):
        for f in ("responsible_boxes", "ancestor_boxes", "point_src_boxes", "multipole_src_boxes"):
            assert np.array_equal(getattr(got_masks[r], f).cpu().numpy(),
                                  getattr(want_masks[r], f)), (r, f)
    want_mp = np.stack([m.multipole_src_boxes for m in want_masks])
    got_mp = torch.stack([m.multipole_src_boxes for m in got_masks])

    ntgt_total = 0
    for r in range(nranks):
        wt, wsrc_idx, wtgt_idx = od.generate_local_tree(rtrav, want_resp[r], want_mp)
        gt, gsrc_idx, gtgt_idx = bd.generate_local_tree(actx, trav, got_resp[r], FakeComm(r, nranks),
                                                        multipole_masks_all_ranks=got_mp)
        assert np.array_equal(gsrc_idx.cpu().numpy(), wsrc_idx)
        assert np.array_equal(gtgt_idx.cpu().numpy(), wtgt_idx)
        ntgt_total += len(wtgt_idx)
        g = actx.to_numpy(gt)
        for f in ("box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
                  "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul",
                  "box_flags", "box_parent_ids", "box_levels", "box_child_ids"):
            a, b = np.asarray(getattr(g, f)), np.asarray(getattr(wt, f))
            assert a.dtype == b.dtype and np.array_equal(a, b), (r, f)
        for ax in range(rtree.dimensions):
            assert bits_equal(g.sources[ax], wt.sources[ax]) and bits_equal(g.targets[ax], wt.targets[ax])
        if rtree.targets_have_extent:
            assert bits_equal(g.target_radii, wt.target_radii)
        assert g.user_source_ids is None and g.sorted_target_ids is None
        for f in ("box_to_user_rank_starts", "box_to_user_rank_lists", "responsible_boxes_mask",
                  "ancestor_mask"):
            assert np.array_equal(np.asarray(getattr(g, f)), wt.extra[f]), (r, f)
        assert np.array_equal(np.asarray(g.responsible_boxes_list), want_resp[r])

        wtrav = od.generate_local_travs(wt, **vkw)
        gtrav = actx.to_numpy(bd.generate_local_travs(actx, gt, tg))
        assert not trav_mismatches(wtrav, gtrav), r
        got_digests = distributed_rank_digests(
            got_resp[r], {f: getattr(got_masks[r], f).cpu().numpy() for f in (
                "responsible_boxes", "ancestor_boxes", "point_src_boxes", "multipole_src_boxes")},
            {**{f: getattr(g, f) for f in (
                "box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
                "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul",
                "box_flags", "box_parent_ids", "box_levels", "box_child_ids",
                "box_to_user_rank_starts", "box_to_user_rank_lists", "responsible_boxes_mask",
                "ancestor_mask", "sources", "targets")},
             "target_radii": g.target_radii if rtree.targets_have_extent else None},
            gsrc_idx.cpu().numpy(), gtgt_idx.cpu().numpy(), gtrav, nb)
        assert digest_mismatches(reference[r], got_digests) == [], r
    assert ntgt_total == rtree.ntargets


@pytest.mark.parametrize("nranks", [1, 4])
@pytest.mark.parametrize("name", sorted(CASES))
def test_sharded_setup_matches_oracle(actx, name, nranks):
    """The scalable setup (no global traversal) yields the reference flow's masks, local tree
    and local traversal; only the colleague CSR is restricted to the rows that are read."""
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    from boxtree_b200 import distributed as bd
    src, tkw, vkw = CASES[name]()
    rtree = build_tree(src, **tkw)
    rtrav = build_traversal(rtree, **vkw)
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in tkw.items()}
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], **dkw)
    tg = FMMTraversalBuilder(actx, **vkw)
    nb = rtree.nboxes
    cost = (1.0 + rtree.box_source_counts_nonchild[:nb]
            + rtree.box_target_counts_nonchild[:nb]).astype(np.float64)
    want_resp, _ = od.partition_work(cost, rtree, nranks)
    want_mp = np.stack([od.get_box_masks(rtrav, want_resp[r]).multipole_src_boxes
                        for r in range(nranks)])

    class Comm(FakeComm):
        def Get_rank(self):  # noqa: N802
            return self.rank

        def allgather_tensor(self, t):
            return actx.from_numpy(want_mp)

    for r in range(nranks):
        wt, wsrc, wtgt = od.generate_local_tree(rtrav, want_resp[r], want_mp)
        wtrav = od.generate_local_travs(wt, **vkw)
        lt, ltrav, sidx, tidx = bd.sharded_setup(actx, tree, tg, Comm(r, nranks))
        g = actx.to_numpy(lt)
        assert np.array_equal(np.asarray(g.responsible_boxes_list), want_resp[r])
        assert np.array_equal(sidx.cpu().numpy(), wsrc) and np.array_equal(tidx.cpu().numpy(), wtgt)
        for f in ("box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
                  "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul",
                  "box_flags"):
            assert np.array_equal(np.asarray(getattr(g, f)), np.asarray(getattr(wt, f))), (r, f)
        for f in ("box_to_user_rank_starts", "box_to_user_rank_lists", "responsible_boxes_mask",
                  "ancestor_mask"):
            assert np.array_equal(np.asarray(getattr(g, f)), wt.extra[f]), (r, f)
        gtrav = actx.to_numpy(ltrav)
        bad = [b for b in trav_mismatches(wtrav, gtrav) if "same_level_non_well_sep" not in b]
        assert not bad, (r, bad[:8])
        # colleague rows of the boxes that are read
        need = np.nonzero(wt.extra["responsible_boxes_mask"] | wt.extra["ancestor_mask"])[0]
        ws, wl = wtrav.same_level_non_well_sep_boxes_starts, wtrav.same_level_non_well_sep_boxes_lists
        gs, gl = gtrav.same_level_non_well_sep_boxes_starts, gtrav.same_level_non_well_sep_boxes_lists
        for b in need[:: max(1, len(need) // 500)]:
            assert np.array_equal(wl[ws[b]:ws[b + 1]], gl[gs[b]:gs[b + 1]]), (r, b)


def test_distributed_setup_with_cost_model_weights(actx):
    """distributed_setup with the reference's choice of partition weights
    (distributed/__init__.py:208-230): FMMCostModel().cost_per_box with unit calibration
    parameters.  One rank: the local tree must hold every particle, the local traversal must be
    the global one."""
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    from boxtree_b200 import distributed as bd
    src = normal_particles(20000, 3, np.float64)
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], max_particles_in_box=30)
    tg = FMMTraversalBuilder(actx)
    level_orders = np.full(tree.nlevels, 4, np.int64)
    lt, ltrav, sidx, tidx, gtrav = bd.distributed_setup(actx, tree, tg, bd.SingleProcessComm(),
                                                        level_orders=level_orders)
    assert int(lt.sources[0].shape[0]) == 20000 and int(lt.targets[0].shape[0]) == 20000
    assert not trav_mismatches(actx.to_numpy(gtrav), actx.to_numpy(ltrav))


@pytest.mark.parametrize("nranks", [1, 2, 3, 5, 8, 64])
@pytest.mark.parametrize("name", sorted(CASES))
def test_default_cost_partition_cuts_on_device(actx, name, nranks):
    """``bt_dist_partition_cuts`` (one scan + one binary search per cut) against the sequential
    accumulation of ``partition.py:81-116`` on the default cost, and against the generic device
    path."""
    from boxtree_b200 import TreeBuilder
    from boxtree_b200 import distributed as bd
    from boxtree_b200.distributed.partition import (partition_segments,
                                                    partition_segments_default_cost,
                                                    partition_segments_device)
    src, tkw, _ = CASES[name]()
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in tkw.items()}
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], **dkw)
    order = bd.get_box_ids_dfs_order(actx, tree)
    got = partition_segments_default_cost(actx, tree, order, nranks)
    cost = (1.0 + tree.box_source_counts_nonchild.double()
            + tree.box_target_counts_nonchild.double())
    cost_host = cost.cpu().numpy()
    want = partition_segments(cost_host[order.cpu().numpy()], nranks,
                              total_workload=np.sum(cost_host))
    assert np.array_equal(got, want)
    assert np.array_equal(got, partition_segments_device(actx, cost, order, nranks))
