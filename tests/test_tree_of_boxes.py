"""TreeOfBoxes inputs (built by the reference's own ``boxtree/tree_of_boxes.py``, committed as
fixtures) through the traversal builders: the reference's ``test_traversal_from_tob``
(``test/test_tree_of_boxes.py:240-270``) with values checked -- known answers from the integer
geometry on the CPU, bit-for-bit parity of the CUDA builder with the oracle on the GPU."""
import numpy as np
import pytest

from oracle.traversal import build_traversal
from tests.invariants import integer_box_coords
from tests.tob_util import NAMES, as_oracle_tree, as_product_tob


def _rows(starts, lists, i):
    return set(np.asarray(lists[starts[i]:starts[i + 1]]).tolist())


@pytest.mark.parametrize("name", NAMES)
def test_oracle_traversal_of_tob_known_answers(name):
    tree = as_oracle_tree(name)
    trav = build_traversal(tree)
    lev, lo, size = integer_box_coords(tree)
    nb = tree.nboxes
    hi = lo + size[:, None]
    adj = np.all((lo[:, None, :] <= hi[None, :, :]) & (lo[None, :, :] <= hi[:, None, :]), axis=2)
    same = lev[:, None] == lev[None, :]
    coll = adj & same & ~np.eye(nb, dtype=bool)
    parent = tree.box_parent_ids
    list2 = same & ~adj & coll[parent][:, parent]
    list2[0, :] = False
    for b in range(nb):
        assert _rows(trav.same_level_non_well_sep_boxes_starts,
                     trav.same_level_non_well_sep_boxes_lists, b) == set(np.nonzero(coll[b])[0])
    tp = trav.target_or_target_parent_boxes
    assert np.array_equal(tp, np.arange(nb))             # every box of a TreeOfBoxes is a target box
    for i, b in enumerate(tp):
        assert _rows(trav.from_sep_siblings_starts, trav.from_sep_siblings_lists, i) == \
            set(np.nonzero(list2[b])[0])
    if "uniform" in name:
        # interior boxes of a uniform grid: 3^d - 1 colleagues, 6^d - 3^d list-2 entries
        d = tree.dimensions
        deepest = np.nonzero(lev == lev.max())[0]
        n = 1 << lev.max()
        interior = [b for b in deepest if np.all((lo[b] >= 2) & (lo[b] <= n - 3))]
        assert interior
        cs = trav.same_level_non_well_sep_boxes_starts
        s2 = trav.from_sep_siblings_starts
        for b in interior:
            assert cs[b + 1] - cs[b] == 3 ** d - 1
            assert s2[b + 1] - s2[b] == 6 ** d - 3 ** d
        # every box is a source box, so list 1 of a deepest box holds the box, its neighbours
        # and every ancestor-level box touching it
        l1s, l1l = trav.neighbor_source_boxes_starts, trav.neighbor_source_boxes_lists
        for i, b in enumerate(trav.target_boxes):
            if lev[b] == lev.max():
                want = set(np.nonzero(adj[b])[0])
                assert _rows(l1s, l1l, i) == want


@pytest.mark.gpu
@pytest.mark.parametrize("n_away", [1, 2])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_traversal_of_tob_matches_oracle(actx, name, n_away):
    from boxtree_b200 import FMMTraversalBuilder
    from tests.parity_util import trav_mismatches
    want = build_traversal(as_oracle_tree(name), well_sep_is_n_away=n_away)
    tob = as_product_tob(name)                           # host numpy, as the reference hands it over
    got, _ = FMMTraversalBuilder(actx, well_sep_is_n_away=n_away)(actx, tob)
    assert not trav_mismatches(want, actx.to_numpy(got))
