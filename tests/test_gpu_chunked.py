"""Large-list mode (``FMMTraversalBuilder.build_in_chunks``): the rows of the pieces are the
rows of the global traversal, every row in exactly one piece."""
import numpy as np
import pytest

from tests.dist_cases import CASES

pytestmark = pytest.mark.gpu


def _rows(starts, lists, boxes):
    return {int(b): lists[starts[i]:starts[i + 1]].tolist() for i, b in enumerate(boxes)}


def _collect(trav, nlevels):
    out = {}
    tb, tp = trav.target_boxes, trav.target_or_target_parent_boxes
    out["l1"] = _rows(trav.neighbor_source_boxes_starts, trav.neighbor_source_boxes_lists, tb)
    out["l2"] = _rows(trav.from_sep_siblings_starts, trav.from_sep_siblings_lists, tp)
    out["l4"] = _rows(trav.from_sep_bigger_starts, trav.from_sep_bigger_lists, tp)
    if trav.from_sep_close_smaller_starts is not None:
        out["l3c"] = _rows(trav.from_sep_close_smaller_starts, trav.from_sep_close_smaller_lists, tb)
        out["l4c"] = _rows(trav.from_sep_close_bigger_starts, trav.from_sep_close_bigger_lists, tb)
    for lev in range(nlevels):
        bl = trav.from_sep_smaller_by_level[lev]
        out[f"l3[{lev}]"] = _rows(bl.starts, bl.lists, trav.target_boxes_sep_smaller_by_source_level[lev])
    return out


@pytest.mark.parametrize("nchunks", [2, 5])
@pytest.mark.parametrize("name", sorted(CASES))
def test_chunked_traversal_rows(actx, name, nchunks):
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    src, tkw, vkw = CASES[name]()
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in tkw.items()}
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], **dkw)
    tg = FMMTraversalBuilder(actx, **vkw)
    whole = _collect(actx.to_numpy(tg(actx, tree)[0]), tree.nlevels)
    pieces = tg.build_in_chunks(actx, tree, nchunks=nchunks)
    assert len(pieces) == nchunks
    seen = {k: {} for k in whole}
    all_boxes = []
    for boxes, piece in pieces:
        all_boxes.append(boxes.cpu().numpy())
        mine = set(all_boxes[-1].tolist())
        for k, rows in _collect(actx.to_numpy(piece), tree.nlevels).items():
            for b, row in rows.items():
                assert b in mine, (k, b)            # only rows of the segment's boxes
                assert b not in seen[k], (k, b)     # every row in exactly one piece
                seen[k][b] = row
    assert np.array_equal(np.sort(np.concatenate(all_boxes)), np.arange(tree.nboxes))
    for k in whole:
        # empty rows of compressed lists are absent on both sides
        assert {b: r for b, r in seen[k].items() if r or not k.startswith("l3[")} == \
            {b: r for b, r in whole[k].items() if r or not k.startswith("l3[")}, k


def test_chunked_traversal_auto_single_piece(actx):
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    src, tkw, vkw = CASES["points"]()
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], **tkw)
    pieces = FMMTraversalBuilder(actx, **vkw).build_in_chunks(actx, tree)
    assert len(pieces) == 1
