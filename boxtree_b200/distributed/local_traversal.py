"""``boxtree/distributed/local_traversal.py:34-62`` on B200."""
from __future__ import annotations


def generate_local_travs(actx, local_tree, traversal_builder, merge_close_lists=False):
    """Traversal of a rank's local tree: multipole formation and upward propagation are
    restricted to the rank's responsible boxes / their ancestors through the builder's
    ``source_boxes_mask`` / ``source_parent_boxes_mask``."""
    local_trav, _ = traversal_builder(
        actx, local_tree,
        source_boxes_mask=local_tree.responsible_boxes_mask,
        source_parent_boxes_mask=local_tree.ancestor_mask)
    if merge_close_lists and local_tree.targets_have_extent:
        local_trav = local_trav.merge_close_lists(actx)
    return local_trav
