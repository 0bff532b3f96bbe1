"""Distributed tree build: one global tree over particles that live on several ranks.

The reference builds the global tree on one rank and broadcasts it
(``boxtree/distributed/__init__.py:185-203``); every rank then holds all box arrays and
fetches its local particles from the root's global arrays
(``distributed/local_tree.py:316-495``).  Here no rank ever holds all particles:

* every rank keeps its slice of the input, computes the same Morton keys (the bounding box
  is all-reduced, min/max are exact) and sorts its slice ONCE;
* the level loop runs on every rank in lock step on identical per-box data: the only
  particle-dependent inputs -- the children's (lower bound, count, non-child count) found by
  binary search in the rank's sorted keys -- are all-reduced per level (``bt_pool.xch``), so
  every rank takes the same split / level-restriction / pruning decisions and ends with the
  global tree's box arrays, bit-identical to the single-GPU build of the concatenated input;
* per-box source counts and particle bounding boxes are all-reduced once (sum, min/max);
* particles stay where they are, in tree order, until the work partition is known; then ONE
  all-to-all (:mod:`boxtree_b200.distributed.exchange`) sends every particle straight to the
  ranks whose local trees need it.

What travels: O(boxes) integers per level and three O(boxes) reductions; particles once.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

from ..tree import Tree


@dataclass(frozen=True)
class DistributedTree(Tree):
    """The global tree of a distributed build as seen by one rank.

    Box arrays (``box_centers``, ``box_child_ids``, ``box_source_starts``, ... -- everything
    indexed by box) are the GLOBAL tree's and identical on all ranks.  ``sources``,
    ``targets``, the radii, ``user_source_ids`` and ``sorted_target_ids`` cover only the
    particles this rank contributed, in tree order; ``user_source_ids`` /
    ``sorted_target_ids`` index the rank's input arrays.  ``local_box_*`` give every box's
    range in those local arrays.  Inside a box's own range the global tree order is rank-major
    (global particle ids are), so local source ``local_box_source_starts[b] + k`` of box *b*
    sits at ``box_source_starts[b] + (own sources of b on lower ranks) + k`` globally."""
    local_box_source_starts: Any = None
    local_box_source_counts_nonchild: Any = None
    local_box_source_counts_cumul: Any = None
    local_box_target_starts: Any = None
    local_box_target_counts_nonchild: Any = None
    local_box_target_counts_cumul: Any = None
    nsources_global: int = 0
    ntargets_global: int = 0
    rank: int = 0
    nranks: int = 1
    #: set by ``build_distributed_tree(..., defer_extents=True)``: the all-reduce of the particle
    #: bounding boxes is still in flight; ``pending.finish()`` (called by
    #: ``distributed_tree_setup``) completes ``box_{source,target}_bounding_box_{min,max}``
    pending: Any = None


def build_distributed_tree(actx, tree_builder, comm, particles, defer_extents=False,
                           **kwargs) -> DistributedTree:
    """Collective over *comm*: *particles* (and ``targets``, radii in *kwargs*) are this rank's
    slice of the global particle set; the global set is the concatenation in rank order.
    Accepts :class:`boxtree_b200.TreeBuilder`'s arguments except refine weights.

    *defer_extents*: return while the all-reduce of the boxes' particle extents is still in
    flight on the backend's stream (``tree.pending``); every other array is final.
    :func:`boxtree_b200.distributed.distributed_tree_setup` overlaps the reduction with the work
    partition and the colleague pass and completes the extents before anything reads them; any
    other consumer calls ``tree.pending.finish()`` first."""
    tree, _ = tree_builder(actx, particles, comm=comm, _defer_extents=bool(defer_extents), **kwargs)
    return tree
