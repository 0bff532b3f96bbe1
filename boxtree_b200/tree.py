"""Output data types of the tree build: ``box_flags_enum``, ``TreeOfBoxes``, ``Tree``.

Field names, dtypes, shapes and aliasing follow the reference
(``boxtree/tree.py:109-145`` flags, ``:154-289`` TreeOfBoxes, ``:298-686`` Tree;
construction at ``boxtree/tree_build.py:1830-1876``).  Arrays are
``torch.Tensor`` s in HBM (or numpy arrays after ``actx.to_numpy``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np


class box_flags_enum:  # noqa: N801  (reference name)
    """Bit field constants of ``Tree.box_flags`` (``boxtree/tree.py:109-145``)."""
    dtype = np.dtype(np.uint8)

    IS_SOURCE_BOX = 1 << 0
    IS_TARGET_BOX = 1 << 1
    IS_SOURCE_OR_TARGET_BOX = IS_SOURCE_BOX | IS_TARGET_BOX
    HAS_SOURCE_CHILD_BOXES = 1 << 2
    HAS_TARGET_CHILD_BOXES = 1 << 3
    HAS_SOURCE_OR_TARGET_CHILD_BOXES = HAS_SOURCE_CHILD_BOXES | HAS_TARGET_CHILD_BOXES
    IS_LEAF_BOX = 1 << 4
    HAS_CHILDREN = HAS_SOURCE_OR_TARGET_CHILD_BOXES


def _len(a) -> int:
    return int(a.shape[0])


@dataclass(frozen=True)
class TreeOfBoxes:
    """A tree of boxes without particles (``boxtree/tree.py:154-289``)."""
    root_extent: Any
    box_centers: Any            # [dim, aligned_nboxes]
    box_parent_ids: Any
    box_child_ids: Any          # [2^dim, aligned_nboxes]
    box_levels: Any
    box_flags: Any
    level_start_box_nrs: Any
    box_id_dtype: Any
    box_level_dtype: Any
    coord_dtype: Any
    sources_have_extent: bool
    targets_have_extent: bool
    extent_norm: Any
    stick_out_factor: Any
    _is_pruned: bool

    @property
    def dimensions(self) -> int:
        return int(self.box_centers.shape[0])

    @property
    def nboxes(self) -> int:
        return int(self.box_centers.shape[1])

    @property
    def aligned_nboxes(self) -> int:
        return int(self.box_child_ids.shape[-1])

    @property
    def nlevels(self) -> int:
        return int(self.box_levels.max()) + 1


@dataclass(frozen=True)
class Tree:
    """A quad/octree of particles sorted into boxes (``boxtree/tree.py:298-686``)."""
    sources_are_targets: bool
    sources_have_extent: bool
    targets_have_extent: bool

    particle_id_dtype: Any
    box_id_dtype: Any
    coord_dtype: Any
    box_level_dtype: Any

    bounding_box: Any           # (bbox_min, bbox_max) host numpy vectors
    root_extent: Any            # host numpy scalar
    stick_out_factor: Any
    extent_norm: Any

    level_start_box_nrs: Any    # [nlevels + 1]

    sources: Any                # object array of dim arrays [nsources]
    targets: Any
    source_radii: Any
    target_radii: Any

    box_source_starts: Any
    box_source_counts_nonchild: Any
    box_source_counts_cumul: Any
    box_target_starts: Any
    box_target_counts_nonchild: Any
    box_target_counts_cumul: Any

    box_parent_ids: Any
    box_child_ids: Any          # [2^dim, aligned_nboxes]
    box_centers: Any            # [dim, aligned_nboxes]
    box_levels: Any
    box_flags: Any

    user_source_ids: Any
    sorted_target_ids: Any

    box_source_bounding_box_min: Any
    box_source_bounding_box_max: Any
    box_target_bounding_box_min: Any
    box_target_bounding_box_max: Any

    _is_pruned: bool

    @property
    def dimensions(self) -> int:
        return len(self.sources)

    @property
    def nboxes(self) -> int:
        return _len(self.box_flags)

    @property
    def nsources(self) -> int:
        return _len(self.sources[0])

    @property
    def ntargets(self) -> int:
        return _len(self.targets[0])

    @property
    def nlevels(self) -> int:
        return _len(self.level_start_box_nrs) - 1

    @property
    def aligned_nboxes(self) -> int:
        return int(self.box_child_ids.shape[-1])

    # debugging aids of the reference (tree.py:633-686); numpy arrays assumed
    def get_box_extent(self, ibox):
        lev = int(self.box_levels[ibox])
        box_size = self.root_extent / (1 << lev)
        extent_low = self.box_centers[:, ibox] - 0.5 * box_size
        return extent_low, extent_low + box_size

    def indices_to_tree_target_order(self, user_indices):
        return self.sorted_target_ids[user_indices]

    def find_box_nr_for_target(self, itarget):
        crit = ((self.box_target_starts <= itarget)
                & (itarget < self.box_target_starts + self.box_target_counts_nonchild))
        return int(np.where(crit)[0])

    def find_box_nr_for_source(self, isource):
        crit = ((self.box_source_starts <= isource)
                & (isource < self.box_source_starts + self.box_source_counts_nonchild))
        return int(np.where(crit)[0])
