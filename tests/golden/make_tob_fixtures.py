"""Generate ``tests/golden/tob_*.npz``: TreeOfBoxes inputs built by the REFERENCE's own host
code (``/root/reference/boxtree/tree_of_boxes.py``, pure numpy), executed in place in this
container.  ``boxtree.tree`` needs pyopencl, so a minimal stand-in with the two names
``tree_of_boxes.py`` imports (``TreeOfBoxes`` as a plain dataclass with the reference's field
names, ``box_flags_enum`` with the reference's bit values, ``boxtree/tree.py:109-145,154-289``)
is put in its place; nothing of the reference is copied into the repository.

    python tests/golden/make_tob_fixtures.py
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from dataclasses import dataclass
from typing import Any

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"


def load_reference_tree_of_boxes():
    pkg = types.ModuleType("boxtree")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "boxtree")]
    sys.modules["boxtree"] = pkg

    class box_flags_enum:  # noqa: N801
        dtype = np.dtype(np.uint8)
        IS_SOURCE_BOX = 1 << 0
        IS_TARGET_BOX = 1 << 1
        IS_SOURCE_OR_TARGET_BOX = 3
        HAS_SOURCE_CHILD_BOXES = 1 << 2
        HAS_TARGET_CHILD_BOXES = 1 << 3
        HAS_SOURCE_OR_TARGET_CHILD_BOXES = 12
        IS_LEAF_BOX = 1 << 4

    @dataclass(frozen=True)
    class TreeOfBoxes:
        root_extent: Any
        box_centers: Any
        box_parent_ids: Any
        box_child_ids: Any
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
        def dimensions(self):
            return self.box_centers.shape[0]

        @property
        def nboxes(self):
            return self.box_centers.shape[1]

        @property
        def nlevels(self):
            return int(max(self.box_levels)) + 1

        @property
        def leaf_boxes(self):
            return np.nonzero(self.box_flags & box_flags_enum.IS_LEAF_BOX)[0]

    tree_mod = types.ModuleType("boxtree.tree")
    tree_mod.TreeOfBoxes = TreeOfBoxes
    tree_mod.box_flags_enum = box_flags_enum
    sys.modules["boxtree.tree"] = tree_mod
    try:
        import pytools  # noqa: F401
    except ImportError:
        stub = types.ModuleType("pytools")

        def single_valued(iterable, equality_pred=None):
            items = list(iterable)
            return items[0]
        stub.single_valued = single_valued
        sys.modules["pytools"] = stub
    mod = importlib.import_module("boxtree.tree_of_boxes")
    assert mod.__file__.startswith(REFERENCE_ROOT)
    return mod


def save(name, tob):
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        root_extent=np.asarray(tob.root_extent), box_centers=tob.box_centers,
        box_parent_ids=tob.box_parent_ids, box_child_ids=tob.box_child_ids,
        box_levels=tob.box_levels, box_flags=tob.box_flags)
    print(name, "nboxes", tob.nboxes, "nlevels", tob.nlevels, "levels dtype", tob.box_levels.dtype,
          "root parent", tob.box_parent_ids[0])


def main():
    tb = load_reference_tree_of_boxes()
    # test/test_tree_of_boxes.py:240-270: 2-D, 3 uniform refinements of [-pi/2, pi/2]^2
    radius = np.pi
    lower = np.zeros(2) - radius / 2
    tob = tb.make_tree_of_boxes_root((lower, lower + radius))
    for _ in range(3):
        tob = tb.uniformly_refine_tree_of_boxes(tob)
    save("tob_uniform_2d_3levels", tb._sort_boxes_by_level(tob))
    # 3-D, 3 uniform refinements of the unit cube
    tob = tb.make_tree_of_boxes_root((np.zeros(3), np.ones(3)))
    for _ in range(3):
        tob = tb.uniformly_refine_tree_of_boxes(tob)
    save("tob_uniform_3d_3levels", tb._sort_boxes_by_level(tob))
    # 2-D adaptive: refine towards the point (0.3, 0.6) five times (test_tree_of_boxes.py style)
    tob = tb.make_tree_of_boxes_root((np.zeros(2), np.ones(2)))
    tob = tb.uniformly_refine_tree_of_boxes(tob)
    pt = np.array([0.3, 0.6])
    for _ in range(5):
        flags = np.zeros(tob.nboxes, bool)
        leaves = tob.leaf_boxes
        half = tob.root_extent / 2 ** (1 + tob.box_levels[leaves].astype(np.float64))
        near = np.all(np.abs(tob.box_centers[:, leaves] - pt[:, None]) <= 1.5 * half, axis=0)
        flags[leaves[near]] = True
        tob = tb.refine_tree_of_boxes(tob, flags)
    save("tob_adaptive_2d", tb._sort_boxes_by_level(tob))


if __name__ == "__main__":
    main()
