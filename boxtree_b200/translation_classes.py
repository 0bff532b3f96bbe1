"""``TranslationClassesBuilder`` / ``RotationClassesBuilder`` (``boxtree/translation_classes.py:
60-438``, ``boxtree/rotation_classes.py:52-200``): every list-2 pair (target box, source box) is
classified by its normalised centre-to-centre vector (translation class) or by the angle of that
vector with the last axis (rotation class).  Per-entry work is elementwise torch on the array
context's stream; the class bookkeeping follows the reference's host loops (numpy scalars in the
coordinate dtype).  Consumer-side utilities of the traversal.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from .array_context import TorchArrayContext


@dataclass(frozen=True)
class TranslationClassesInfo:
    """``translation_classes.py:199-245``."""
    traversal: Any
    from_sep_siblings_translation_classes: Any
    from_sep_siblings_translation_class_to_distance_vector: Any
    from_sep_siblings_translation_classes_level_starts: Any

    @property
    def nfrom_sep_siblings_translation_classes(self):
        return len(self.from_sep_siblings_translation_class_to_distance_vector)


@dataclass(frozen=True)
class RotationClassesInfo:
    """``rotation_classes.py:52-86``."""
    from_sep_siblings_rotation_classes: Any
    from_sep_siblings_rotation_class_to_angle: Any

    @property
    def nfrom_sep_siblings_rotation_classes(self):
        return len(self.from_sep_siblings_rotation_class_to_angle)


class TranslationClassesBuilder:
    def __init__(self, array_context: TorchArrayContext) -> None:
        assert isinstance(array_context, TorchArrayContext)
        self._setup_actx = array_context

    @staticmethod
    def ntranslation_classes_per_level(well_sep_is_n_away: int, dimensions: int) -> int:
        return (4 * well_sep_is_n_away + 3) ** dimensions

    def translation_class_to_normalized_vector(self, well_sep_is_n_away, dimensions, cls):
        """``translation_classes.py:302-318``: the inverse of ``get_translation_class``."""
        assert 0 <= cls < self.ntranslation_classes_per_level(well_sep_is_n_away, dimensions)
        result = np.zeros(dimensions, dtype=np.int32)
        shift = 2 * well_sep_is_n_away + 1
        base = 4 * well_sep_is_n_away + 3
        for i in range(dimensions):
            result[i] = cls % base - shift
            cls //= base
        return result

    def compute_translation_classes(self, actx, trav, tree, wait_for, is_translation_per_level):
        """``TRANSLATION_CLASS_FINDER_TEMPLATE`` (``translation_classes.py:60-196, 320-366``).
        Returns ``(event, translation_class_is_used, translation_classes_lists)``."""
        assert isinstance(actx, TorchArrayContext)
        n = int(trav.well_sep_is_n_away)
        dims = int(tree.dimensions)
        nper = self.ntranslation_classes_per_level(n, dims)
        if not nper <= 1 + np.iinfo(np.int32).max:
            raise ValueError("would overflow")
        ncls = nper * (int(tree.nlevels) if is_translation_per_level else 1)
        from . import _cabi
        from ._cabi import check, dptr
        lib = _cabi.load()
        with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
            starts = trav.from_sep_siblings_starts
            src = trav.from_sep_siblings_lists
            npairs = int(src.shape[0])
            nrows = int(starts.shape[0]) - 1
            cls = actx.empty(max(npairs, 1), np.int32)
            used = actx.zeros(ncls, np.int32)
            err = actx.zeros(1, np.int32)
            check(lib.bt_translation_classes(
                _cabi.dtype_code(tree.coord_dtype), dims, nrows,
                dptr(trav.target_or_target_parent_boxes), dptr(starts), dptr(src),
                dptr(tree.box_centers), int(tree.aligned_nboxes), dptr(tree.box_levels),
                float(tree.root_extent), n, int(bool(is_translation_per_level)), nper, npairs,
                dptr(cls), dptr(used), dptr(err), actx.stream_handle), "bt_translation_classes")
            if int(err.item()):
                raise ValueError("could not compute translation classes")
            cls = cls[:npairs]
            evt = torch.cuda.Event()
            evt.record(actx.stream)
        return evt, used, cls

    def __call__(self, actx, trav, tree, wait_for=None, is_translation_per_level=True):
        """``translation_classes.py:368-436``: ``(TranslationClassesInfo, event)``."""
        evt, used_dev, cls_lists = self.compute_translation_classes(
            actx, trav, tree, wait_for, is_translation_per_level)
        n = int(trav.well_sep_is_n_away)
        dims = int(tree.dimensions)
        used = used_dev.cpu().numpy()
        coord_dtype = np.dtype(tree.coord_dtype)
        root_extent = coord_dtype.type(tree.root_extent)
        used_map = np.full(len(used), -1, dtype=np.int32)
        # the reference leaves the columns past the used classes uninitialised; zeros here
        distances = np.zeros((dims, len(used)), dtype=coord_dtype)
        nper = self.ntranslation_classes_per_level(n, dims)
        nlevels = int(tree.nlevels)
        level_starts = np.empty(nlevels + 1, dtype=np.int32)
        count = 0
        prev_level = -1
        for i, u in enumerate(used):
            cls_without_level = i % nper
            level = i // nper
            if prev_level != level:
                level_starts[level] = count
                prev_level = level
            if not u:
                continue
            used_map[i] = count
            unit_vector = self.translation_class_to_normalized_vector(n, dims, cls_without_level)
            distances[:, count] = unit_vector * root_extent / (1 << level)
            count += 1
        if not is_translation_per_level:
            level_starts[1:] = count          # one level's worth of classes: the loop sets entry 0 only
        level_starts[nlevels] = count
        with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
            from . import _cabi
            from ._cabi import check, dptr
            used_map_dev = actx.from_numpy(used_map)
            check(_cabi.load().bt_remap_classes(int(cls_lists.shape[0]), dptr(used_map_dev),
                                                dptr(cls_lists), actx.stream_handle),
                  "bt_remap_classes")
            info = TranslationClassesInfo(
                traversal=trav, from_sep_siblings_translation_classes=cls_lists,
                from_sep_siblings_translation_class_to_distance_vector=actx.from_numpy(distances),
                from_sep_siblings_translation_classes_level_starts=actx.from_numpy(level_starts))
        return actx.freeze(info), evt


class RotationClassesBuilder:
    def __init__(self, array_context: TorchArrayContext) -> None:
        assert isinstance(array_context, TorchArrayContext)
        self._setup_actx = array_context
        self.tcb = TranslationClassesBuilder(array_context)

    @staticmethod
    def vec_gcd(vec) -> int:
        result = abs(int(vec[0]))
        for elem in vec[1:]:
            result = math.gcd(result, abs(int(elem)))
        return result

    def compute_rotation_classes(self, well_sep_is_n_away, dimensions, used_translation_classes):
        """``rotation_classes.py:114-163``."""
        angle_to_rot_class = {}
        angles = []
        nper = self.tcb.ntranslation_classes_per_level(well_sep_is_n_away, dimensions)
        translation_class_to_rot_class = np.full(nper, -1, dtype=np.int32)
        for cls in used_translation_classes:
            vec = self.tcb.translation_class_to_normalized_vector(well_sep_is_n_away, dimensions, cls)
            vec //= self.vec_gcd(vec)
            norm = np.linalg.norm(vec)
            assert norm != 0
            angle = np.arccos(vec[-1] / norm)
            if angle in angle_to_rot_class:
                rot_class = angle_to_rot_class[angle]
            else:
                rot_class = len(angles)
                angle_to_rot_class[angle] = rot_class
                angles.append(angle)
            translation_class_to_rot_class[cls] = rot_class
        return translation_class_to_rot_class, angles

    def __call__(self, actx, trav, tree, wait_for=None):
        """``rotation_classes.py:165-199``: ``(RotationClassesInfo, event)``."""
        evt, used, cls_lists = self.tcb.compute_translation_classes(actx, trav, tree, wait_for, False)
        d, n = int(tree.dimensions), int(trav.well_sep_is_n_away)
        used_classes = np.flatnonzero(used.cpu().numpy())
        cls_to_rot, angles = self.compute_rotation_classes(n, d, used_classes)
        assert len(angles) <= 2 ** (d - 1) * (2 * n + 1) ** d
        with torch.cuda.stream(actx.stream):
            rot_lists = actx.from_numpy(cls_to_rot)[cls_lists.long()]
            info = RotationClassesInfo(
                from_sep_siblings_rotation_classes=rot_lists,
                from_sep_siblings_rotation_class_to_angle=actx.from_numpy(np.array(angles)))
        return actx.freeze(info), evt
