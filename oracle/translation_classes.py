"""CPU restatement of the translation / rotation class builders (test infrastructure only):
``/root/reference/boxtree/translation_classes.py:60-196, 302-436`` and
``boxtree/rotation_classes.py:100-199`` with the per-entry kernel written as a Python loop."""
from __future__ import annotations

import math

import numpy as np


def _nper(n, d):
    return (4 * n + 3) ** d


def _class_to_vector(n, d, cls):                                   # :302-318
    result = np.zeros(d, dtype=np.int32)
    shift, base = 2 * n + 1, 4 * n + 3
    for i in range(d):
        result[i] = cls % base - shift
        cls //= base
    return result


def compute_translation_classes(trav, tree, per_level):
    n, d = trav.well_sep_is_n_away, tree.dimensions
    ct = np.dtype(tree.coord_dtype).type
    nper = _nper(n, d)
    ncls = nper * (tree.nlevels if per_level else 1)
    lists, starts = trav.from_sep_siblings_lists, trav.from_sep_siblings_starts
    tp = trav.target_or_target_parent_boxes
    out = np.zeros(len(lists), np.int32)
    used = np.zeros(ncls, np.int32)
    root_extent = ct(tree.root_extent)
    for itgt in range(len(tp)):
        tbox = tp[itgt]
        for i in range(starts[itgt], starts[itgt + 1]):            # :150-196
            sbox = lists[i]
            level = int(tree.box_levels[sbox])
            if level != tree.box_levels[tbox]:
                raise ValueError("could not compute translation classes")
            diam = ct(2) * (root_extent * ct(1) / ct(1 << (level + 1)))
            cls, mult = 0, 1
            for a in range(d):
                v = int(np.rint((tree.box_centers[a, tbox] - tree.box_centers[a, sbox]) / diam))
                if not (-(2 * n + 1) <= v <= 2 * n + 1):
                    raise ValueError("could not compute translation classes")
                cls += (2 * n + 1 + v) * mult
                mult *= 4 * n + 3
            if per_level:
                cls += level * nper
            out[i] = cls
            used[cls] = 1
    return used, out


def translation_classes(trav, tree, per_level=True):
    n, d = trav.well_sep_is_n_away, tree.dimensions
    coord_dtype = np.dtype(tree.coord_dtype)
    used, lists = compute_translation_classes(trav, tree, per_level)
    nper = _nper(n, d)
    used_map = np.full(len(used), -1, np.int32)
    distances = np.zeros((d, len(used)), coord_dtype)
    level_starts = np.empty(tree.nlevels + 1, np.int32)
    count, prev_level = 0, -1
    root_extent = coord_dtype.type(tree.root_extent)
    for i, u in enumerate(used):                                    # :396-420
        level = i // nper
        if prev_level != level:
            level_starts[level] = count
            prev_level = level
        if not u:
            continue
        used_map[i] = count
        distances[:, count] = _class_to_vector(n, d, i % nper) * root_extent / (1 << level)
        count += 1
    if not per_level:
        level_starts[1:] = count
    level_starts[tree.nlevels] = count
    return used_map[lists], distances, level_starts


def rotation_classes(trav, tree):
    n, d = trav.well_sep_is_n_away, tree.dimensions
    used, lists = compute_translation_classes(trav, tree, False)
    angle_to_class, angles = {}, []
    cls_to_rot = np.full(_nper(n, d), -1, np.int32)
    for cls in np.flatnonzero(used):                                # rotation_classes.py:114-163
        vec = _class_to_vector(n, d, cls)
        g = abs(int(vec[0]))
        for e in vec[1:]:
            g = math.gcd(g, abs(int(e)))
        vec //= g
        angle = np.arccos(vec[-1] / np.linalg.norm(vec))
        if angle not in angle_to_class:
            angle_to_class[angle] = len(angles)
            angles.append(angle)
        cls_to_rot[cls] = angle_to_class[angle]
    return cls_to_rot[lists], np.array(angles)
