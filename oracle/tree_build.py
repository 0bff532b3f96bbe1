"""CPU oracle: restatement of ``boxtree.tree_build.TreeBuilder.__call__``.

TEST INFRASTRUCTURE ONLY -- never imported by ``boxtree_b200``.  PARITY PINNED
against the reference's own ``TreeBuilder`` executed on the CPU (``tests/refexec``,
``tests/test_refexec.py``; see ``oracle_tree.c``).  This driver restates the host control flow of
``/root/reference/boxtree/tree_build.py:145-1878`` step by step (same arrays,
same level loop, same reallocation/renumbering bookkeeping) on numpy arrays,
calling the C restatements of the reference's OpenCL kernels.

Scalar arithmetic on the bounding box follows NumPy 2 (NEP 50) promotion, i.e.
``root_extent`` is computed in the coordinate dtype.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from ._lib import cint, coord_arg, i32, i64, lib_for, ptr, ptr_array


class MaxLevelsExceeded(RuntimeError):
    pass


ROOT_EXTENT_STRETCH_FACTOR = 1e-4   # tree_build.py:101
EXTENT_NORM_CODE = {None: 0, "linf": 1, "l2": 2}


@dataclass
class OracleTree:
    """Plain-numpy mirror of ``boxtree.tree.Tree`` (tree.py:298-590)."""
    sources_are_targets: bool
    sources_have_extent: bool
    targets_have_extent: bool
    particle_id_dtype: Any
    box_id_dtype: Any
    coord_dtype: Any
    box_level_dtype: Any
    bounding_box: tuple
    root_extent: Any
    stick_out_factor: Any
    extent_norm: Any
    level_start_box_nrs: np.ndarray
    sources: list
    targets: list
    source_radii: Any
    target_radii: Any
    box_source_starts: np.ndarray
    box_source_counts_nonchild: np.ndarray
    box_source_counts_cumul: np.ndarray
    box_target_starts: np.ndarray
    box_target_counts_nonchild: np.ndarray
    box_target_counts_cumul: np.ndarray
    box_parent_ids: np.ndarray
    box_child_ids: np.ndarray
    box_centers: np.ndarray
    box_levels: np.ndarray
    box_flags: np.ndarray
    user_source_ids: np.ndarray
    sorted_target_ids: np.ndarray
    box_source_bounding_box_min: np.ndarray
    box_source_bounding_box_max: np.ndarray
    box_target_bounding_box_min: np.ndarray
    box_target_bounding_box_max: np.ndarray
    _is_pruned: bool
    extra: dict = field(default_factory=dict)

    @property
    def dimensions(self):
        return len(self.sources)

    @property
    def nboxes(self):
        return len(self.box_flags)

    @property
    def nsources(self):
        return len(self.sources[0])

    @property
    def ntargets(self):
        return len(self.targets[0])

    @property
    def nlevels(self):
        return len(self.level_start_box_nrs) - 1

    @property
    def aligned_nboxes(self):
        return self.box_child_ids.shape[-1]


def _realloc(new_len, ary):
    """tools.py:56-78 (realloc_array): zero-filled, old bytes copied to the front."""
    res = np.zeros((new_len,) + ary.shape[1:], ary.dtype)
    res[:len(ary)] = ary
    return res


def _gappy_copy_and_map(new_len, ary, src_indices=None, dst_indices=None,
                        map_values=None, rng=None):
    """tools.py:417-534 (GappyCopyAndMapKernel)."""
    res = np.zeros((new_len,) + ary.shape[1:], ary.dtype)
    if rng is None:
        rng = len(src_indices) if src_indices is not None else len(dst_indices)
    idx = np.arange(rng)
    val = ary[src_indices[:rng]] if src_indices is not None else ary[:rng]
    if map_values is not None:
        val = map_values[val]
    if dst_indices is not None:
        res[dst_indices[:rng]] = val
    else:
        res[idx] = val
    return res


def build_tree(particles, kind="adaptive", max_particles_in_box=None,
               targets=None, source_radii=None, target_radii=None,
               stick_out_factor=None, refine_weights=None,
               max_leaf_refine_weight=None, extent_norm=None, bbox=None,
               nboxes_guess=None, lr_lookbehind=1, skip_prune=False,
               debug=False) -> OracleTree:
    # {{{ input processing -- tree_build.py:225-295
    if kind not in ["adaptive", "adaptive-level-restricted", "non-adaptive"]:
        raise ValueError(f"unknown tree kind: '{kind}'")
    particles = [np.ascontiguousarray(p) for p in particles]
    dimensions = len(particles)
    sources_are_targets = targets is None
    sources_have_extent = source_radii is not None
    targets_have_extent = target_radii is not None
    if extent_norm is None:
        extent_norm = "linf"
    if extent_norm not in ["linf", "l2"]:
        raise ValueError(f"unexpected value of 'extent_norm': {extent_norm}")
    srcntgts_extent_norm = extent_norm
    srcntgts_have_extent = sources_have_extent or targets_have_extent
    if not srcntgts_have_extent:
        srcntgts_extent_norm = None
    if srcntgts_extent_norm and targets is None:
        raise ValueError("must specify targets when specifying any kind of radii")

    particle_id_dtype = np.dtype(np.int32)
    box_id_dtype = np.dtype(np.int32)
    coord_dtype = particles[0].dtype
    assert all(p.dtype == coord_dtype for p in particles)
    if targets is None:
        nsrcntgts = len(particles[0])
        nsources = nsrcntgts
        ntargets = nsrcntgts
    else:
        targets = [np.ascontiguousarray(t) for t in targets]
        nsources = len(particles[0])
        ntargets = len(targets[0])
        nsrcntgts = nsources + ntargets
    if source_radii is not None:
        if source_radii.shape != (nsources,):
            raise ValueError("'source_radii' has an invalid shape")
        if source_radii.dtype != coord_dtype:
            raise TypeError("dtypes of 'particles' and 'source_radii' must agree")
    if target_radii is not None:
        if target_radii.shape != (ntargets,):
            raise ValueError("'target_radii' has an invalid shape")
        if target_radii.dtype != coord_dtype:
            raise TypeError("dtypes of 'particles' and 'target_radii' must agree")
    if sources_have_extent or targets_have_extent:
        if stick_out_factor is None:
            raise ValueError("if sources or targets have extent, "
                             "'stick_out_factor' must be explicitly specified")
    else:
        stick_out_factor = 0
    # }}}

    lib = lib_for(coord_dtype)
    level_restrict = kind == "adaptive-level-restricted"
    adaptive = kind != "non-adaptive"
    have_ext = int(srcntgts_have_extent)
    norm_code = EXTENT_NORM_CODE[srcntgts_extent_norm]
    nb = 2**dimensions
    mbc_w = have_ext + 2 * nb

    # {{{ combine sources and targets -- tree_build.py:328-388
    if targets is None:
        srcntgts = [p.copy() for p in particles]
        srcntgt_radii = None
    else:
        if targets[0].dtype != coord_dtype:
            raise TypeError("sources and targets coordinates must have same dtype")

        def combine(a1, a2):
            dt = a1.dtype if a1 is not None else a2.dtype
            res = np.zeros(nsrcntgts, dt)
            if a1 is not None and a1.nbytes:
                res[:len(a1)] = a1
            if a2 is not None and a2.nbytes:
                res[nsources:] = a2
            return res
        srcntgts = [combine(s, t) for s, t in zip(particles, targets)]
        srcntgt_radii = combine(source_radii, target_radii) if srcntgts_have_extent else None
    user_srcntgt_ids = np.arange(nsrcntgts, dtype=particle_id_dtype)
    # }}}

    # {{{ refine weights -- tree_build.py:405-452
    spec_max = max_particles_in_box is not None
    spec_w = refine_weights is not None and max_leaf_refine_weight is not None
    if spec_max and spec_w:
        raise ValueError("may only specify one of 'max_particles_in_box' and "
                         "'refine_weights'/'max_leaf_refine_weight")
    elif not spec_max and not spec_w:
        raise ValueError("must specify either 'max_particles_in_box' or "
                         "'refine_weights'/'max_leaf_refine_weight'")
    elif spec_max:
        refine_weights = np.ones(nsrcntgts, np.int32)
        max_leaf_refine_weight = max_particles_in_box
    else:
        if refine_weights.dtype != np.int32:
            raise TypeError("'refine_weights' must have dtype 'int32'")
        refine_weights = np.ascontiguousarray(refine_weights)
    if max_leaf_refine_weight <= 0:
        raise ValueError("'max_leaf_refine_weight' must be positive")
    if nsrcntgts and max_leaf_refine_weight < refine_weights.max():
        raise ValueError("entries of 'refine_weights' cannot exceed 'max_leaf_refine_weight'")
    if nsrcntgts and refine_weights.min() < 0:
        raise ValueError("all entries of 'refine_weights' must be nonnegative")
    total_refine_weight = int(np.sum(refine_weights, dtype=np.int64))
    max_leaf_refine_weight = int(max_leaf_refine_weight)
    # }}}

    # {{{ bounding box -- tree_build.py:458-508, bounding_box.py
    auto_min = np.empty(dimensions, coord_dtype)
    auto_max = np.empty(dimensions, coord_dtype)
    lib.orc_bounding_box(cint(dimensions), i64(nsrcntgts), ptr_array(srcntgts),
                         ptr(srcntgt_radii), ptr(auto_min), ptr(auto_max))
    if bbox is None:
        root_extent = max(auto_max[i] - auto_min[i] for i in range(dimensions)) \
            * (1 + ROOT_EXTENT_STRETCH_FACTOR)
        bbox_min = auto_min.copy()
        bbox_max = bbox_min + root_extent
    else:
        bbox = np.asarray(bbox)
        assert len(bbox) == dimensions
        bbox_min = np.empty(dimensions, coord_dtype)
        bbox_max = np.empty(dimensions, coord_dtype)
        for i in range(dimensions):
            bbox_min[i] = bbox[i][0]
            bbox_max[i] = bbox[i][1]
            assert bbox_min[i] < bbox_max[i]
            assert bbox_min[i] <= auto_min[i]
            assert bbox_max[i] >= auto_max[i]
        exts = bbox_max - bbox_min
        for ext in exts:
            assert abs(ext - exts[0]) < 1e-15
        root_extent = exts[0]
    root_extent = coord_dtype.type(root_extent)
    # }}}

    # {{{ allocate -- tree_build.py:516-631
    morton_bin_counts = np.zeros((nsrcntgts, mbc_w), np.int32)
    morton_nrs = np.zeros(nsrcntgts, np.int8)
    box_start_flags = np.zeros(nsrcntgts, np.int8)
    srcntgt_box_ids = np.zeros(nsrcntgts, box_id_dtype)
    if nboxes_guess is None:
        nboxes_guess = int(nb * ((max_leaf_refine_weight + total_refine_weight - 1)
                                 // max_leaf_refine_weight))
    assert nboxes_guess > 0
    split_box_ids = np.zeros(nboxes_guess, box_id_dtype)
    box_morton_bin_counts = np.zeros((nboxes_guess, mbc_w), np.int32)
    box_srcntgt_starts = np.zeros(nboxes_guess, particle_id_dtype)
    box_parent_ids = np.zeros(nboxes_guess, box_id_dtype)
    box_child_ids = [np.zeros(nboxes_guess, box_id_dtype) for _ in range(nb)]
    box_centers = [np.zeros(nboxes_guess, coord_dtype) for _ in range(dimensions)]
    for d in range(dimensions):
        box_centers[d][0] = bbox_min[d] + (bbox_max[d] - bbox_min[d]) / 2
    box_levels = np.zeros(nboxes_guess, np.uint8)
    box_srcntgt_counts_cumul = np.zeros(nboxes_guess, particle_id_dtype)
    box_srcntgt_counts_cumul[0] = nsrcntgts
    box_has_children = np.zeros(nboxes_guess, np.int32)
    force_split_box = np.zeros(nboxes_guess if level_restrict else 0, np.int32)
    nlevels_max = 2 * (np.finfo(coord_dtype).nmant + 1)
    level_start_box_nrs_dev = np.zeros(nlevels_max, box_id_dtype)
    level_used_box_counts_dev = np.zeros(nlevels_max, box_id_dtype)
    have_oversize_split_box = np.zeros((), np.int32)
    have_upper_level_split_box = np.zeros((), np.int32)
    # }}}

    # {{{ level loop -- tree_build.py:653-1276
    level_start_box_nrs = [0, 1]
    level_start_box_nrs_dev[0] = 0
    level_start_box_nrs_dev[1] = 1
    level_used_box_counts = [1]
    level_used_box_counts_dev[0] = 1
    level_leaf_counts = np.array([1])
    level = 1 if total_refine_weight > max_leaf_refine_weight else 0
    final_level_restrict_iteration = False
    stats = {"lr_boxes_split": [], "level_iterations": 0}

    coords_pp = ptr_array(srcntgts)
    sof = coord_arg(coord_dtype, stick_out_factor)

    while level:
        stats["level_iterations"] += 1
        if level + 1 >= nlevels_max:
            raise MaxLevelsExceeded("Level count exceeded number of significant "
                                    "bits in coordinate dtype.")

        # morton count scan -- tree_build.py:732
        lib.orc_morton_count_scan(
            cint(dimensions), cint(norm_code), i64(nsrcntgts),
            ptr(morton_bin_counts), ptr(morton_nrs), ptr(box_start_flags),
            ptr(srcntgt_box_ids), ptr(box_morton_bin_counts), ptr(refine_weights),
            ptr(box_srcntgt_counts_cumul), ptr(box_levels), ptr(bbox_min), ptr(bbox_max),
            ptr(user_srcntgt_ids), coords_pp, ptr(srcntgt_radii), sof)

        # split box id scan -- tree_build.py:740-759
        lib.orc_split_box_id_scan(
            cint(dimensions), cint(have_ext), cint(adaptive), cint(level_restrict),
            i64(level_start_box_nrs[level]),
            ptr(box_srcntgt_counts_cumul), ptr(box_morton_bin_counts),
            i32(max_leaf_refine_weight), ptr(box_levels), ptr(level_start_box_nrs_dev),
            ptr(level_used_box_counts_dev),
            ptr(force_split_box) if level_restrict else C.c_void_p(0), cint(level),
            ptr(box_has_children), ptr(split_box_ids),
            C.c_void_p(have_oversize_split_box.ctypes.data))

        # tree_build.py:762-786
        new_level_used_box_counts = [1]
        for level_start_box_id in level_start_box_nrs[1:]:
            last_box_on_prev_level = level_start_box_id - 1
            new_level_used_box_counts.append(
                int(split_box_ids[last_box_on_prev_level]) - level_start_box_id)
        level_used_box_counts_diff = (np.array(new_level_used_box_counts)
                                      - np.append(level_used_box_counts, [0]))
        new_level_leaf_counts = (level_leaf_counts
                                 + level_used_box_counts_diff[:-1]
                                 - level_used_box_counts_diff[1:] // 2**dimensions)
        new_level_leaf_counts = np.append(new_level_leaf_counts,
                                          [level_used_box_counts_diff[-1]])

        # tree_build.py:804-827
        curr_upper_level_lengths = np.diff(level_start_box_nrs)
        minimal_upper_level_lengths = np.max(
            [new_level_used_box_counts[:-1], curr_upper_level_lengths], axis=0)
        minimal_new_level_length = new_level_used_box_counts[-1]
        if level_restrict and int(have_oversize_split_box):
            minimal_new_level_length += sum(
                2**(lev * dimensions) * new_level_leaf_counts[level - lev]
                for lev in range(1, 1 + min(level, lr_lookbehind)))
        nboxes_minimal = int(sum(minimal_upper_level_lengths) + minimal_new_level_length)
        needs_renumbering = (curr_upper_level_lengths < minimal_upper_level_lengths).any()

        # tree_build.py:831-908
        if needs_renumbering:
            assert level_restrict
            upper_level_padding = np.zeros(level, dtype=int)
            for ulevel in range(level):
                upper_level_padding[ulevel] = sum(
                    2**(lev * dimensions) * new_level_leaf_counts[ulevel - lev]
                    for lev in range(1, 1 + min(ulevel, lr_lookbehind)))
            new_upper_level_unused_box_counts = np.max(
                [upper_level_padding,
                 minimal_upper_level_lengths - new_level_used_box_counts[:-1]], axis=0)
            new_level_start_box_nrs = np.empty(level + 1, dtype=int)
            new_level_start_box_nrs[0] = 0
            new_level_start_box_nrs[1:] = np.cumsum(
                np.array(new_level_used_box_counts[:-1]) + new_upper_level_unused_box_counts)
            assert not (np.array(level_start_box_nrs) == new_level_start_box_nrs).all()

            old_box_count = level_start_box_nrs[-1]
            dst_box_id = np.zeros(old_box_count, box_id_dtype)
            for level_start, new_level_start, level_len in zip(
                    level_start_box_nrs[:-1], new_level_start_box_nrs[:-1],
                    curr_upper_level_lengths):
                dst_box_id[level_start:level_start + level_len] = np.arange(
                    new_level_start, new_level_start + level_len, dtype=box_id_dtype)

            def realloc_array(new_len, ary):
                return _gappy_copy_and_map(new_len, ary, dst_indices=dst_box_id,
                                           rng=old_box_count)

            def realloc_and_renumber_array(new_len, ary):
                return _gappy_copy_and_map(new_len, ary, dst_indices=dst_box_id,
                                           map_values=dst_box_id, rng=old_box_count)

            renumber = True
            level_start_box_nrs = [int(x) for x in new_level_start_box_nrs]
            level_start_box_nrs_dev[:level + 1] = np.array(new_level_start_box_nrs,
                                                           dtype=box_id_dtype)
            level_start_box_nrs_updated = True
            nboxes_new = level_start_box_nrs[-1] + minimal_new_level_length
        else:
            realloc_array = _realloc
            realloc_and_renumber_array = _realloc
            renumber = False
            level_start_box_nrs_updated = False
            nboxes_new = nboxes_minimal

        # tree_build.py:914-1005
        if level_start_box_nrs_updated or nboxes_new > nboxes_guess:
            while nboxes_guess < nboxes_new:
                nboxes_guess *= 2
            split_box_ids = np.zeros(nboxes_guess, box_id_dtype)
            box_morton_bin_counts = realloc_array(nboxes_guess, box_morton_bin_counts)
            if level_restrict:
                force_split_box = realloc_array(nboxes_guess, force_split_box)
            box_srcntgt_starts = realloc_array(nboxes_guess, box_srcntgt_starts)
            box_srcntgt_counts_cumul = realloc_array(nboxes_guess, box_srcntgt_counts_cumul)
            box_has_children = realloc_array(nboxes_guess, box_has_children)
            box_centers = [realloc_array(nboxes_guess, a) for a in box_centers]
            box_child_ids = [realloc_and_renumber_array(nboxes_guess, a)
                             for a in box_child_ids]
            box_parent_ids = realloc_and_renumber_array(nboxes_guess, box_parent_ids)
            if not level_start_box_nrs_updated:
                box_levels = realloc_array(nboxes_guess, box_levels)
            else:
                box_levels = np.zeros(nboxes_guess, np.uint8)
                for box_level, (ls, le) in enumerate(
                        zip(level_start_box_nrs[:-1], level_start_box_nrs[1:])):
                    box_levels[ls:le] = box_level
            if level_start_box_nrs_updated and renumber:
                srcntgt_box_ids = dst_box_id[srcntgt_box_ids]
            stats["reallocs"] = stats.get("reallocs", 0) + 1
            continue  # retry the level

        assert (level_start_box_nrs[-1] != nboxes_new or srcntgts_have_extent
                or final_level_restrict_iteration)
        if level_start_box_nrs[-1] == nboxes_new:
            if srcntgts_have_extent and not final_level_restrict_iteration:
                level -= 1
                break
            assert final_level_restrict_iteration

        # tree_build.py:1027-1038
        level_start_box_nrs.append(int(nboxes_new))
        level_start_box_nrs_dev[level + 1] = nboxes_new
        level_used_box_counts = list(new_level_used_box_counts)
        level_used_box_counts_dev[:level + 1] = np.array(level_used_box_counts,
                                                         dtype=box_id_dtype)
        level_leaf_counts = new_level_leaf_counts
        if debug:
            for ls, ln, lc in zip(level_start_box_nrs[:-1], level_used_box_counts,
                                  level_leaf_counts):
                if ln == 0:
                    assert lc == 0
                    continue
                assert lc == ln - int(np.sum(box_has_children[ls:ls + ln]))

        # box splitter -- tree_build.py:1064-1085
        child_pp = ptr_array(box_child_ids)
        center_pp = ptr_array(box_centers)
        lib.orc_box_splitter(
            cint(dimensions), cint(have_ext), cint(level_restrict),
            i64(level_start_box_nrs[-1]), cint(level),
            ptr(box_morton_bin_counts), ptr(box_start_flags), ptr(split_box_ids),
            ptr(box_srcntgt_starts), ptr(box_srcntgt_counts_cumul), ptr(box_parent_ids),
            ptr(box_levels), ptr(box_has_children),
            ptr(force_split_box) if level_restrict else C.c_void_p(0),
            coord_arg(coord_dtype, root_extent), child_pp, center_pp)
        last_used_box = level_start_box_nrs[-2] + level_used_box_counts[-1]
        box_levels[last_used_box:level_start_box_nrs[-1]] = level
        if debug:
            assert np.all(box_levels[level_start_box_nrs[-2]:level_start_box_nrs[-1]]
                          == level)
            assert np.all(box_srcntgt_starts < max(nsrcntgts, 1))

        # particle renumberer -- tree_build.py:1101-1121
        new_user_srcntgt_ids = np.zeros_like(user_srcntgt_ids)
        new_srcntgt_box_ids = np.zeros_like(srcntgt_box_ids)
        lib.orc_particle_renumberer(
            cint(dimensions), cint(have_ext), cint(level_restrict), i64(nsrcntgts),
            cint(level), ptr(morton_bin_counts), ptr(morton_nrs), ptr(srcntgt_box_ids),
            ptr(split_box_ids), ptr(box_morton_bin_counts), ptr(box_srcntgt_starts),
            ptr(box_levels), ptr(user_srcntgt_ids), ptr(box_has_children),
            ptr(force_split_box) if level_restrict else C.c_void_p(0),
            ptr(new_user_srcntgt_ids), ptr(new_srcntgt_box_ids))
        user_srcntgt_ids = new_user_srcntgt_ids
        srcntgt_box_ids = new_srcntgt_box_ids

        # tree_build.py:1127-1143
        if final_level_restrict_iteration:
            assert int(have_oversize_split_box) == 0
            assert level_used_box_counts[-1] == 0
            del level_used_box_counts[-1]
            del level_start_box_nrs[-1]
            level -= 1
            break

        # level restriction -- tree_build.py:1145-1224
        if level_restrict:
            force_split_box[:] = 0
            did_upper_level_split = False
            boxes_split = []
            for upper_level, upper_level_start, upper_level_box_count in zip(
                    range(level - 2, 0, -1),
                    level_start_box_nrs[-4::-1],
                    level_used_box_counts[-3::-1]):
                have_upper_level_split_box[...] = 0
                lib.orc_level_restrict(
                    cint(dimensions), cint(upper_level), coord_arg(coord_dtype, root_extent),
                    i64(upper_level_start), i64(upper_level_box_count),
                    ptr(box_has_children), ptr(force_split_box),
                    C.c_void_p(have_upper_level_split_box.ctypes.data),
                    ptr_array(box_child_ids), ptr_array(box_centers))
                boxes_split.append(int(np.sum(force_split_box[
                    upper_level_start:upper_level_start + upper_level_box_count])))
                if int(have_upper_level_split_box) == 0:
                    break
                did_upper_level_split = True
            stats["lr_boxes_split"].append(boxes_split)
            if int(have_oversize_split_box) == 0 and did_upper_level_split:
                final_level_restrict_iteration = True
                level += 1
                continue

        if not int(have_oversize_split_box):
            break
        level += 1
        have_oversize_split_box[...] = 0
    # }}}

    nboxes = level_start_box_nrs[-1]

    # {{{ nonchild counts -- tree_build.py:1288-1305
    if srcntgts_have_extent:
        box_srcntgt_counts_nonchild = np.zeros(nboxes, particle_id_dtype)
        assert len(level_start_box_nrs) >= 2
        lib.orc_extract_nonchild_srcntgt_count(
            cint(dimensions), i64(nboxes), ptr(box_morton_bin_counts),
            ptr(box_srcntgt_counts_cumul), i32(level_start_box_nrs[-2]),
            ptr(box_srcntgt_counts_nonchild))
        if debug:
            assert np.all(box_srcntgt_counts_nonchild <= box_srcntgt_counts_cumul[:nboxes])
    # }}}

    # {{{ prune -- tree_build.py:1330-1456
    prune_empty_leaves = not skip_prune
    if prune_empty_leaves:
        src_box_id = np.zeros(nboxes, box_id_dtype)
        dst_box_id = np.zeros(nboxes, box_id_dtype)
        npp = np.zeros((), box_id_dtype)
        lib.orc_find_prune_indices(i64(nboxes), ptr(box_srcntgt_counts_cumul),
                                   ptr(src_box_id), ptr(dst_box_id),
                                   C.c_void_p(npp.ctypes.data))
        nboxes_post_prune = int(npp)
        should_prune = True
    elif level_restrict:
        src_box_id = np.zeros(nboxes, box_id_dtype)
        dst_box_id = np.zeros(nboxes, box_id_dtype)
        new_level_start_box_nrs = np.zeros(len(level_start_box_nrs), dtype=int)
        new_level_start_box_nrs[1:] = np.cumsum(level_used_box_counts)
        for ls, nls, used in zip(level_start_box_nrs[:-1], new_level_start_box_nrs[:-1],
                                 level_used_box_counts):
            src_box_id[nls:nls + used] = np.arange(ls, ls + used, dtype=box_id_dtype)
            dst_box_id[ls:ls + used] = np.arange(nls, nls + used, dtype=box_id_dtype)
        nboxes_post_prune = int(new_level_start_box_nrs[-1])
        should_prune = True
    else:
        should_prune = False

    if should_prune:
        def prune_empty(ary, map_values=None):
            return _gappy_copy_and_map(nboxes_post_prune, ary, src_indices=src_box_id,
                                       map_values=map_values, rng=nboxes_post_prune)
        box_srcntgt_starts = prune_empty(box_srcntgt_starts)
        box_srcntgt_counts_cumul = prune_empty(box_srcntgt_counts_cumul)
        if debug and prune_empty_leaves:
            assert np.all(box_srcntgt_counts_cumul > 0)
        srcntgt_box_ids = dst_box_id[srcntgt_box_ids]
        box_parent_ids = prune_empty(box_parent_ids, map_values=dst_box_id)
        box_levels = prune_empty(box_levels)
        if srcntgts_have_extent:
            box_srcntgt_counts_nonchild = prune_empty(box_srcntgt_counts_nonchild)
        box_has_children = prune_empty(box_has_children)
        box_child_ids = [prune_empty(a, map_values=dst_box_id) for a in box_child_ids]
        box_centers = [prune_empty(a) for a in box_centers]
        lib.orc_find_level_box_counts(i64(nboxes_post_prune), ptr(box_levels),
                                      ptr(level_used_box_counts_dev))
        nlevels = len(level_used_box_counts)
        level_used_box_counts = level_used_box_counts_dev[:nlevels].copy()
        level_start_box_nrs = [0]
        level_start_box_nrs.extend(np.cumsum(level_used_box_counts))
    else:
        nboxes_post_prune = nboxes
    level_start_box_nrs = np.array(level_start_box_nrs, box_id_dtype)
    # }}}

    # {{{ source/target split -- tree_build.py:1464-1561
    if targets is None:
        user_source_ids = user_srcntgt_ids
        sorted_target_ids = np.zeros(nsrcntgts, particle_id_dtype)   # tools.py:81-109
        sorted_target_ids[user_srcntgt_ids] = np.arange(nsrcntgts, dtype=particle_id_dtype)
        box_source_starts = box_target_starts = box_srcntgt_starts[:nboxes_post_prune]
        box_source_counts_cumul = box_target_counts_cumul = \
            box_srcntgt_counts_cumul[:nboxes_post_prune]
        if srcntgts_have_extent:
            box_source_counts_nonchild = box_target_counts_nonchild = \
                box_srcntgt_counts_nonchild
    else:
        source_numbers = np.zeros(nsrcntgts, particle_id_dtype)
        lib.orc_source_counter(i64(nsrcntgts), ptr(user_srcntgt_ids), i32(nsources),
                               ptr(source_numbers))
        user_source_ids = np.zeros(nsources, particle_id_dtype)
        srcntgt_target_ids = np.zeros(ntargets, particle_id_dtype)
        sorted_target_ids = np.zeros(ntargets, particle_id_dtype)
        box_source_starts = np.zeros(nboxes_post_prune, particle_id_dtype)
        box_source_counts_cumul = np.zeros(nboxes_post_prune, particle_id_dtype)
        box_target_starts = np.zeros(nboxes_post_prune, particle_id_dtype)
        box_target_counts_cumul = np.zeros(nboxes_post_prune, particle_id_dtype)
        if srcntgts_have_extent:
            box_source_counts_nonchild = np.zeros(nboxes_post_prune, particle_id_dtype)
            box_target_counts_nonchild = np.zeros(nboxes_post_prune, particle_id_dtype)
        lib.orc_source_and_target_index_finder(
            cint(have_ext), i64(nsrcntgts), ptr(user_srcntgt_ids), i32(nsources),
            ptr(srcntgt_box_ids), ptr(box_parent_ids), ptr(box_srcntgt_starts),
            ptr(box_srcntgt_counts_cumul), ptr(source_numbers),
            ptr(box_srcntgt_counts_nonchild) if srcntgts_have_extent else C.c_void_p(0),
            ptr(user_source_ids), ptr(srcntgt_target_ids), ptr(sorted_target_ids),
            ptr(box_source_starts), ptr(box_source_counts_cumul),
            ptr(box_target_starts), ptr(box_target_counts_cumul),
            ptr(box_source_counts_nonchild) if srcntgts_have_extent else C.c_void_p(0),
            ptr(box_target_counts_nonchild) if srcntgts_have_extent else C.c_void_p(0))
        if srcntgts_have_extent and debug:
            assert np.all(box_srcntgt_counts_nonchild
                          == box_source_counts_nonchild + box_target_counts_nonchild)
        if debug:
            assert np.all(box_source_counts_cumul + box_target_counts_cumul
                          == box_srcntgt_counts_cumul[:nboxes_post_prune])
    # }}}

    # {{{ permute -- tree_build.py:1571-1620
    if targets is None:
        sources = targets_out = [c[user_srcntgt_ids] for c in srcntgts]
        out_source_radii = out_target_radii = None
    else:
        sources = [c[user_source_ids] for c in srcntgts]
        targets_out = [c[srcntgt_target_ids] for c in srcntgts]
        out_source_radii = out_target_radii = None
        if srcntgt_radii is not None:
            out_source_radii = srcntgt_radii[user_source_ids]
            out_target_radii = srcntgt_radii[srcntgt_target_ids]
    # }}}

    nlevels = len(level_start_box_nrs) - 1
    assert nlevels == len(level_used_box_counts)
    assert level + 1 == nlevels, (level + 1, nlevels)
    if debug and nboxes_post_prune:
        assert int(np.max(box_levels[:nboxes_post_prune])) + 1 == nlevels

    # {{{ pack child ids / centres -- tree_build.py:1641-1659
    aligned_nboxes = ((nboxes_post_prune + 31) // 32) * 32
    box_child_ids_new = np.zeros((nb, aligned_nboxes), box_id_dtype)
    box_centers_new = np.zeros((dimensions, aligned_nboxes), coord_dtype)
    for m in range(nb):
        box_child_ids_new[m, :nboxes_post_prune] = box_child_ids[m][:nboxes_post_prune]
    for d in range(dimensions):
        box_centers_new[d, :nboxes_post_prune] = box_centers[d][:nboxes_post_prune]
    box_child_ids = box_child_ids_new
    box_centers = box_centers_new
    # }}}

    # {{{ box flags -- tree_build.py:1666-1723
    box_flags = np.zeros(nboxes_post_prune, np.uint8)
    if not srcntgts_have_extent:
        box_source_counts_nonchild = np.zeros(nboxes_post_prune, particle_id_dtype)
        if sources_are_targets:
            box_target_counts_nonchild = box_source_counts_nonchild
        else:
            box_target_counts_nonchild = np.zeros(nboxes_post_prune, particle_id_dtype)
    box_source_counts_nonchild = np.ascontiguousarray(box_source_counts_nonchild)
    lib.orc_box_info(
        cint(have_ext), cint(sources_are_targets), i64(nboxes_post_prune),
        ptr(box_srcntgt_counts_cumul), ptr(np.ascontiguousarray(box_source_counts_cumul)),
        ptr(np.ascontiguousarray(box_target_counts_cumul)), ptr(box_has_children),
        ptr(box_source_counts_nonchild), ptr(box_target_counts_nonchild), ptr(box_flags))
    # }}}

    # {{{ box extents -- tree_build.py:1730-1802
    bb_src_min = np.zeros((dimensions, aligned_nboxes), coord_dtype)
    bb_src_max = np.zeros((dimensions, aligned_nboxes), coord_dtype)
    if sources_are_targets:
        bb_tgt_min, bb_tgt_max = bb_src_min, bb_src_max
    else:
        bb_tgt_min = np.zeros((dimensions, aligned_nboxes), coord_dtype)
        bb_tgt_max = np.zeros((dimensions, aligned_nboxes), coord_dtype)
    bogus_radii = np.zeros(1, coord_dtype)
    box_source_starts = np.ascontiguousarray(box_source_starts)
    box_target_starts = np.ascontiguousarray(box_target_starts)
    for lev in range(nlevels - 1, -1, -1):
        start, stop = (int(x) for x in level_start_box_nrs[lev:lev + 2])
        rounds = [(False, sources_have_extent, bb_src_min, bb_src_max, box_source_starts,
                   box_source_counts_nonchild,
                   out_source_radii if sources_have_extent else bogus_radii, sources),
                  (sources_are_targets, targets_have_extent, bb_tgt_min, bb_tgt_max,
                   box_target_starts, box_target_counts_nonchild,
                   out_target_radii if targets_have_extent else bogus_radii, targets_out)]
        for skip, enable_radii, bmin, bmax, pstarts, pcounts, pradii, parts in rounds:
            if skip:
                continue
            parts = [np.ascontiguousarray(p) for p in parts]
            lib.orc_box_extents(
                cint(dimensions), cint(have_ext), i64(start), i64(stop), i64(aligned_nboxes),
                ptr(box_child_ids), ptr(box_centers), ptr(pstarts), ptr(pcounts),
                ptr_array(parts), ptr(np.ascontiguousarray(pradii)), cint(enable_radii),
                ptr(bmin), ptr(bmax))
    # }}}

    stats["nboxes_pre_prune"] = int(nboxes)
    return OracleTree(
        sources_are_targets=sources_are_targets,
        sources_have_extent=sources_have_extent,
        targets_have_extent=targets_have_extent,
        particle_id_dtype=particle_id_dtype, box_id_dtype=box_id_dtype,
        coord_dtype=coord_dtype, box_level_dtype=np.dtype(np.uint8),
        bounding_box=(bbox_min, bbox_max), root_extent=root_extent,
        stick_out_factor=stick_out_factor, extent_norm=srcntgts_extent_norm,
        level_start_box_nrs=level_start_box_nrs,
        sources=sources, targets=targets_out,
        source_radii=out_source_radii if sources_have_extent else None,
        target_radii=out_target_radii if targets_have_extent else None,
        box_source_starts=box_source_starts,
        box_source_counts_nonchild=box_source_counts_nonchild,
        box_source_counts_cumul=np.ascontiguousarray(box_source_counts_cumul),
        box_target_starts=box_target_starts,
        box_target_counts_nonchild=box_target_counts_nonchild,
        box_target_counts_cumul=np.ascontiguousarray(box_target_counts_cumul),
        box_parent_ids=box_parent_ids[:nboxes_post_prune],
        box_child_ids=box_child_ids, box_centers=box_centers,
        box_levels=box_levels[:nboxes_post_prune], box_flags=box_flags,
        user_source_ids=user_source_ids, sorted_target_ids=sorted_target_ids,
        box_source_bounding_box_min=bb_src_min, box_source_bounding_box_max=bb_src_max,
        box_target_bounding_box_min=bb_tgt_min, box_target_bounding_box_max=bb_tgt_max,
        _is_pruned=prune_empty_leaves, extra=stats)
