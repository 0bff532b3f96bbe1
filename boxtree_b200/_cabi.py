"""ctypes binding of ``libboxtree_b200.so`` (see ``include/boxtree_b200.h``).

The product path has no CPU fallback: importing this module without the built
library, or calling into it without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# BT_LIB_PATH selects another build of the same sources (kernel tuning experiments)
LIB_PATH = os.environ.get("BT_LIB_PATH") or os.path.join(HERE, "libboxtree_b200.so")

BT_F32, BT_F64 = 0, 1

# control block slots (include/boxtree_b200.h)
CTL_NBOXES = 0
CTL_NSPLIT = 1
CTL_OVERSIZE = 2
CTL_OVERFLOW = 3
CTL_NSPLIT_REGULAR = 4
CTL_COMMITTED = 5
CTL_NBOXES_FINAL = 6
CTL_NBIG = 7
CTL_NHUGE = 8
CTL_LR_FOUND = 16
CTL_SIZE = 64
STEP_DECIDE, STEP_CREATE, STEP_COMMIT, STEP_ALL = 1, 2, 4, 7

vp = C.c_void_p


class bt_particles(C.Structure):
    _fields_ = [("sources", vp * 3), ("targets", vp * 3), ("source_radii", vp),
                ("target_radii", vp), ("nsources", C.c_int64), ("ntargets", C.c_int64)]


class bt_pool(C.Structure):
    _fields_ = [("start", vp), ("count", vp), ("level", vp), ("parent", vp), ("child0", vp),
                ("has_children", vp), ("force_split", vp), ("nonchild", vp), ("center", vp * 3),
                ("capacity", C.c_int32), ("gstart", vp), ("gcount", vp), ("gnonchild", vp),
                ("xch", vp)]


class bt_box_out(C.Structure):
    _fields_ = [("box_start", vp), ("box_count", vp), ("box_nonchild", vp), ("box_levels", vp),
                ("box_parent_ids", vp), ("box_child_ids", vp), ("box_centers", vp),
                ("has_children", vp), ("real_children", vp), ("local_start", vp),
                ("local_count", vp), ("local_nonchild", vp)]


class bt_tree_view(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nboxes", C.c_int32), ("aligned_nboxes", C.c_int32),
                ("nlevels", C.c_int32), ("root_extent", C.c_double), ("box_centers", vp),
                ("box_levels", vp), ("box_child_ids", vp), ("box_flags", vp),
                ("box_parent_ids", vp), ("well_sep_is_n_away", C.c_int32),
                ("box_child_ids_t", vp)]


class bt_list_args(C.Structure):
    _fields_ = [("row_boxes", vp), ("coll_starts", vp), ("coll_lists", vp),
                ("stick_out_factor", C.c_double), ("with_extent", C.c_int32), ("row_mask", vp)]


class bt_list3_args(C.Structure):
    _fields_ = [("target_boxes", vp), ("coll_starts", vp), ("coll_lists", vp),
                ("stick_out_factor", C.c_double), ("targets_have_extent", C.c_int32),
                ("sources_have_extent", C.c_int32), ("crit", C.c_int32),
                ("box_target_bounding_box_min", vp), ("box_target_bounding_box_max", vp),
                ("box_source_counts_cumul", vp), ("min_nsources_cumul", C.c_int32)]


class bt_heavy_ws(C.Structure):
    _fields_ = [("walk_budget", C.c_int32), ("row_heavy", vp), ("heavy_rows", vp), ("hctl", vp),
                ("heavy_total", vp), ("frontier", vp * 2), ("frontier_cap", C.c_int64),
                ("dfs_rank", vp), ("ekeys", vp * 2), ("evals", vp * 2), ("ecap", C.c_int64),
                ("row_mask", vp), ("stage", vp), ("stage_cap", C.c_int32), ("stage_count", vp),
                ("dfs_order", vp), ("subtree_size", vp), ("hrow_base", vp), ("hplan", vp),
                ("seg_stride", C.c_int32), ("hseg_rank", vp), ("hseg_prefix", vp), ("hseg_kind", vp),
                ("hseg_n", vp), ("hmap", vp), ("hmap_cap", C.c_int64), ("chunk_cnt", vp), ("hctx", vp)]


HCTL_NWALK = 3
HCTL_SIZE = 64
STEP_DECIDE, STEP_CREATE, STEP_COMMIT, STEP_ALL = 1, 2, 4, 7
HCTL_NHEAVY = 0
HCTL_OVERFLOW = 1

# every exported symbol of include/boxtree_b200.h with its argument types
_i, _i64, _d = C.c_int, C.c_int64, C.c_double
_P = C.POINTER
SIGNATURES = {
    "bt_max_key_level": [_i],
    "bt_max_tree_level": [_i],
    "bt_bounding_box": [_i, _i, _P(bt_particles), vp, vp],
    "bt_make_keys": [_i, _i, _P(bt_particles), _P(_d), _P(_d), _i, _d, _i, vp, vp, vp, vp],
    "bt_sort_particles_deep": [_i64, _i, _i, vp, vp, vp, vp, vp, vp, vp],
    "bt_sort_particles": [_i64, _i, _i, _i, vp, vp, vp, vp, _P(_i), vp],
    "bt_weight_prefix": [_i64, vp, vp, vp, vp],
    "bt_pool_init": [_i, _i, _P(bt_pool), _i64, _i, vp, _P(_d), vp, vp],
    "bt_level_step": [_i, _i, _P(bt_pool), vp, vp, vp, vp, vp, _i, _i, _i, _i, _i, _i, _i, _i,
                      _d, _i, _i, vp, vp],
    "bt_level_restrict": [_i, _i, _P(bt_pool), vp, _i, _i, _d, vp],
    "bt_finalize_numbering": [_i, _i, _P(bt_pool), _i, _i, _i, vp, vp, vp, vp, vp],
    "bt_gather_boxes": [_i, _i, _P(bt_pool), _i, vp, vp, _i, _i, _P(bt_box_out), vp],
    "bt_leaf_fixup": [_i, vp, vp, vp, vp, vp, vp, _i, vp, vp],
    "bt_sort_u32_segment": [_i64, vp, vp],
    "bt_split_sources_targets": [_i64, _i64, vp, vp, vp, vp, vp, vp],
    "bt_reverse_index": [_i64, vp, vp, vp],
    "bt_permute": [_i, _i, _P(bt_particles), vp, vp, _i64, _P(vp), vp, vp],
    "bt_box_info": [_i, _i, _i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_box_extents": [_i, _i, _i, _i, _i, _P(C.c_int32), vp, vp, vp, vp, _P(vp), vp, vp, vp, vp],
    "bt_box_extents_phase": [_i, _i, _i, _i, _i, _P(C.c_int32), vp, vp, vp, vp, _P(vp), vp, vp, vp,
                             _i, vp],
    "bt_box_info_local": [_i, _i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_box_info_global": [_i, _i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_trav_box_list": [_i, _i, vp, vp, vp, vp, vp],
    "bt_trav_level_starts": [_i, vp, vp, _i, vp, vp],
    "bt_trav_build_list": [_i, _i, _i, _P(bt_tree_view), _P(bt_list_args), _i, vp, vp, vp, vp,
                           vp, vp],
    "bt_trav_mark_rows": [_i, _P(bt_tree_view), _P(bt_list_args), vp, vp, vp],
    "bt_trav_list3": [_i, _i, _P(bt_tree_view), _P(bt_list3_args), _i, vp, vp, vp, vp,
                      _P(bt_heavy_ws), _i64, vp],
    "bt_trav_list1": [_i, _i, _P(bt_tree_view), vp, _i, vp, vp, vp, _P(bt_heavy_ws), _i64, vp],
    "bt_trav_dfs_rank": [_i, _i, _i, _i, vp, vp, vp, vp, vp],
    "bt_trav_transpose_children": [_i, _i, vp, vp, vp],
    "bt_trav_colleagues": [_i, _i, _P(bt_tree_view), vp, vp, vp, _i, vp, vp, vp, vp, vp, vp, _i,
                           vp, vp],
    "bt_trav_list2_fill_masked": [_i, _i, vp, vp, vp, vp, vp, vp, _i, vp, vp, vp],
    "bt_trav_list2_starts": [_i, vp, vp, vp, vp, vp],
    "bt_trav_list3_compress": [_i, _i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_trav_list13": [_i, _i, _P(bt_tree_view), _P(bt_list3_args), vp, _i, vp, vp, vp, vp,
                       _P(bt_heavy_ws), _i64, _i, _i, vp],
    "bt_trav_merge_lists": [_i, _i, vp, _i, _P(vp), _P(vp), vp, vp, vp, vp],
    "bt_area_query": [_i, _i, _P(bt_tree_view), vp, vp, _i, _P(vp), vp, _P(_d), vp, vp, vp, vp],
    "bt_gather_i32": [_i64, vp, vp, vp, vp],
    "bt_csr_row_sums": [_i, _i, vp, vp, vp, vp, vp, _i, _d, vp],
    "bt_range_sums_i64": [_i, vp, vp, vp, vp, vp],
    "bt_add_to_ranges_i64": [_i, vp, vp, vp, vp, vp, vp],
    "bt_fmm_upward_i64": [_i, _i, vp, vp, _i, vp, vp],
    "bt_fmm_downward_i64": [_i, vp, vp, vp, vp],
    "bt_gather_i64": [_i64, vp, vp, vp, vp],
    "bt_gather_coords": [_i, _i64, vp, vp, vp, vp],
    "bt_widen_i32": [_i, _i64, vp, vp, vp],
    "bt_filter_targets_user_order": [_i, _i, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_filter_targets_tree_order": [_i, _i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_link_point_sources": [_i, _i, _i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_translation_classes": [_i, _i, _i, vp, vp, vp, vp, _i, vp, _d, _i, _i, _i, _i64, vp, vp, vp,
                               vp],
    "bt_remap_classes": [_i64, vp, vp, vp],
    "bt_dist_dfs_order": [_i, _i, _i, _i, vp, vp, vp, vp, vp, vp],
    "bt_dist_partition_cuts": [_i, _i, vp, vp, vp, vp, vp],
    "bt_dist_mask_from_list": [_i, vp, vp, vp],
    "bt_dist_ancestor_mask": [_i, vp, vp, vp, vp],
    "bt_dist_add_list_boxes": [_i, vp, vp, vp, vp, vp, vp, vp],
    "bt_dist_particle_mask": [_i, vp, vp, vp, vp, vp],
    "bt_dist_mask_scan": [_i64, vp, vp, vp],
    "bt_dist_fetch_local_particles": [_i, _i, _i64, vp, vp, _P(vp), vp, _P(vp), vp, vp, vp],
    "bt_dist_local_lists": [_i, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_dist_modify_target_flags": [_i, vp, vp, vp, vp],
    "bt_dist_box_to_user_rank": [_i, _i, _i, vp, vp, vp, vp, vp],
    "bt_dist_restrict_target_flags": [_i, vp, vp, vp, vp, vp, vp],
    "bt_dist_corner_flags": [_i, vp, vp, vp, vp, vp, vp, vp],
    "bt_dist_mark_list_boxes": [_i64, vp, vp, vp],
    "bt_dist_mask_bits": [_i, _i, _i, vp, vp, vp],
    "bt_dist_pack_ntiles": [_i64],
    "bt_dist_pack_count": [_i, _i, _i64, vp, vp, vp, vp, vp, vp, vp, vp],
    "bt_dist_pack_records": [_i, _i, _i, _i64, vp, vp, vp, _P(vp), vp, vp, vp, vp],
    "bt_dist_compact_index": [_i, vp, vp, vp, vp],
    "bt_dist_unpack_records": [_i, _i, _i, _i64, _i, vp, _P(_i64), vp, _i, vp, vp, vp, _P(vp), vp,
                               vp, vp],
    "bt_dist_box_to_user_rank_bits": [_i, _i, _i, _i, vp, vp, vp, vp, vp],
    "bt_dist_local_ranges": [_i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
}

_lib = None


class BoxtreeB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BoxtreeB200Error(
                f"{LIB_PATH} is missing: build it with `python -m boxtree_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        lib.bt_launch_count.argtypes = []
        lib.bt_launch_count.restype = C.c_longlong
        lib.bt_prof_enable.argtypes = [C.c_int]
        lib.bt_prof_enable.restype = None
        lib.bt_prof_reset.argtypes = []
        lib.bt_prof_reset.restype = None
        lib.bt_prof_report.argtypes = [C.c_char_p, C.c_int]
        lib.bt_prof_report.restype = C.c_int
        lib.bt_set_walk_mode.argtypes = [C.c_int]
        lib.bt_set_walk_mode.restype = None
        lib.bt_get_walk_mode.argtypes = []
        lib.bt_get_walk_mode.restype = C.c_int
        if os.environ.get("BT_WALK_MODE"):
            lib.bt_set_walk_mode(int(os.environ["BT_WALK_MODE"]))
        _lib = lib
    return _lib


def launch_count() -> int:
    """Number of kernels this library has launched so far in this process."""
    return int(load().bt_launch_count())


def profile_report() -> dict:
    """``{scope: (calls, total_ms)}`` of the scopes recorded since ``bt_prof_reset``;
    the stream must be synchronised."""
    lib = load()
    need = lib.bt_prof_report(None, 0)
    buf = C.create_string_buffer(need + 16)
    lib.bt_prof_report(buf, need + 16)
    out = {}
    for line in buf.value.decode().splitlines():
        name, calls, ms = line.split("\t")
        out[name] = (int(calls), float(ms))
    return out


def check(code: int, what: str) -> None:
    if code != 0:
        raise BoxtreeB200Error(f"{what} failed with status {code}"
                               + (" (CUDA error)" if code < 10000 else ""))


def dtype_code(np_dtype) -> int:
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float32:
        return BT_F32
    if np_dtype == np.float64:
        return BT_F64
    raise TypeError(f"unsupported coordinate dtype: {np_dtype}")


def dptr(t):
    """Device pointer of a torch tensor (or NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return t.data_ptr()


def darray(values):
    arr = (C.c_double * 3)()
    for k, v in enumerate(values):
        arr[k] = float(v)
    return arr


def ptr_array(tensors):
    arr = (vp * max(len(tensors), 1))()
    for k, t in enumerate(tensors):
        arr[k] = dptr(t)
    return arr
