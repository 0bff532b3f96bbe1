"""Shared helpers for the parity tests: input recipes and bit-exact comparison of a
CUDA result (moved to numpy) against the CPU oracle."""
from __future__ import annotations

import numpy as np

TREE_INT_FIELDS = [
    "level_start_box_nrs", "box_source_starts", "box_source_counts_nonchild",
    "box_source_counts_cumul", "box_target_starts", "box_target_counts_nonchild",
    "box_target_counts_cumul", "box_parent_ids", "box_levels", "box_flags",
    "user_source_ids", "sorted_target_ids",
]
TREE_PADDED_INT_FIELDS = ["box_child_ids"]
TREE_PADDED_FLOAT_FIELDS = [
    "box_centers", "box_source_bounding_box_min", "box_source_bounding_box_max",
    "box_target_bounding_box_min", "box_target_bounding_box_max",
]
TRAV_FIELDS = [
    "source_boxes", "target_boxes", "level_start_source_box_nrs", "level_start_target_box_nrs",
    "source_parent_boxes", "level_start_source_parent_box_nrs", "target_or_target_parent_boxes",
    "level_start_target_or_target_parent_box_nrs", "same_level_non_well_sep_boxes_starts",
    "same_level_non_well_sep_boxes_lists", "neighbor_source_boxes_starts",
    "neighbor_source_boxes_lists", "from_sep_siblings_starts", "from_sep_siblings_lists",
    "from_sep_close_smaller_starts", "from_sep_close_smaller_lists", "from_sep_bigger_starts",
    "from_sep_bigger_lists", "from_sep_close_bigger_starts", "from_sep_close_bigger_lists",
]


def bits_equal(a, b):
    """Bitwise equality of two float arrays (0 ULP, distinguishes -0.0 / NaN payloads)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


def tree_mismatches(ref, got):
    """Names of Tree fields of *got* (numpy) that differ from the oracle's *ref*."""
    bad = []
    nb = ref.nboxes
    if got.nboxes != nb:
        return [f"nboxes {got.nboxes} != {nb}"]
    for name in ("sources_are_targets", "sources_have_extent", "targets_have_extent",
                 "extent_norm", "_is_pruned"):
        if getattr(ref, name) != getattr(got, name):
            bad.append(name)
    if np.dtype(ref.coord_dtype) != np.dtype(got.coord_dtype):
        bad.append("coord_dtype")
    if not bits_equal(np.asarray(ref.root_extent), np.asarray(got.root_extent)):
        bad.append("root_extent")
    for k in range(2):
        if not bits_equal(ref.bounding_box[k], got.bounding_box[k]):
            bad.append(f"bounding_box[{k}]")
    for name in TREE_INT_FIELDS:
        r, g = np.asarray(getattr(ref, name)), np.asarray(getattr(got, name))
        if name.startswith("box_"):
            r, g = r[:nb], g[:nb]
        if r.dtype != g.dtype or not np.array_equal(r, g):
            bad.append(name)
    for name in TREE_PADDED_INT_FIELDS:
        r, g = np.asarray(getattr(ref, name)), np.asarray(getattr(got, name))
        if r.shape != g.shape or r.dtype != g.dtype or not np.array_equal(r, g):
            bad.append(name)
    for name in TREE_PADDED_FLOAT_FIELDS:
        if not bits_equal(getattr(ref, name), getattr(got, name)):
            bad.append(name)
    for name in ("sources", "targets"):
        for ax in range(ref.dimensions):
            if not bits_equal(getattr(ref, name)[ax], getattr(got, name)[ax]):
                bad.append(f"{name}[{ax}]")
    for name in ("source_radii", "target_radii"):
        r, g = getattr(ref, name), getattr(got, name)
        if (r is None) != (g is None) or (r is not None and not bits_equal(r, g)):
            bad.append(name)
    return bad


def trav_mismatches(ref, got):
    bad = []
    for name in TRAV_FIELDS:
        r, g = getattr(ref, name), getattr(got, name)
        if (r is None) != (g is None):
            bad.append(name + " (None-ness)")
            continue
        if r is None:
            continue
        r, g = np.asarray(r), np.asarray(g)
        if r.dtype != g.dtype or r.shape != g.shape or not np.array_equal(r, g):
            bad.append(name)
    if len(ref.from_sep_smaller_by_level) != len(got.from_sep_smaller_by_level):
        bad.append("from_sep_smaller_by_level (length)")
    else:
        for lev, (r, g) in enumerate(zip(ref.from_sep_smaller_by_level,
                                         got.from_sep_smaller_by_level)):
            for name in ("starts", "lists", "nonempty_indices", "compressed_indices"):
                a, b = np.asarray(getattr(r, name)), np.asarray(getattr(g, name))
                if a.dtype != b.dtype or a.shape != b.shape or not np.array_equal(a, b):
                    bad.append(f"from_sep_smaller_by_level[{lev}].{name}")
            for name in ("count", "num_nonempty_lists"):
                if int(getattr(r, name)) != int(getattr(g, name)):
                    bad.append(f"from_sep_smaller_by_level[{lev}].{name}")
            a = np.asarray(ref.target_boxes_sep_smaller_by_source_level[lev])
            b = np.asarray(got.target_boxes_sep_smaller_by_source_level[lev])
            if a.dtype != b.dtype or not np.array_equal(a, b):
                bad.append(f"target_boxes_sep_smaller_by_source_level[{lev}]")
    return bad


# ---- digests ---------------------------------------------------------------

def _digest(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(f"{a.dtype.str}{a.shape}".encode())
        h.update(a.tobytes())
    return h.hexdigest()[:12]


def tree_digests(tree):
    """Short sha256 per ``Tree`` field (dtype, shape and bytes), the fields and slicing of
    ``tree_mismatches``.  Used to compare against reference-generated values that cannot travel
    to the GPU box as arrays (``tests/golden/refexec_digests.json``)."""
    nb = tree.nboxes
    out = {"nboxes": int(nb), "nlevels": int(tree.nlevels),
           "root_extent": _digest(np.asarray(tree.root_extent)),
           "bounding_box": _digest(*[np.asarray(b) for b in tree.bounding_box])}
    for name in TREE_INT_FIELDS:
        a = np.asarray(getattr(tree, name))
        out[name] = _digest(a[:nb] if name.startswith("box_") else a)
    for name in TREE_PADDED_INT_FIELDS + TREE_PADDED_FLOAT_FIELDS:
        out[name] = _digest(np.asarray(getattr(tree, name)))
    for name in ("sources", "targets"):
        out[name] = _digest(*[np.asarray(x) for x in getattr(tree, name)])
    for name in ("source_radii", "target_radii"):
        v = getattr(tree, name)
        out[name] = None if v is None else _digest(np.asarray(v))
    return out


def trav_digests(trav):
    out = {}
    for name in TRAV_FIELDS:
        v = getattr(trav, name)
        out[name] = None if v is None else _digest(np.asarray(v))
    by_level = []
    for lev, bl in enumerate(trav.from_sep_smaller_by_level):
        by_level.append(_digest(
            np.asarray(bl.starts), np.asarray(bl.lists), np.asarray(bl.nonempty_indices),
            np.asarray(bl.compressed_indices),
            np.asarray([int(bl.count), int(bl.num_nonempty_lists)]),
            np.asarray(trav.target_boxes_sep_smaller_by_source_level[lev])))
    out["from_sep_smaller_by_level"] = by_level
    return out


DIST_MASK_FIELDS = ("responsible_boxes", "ancestor_boxes", "point_src_boxes", "multipole_src_boxes")
DIST_LOCAL_TREE_FIELDS = (
    "box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
    "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul", "box_flags",
    "box_parent_ids", "box_levels", "box_child_ids", "box_to_user_rank_starts",
    "box_to_user_rank_lists", "responsible_boxes_mask", "ancestor_mask")


def distributed_rank_digests(resp, masks, local_tree, src_idx, tgt_idx, local_trav, nboxes):
    """Digests of one rank's distributed-setup outputs.  *masks*, *local_tree*: mappings
    field -> numpy array (``sources`` / ``targets``: lists of arrays, ``target_radii`` or None)."""
    out = {"responsible_boxes_list": _digest(np.asarray(resp)),
           "src_idx": _digest(np.asarray(src_idx)), "tgt_idx": _digest(np.asarray(tgt_idx))}
    for f in DIST_MASK_FIELDS:
        out["mask." + f] = _digest(np.asarray(masks[f]))
    for f in DIST_LOCAL_TREE_FIELDS:
        a = np.asarray(local_tree[f])
        if f in ("box_parent_ids", "box_levels", "box_flags") or f.startswith("box_source") \
                or f.startswith("box_target"):
            a = a[:nboxes]
        out["tree." + f] = _digest(a)
    out["tree.sources"] = _digest(*[np.asarray(x) for x in local_tree["sources"]])
    out["tree.targets"] = _digest(*[np.asarray(x) for x in local_tree["targets"]])
    tr = local_tree.get("target_radii")
    out["tree.target_radii"] = None if tr is None else _digest(np.asarray(tr))
    for k, v in trav_digests(local_trav).items():
        out["trav." + k] = v
    return out


def digest_mismatches(ref: dict, got: dict):
    return [k for k in ref if ref[k] != got.get(k)]


def reference_case_key(case, quick):
    """Key of a ``tests/gpu_sweep.py`` case in ``tests/golden/refexec_digests.json``."""
    return (f"{'quick' if quick else 'full'}:{case['dims']}d-{np.dtype(case['dtype']).name}-"
            f"{case['name']}-n{case['n']}")


_REFERENCE_DIGESTS = None


def reference_digests():
    """Per-field digests of what the REFERENCE ITSELF produced for the sweep cases
    (``tests/golden/make_refexec_golden.py``)."""
    global _REFERENCE_DIGESTS
    if _REFERENCE_DIGESTS is None:
        import json
        import os
        path = os.path.join(os.path.dirname(__file__), "golden", "refexec_digests.json")
        with open(path) as f:
            _REFERENCE_DIGESTS = json.load(f)
    return _REFERENCE_DIGESTS


# ---- input recipes ----------------------------------------------------------

def normal_particles(n, dims, dtype, seed=15):
    """``make_normal_particle_array`` (boxtree/tools.py:114-119)."""
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(n).astype(dtype) for _ in range(dims)]


def uniform_particles(n, dims, dtype, seed=15):
    """``default_rng(seed).random((dims, n))``."""
    pts = np.random.default_rng(seed).random((dims, n))
    return [np.ascontiguousarray(pts[i]).astype(dtype) for i in range(dims)]


def plummer_particles(n, dtype, seed=15, a=1.0):
    """Plummer sphere, r <= 32 a, generated in fp64 then cast."""
    rng = np.random.default_rng(seed)
    out = np.empty((3, 0))
    while out.shape[1] < n:
        m = int((n - out.shape[1]) * 1.05) + 16
        u = rng.random(m)
        r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        keep = r <= 32 * a
        r = r[keep]
        v = rng.standard_normal((3, len(r)))
        v /= np.linalg.norm(v, axis=0)
        out = np.concatenate([out, v * r], axis=1)
    out = out[:, :n]
    return [np.ascontiguousarray(out[i]).astype(dtype) for i in range(3)]


def config3_inputs(nsources, ntargets, dtype=np.float64):
    """BASELINE.md config 3 recipe at a chosen size."""
    s = np.random.default_rng(12).random((3, nsources))
    t = np.random.default_rng(19).random((3, ntargets))
    radii = 2 ** np.random.default_rng(13).uniform(-14, -4, ntargets)
    return ([np.ascontiguousarray(s[i]).astype(dtype) for i in range(3)],
            [np.ascontiguousarray(t[i]).astype(dtype) for i in range(3)],
            radii.astype(dtype))


# ---- argument validation ----------------------------------------------------

def error_cases():
    """``(name, particles, TreeBuilder kwargs, traversal ctor kwargs or None, exception name)``:
    invalid calls and the exception type the REFERENCE raises for them (checked against the
    reference's own code in ``tests/test_refexec.py``, against the CUDA path in
    ``tests/test_gpu_parity.py``).  Arrays are numpy; ``targets`` is a list of arrays."""
    src = normal_particles(100, 2, np.float64)
    ones = np.ones(100)
    pts = np.array([0.5] * 11 + [1.0])
    return [
        ("unknown kind", src, dict(kind="bogus", max_particles_in_box=10), None, "ValueError"),
        ("no refinement criterion", src, dict(), None, "ValueError"),
        ("both criteria", src, dict(max_particles_in_box=10, refine_weights=np.ones(100, np.int32),
                                    max_leaf_refine_weight=5), None, "ValueError"),
        ("radii without targets", src, dict(source_radii=ones, max_particles_in_box=10), None,
         "ValueError"),
        ("radii without stick_out_factor", src, dict(targets=src, target_radii=ones,
                                                     max_particles_in_box=10), None, "ValueError"),
        ("radii dtype", src, dict(targets=src, target_radii=np.ones(100, np.float32),
                                  stick_out_factor=0.1, max_particles_in_box=10), None, "TypeError"),
        ("weights dtype", src, dict(refine_weights=np.ones(100, np.int64),
                                    max_leaf_refine_weight=5), None, "TypeError"),
        ("weight above the leaf maximum", src, dict(refine_weights=np.full(100, 7, np.int32),
                                                    max_leaf_refine_weight=5), None, "ValueError"),
        ("radii shape", src, dict(targets=src, target_radii=np.ones(99), stick_out_factor=0.1,
                                  max_particles_in_box=10), None, "ValueError"),
        ("unknown extent norm", src, dict(targets=src, target_radii=ones, stick_out_factor=0.1,
                                          extent_norm="l7", max_particles_in_box=10), None,
         "ValueError"),
        ("coincident points", [pts, pts.copy()], dict(max_particles_in_box=10), None,
         "MaxLevelsExceeded"),
        ("traversal of an unpruned tree", src, dict(max_particles_in_box=10, skip_prune=True), {},
         "ValueError"),
        ("unknown from_sep_smaller_crit", src, dict(max_particles_in_box=10),
         dict(from_sep_smaller_crit="bogus"), "ValueError"),
        ("static_linf with l2 extents", src, dict(targets=src, target_radii=np.full(100, 1e-3),
                                                  stick_out_factor=0.1, extent_norm="l2",
                                                  max_particles_in_box=10),
         dict(from_sep_smaller_crit="static_linf"), "ValueError"),
        ("traversal with source extents", src, dict(targets=src, source_radii=np.full(100, 1e-3),
                                                    stick_out_factor=0.1, max_particles_in_box=10),
         {}, "NotImplementedError"),
    ]
