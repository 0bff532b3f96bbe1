"""Regenerates the committed golden fixtures of the BASELINE configurations (reduced sizes).

    python tests/golden/make_golden.py

Where ``/root/reference`` is mounted (this container) the arrays come from the REFERENCE ITSELF:
its ``TreeBuilder`` and ``FMMTraversalBuilder`` executed on the CPU through ``tests/refexec``
(host code and kernel text unmodified; pyopencl's kernel-launch classes restated).  The oracle
must reproduce every byte, or the script stops.  Elsewhere the script falls back to the oracle.

* ``config1_2d_1e4.npz``: every Tree / FMMTraversalInfo array of BASELINE config 1
  (2-D, 1e4 uniform fp64 particles, max 30 per box, adaptive, sources are targets).
* ``digests.json``: sha256 of every output array for a list of larger / other cases.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.traversal import build_traversal  # noqa: E402
from oracle.tree_build import build_tree  # noqa: E402
from tests.parity_util import (TRAV_FIELDS, TREE_INT_FIELDS, TREE_PADDED_FLOAT_FIELDS,  # noqa: E402
                               TREE_PADDED_INT_FIELDS, config3_inputs, normal_particles,
                               plummer_particles, uniform_particles)


def flatten(tree, trav):
    """name -> numpy array for every output array of the path."""
    out = {}
    nb = tree.nboxes
    for name in TREE_INT_FIELDS:
        a = np.asarray(getattr(tree, name))
        out["tree." + name] = a[:nb] if name.startswith("box_") else a
    for name in TREE_PADDED_INT_FIELDS + TREE_PADDED_FLOAT_FIELDS:
        out["tree." + name] = np.asarray(getattr(tree, name))
    for ax in range(tree.dimensions):
        out[f"tree.sources[{ax}]"] = np.asarray(tree.sources[ax])
        out[f"tree.targets[{ax}]"] = np.asarray(tree.targets[ax])
    out["tree.root_extent"] = np.asarray(tree.root_extent).reshape(1)
    out["tree.bounding_box"] = np.stack([np.asarray(b) for b in tree.bounding_box])
    if trav is not None:
        for name in TRAV_FIELDS:
            v = getattr(trav, name)
            if v is not None:
                out["trav." + name] = np.asarray(v)
        for lev, bl in enumerate(trav.from_sep_smaller_by_level):
            out[f"trav.from_sep_smaller_by_level[{lev}].starts"] = np.asarray(bl.starts)
            out[f"trav.from_sep_smaller_by_level[{lev}].lists"] = np.asarray(bl.lists)
            out[f"trav.from_sep_smaller_by_level[{lev}].nonempty_indices"] = \
                np.asarray(bl.nonempty_indices)
            out[f"trav.target_boxes_sep_smaller_by_source_level[{lev}]"] = \
                np.asarray(trav.target_boxes_sep_smaller_by_source_level[lev])
    return out


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def digest_cases():
    """name -> (sources, tree kwargs, traversal kwargs)"""
    cases = {}
    cases["config1_2d_1e4"] = (uniform_particles(10_000, 2, np.float64), dict(
        max_particles_in_box=30), {})
    cases["config2_3d_1e5"] = (uniform_particles(100_000, 3, np.float64), dict(
        max_particles_in_box=30), {})
    s, t, r = config3_inputs(50_000, 50_000)
    cases["config3_3d_1e5"] = (s, dict(max_particles_in_box=30, targets=t, target_radii=r,
                                       stick_out_factor=0.25, extent_norm="linf",
                                       kind="adaptive-level-restricted"), {})
    s, t, r = config3_inputs(20_000, 20_000)        # what __graft_entry__.smoke() builds
    cases["config3_3d_4e4_smoke"] = (s, dict(max_particles_in_box=30, targets=t, target_radii=r,
                                             stick_out_factor=0.25, extent_norm="linf",
                                             kind="adaptive-level-restricted"), {})
    cases["config4_plummer_1e5_f32"] = (plummer_particles(100_000, np.float32), dict(
        max_particles_in_box=30), {})
    cases["normal_3d_f32_2away"] = (normal_particles(30_000, 3, np.float32), dict(
        max_particles_in_box=30), dict(well_sep_is_n_away=2))
    return cases


def produce(src, tkw, vkw):
    """The reference's own run when it can be executed here, checked against the oracle."""
    tree = build_tree(src, **tkw)
    trav = build_traversal(tree, **vkw)
    flat = flatten(tree, trav)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
    import refexec
    if refexec.available():
        from refexec.run import reference_traversal, reference_tree
        rtree = reference_tree(src, **tkw)
        rflat = flatten(rtree, reference_traversal(rtree, **vkw))
        assert set(rflat) == set(flat)
        for k in rflat:
            assert rflat[k].dtype == flat[k].dtype and rflat[k].shape == flat[k].shape, k
            assert np.array_equal(rflat[k].view(np.uint8), flat[k].view(np.uint8)), k
        return rtree, rflat, "reference"
    return tree, flat, "oracle"


def main():
    src, tkw, vkw = digest_cases()["config1_2d_1e4"]
    _, flat, origin = produce(src, tkw, vkw)
    np.savez_compressed(os.path.join(HERE, "config1_2d_1e4.npz"), **flat)
    digests = {}
    for name, (src, tkw, vkw) in digest_cases().items():
        tree, flat, origin = produce(src, tkw, vkw)
        digests[name] = {k: digest(v) for k, v in flat.items()}
        print(name, "nboxes", tree.nboxes, "nlevels", tree.nlevels, "from the", origin)
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
