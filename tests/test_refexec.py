"""The oracle pinned against the REFERENCE ITSELF.

``tests/refexec`` executes the reference's ``TreeBuilder.__call__`` and
``FMMTraversalBuilder.__call__`` -- host code and kernel templates unmodified, from where they lie
under ``/root/reference`` -- on the CPU, with stand-ins for the third-party modules that are not
installed (pyopencl, mako, pytools, arraycontext, cgen, pymbolic).  Besides unit tests of the
stand-ins themselves, two kinds of test:

* live (skipped where ``/root/reference`` is absent): the reference runs here and its arrays are
  compared with the oracle's, bit for bit including dtypes;
* committed: ``tests/golden/refexec_digests.json`` / ``refexec_*.npz`` hold what the reference
  produced for every sweep case (``tests/golden/make_refexec_golden.py``); the oracle must
  reproduce them.  (The CUDA path is held to the same files in ``tests/test_gpu_parity.py`` and
  ``tests/test_gpu_distributed.py``.)

Also live: the distributed setup, area queries, particle filters, point-source linking, translation
and rotation classes, ``merge_close_lists``, the device cost model, the error table, and a selection
of the reference's own tests.  ``tests/refexec/fuzz.py`` repeats the comparisons on random cases.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))

from oracle.traversal import build_traversal                    # noqa: E402
from oracle.tree_build import MaxLevelsExceeded, build_tree     # noqa: E402
from tests.gpu_sweep import make_cases, make_inputs             # noqa: E402
from tests.parity_util import (digest_mismatches, reference_case_key,  # noqa: E402
                               reference_digests, trav_digests, trav_mismatches, tree_digests,
                               tree_mismatches)

import refexec                                                   # noqa: E402
from refexec.minimako import Template                            # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_reference = pytest.mark.skipif(not refexec.available(),
                                     reason="/root/reference is not mounted")


def _trav_kwargs(case):
    tkw = dict(case.get("trav") or {})
    ctor = {k: tkw.pop(k) for k in ("well_sep_is_n_away", "from_sep_smaller_crit") if k in tkw}
    return {**ctor, **tkw}


# {{{ the template renderer

def test_minimako_subset():
    t = Template(r"""
<%def name="decl(name, n=2)">
    int ${name}[${n}];
</%def>
## a comment line
%if flag:
    ${decl("a")}
%elif other:
    nothing
%else:
    ${decl("b", n=3)}
%endif
%for i in range(2):
    x${i} = ${ {"k": i}["k"] + 1 };
%endfor
<% y = 5 %>
#define M(a) \
    (a + ${y})
100 %% 7
""", strict_undefined=True)
    out = t.render(flag=False, other=False)
    assert "int b[3];" in out and "int a[2]" not in out and "nothing" not in out
    assert "x0 = 1;" in out and "x1 = 2;" in out
    assert "#define M(a)     (a + 5)" in out
    assert "comment" not in out
    with pytest.raises(NameError):
        Template("${missing}").render()

# }}}


# {{{ the kernel-launch stand-ins (pyopencl's published semantics, restated in fakecl.py)

def test_fakecl_list_of_lists_builder():
    from refexec import fakecl
    knl = fakecl.ListOfListsBuilder(
        None, [("mult", np.int32), ("big", np.int64)], """
        void generate(LIST_ARG_DECL USER_ARG_DECL index_type i)
        {
            for (int j = 0; j < reps[i]; ++j) APPEND_mult(i * scale);
            if (reps[i] > 2) APPEND_big(1000 + i);
        }""", [fakecl.VectorArg(np.int32, "reps"), fakecl.ScalarArg(np.int32, "scale")],
        eliminate_empty_output_lists=["big"])
    reps = fakecl.Array(np.array([2, 0, 3, 1, 0, 4], np.int32))
    result, _ = knl(None, 6, reps, 10)
    m, b = result["mult"], result["big"]
    assert m.count == 10 and np.array_equal(m.starts._a, [0, 2, 2, 5, 6, 6, 10])
    assert np.array_equal(m.lists._a, [0, 0, 20, 20, 20, 30, 50, 50, 50, 50])
    assert m.starts.dtype == np.int32 and m.lists.dtype == np.int32
    assert b.count == 2 and b.num_nonempty_lists == 2 and b.lists.dtype == np.int64
    assert np.array_equal(b.starts._a, [0, 1, 2]) and np.array_equal(b.lists._a, [1002, 1005])
    assert np.array_equal(b.nonempty_indices._a, [2, 5])
    assert np.array_equal(b.compressed_indices._a, [0, 0, 0, 1, 1, 1, 2])
    result, _ = knl(None, 6, reps, 10, omit_lists=("mult",))
    assert result["mult"].lists is None and result["big"].count == 2


def test_fakecl_scan_elementwise_reduction():
    from refexec import fakecl
    scan = fakecl.GenericScanKernel(
        None, np.int32, arguments="int *x, int *seg, int *excl, int *incl, int *total",
        input_expr="x[i]", scan_expr="across_seg_boundary ? b : a + b", neutral="0",
        is_segment_start_expr="seg[i]",
        output_statement="excl[i] = prev_item; incl[i] = item; if (i == N - 1) *total = last_item;")
    x = fakecl.Array(np.array([3, 1, 4, 1, 5, 9], np.int32))
    seg = fakecl.Array(np.array([0, 0, 1, 0, 0, 1], np.int32))
    excl, incl = fakecl.Array(np.zeros(6, np.int32)), fakecl.Array(np.zeros(6, np.int32))
    total = fakecl.Array(np.zeros(1, np.int32))
    scan(x, seg, excl, incl, total)
    assert np.array_equal(incl._a, [3, 4, 4, 5, 10, 9])
    assert np.array_equal(excl._a, [0, 3, 0, 4, 5, 0]) and total._a[0] == 9
    elwise = fakecl.ElementwiseKernel(None, "double *y, double a, int *src",
                                      "if (src[i] < 0) PYOPENCL_ELWISE_CONTINUE; y[i] = a * src[i]")
    y = fakecl.Array(np.full(6, -1.0))
    elwise(y, 0.5, fakecl.Array(np.array([2, -1, 4, 6, 8, 10], np.int32)), range=slice(1, 5))
    assert np.array_equal(y._a, [-1.0, -1.0, 2.0, 3.0, 4.0, -1.0])
    red = fakecl.ReductionKernel(None, np.float32, neutral="-FLT_MAX", reduce_expr="fmax(a, b)",
                                 map_expr="v[i] * v[i]", arguments="float *v")
    assert red(fakecl.Array(np.array([1, -3, 2], np.float32))).get() == np.float32(9)
    starts, lists, _ = fakecl.KeyValueSorter(None)(
        None, fakecl.Array(np.array([2, 0, 2, 1, 0])), fakecl.Array(np.arange(5)), 4, np.int32)
    assert np.array_equal(starts._a, [0, 2, 3, 5, 5]) and np.array_equal(lists._a, [1, 4, 3, 0, 2])


def test_cl_shim_keeps_float_arithmetic_in_float():
    """The finding of the first fuzz run: ``sqrt(float)`` must not be evaluated in double."""
    from refexec import fakecl
    knl = fakecl.ElementwiseKernel(
        None, "float *x, float *out, int *size",
        "out[i] = sqrt(x[i]) - x[i]; size[i] = sizeof(sqrt(x[i])) * 100 + sizeof(fmin(x[i], x[i])) "
        "* 10 + sizeof(rint(x[i]));")
    x = np.array([0.0034957202, 2.0, 1e-3], np.float32)
    out, size = fakecl.Array(np.zeros(3, np.float32)), fakecl.Array(np.zeros(3, np.int32))
    knl(fakecl.Array(x), out, size)
    assert np.all(size._a == 444)
    assert np.array_equal(out._a, np.sqrt(x) - x)

# }}}


# {{{ live: the reference runs here

_LIVE = [c for c in make_cases(quick=True) if (c["dims"], np.dtype(c["dtype"]).name, c["name"]) in {
    (2, "float64", "adaptive"), (3, "float32", "lr"), (3, "float64", "src-tgt"),
    (2, "float32", "weights"), (3, "float64", "ext-l2-static_l2-n2"), (3, "float32", "ext-lr-n1"),
    (2, "float64", "ext-minsrc"), (3, "float64", "nsep3"), (2, "float64", "coincident"),
    (3, "float64", "non-adaptive"), (2, "float64", "user-bbox"), (3, "float32", "tiny-guess")}]


@needs_reference
@pytest.mark.parametrize(
    "case", _LIVE, ids=[f"{c['dims']}d-{np.dtype(c['dtype']).name}-{c['name']}" for c in _LIVE])
def test_reference_run_matches_oracle(case):
    from refexec.run import reference_traversal, reference_tree
    src, kw = make_inputs(case)
    if case.get("expect_max_levels"):
        with pytest.raises(MaxLevelsExceeded):
            build_tree(src, **kw)
        with pytest.raises(Exception) as ei:
            reference_tree(src, **kw)
        assert type(ei.value).__name__ == "MaxLevelsExceeded"
        return
    ref_tree = reference_tree(src, **kw)
    tree = build_tree(src, **kw)
    assert tree_mismatches(ref_tree, tree) == []
    if case.get("trav", {}) is None:
        return
    tkw = _trav_kwargs(case)
    # the reference's traversal of the reference's tree vs the oracle's of the oracle's
    ref_trav = reference_traversal(ref_tree, **tkw)
    trav = build_traversal(tree, **tkw)
    assert trav_mismatches(ref_trav, trav) == []
    assert trav.from_sep_siblings_lists.size > 0
    # the comparison is not vacuous
    broken = np.array(trav.neighbor_source_boxes_lists, copy=True)
    broken[0] ^= 1
    trav.neighbor_source_boxes_lists = broken
    assert trav_mismatches(ref_trav, trav) == ["neighbor_source_boxes_lists"]


@needs_reference
@pytest.mark.parametrize("n_away", [1, 2])
def test_reference_merge_close_lists_matches_oracle(n_away):
    """``FMMTraversalInfo.merge_close_lists`` (the reference's ``_ListMerger`` kernels)."""
    from oracle.traversal import merge_close_lists
    from refexec.run import reference_traversal
    from tests.parity_util import config3_inputs
    src, tgt, radii = config3_inputs(3000, 3000)
    tree = build_tree(src, targets=tgt, target_radii=radii * 8, stick_out_factor=0.25,
                      max_particles_in_box=20, extent_norm="linf")
    ref = reference_traversal(tree, well_sep_is_n_away=n_away, merge_close_lists=True)
    got = merge_close_lists(build_traversal(tree, well_sep_is_n_away=n_away))
    assert ref.from_sep_close_smaller_starts is None and ref.from_sep_close_bigger_lists is None
    assert trav_mismatches(ref, got) == []
    plain = build_traversal(tree, well_sep_is_n_away=n_away)
    assert len(got.neighbor_source_boxes_lists) > len(plain.neighbor_source_boxes_lists)


@needs_reference
def test_reference_run_with_box_masks():
    """``source_boxes_mask`` / ``source_parent_boxes_mask`` (the distributed code's entry)."""
    from refexec.run import reference_traversal
    rng = np.random.default_rng(5)
    src = [rng.standard_normal(3000) for _ in range(3)]
    tree = build_tree(src, max_particles_in_box=20)
    sm = (rng.random(tree.nboxes) < 0.6).astype(np.int8)
    pm = (rng.random(tree.nboxes) < 0.7).astype(np.int8)
    ref = reference_traversal(tree, source_boxes_mask=sm, source_parent_boxes_mask=pm)
    got = build_traversal(tree, source_boxes_mask=sm, source_parent_boxes_mask=pm)
    assert trav_mismatches(ref, got) == []

# }}}


# {{{ the reference's own tests, executed through refexec

@needs_reference
def test_reference_own_tests_pass_under_refexec():
    """A small selection of ``/root/reference/test`` (the whole files are run by
    ``tests/refexec/run_reference_tests.py``; log in ``profiles/``): the reference's assertions
    about its own trees and traversals hold when its kernels are executed by the stand-ins."""
    from refexec.run_reference_tests import run
    proc = run(["/root/reference/test/test_traversal.py::test_tree_connectivity",
                "/root/reference/test/test_tree.py::test_bounding_box",
                "/root/reference/test/test_tree.py::test_leaves_to_balls_query",
                "-k", "not 3"])
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert " passed" in proc.stdout and "failed" not in proc.stdout

# }}}


# {{{ argument validation

@needs_reference
def test_reference_raises_what_the_error_table_says():
    """``tests.parity_util.error_cases`` (the table the CUDA path is held to) against the
    reference's own ``TreeBuilder`` / ``FMMTraversalBuilder``; and the oracle agrees."""
    from refexec.run import Session
    from tests.parity_util import error_cases
    for name, particles, tkw, vkw, exc_name in error_cases():
        with Session() as s:
            with pytest.raises(Exception) as ei:
                tree = s.tree(particles, **tkw)
                assert vkw is not None, f"{name}: the reference built the tree"
                s.traversal(tree, **dict(vkw))
            assert type(ei.value).__name__ == exc_name, (name, repr(ei.value))
        with pytest.raises(Exception) as ei:
            tree = build_tree(particles, **tkw)
            assert vkw is not None, f"{name}: the oracle built the tree"
            build_traversal(tree, **dict(vkw))
        assert type(ei.value).__name__ == exc_name, (name, "oracle", repr(ei.value))

# }}}


# {{{ distributed setup

from refexec.compare import oracle_distributed_digests as _oracle_distributed_digests  # noqa: E402


@pytest.mark.parametrize("nranks", [1, 3, 4])
@pytest.mark.parametrize("name", ["points", "points2d-f32-2away", "config3"])
def test_oracle_distributed_setup_matches_reference_digests(name, nranks):
    """``tests/golden/refexec_distributed_digests.json``: partition, box masks, local trees,
    particle indices and local traversals of every rank as the reference's own
    ``boxtree/distributed`` code produced them."""
    import json
    from tests.dist_cases import CASES
    with open(os.path.join(GOLDEN, "refexec_distributed_digests.json")) as f:
        want = json.load(f)[f"{name}:{nranks}"]
    got = _oracle_distributed_digests(*CASES[name](), nranks)
    assert len(got) == len(want) == nranks
    for r in range(nranks):
        assert digest_mismatches(want[r], got[r]) == [], r


@needs_reference
def test_reference_distributed_setup_live():
    """The reference's distributed setup runs here (3 ranks as threads) and the oracle matches."""
    from refexec.run import reference_distributed_setup
    from tests.dist_cases import box_cost
    from tests.parity_util import DIST_LOCAL_TREE_FIELDS, distributed_rank_digests
    src, tgt, radii = __import__("tests.parity_util", fromlist=["x"]).config3_inputs(4000, 4000)
    tkw = dict(max_particles_in_box=30, targets=tgt, target_radii=radii, stick_out_factor=0.25,
               extent_norm="linf", kind="adaptive-level-restricted")
    tree, trav, ranks = reference_distributed_setup(src, tkw, {}, 3, box_cost)
    assert tree_mismatches(tree, build_tree(src, **tkw)) == []
    want = [distributed_rank_digests(r["responsible_boxes_list"], r["masks"], r["local_tree"],
                                     r["src_idx"], r["tgt_idx"], r["local_trav"], tree.nboxes)
            for r in ranks]
    got = _oracle_distributed_digests(src, tkw, {}, 3)
    for r in range(3):
        assert digest_mismatches(want[r], got[r]) == [], r
    assert sum(len(r["tgt_idx"]) for r in ranks) == 4000
    assert set(DIST_LOCAL_TREE_FIELDS) <= set(ranks[0]["local_tree"])

# }}}


# {{{ the widened rows (N1-N4), live

def _same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and \
        np.array_equal(a.view(np.uint8), b.view(np.uint8))


@needs_reference
@pytest.mark.parametrize("dims,dtype,kind", [(2, np.float64, "adaptive"),
                                             (3, np.float32, "adaptive-level-restricted"),
                                             (3, np.float64, "adaptive")])
def test_reference_area_queries_match_oracle(dims, dtype, kind):
    """PeerListFinder, AreaQueryBuilder, LeavesToBallsLookupBuilder, SpaceInvaderQueryBuilder of
    ``boxtree/area_query.py`` executed here vs the oracle's restatement."""
    from oracle import traversal as ot
    from refexec.run import reference_area_queries
    from tests.parity_util import normal_particles
    tree = build_tree(normal_particles(4000, dims, dtype), max_particles_in_box=20, kind=kind)
    rng = np.random.default_rng(7)
    centers = [rng.normal(size=500).astype(dtype) for _ in range(dims)]
    radii = (0.3 * 2 ** rng.uniform(-6, 0, 500)).astype(dtype)
    ref = reference_area_queries(tree, centers, radii)
    got = dict(zip(("peer_list_starts", "peer_lists"), ot.find_peer_lists(tree)))
    got.update(zip(("leaves_near_ball_starts", "leaves_near_ball_lists"),
                   ot.area_query(tree, centers, radii)))
    got.update(zip(("balls_near_box_starts", "balls_near_box_lists"),
                   ot.leaves_to_balls(tree, centers, radii)))
    got["outer_space_invader_dists"] = ot.space_invader_query(tree, centers, radii)
    assert ref["leaves_near_ball_lists"].size > 500
    for k, v in ref.items():
        assert _same(v, got[k]), k


@needs_reference
def test_reference_particle_filter_matches_oracle():
    from oracle import particle_filter as opf
    from refexec.run import reference_particle_filter
    from tests.parity_util import normal_particles
    src = normal_particles(5000, 3, np.float64, seed=12)
    tgt = normal_particles(4000, 3, np.float64, seed=19)
    tkw = dict(targets=tgt, max_particles_in_box=30)
    flags = (np.random.default_rng(5).random(4000) < 0.3).astype(np.int8)
    ref = reference_particle_filter(src, tkw, flags)
    tree = build_tree(src, **tkw)
    n, starts, lists = opf.filter_target_lists_in_user_order(tree, flags)
    assert n == ref["user.nfiltered_targets"]
    assert _same(starts, ref["user.target_starts"]) and _same(lists, ref["user.target_lists"])
    n, bstart, bcount, targets, unfiltered = opf.filter_target_lists_in_tree_order(tree, flags)
    assert n == ref["tree.nfiltered_targets"]
    # boxes that start at ntargets (empty, at the very end): the reference's index kernel reads
    # filtered_from_unfiltered_target_index[ntargets], one past the end of the array
    # (tree_build_kernels.py:1994-1996), so its value there is undefined; ours is nfiltered
    defined = np.asarray(tree.box_target_starts) < tree.ntargets
    assert np.array_equal(bstart[defined], ref["tree.box_target_starts"][defined])
    assert np.all(bstart[~defined] == n)
    assert _same(bcount, ref["tree.box_target_counts_nonchild"])
    assert _same(unfiltered, ref["tree.unfiltered_from_filtered_target_indices"])
    assert all(_same(a, b) for a, b in zip(targets, ref["tree.targets"]))


@needs_reference
def test_reference_link_point_sources_matches_oracle():
    from oracle import particle_filter as opf
    from refexec.run import reference_link_point_sources
    from tests.parity_util import normal_particles
    ns = 3000
    src = normal_particles(ns, 3, np.float64)
    rng = np.random.default_rng(7)
    radii = 0.05 * 2 ** rng.uniform(-10, 0, ns)
    tkw = dict(max_particles_in_box=30, targets=normal_particles(2000, 3, np.float64, seed=19),
               source_radii=radii, stick_out_factor=0.25, extent_norm="linf")
    counts = rng.integers(1, 5, ns)
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    owner = np.repeat(np.arange(ns), counts)
    pts = [src[a][owner] + radii[owner] * rng.uniform(-1, 1, len(owner)) for a in range(3)]
    ref = reference_link_point_sources(src, tkw, starts, pts)
    want = opf.link_point_sources(build_tree(src, **tkw), starts, pts)
    assert want["npoint_sources"] == ref["npoint_sources"] == len(owner)
    for k, v in want.items():
        if k == "point_sources":
            assert all(_same(a, b) for a, b in zip(v, ref[k]))
        elif k != "npoint_sources":
            assert _same(v, ref[k]), k


@needs_reference
@pytest.mark.parametrize("dims,dtype,n_away,per_level", [
    (2, np.float64, 1, True), (3, np.float64, 2, True), (3, np.float32, 1, False)])
def test_reference_translation_classes_match_oracle(dims, dtype, n_away, per_level):
    from oracle import translation_classes as otc
    from refexec.run import reference_translation_classes
    from tests.parity_util import normal_particles
    src = normal_particles(4000, dims, dtype)
    ref = reference_translation_classes(src, dict(max_particles_in_box=30),
                                        dict(well_sep_is_n_away=n_away), per_level)
    tree = build_tree(src, max_particles_in_box=30)
    trav = build_traversal(tree, well_sep_is_n_away=n_away)
    classes, distances, level_starts = otc.translation_classes(trav, tree, per_level)
    assert _same(classes, ref["from_sep_siblings_translation_classes"])
    ref_starts = ref["from_sep_siblings_translation_classes_level_starts"]
    ref_dist = ref["from_sep_siblings_translation_class_to_distance_vector"]
    count = int(level_starts[-1])
    assert count == int(ref_starts[-1]) and ref_starts[0] == level_starts[0] == 0
    # the reference allocates both with np.empty: only the used classes' columns, and (when the
    # classes are not per level) only the first and last start, are ever written
    assert _same(distances[:, :count], ref_dist[:, :count])
    if per_level:
        assert _same(level_starts, ref_starts)


@needs_reference
@pytest.mark.parametrize("n_away", [1, 2])
def test_reference_rotation_classes_match_oracle(n_away):
    from oracle import translation_classes as otc
    from refexec.run import reference_rotation_classes
    from tests.parity_util import normal_particles
    src = normal_particles(4000, 3, np.float64)
    ref = reference_rotation_classes(src, dict(max_particles_in_box=30),
                                     dict(well_sep_is_n_away=n_away))
    tree = build_tree(src, max_particles_in_box=30)
    classes, angles = otc.rotation_classes(build_traversal(tree, well_sep_is_n_away=n_away), tree)
    assert _same(classes, ref["from_sep_siblings_rotation_classes"])
    assert _same(angles, ref["from_sep_siblings_rotation_class_to_angle"])



@needs_reference
@pytest.mark.parametrize("factory", ["make_pde_aware_translation_cost_model",
                                     "make_taylor_translation_cost_model"])
@pytest.mark.parametrize("name", ["3d-points", "2d-points-2away", "3d-config3"])
def test_reference_device_cost_model_matches_the_golden(name, factory):
    """``tests/golden/cost_model.json`` was written by the reference's ``_PythonFMMCostModel``
    on the oracle's traversal; here the reference's device ``FMMCostModel`` (its OpenCL kernels,
    executed through refexec on the reference's own tree and traversal) must give the same."""
    import json
    from refexec.run import reference_cost_model
    from tests.golden.make_cost_golden import CALIBRATION, cases, level_to_order
    with open(os.path.join(GOLDEN, "cost_model.json")) as f:
        want = json.load(f)[f"{name}/{factory}"]
    src, tkw, vkw = cases()[name]
    nlevels = build_tree(src, **tkw).nlevels
    per_box, per_stage = reference_cost_model(src, tkw, vkw, factory, level_to_order(nlevels),
                                              CALIBRATION)
    assert len(per_box) == want["nboxes"]
    assert np.allclose(per_box[:64], want["per_box_head"], rtol=1e-12, atol=0)
    assert np.allclose(per_box[::97], want["per_box_every_97th"], rtol=1e-12, atol=0)
    assert np.isclose(float(np.sum(per_box)), want["per_box_sum"], rtol=1e-12)
    assert set(per_stage) == set(want["per_stage"])
    for k, v in want["per_stage"].items():
        assert np.isclose(per_stage[k], v, rtol=1e-12), k

# }}}


# {{{ committed outputs of the reference

_QUICK = make_cases(quick=True)


@pytest.mark.parametrize(
    "case", _QUICK, ids=[f"{c['dims']}d-{np.dtype(c['dtype']).name}-{c['name']}" for c in _QUICK])
def test_oracle_matches_reference_digests(case):
    ref = reference_digests()[reference_case_key(case, quick=True)]
    src, kw = make_inputs(case)
    if "error" in ref:
        assert ref["error"] == "MaxLevelsExceeded"
        with pytest.raises(MaxLevelsExceeded):
            build_tree(src, **kw)
        return
    tree = build_tree(src, **kw)
    assert digest_mismatches(ref["tree"], tree_digests(tree)) == []
    if "trav" in ref:
        trav = build_traversal(tree, **_trav_kwargs(case))
        assert digest_mismatches(ref["trav"], trav_digests(trav)) == []


def test_full_size_digest_file_matches_its_case_table():
    """What ``tests/test_gpu_parity.py::test_full_size_matches_reference_run`` parametrises over."""
    import json
    from tests.golden.make_full_size_golden import full_size_cases
    with open(os.path.join(GOLDEN, "full_size_digests.json")) as f:
        d = json.load(f)
    assert set(d) == set(full_size_cases())
    for entry in d.values():
        assert entry["_nboxes"] > 0 and entry["_nlevels"] > 0
        assert {"tree.box_child_ids", "trav.from_sep_siblings_lists",
                "trav.neighbor_source_boxes_lists"} <= set(entry)


def test_reference_digests_cover_both_sweeps():
    d = reference_digests()
    for quick in (True, False):
        for case in make_cases(quick):
            entry = d[reference_case_key(case, quick)]
            if case.get("expect_max_levels"):
                assert entry == {"error": "MaxLevelsExceeded"}
            elif case["dims"] == 1 and "lr" in case["name"]:
                # boxtree/tree_build_kernels.py:825-915 writes `box_center.x` on what is a scalar
                # in 1-D: the reference cannot build 1-D level-restricted trees at all
                assert entry == {"error": "RuntimeError"}
            else:
                assert "tree" in entry and (("trav" in entry) == (case.get("trav", {}) is not None))


@pytest.mark.parametrize("stem,index", [("refexec_2d_f64_adaptive", (2, "float64", "adaptive")),
                                        ("refexec_3d_f32_ext_lr", (3, "float32", "ext-lr-n1")),
                                        ("refexec_3d_f64_nsep2", (3, "float64", "nsep2"))])
def test_oracle_matches_reference_arrays(stem, index):
    """Three cases with every array the reference produced."""
    case = next(c for c in _QUICK if (c["dims"], np.dtype(c["dtype"]).name, c["name"]) == index)
    ref = np.load(os.path.join(GOLDEN, stem + ".npz"))
    src, kw = make_inputs(case)
    tree = build_tree(src, **kw)
    trav = build_traversal(tree, **_trav_kwargs(case))
    checked = 0
    for key in ref.files:
        kind, name = key.split(".", 1)
        if kind == "tree":
            got = getattr(tree, name)
            got = np.stack(got) if isinstance(got, (list, tuple)) else np.asarray(got)
        elif name.startswith("from_sep_smaller_by_level."):
            _, lev, fld = name.split(".")
            got = np.asarray(getattr(trav.from_sep_smaller_by_level[int(lev)], fld))
        elif name.startswith("target_boxes_sep_smaller_by_source_level."):
            got = np.asarray(trav.target_boxes_sep_smaller_by_source_level[int(name.split(".")[1])])
        else:
            got = np.asarray(getattr(trav, name))
        want = ref[key]
        assert got.dtype == want.dtype and got.shape == want.shape, key
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), key
        checked += 1
    assert checked > 40

# }}}
