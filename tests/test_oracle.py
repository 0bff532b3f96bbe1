"""CPU tests of the oracle (the restatement of the reference algorithm): the reference's own
structural tests, the constant-one FMM completeness test, an independent brute-force list
definition, the committed golden fixtures, and error behaviour."""
import json
import os

import numpy as np
import pytest

from oracle.fmm import constant_one_fmm
from oracle.traversal import build_traversal, merge_close_lists
from oracle.tree_build import MaxLevelsExceeded, build_tree
from tests.golden.make_golden import digest, digest_cases, flatten
from tests.invariants import check_against_brute_force, check_traversal, check_tree
from tests.parity_util import normal_particles

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", [
    dict(n=4, max_particles_in_box=30),                      # test_tree.py:236
    dict(n=50, max_particles_in_box=30),                     # :247
    dict(n=20000, max_particles_in_box=30, skip_prune=True),  # :258
    dict(n=20000, max_particles_in_box=30, nboxes_guess=5),   # :270
    dict(n=20000, max_particles_in_box=5),                   # :282
    dict(n=20000, max_particles_in_box=30),                  # :294
    dict(n=5000, max_particles_in_box=30, kind="non-adaptive"),  # :325
    dict(n=20000, max_particles_in_box=30, kind="adaptive-level-restricted"),
])
def test_tree_invariants(dims, dtype, case):
    case = dict(case)
    n = case.pop("n")
    src = normal_particles(n, dims, dtype)
    tree = build_tree(src, debug=True, **case)
    check_tree(tree, src, max_particles_in_box=case["max_particles_in_box"])


@pytest.mark.parametrize("dims", [2, 3])
def test_tree_refine_weights(dims):   # test_tree.py:305
    n = 20000
    src = normal_particles(n, dims, np.float64)
    w = np.random.default_rng(10).integers(1, 10, n, dtype=np.int32)
    tree = build_tree(src, refine_weights=w, max_leaf_refine_weight=100, debug=True)
    check_tree(tree, src, refine_weights=w, max_leaf_refine_weight=100)


@pytest.mark.parametrize("dims", [2, 3])
def test_source_target_tree(dims):    # test_tree.py:341
    src = normal_particles(20000, dims, np.float64, seed=12)
    tgt = normal_particles(30000, dims, np.float64, seed=19)
    tree = build_tree(src, targets=tgt, max_particles_in_box=10, debug=True)
    check_tree(tree, src, unsorted_targets=tgt)
    nb = tree.nboxes
    assert np.all(tree.box_source_counts_cumul[:nb] + tree.box_target_counts_cumul[:nb] > 0)


def test_max_levels_exceeded():       # test_tree.py:1103-1112
    pts = [np.zeros(11), np.zeros(11)]
    pts[0][0] = 1.0
    with pytest.raises(MaxLevelsExceeded):
        build_tree([np.array([0.5] * 11 + [1.0]), np.array([0.5] * 11 + [1.0])],
                   max_particles_in_box=10)


def test_argument_errors():
    src = normal_particles(100, 2, np.float64)
    with pytest.raises(ValueError):
        build_tree(src, kind="bogus", max_particles_in_box=10)
    with pytest.raises(ValueError):
        build_tree(src)
    with pytest.raises(ValueError):
        build_tree(src, max_particles_in_box=10, refine_weights=np.ones(100, np.int32),
                   max_leaf_refine_weight=5)
    with pytest.raises(ValueError):
        build_tree(src, source_radii=np.ones(100), max_particles_in_box=10)
    with pytest.raises(ValueError):
        build_tree(src, targets=src, target_radii=np.ones(100), max_particles_in_box=10)
    tree = build_tree(src, max_particles_in_box=10, skip_prune=True)
    with pytest.raises(ValueError):
        build_traversal(tree)


@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("sources_are_targets", [True, False])
def test_traversal_connectivity(dims, sources_are_targets):   # test_traversal.py:58
    src = normal_particles(20000, dims, np.float64)
    tgt = None if sources_are_targets else normal_particles(30000, dims, np.float64, seed=18)
    tree = build_tree(src, targets=tgt, max_particles_in_box=30, debug=True)
    trav = build_traversal(tree)
    check_traversal(tree, trav, sources_are_targets)


@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("kind", ["adaptive", "adaptive-level-restricted"])
def test_lists_against_brute_force(dims, kind):
    src = normal_particles(2500, dims, np.float64)
    tree = build_tree(src, max_particles_in_box=10, kind=kind)
    check_against_brute_force(tree, build_traversal(tree))


FMM_CASES = [   # modelled on test/test_fmm.py:141-164 (dims, nsources, ntargets, extents, ...)
    dict(dims=1, ns=3000, nt=0),
    dict(dims=2, ns=5000, nt=0),
    dict(dims=3, ns=5000, nt=0),
    dict(dims=2, ns=5000, nt=4000),
    dict(dims=3, ns=5000, nt=4000),
    dict(dims=2, ns=5000, nt=4000, radii=True, norm="linf", crit="static_linf"),
    dict(dims=2, ns=5000, nt=4000, radii=True, norm="linf", crit="precise_linf"),
    dict(dims=2, ns=5000, nt=4000, radii=True, norm="l2", crit="static_l2"),
    dict(dims=3, ns=5000, nt=4000, radii=True, norm="linf", crit="static_linf"),
    dict(dims=3, ns=5000, nt=4000, radii=True, norm="linf", crit="precise_linf"),
    dict(dims=3, ns=5000, nt=4000, radii=True, norm="l2", crit="static_l2"),
    dict(dims=3, ns=5000, nt=4000, radii=True, norm="l2", crit="precise_linf"),
    dict(dims=3, ns=5000, nt=4000, radii=True, norm="linf", crit=None,
         kind="adaptive-level-restricted"),
    dict(dims=3, ns=5000, nt=0, kind="adaptive-level-restricted"),
]


@pytest.mark.parametrize("well_sep_is_n_away", [1, 2])
@pytest.mark.parametrize("case", FMM_CASES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_constant_one_fmm(case, well_sep_is_n_away, dtype):   # test/test_fmm.py:166-391
    dims, ns, nt = case["dims"], case["ns"], case["nt"]
    src = normal_particles(ns, dims, dtype, seed=12)
    kw = dict(max_particles_in_box=30, kind=case.get("kind", "adaptive"))
    if nt:
        kw["targets"] = [0.7 * t for t in normal_particles(nt, dims, dtype, seed=19)]
        if case.get("radii"):
            rng = np.random.default_rng(13)
            kw.update(target_radii=(0.1 * 2 ** rng.uniform(-10, 0, nt)).astype(dtype),
                      stick_out_factor=0.25, extent_norm=case["norm"])
    tree = build_tree(src, debug=True, **kw)
    trav = build_traversal(tree, well_sep_is_n_away=well_sep_is_n_away,
                           from_sep_smaller_crit=case.get("crit"))
    weights = np.random.default_rng(1).integers(1, 5, ns).astype(np.float64)
    assert np.all(constant_one_fmm(tree, trav, weights) == weights.sum())
    if case.get("radii"):
        assert np.all(constant_one_fmm(tree, merge_close_lists(trav), weights) == weights.sum())


def test_min_nsources_threshold():    # test/test_fmm.py:618-665
    src = normal_particles(5000, 3, np.float64, seed=12)
    tgt = normal_particles(4000, 3, np.float64, seed=19)
    radii = (0.05 * 2 ** np.random.default_rng(13).uniform(-10, 0, 4000))
    tree = build_tree(src, targets=tgt, target_radii=radii, stick_out_factor=0.25,
                      max_particles_in_box=30)
    trav = build_traversal(tree, _from_sep_smaller_min_nsources_cumul=40)
    assert np.all(constant_one_fmm(tree, trav, np.ones(5000)) == 5000)


def test_golden_config1():
    src, tkw, vkw = digest_cases()["config1_2d_1e4"]
    tree = build_tree(src, **tkw)
    trav = build_traversal(tree, **vkw)
    got = flatten(tree, trav)
    want = np.load(os.path.join(GOLDEN, "config1_2d_1e4.npz"))
    assert set(got) == set(want.files)
    for k in want.files:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), k


def test_golden_digests():
    want = json.load(open(os.path.join(GOLDEN, "digests.json")))
    for name, (src, tkw, vkw) in digest_cases().items():
        tree = build_tree(src, **tkw)
        trav = build_traversal(tree, **vkw)
        got = {k: digest(v) for k, v in flatten(tree, trav).items()}
        assert got == want[name], name


def test_particle_list_filter_restatement():
    """ParticleListFilter (boxtree/tree.py:1040-1239), the properties of test/test_tree.py's
    filter tests: the filtered lists hold exactly the flagged targets, box by box."""
    from oracle import particle_filter as opf
    src = normal_particles(5000, 3, np.float64, seed=12)
    tgt = normal_particles(4000, 3, np.float64, seed=19)
    tree = build_tree(src, targets=tgt, max_particles_in_box=30)
    flags = (np.random.default_rng(5).random(4000) < 0.3).astype(np.int8)
    n, starts, lists = opf.filter_target_lists_in_user_order(tree, flags)
    assert n == flags.sum() and sorted(lists) == sorted(np.nonzero(flags)[0])
    user_target_ids = np.empty(4000, np.int64)
    user_target_ids[tree.sorted_target_ids] = np.arange(4000)
    for ibox in range(tree.nboxes):
        s, c = tree.box_target_starts[ibox], tree.box_target_counts_nonchild[ibox]
        want = [u for u in user_target_ids[s:s + c] if flags[u]]
        assert list(lists[starts[ibox]:starts[ibox + 1]]) == want
    n2, bstart, bcount, ftargets, ufi = opf.filter_target_lists_in_tree_order(tree, flags)
    assert n2 == n and bcount.sum() == n
    for ibox in range(tree.nboxes):
        s, c = tree.box_target_starts[ibox], tree.box_target_counts_nonchild[ibox]
        mine = [j for j in range(s, s + c) if flags[user_target_ids[j]]]
        assert list(ufi[bstart[ibox]:bstart[ibox] + bcount[ibox]]) == mine
    for ax in range(3):
        assert np.array_equal(ftargets[ax], tree.targets[ax][ufi])


def test_link_point_sources_restatement():
    """link_point_sources (boxtree/tree.py:773-955): every box's point sources are contiguous
    and are exactly those owned by the box's sources (test/test_tree.py's point-source test)."""
    from oracle import particle_filter as opf
    ns = 3000
    src = normal_particles(ns, 3, np.float64)
    tgt = normal_particles(2000, 3, np.float64, seed=19)
    rng = np.random.default_rng(7)
    radii = 0.05 * 2 ** rng.uniform(-10, 0, ns)
    tree = build_tree(src, targets=tgt, max_particles_in_box=30, source_radii=radii,
                      stick_out_factor=0.25, extent_norm="linf")
    counts = rng.integers(1, 5, ns)
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    owner = np.repeat(np.arange(ns), counts)
    pts = [src[a][owner] for a in range(3)]
    w = opf.link_point_sources(tree, starts, pts)
    assert w["npoint_sources"] == starts[-1]
    assert sorted(w["user_point_source_ids"]) == list(range(starts[-1]))
    assert w["box_point_source_counts_cumul"][0] == starts[-1]
    for b in range(tree.nboxes):
        s, c = tree.box_source_starts[b], tree.box_source_counts_nonchild[b]
        ps, pc = w["box_point_source_starts"][b], w["box_point_source_counts_nonchild"][b]
        assert set(owner[w["user_point_source_ids"][ps:ps + pc]]) == set(tree.user_source_ids[s:s + c])
    for a in range(3):
        assert np.array_equal(w["point_sources"][a], pts[a][w["user_point_source_ids"]])


@pytest.mark.parametrize("dims,kind", [(2, "adaptive"), (3, "adaptive-level-restricted")])
def test_peer_lists_restatement_against_definition(dims, kind):
    """Peer lists (boxtree/area_query.py:393-475) against the definition quoted in its docstring
    (adjacent, at least as large, no child with both properties), on integer box coordinates."""
    from oracle.traversal import find_peer_lists
    from tests.invariants import integer_box_coords
    tree = build_tree(normal_particles(2500, dims, np.float64), max_particles_in_box=10, kind=kind)
    st, li = find_peer_lists(tree)
    lev, lo, size = integer_box_coords(tree)
    hi = lo + size[:, None]
    nb = tree.nboxes
    adj = np.all((lo[:, None, :] <= hi[None, :, :]) & (lo[None, :, :] <= hi[:, None, :]), axis=2)
    for b in range(nb):
        cand = adj[b] & (lev <= lev[b])
        want = {k for k in np.nonzero(cand)[0]
                if not any(cand[c] for c in tree.box_child_ids[:, k] if c)}
        assert set(li[st[b]:st[b + 1]]) == want, b


def _random_balls(tree, n, seed=3):
    rng = np.random.default_rng(seed)
    lo, hi = tree.bounding_box
    ext = float(tree.root_extent)
    centers = [rng.uniform(lo[a] - 0.2 * ext, lo[a] + 1.2 * ext, n).astype(tree.coord_dtype)
               for a in range(tree.dimensions)]             # some centres outside the bounding box
    radii = (ext * 2 ** rng.uniform(-9, -0.5, n)).astype(tree.coord_dtype)
    return centers, radii


@pytest.mark.parametrize("dims,kind", [(2, "adaptive"), (3, "adaptive-level-restricted")])
def test_area_query_restatement_against_definition(dims, kind):
    """AreaQueryBuilder (boxtree/area_query.py:168-392): the leaves found through the guiding box's
    peers are exactly the leaves whose box overlaps the ball (test/test_tree.py:645-700)."""
    from oracle.traversal import area_query
    tree = build_tree(normal_particles(3000, dims, np.float64), max_particles_in_box=10, kind=kind)
    centers, radii = _random_balls(tree, 400)
    starts, lists = area_query(tree, centers, radii)
    nb = tree.nboxes
    leaf = (tree.box_flags[:nb] & 12) == 0
    half = float(tree.root_extent) / 2.0 ** (1 + tree.box_levels[:nb].astype(np.float64))
    for i in range(len(radii)):
        dist = np.max(np.abs(tree.box_centers[:, :nb] - np.array([c[i] for c in centers])[:, None]), axis=0)
        want = set(np.nonzero(leaf & (dist <= half + radii[i]))[0])
        got = lists[starts[i]:starts[i + 1]]
        assert len(set(got)) == len(got) and set(got) == want, i


@pytest.mark.parametrize("dims,n_away", [(2, 1), (3, 1), (3, 2)])
def test_translation_and_rotation_classes_restatement(dims, n_away):
    """translation_classes.py / rotation_classes.py: the class of every list-2 pair maps back to
    the pair's centre-to-centre vector (test/test_traversal.py's translation-class test) and
    pairs with parallel vectors share a rotation class."""
    from oracle import translation_classes as otc
    tree = build_tree(normal_particles(3000, dims, np.float64), max_particles_in_box=15)
    trav = build_traversal(tree, well_sep_is_n_away=n_away)
    cls, dist, level_starts = otc.translation_classes(trav, tree)
    tp, st, li = trav.target_or_target_parent_boxes, trav.from_sep_siblings_starts, \
        trav.from_sep_siblings_lists
    rows = np.repeat(np.arange(len(tp)), np.diff(st))
    want = tree.box_centers[:, tp[rows]] - tree.box_centers[:, li]
    assert np.allclose(dist[:, cls], want, rtol=0, atol=1e-12 * float(tree.root_extent))
    lev = tree.box_levels[li].astype(int)
    assert np.all((level_starts[lev] <= cls) & (cls < level_starts[lev + 1]))
    rot, angles = otc.rotation_classes(trav, tree)
    cosang = want[-1] / np.linalg.norm(want, axis=0)
    assert np.allclose(np.cos(angles[rot]), cosang, atol=1e-12)
