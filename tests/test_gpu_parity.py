"""GPU parity tests: the CUDA path (through the public API, which calls the C ABI) against
the CPU oracle, the committed golden fixtures, and size-independent properties at the
BASELINE sizes.  Bar: bit-exact integers, 0-ULP floats."""
import json
import os

import numpy as np
import pytest

from tests.golden.make_golden import digest, digest_cases, flatten
from tests.gpu_sweep import make_cases, run_case
from tests.invariants import check_against_brute_force, check_traversal, check_tree
from tests.parity_util import config3_inputs, normal_particles, uniform_particles

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

_CASES = make_cases(quick=False)


@pytest.fixture(scope="module")
def builders(actx):
    from boxtree_b200 import TreeBuilder
    return TreeBuilder(actx), {}


@pytest.mark.parametrize(
    "case", _CASES, ids=[f"{c['dims']}d-{np.dtype(c['dtype']).name}-{c['name']}" for c in _CASES])
def test_parity_sweep(actx, builders, case):
    """CUDA path == oracle (arrays) AND == the reference's own run (digests of
    ``tests/golden/refexec_digests.json``, made by executing /root/reference's TreeBuilder and
    FMMTraversalBuilder on the CPU, see ``tests/refexec``)."""
    from tests.parity_util import reference_case_key, reference_digests
    tb, travs = builders
    ref = reference_digests()[reference_case_key(case, quick=False)]
    if "error" in ref:
        # the reference gives up on coincident points; and its level-restriction kernel does not
        # compile in 1-D (component access on a scalar coord_vec_t), so those cases have no
        # reference output and are held to the oracle only
        assert (ref["error"] == "MaxLevelsExceeded" and case.get("expect_max_levels")) or \
            (case["dims"] == 1 and "lr" in case["name"])
        ref = None
    bad = run_case(dict(case), actx, tb, travs, ref_digests=ref)
    assert not bad, bad[:10]


_HEAVY_CASES = [c for c in _CASES if c["name"] in (
    "adaptive", "src-tgt", "lr", "ext-linf-precise_linf-n1", "ext-l2-static_l2-n2", "ext-lr-n1",
    "ext-lr-n2", "ext-minsrc", "two-level", "single-box")]


@pytest.mark.parametrize("heavy_by_sort", [False, True])
@pytest.mark.parametrize("budget", [1, 8, 100])
@pytest.mark.parametrize(
    "case", _HEAVY_CASES,
    ids=[f"{c['dims']}d-{np.dtype(c['dtype']).name}-{c['name']}" for c in _HEAVY_CASES])
def test_parity_heavy_row_path(actx, builders, case, budget, heavy_by_sort, monkeypatch):
    """Force (almost) every row of lists 1 and 3 through the grid-wide heavy-row path, with the
    position map (default) and with the radix sort."""
    from boxtree_b200 import _cabi
    lib = _cabi.load()
    monkeypatch.setenv("BT_WALK_BUDGET", str(budget))
    tb, travs = builders
    case = dict(case)
    default_mode = lib.bt_get_walk_mode()
    try:
        if heavy_by_sort:
            lib.bt_set_walk_mode(default_mode | 512)
        bad = run_case(case, actx, tb, travs)
    finally:
        lib.bt_set_walk_mode(default_mode)
    assert not bad, bad[:10]
    if case["n"] > 1000 and case["dims"] >= 2 and budget <= 8:
        st = case["_trav_stats"]      # the fused list-1+3 walk reports its heavy rows under list 3
        assert st["heavy_rows_list1"] > 0 or st["heavy_rows_list3"] > 0


@pytest.mark.parametrize("stride", [0, 3, 16])
@pytest.mark.parametrize(
    "case", _HEAVY_CASES,
    ids=[f"{c['dims']}d-{np.dtype(c['dtype']).name}-{c['name']}" for c in _HEAVY_CASES])
def test_parity_stage_stride(actx, builders, case, stride, monkeypatch):
    """Fused list-1+3 walk: staging off (every row walked twice), and strides so small that
    most rows overflow their staging area and are walked again by the fill pass."""
    monkeypatch.setenv("BT_STAGE_STRIDE", str(stride))
    tb, travs = builders
    case = dict(case)
    bad = run_case(case, actx, tb, travs)
    assert not bad, bad[:10]
    if case["n"] > 1000 and case["dims"] >= 2:
        assert case["_trav_stats"]["rewalked_rows_list13"] > 0


@pytest.mark.parametrize("mode", [0, 1 | 2 | 4 | 16 | 32, 64 | 2 | 8 | 32, 64 | 1 | 4 | 16, 64 | 128,
                                  64 | 128 | 256])
@pytest.mark.parametrize(
    "case", _HEAVY_CASES,
    ids=[f"{c['dims']}d-{np.dtype(c['dtype']).name}-{c['name']}" for c in _HEAVY_CASES])
def test_parity_all_walk_mappings(actx, builders, case, mode):
    """Every builder in its one-thread-per-row mapping (0) and its cooperative mapping."""
    from boxtree_b200 import _cabi
    lib = _cabi.load()
    tb, travs = builders
    default_mode = lib.bt_get_walk_mode()
    try:
        lib.bt_set_walk_mode(mode)
        bad = run_case(dict(case), actx, tb, travs)
    finally:
        lib.bt_set_walk_mode(default_mode)
    assert not bad, bad[:10]


def _build_tree_only(actx, src, tkw):
    from boxtree_b200 import TreeBuilder
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in tkw.items()}
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], **dkw)
    return tree


def _build(actx, src, tkw, vkw):
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
               [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in tkw.items()}
    tree, _ = TreeBuilder(actx)(actx, [actx.from_numpy(s) for s in src], **dkw)
    ctor = {k: vkw[k] for k in ("well_sep_is_n_away", "from_sep_smaller_crit") if k in vkw}
    trav, _ = FMMTraversalBuilder(actx, **ctor)(actx, tree)
    return tree, trav


@pytest.mark.parametrize("name", ["config1_2d_1e4", "config3_3d_1e5", "config4_plummer_1e5_f32",
                                  "normal_3d_f32_2away"])
@pytest.mark.parametrize("key_depth", [-1, 3, 9])
def test_key_depth_does_not_change_the_tree(actx, name, key_depth):
    """The sort key may resolve fewer levels than the tree needs at first (the build retries with
    more) or be a two-word key right away (``_key_depth=-1``, what trees deeper than 19 / 28
    levels use): every tree array must still have the digest of the reference's own run."""
    want = json.load(open(os.path.join(GOLDEN, "digests.json")))[name]
    src, tkw, vkw = digest_cases()[name]
    tree, trav = _build(actx, src, dict(tkw, _key_depth=key_depth), vkw)
    got = {k: digest(v) for k, v in flatten(actx.to_numpy(tree), actx.to_numpy(trav)).items()}
    assert got == want


def test_lr_overflow_retry_keeps_the_split_list(actx):
    """Level-restricted build whose splits exceed the pool (negative ``_lr_slack``): the
    retry after CTL_OVERFLOW repeats the child creation with the split list of the decide scan,
    which must survive the pool's reallocation."""
    from boxtree_b200 import TreeBuilder
    from oracle.tree_build import build_tree
    from tests.parity_util import tree_mismatches
    src, tgt, radii = config3_inputs(20000, 20000)
    kw = dict(max_particles_in_box=30, stick_out_factor=0.25, extent_norm="linf",
              kind="adaptive-level-restricted")
    tb = TreeBuilder(actx)
    tree, _ = tb(actx, [actx.from_numpy(x) for x in src], targets=[actx.from_numpy(x) for x in tgt],
                 target_radii=actx.from_numpy(radii), _lr_slack=-10**7, nboxes_guess=2, **kw)
    assert tb.last_stats["reallocs"] > 0
    ref = build_tree(src, targets=tgt, target_radii=radii, **kw)
    assert not tree_mismatches(ref, actx.to_numpy(tree))


def test_golden_config1(actx):
    src, tkw, vkw = digest_cases()["config1_2d_1e4"]
    tree, trav = _build(actx, src, tkw, vkw)
    got = flatten(actx.to_numpy(tree), actx.to_numpy(trav))
    want = np.load(os.path.join(GOLDEN, "config1_2d_1e4.npz"))
    assert set(got) == set(want.files)
    for k in want.files:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        assert np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), k


@pytest.mark.parametrize("name", sorted(digest_cases()))
def test_golden_digests(actx, name):
    want = json.load(open(os.path.join(GOLDEN, "digests.json")))[name]
    src, tkw, vkw = digest_cases()[name]
    tree, trav = _build(actx, src, tkw, vkw)
    got = {k: digest(v) for k, v in flatten(actx.to_numpy(tree), actx.to_numpy(trav)).items()}
    assert got == want


def test_config2_full_size_parity(actx):
    """BASELINE config 2 at full size (3-D, 1e6 uniform fp64): every array vs the oracle."""
    from oracle.traversal import build_traversal
    from oracle.tree_build import build_tree
    from tests.parity_util import trav_mismatches, tree_mismatches
    src = uniform_particles(1_000_000, 3, np.float64)
    tree, trav = _build(actx, src, dict(max_particles_in_box=30), {})
    rt = build_tree(src, max_particles_in_box=30)
    assert not tree_mismatches(rt, actx.to_numpy(tree))
    assert not trav_mismatches(build_traversal(rt), actx.to_numpy(trav))


def _full_size_names():
    with open(os.path.join(GOLDEN, "full_size_digests.json")) as f:
        return sorted(json.load(f))


@pytest.mark.parametrize("name", _full_size_names())
def test_full_size_matches_reference_run(actx, name):
    """The bench workloads at FULL size (1e7 points) against the REFERENCE ITSELF: every Tree and
    FMMTraversalInfo array has the sha256 that the reference's own TreeBuilder and
    FMMTraversalBuilder produced when executed on the CPU (tests/golden/make_full_size_golden.py,
    tests/refexec)."""
    import gc
    import torch
    from tests.golden.make_full_size_golden import full_size_cases
    with open(os.path.join(GOLDEN, "full_size_digests.json")) as f:
        want = json.load(f)[name]
    src, tkw, vkw = full_size_cases()[name]()
    tree, trav = _build(actx, src, tkw, vkw)
    assert int(tree.nboxes) == want["_nboxes"] and int(tree.nlevels) == want["_nlevels"]
    flat = flatten(actx.to_numpy(tree), actx.to_numpy(trav))
    del tree, trav
    gc.collect()
    torch.cuda.empty_cache()
    assert set(flat) == {k for k in want if not k.startswith("_")}
    bad = [k for k, v in flat.items() if digest(v) != want[k]]
    assert bad == []


def test_config3_properties_full_size(actx):
    """BASELINE config 3 at full size (1e7 points): size-independent properties on device."""
    import torch
    ns = nt = 5_000_000
    src, tgt, radii = config3_inputs(ns, nt)
    tkw = dict(max_particles_in_box=30, targets=tgt, target_radii=radii, stick_out_factor=0.25,
               extent_norm="linf", kind="adaptive-level-restricted")
    tree, trav = _build(actx, src, tkw, {})
    dsrc = [actx.from_numpy(s) for s in src]
    dtgt = [actx.from_numpy(t) for t in tgt]
    usi = tree.user_source_ids.long()
    sti = tree.sorted_target_ids.long()
    # orderings are permutations and the permuted coordinates are exact copies
    assert int(torch.bincount(usi, minlength=ns).max()) == 1 and usi.numel() == ns
    assert int(torch.bincount(sti, minlength=nt).max()) == 1 and sti.numel() == nt
    for ax in range(3):
        assert torch.equal(tree.sources[ax], dsrc[ax][usi])
        assert torch.equal(tree.targets[ax][sti], dtgt[ax])
    assert torch.equal(tree.target_radii[sti], actx.from_numpy(radii))
    nb = tree.nboxes
    lev = tree.box_levels.long()
    par = tree.box_parent_ids.long()
    assert bool((lev[par[1:]] + 1 == lev[1:]).all())
    assert bool((lev[1:] >= lev[:-1]).all())                       # level-major numbering
    ls = tree.level_start_box_nrs.long()
    assert int(ls[0]) == 0 and int(ls[-1]) == nb
    # counts: root holds everything; nonchild + children's cumul == cumul
    assert int(tree.box_source_counts_cumul[0]) == ns and int(tree.box_target_counts_cumul[0]) == nt
    ch = tree.box_child_ids[:, :nb].long()
    for cum, non in ((tree.box_source_counts_cumul, tree.box_source_counts_nonchild),
                     (tree.box_target_counts_cumul, tree.box_target_counts_nonchild)):
        kid = torch.where(ch != 0, cum.long()[ch], torch.zeros_like(ch)).sum(0)
        assert bool((non.long() + kid == cum.long()).all())
    # 2:1 balance of the level-restricted tree through list 1 (adjacent leaves differ by <= 1
    # level is only promised for boxes without own extent-particles; check colleagues instead):
    cs, cl = trav.same_level_non_well_sep_boxes_starts.long(), \
        trav.same_level_non_well_sep_boxes_lists.long()
    rows = torch.repeat_interleave(torch.arange(nb, device=cs.device), cs[1:] - cs[:-1])
    assert bool((lev[rows] == lev[cl]).all())
    # list 2: same level, never adjacent
    tp = trav.target_or_target_parent_boxes.long()
    s2, l2 = trav.from_sep_siblings_starts.long(), trav.from_sep_siblings_lists.long()
    r2 = tp[torch.repeat_interleave(torch.arange(tp.numel(), device=s2.device), s2[1:] - s2[:-1])]
    assert bool((lev[r2] == lev[l2]).all())
    cen = tree.box_centers[:, :nb]
    dist = (cen[:, r2] - cen[:, l2]).abs().amax(0)
    size = float(tree.root_extent) * 0.5 ** lev[r2].double()
    assert bool((dist > 1.5 * size).all())
    # completeness of the interaction lists at full size: every target hears every source once
    from boxtree_b200.constant_one import constant_one_fmm
    w = torch.randint(1, 5, (ns,), device=cs.device)
    assert bool((constant_one_fmm(tree, trav, w) == w.sum()).all())
    assert bool((constant_one_fmm(tree, trav.merge_close_lists(actx), w) == w.sum()).all())


@pytest.mark.parametrize("spec", ["uniform:4000000:f64", "plummer:4000000:f32"])
def test_constant_one_fmm_on_device_large(actx, spec):
    """Constant-one FMM evaluated on the device on multi-million-point trees."""
    import torch
    from boxtree_b200.constant_one import constant_one_fmm
    from tests.perf_probe import make
    src, kw = make(spec)
    tree, trav = _build(actx, src, kw, {})
    w = torch.randint(1, 5, (len(src[0]),), device=tree.box_flags.device)
    assert bool((constant_one_fmm(tree, trav, w) == w.sum()).all())


def test_invariants_and_brute_force_on_cuda_result(actx):
    src = normal_particles(2500, 3, np.float64)
    tree, trav = _build(actx, src, dict(max_particles_in_box=10), {})
    tree, trav = actx.to_numpy(tree), actx.to_numpy(trav)
    check_tree(tree, src, max_particles_in_box=10)
    check_traversal(tree, trav, True)
    check_against_brute_force(tree, trav)


def test_constant_one_fmm_on_cuda_result(actx):
    from oracle.fmm import constant_one_fmm
    src, tgt, radii = config3_inputs(30000, 30000)
    tkw = dict(max_particles_in_box=30, targets=tgt, target_radii=radii, stick_out_factor=0.25,
               extent_norm="linf", kind="adaptive-level-restricted")
    tree, trav = _build(actx, src, tkw, {})
    merged = trav.merge_close_lists(actx)
    tree, trav, merged = actx.to_numpy(tree), actx.to_numpy(trav), actx.to_numpy(merged)
    w = np.random.default_rng(1).integers(1, 5, 30000).astype(np.float64)
    assert np.all(constant_one_fmm(tree, trav, w) == w.sum())
    assert np.all(constant_one_fmm(tree, merged, w) == w.sum())


def test_device_constant_one_fmm_matches_host(actx):
    """The device evaluation (boxtree_b200.constant_one) against the oracle's host evaluation."""
    import torch
    from boxtree_b200.constant_one import constant_one_fmm as dev_fmm
    from oracle.fmm import constant_one_fmm
    src, tgt, radii = config3_inputs(30000, 30000)
    tkw = dict(max_particles_in_box=30, targets=tgt, target_radii=radii, stick_out_factor=0.25,
               extent_norm="linf", kind="adaptive-level-restricted")
    tree, trav = _build(actx, src, tkw, {})
    w = np.random.default_rng(2).integers(1, 9, 30000)
    got = dev_fmm(tree, trav, torch.from_numpy(w)).cpu().numpy()
    want = constant_one_fmm(actx.to_numpy(tree), actx.to_numpy(trav), w.astype(np.float64))
    assert np.array_equal(got, want.astype(np.int64)) and np.all(got == w.sum())


@pytest.mark.parametrize("with_targets", [False, True])
def test_particle_list_filter(actx, with_targets):
    """ParticleListFilter (boxtree/tree.py:1040-1239) against the oracle's loops."""
    from boxtree_b200 import ParticleListFilter
    from oracle import particle_filter as opf
    src = normal_particles(20000, 3, np.float64)
    tkw = dict(max_particles_in_box=30)
    if with_targets:
        _, tgt, radii = config3_inputs(10, 15000)
        tkw.update(targets=tgt, target_radii=radii, stick_out_factor=0.25, extent_norm="linf")
    tree, _ = _build(actx, src, tkw, {})
    ht = actx.to_numpy(tree)
    flags = (np.random.default_rng(5).random(ht.ntargets) < 0.3).astype(np.int8)
    plf = ParticleListFilter(actx)
    got = actx.to_numpy(plf.filter_target_lists_in_user_order(actx, tree, actx.from_numpy(flags)))
    n, starts, lists = opf.filter_target_lists_in_user_order(ht, flags)
    assert got.nfiltered_targets == n == int(flags.sum())
    assert np.array_equal(got.target_starts, starts) and got.target_starts.dtype == np.int32
    assert np.array_equal(got.target_lists, lists) and got.target_lists.dtype == np.int32
    got = actx.to_numpy(plf.filter_target_lists_in_tree_order(actx, tree, actx.from_numpy(flags)))
    n, bstart, bcount, targets, ufi = opf.filter_target_lists_in_tree_order(ht, flags)
    assert got.nfiltered_targets == n
    assert np.array_equal(got.box_target_starts, bstart)
    assert np.array_equal(got.box_target_counts_nonchild, bcount)
    assert np.array_equal(got.unfiltered_from_filtered_target_indices, ufi)
    for a, b in zip(got.targets, targets):
        assert np.array_equal(a, b)


def test_link_point_sources(actx):
    """link_point_sources (boxtree/tree.py:773-955) against the oracle's loops."""
    from boxtree_b200 import link_point_sources
    from oracle import particle_filter as opf
    ns = 6000
    src = normal_particles(ns, 3, np.float64)
    rng = np.random.default_rng(7)
    radii = 0.05 * 2 ** rng.uniform(-10, 0, ns)
    tgt = normal_particles(4000, 3, np.float64, seed=19)
    tree = _build_tree_only(actx, src, dict(max_particles_in_box=30, targets=tgt, source_radii=radii,
                                            stick_out_factor=0.25, extent_norm="linf"))
    counts = rng.integers(1, 5, ns)
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    npts = int(starts[-1])
    owner = np.repeat(np.arange(ns), counts)
    pts = [src[a][owner] + radii[owner] * rng.uniform(-1, 1, npts) for a in range(3)]
    got = actx.to_numpy(link_point_sources(actx, tree, actx.from_numpy(starts),
                                           [actx.from_numpy(p) for p in pts]))
    want = opf.link_point_sources(actx.to_numpy(tree), starts, pts)
    assert got.npoint_sources == want["npoint_sources"] == npts
    for k in ("point_source_starts", "point_source_counts", "user_point_source_ids",
              "box_point_source_starts", "box_point_source_counts_nonchild",
              "box_point_source_counts_cumul"):
        a, b = np.asarray(getattr(got, k)), want[k]
        assert a.dtype == b.dtype and np.array_equal(a, b), k
    for a, b in zip(got.point_sources, want["point_sources"]):
        assert np.array_equal(a, b)
    assert int(np.asarray(got.box_point_source_counts_cumul)[0]) == npts
    t2 = _build_tree_only(actx, src, dict(max_particles_in_box=30))
    with pytest.raises(ValueError):
        link_point_sources(actx, t2, actx.from_numpy(starts), [actx.from_numpy(p) for p in pts])


@pytest.mark.parametrize("dims,dtype,kind", [(1, np.float64, "adaptive"), (2, np.float32, "adaptive"),
                                             (3, np.float64, "adaptive-level-restricted"),
                                             (3, np.float32, "adaptive")])
def test_peer_lists(actx, dims, dtype, kind):
    """PeerListFinder (boxtree/area_query.py:1057-1188) against the oracle, array for array."""
    from boxtree_b200 import PeerListFinder
    from oracle.traversal import find_peer_lists
    src = normal_particles(20000, dims, dtype)
    tree = _build_tree_only(actx, src, dict(max_particles_in_box=10, kind=kind))
    got, _ = PeerListFinder(actx)(actx, tree)
    got = actx.to_numpy(got)
    st, li = find_peer_lists(actx.to_numpy(tree))
    assert np.array_equal(got.peer_list_starts, st) and got.peer_list_starts.dtype == np.int32
    assert np.array_equal(got.peer_lists, li) and got.peer_lists.dtype == np.int32


@pytest.mark.parametrize("dims,dtype,kind", [(2, np.float64, "adaptive"), (3, np.float32, "adaptive"),
                                             (3, np.float64, "adaptive-level-restricted")])
def test_area_query(actx, dims, dtype, kind):
    """AreaQueryBuilder (boxtree/area_query.py:657-807) against the oracle, array for array."""
    from boxtree_b200 import AreaQueryBuilder
    from oracle.traversal import area_query
    from tests.test_oracle import _random_balls
    src = normal_particles(20000, dims, dtype)
    tree = _build_tree_only(actx, src, dict(max_particles_in_box=10, kind=kind))
    htree = actx.to_numpy(tree)
    centers, radii = _random_balls(htree, 3000)
    got, _ = AreaQueryBuilder(actx)(actx, tree, [actx.from_numpy(c) for c in centers],
                                    actx.from_numpy(radii))
    got = actx.to_numpy(got)
    starts, lists = area_query(htree, centers, radii)
    assert np.array_equal(got.leaves_near_ball_starts, starts)
    assert np.array_equal(got.leaves_near_ball_lists, lists)
    with pytest.raises(TypeError):
        other = np.float32 if dtype == np.float64 else np.float64
        AreaQueryBuilder(actx)(actx, tree, [actx.from_numpy(c) for c in centers],
                               actx.from_numpy(radii.astype(other)))


@pytest.mark.parametrize("dims,dtype", [(2, np.float64), (3, np.float32)])
def test_leaves_to_balls_and_space_invaders(actx, dims, dtype):
    """LeavesToBallsLookupBuilder and SpaceInvaderQueryBuilder (boxtree/area_query.py:810-1048)."""
    from boxtree_b200 import LeavesToBallsLookupBuilder, SpaceInvaderQueryBuilder
    from oracle.traversal import leaves_to_balls, space_invader_query
    from tests.test_oracle import _random_balls
    src = normal_particles(8000, dims, dtype)
    tree = _build_tree_only(actx, src, dict(max_particles_in_box=10))
    htree = actx.to_numpy(tree)
    centers, radii = _random_balls(htree, 600)
    dc, dr = [actx.from_numpy(c) for c in centers], actx.from_numpy(radii)
    got, _ = LeavesToBallsLookupBuilder(actx)(actx, tree, dc, dr)
    got = actx.to_numpy(got)
    st, li = leaves_to_balls(htree, centers, radii)
    assert np.array_equal(got.balls_near_box_starts, st) and np.array_equal(got.balls_near_box_lists, li)
    sq, _ = SpaceInvaderQueryBuilder(actx)(actx, tree, dc, dr)
    sq = actx.to_numpy(sq)
    want = space_invader_query(htree, centers, radii)
    assert sq.dtype == want.dtype and np.array_equal(sq, want)


@pytest.mark.parametrize("dims", [2, 3])
def test_level_restriction_through_area_query(actx, dims):
    """test/test_tree.py:900-974: in a level-restricted tree the leaves found near every leaf (a
    ball slightly larger than the leaf) differ from it by at most one level."""
    from boxtree_b200 import AreaQueryBuilder
    nparticles = 10 ** 5
    rng = np.random.default_rng(15)
    src = [rng.normal(size=nparticles) for _ in range(dims)]
    tree = _build_tree_only(actx, src, dict(max_particles_in_box=30, kind="adaptive-level-restricted",
                                            nboxes_guess=10))
    ht = actx.to_numpy(tree)
    nb = ht.nboxes
    leaf_boxes = np.nonzero((ht.box_flags[:nb] & 12) == 0)[0]
    leaf_radii = float(ht.root_extent) / 2.0 ** (1 + ht.box_levels[leaf_boxes].astype(np.float64))
    leaf_centers = [np.ascontiguousarray(ht.box_centers[a, leaf_boxes]) for a in range(dims)]
    ball_radii = np.min(leaf_radii) / 2 + leaf_radii
    aq, _ = AreaQueryBuilder(actx)(actx, tree, [actx.from_numpy(c) for c in leaf_centers],
                                   actx.from_numpy(ball_radii))
    aq = actx.to_numpy(aq)
    st, li = aq.leaves_near_ball_starts, aq.leaves_near_ball_lists
    rows = np.repeat(np.arange(len(leaf_boxes)), np.diff(st))
    diff = np.abs(ht.box_levels[li].astype(int) - ht.box_levels[leaf_boxes[rows]].astype(int))
    assert np.all(diff <= 1)
    assert np.all(np.diff(st) >= 1)                       # every leaf finds at least itself


@pytest.mark.parametrize("dims,dtype,n_away", [(2, np.float64, 1), (3, np.float64, 2), (3, np.float32, 1)])
def test_translation_and_rotation_classes(actx, dims, dtype, n_away):
    """TranslationClassesBuilder / RotationClassesBuilder (boxtree/translation_classes.py,
    boxtree/rotation_classes.py) against the oracle."""
    from boxtree_b200 import RotationClassesBuilder, TranslationClassesBuilder
    from oracle import translation_classes as otc
    from oracle.traversal import build_traversal
    src = normal_particles(8000, dims, dtype)
    tree, trav = _build(actx, src, dict(max_particles_in_box=15), dict(well_sep_is_n_away=n_away))
    ht = actx.to_numpy(tree)
    htrav = build_traversal(ht, well_sep_is_n_away=n_away)
    for per_level in (True, False):
        got, _ = TranslationClassesBuilder(actx)(actx, trav, tree, is_translation_per_level=per_level)
        got = actx.to_numpy(got)
        cls, dist, ls = otc.translation_classes(htrav, ht, per_level)
        assert np.array_equal(got.from_sep_siblings_translation_classes, cls)
        assert got.from_sep_siblings_translation_classes.dtype == np.int32
        d = got.from_sep_siblings_translation_class_to_distance_vector
        assert d.dtype == dist.dtype and np.array_equal(d, dist)
        assert np.array_equal(got.from_sep_siblings_translation_classes_level_starts, ls)
    got, _ = RotationClassesBuilder(actx)(actx, trav, tree)
    got = actx.to_numpy(got)
    rot, angles = otc.rotation_classes(htrav, ht)
    assert np.array_equal(got.from_sep_siblings_rotation_classes, rot)
    assert np.array_equal(got.from_sep_siblings_rotation_class_to_angle, angles)


def test_cost_model_against_reference_golden(actx):
    """FMMCostModel (boxtree/cost.py) against tests/golden/cost_model.json, which holds the output of
    the reference's own _PythonFMMCostModel run on the same seeded inputs
    (tests/golden/make_cost_golden.py).  Tolerance: 1e-12 relative (float64 sums of products of
    integer counts and cost factors, accumulated in a different order)."""
    import boxtree_b200
    from tests.golden.make_cost_golden import CALIBRATION, cases, level_to_order
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cost_model.json")))
    for name, (src, tkw, vkw) in cases().items():
        tree, trav = _build(actx, src, tkw, vkw)
        for factory in ("make_pde_aware_translation_cost_model", "make_taylor_translation_cost_model"):
            want = golden[f"{name}/{factory}"]
            model = boxtree_b200.FMMCostModel(getattr(boxtree_b200, factory))
            lto = level_to_order(tree.nlevels)
            per_box = model.cost_per_box(actx, trav, lto, dict(CALIBRATION)).cpu().numpy()
            assert per_box.shape == (want["nboxes"],) and per_box.dtype == np.float64
            assert np.allclose(per_box[:64], want["per_box_head"], rtol=1e-12, atol=0)
            assert np.allclose(per_box[::97], want["per_box_every_97th"], rtol=1e-12, atol=0)
            assert np.isclose(per_box.sum(), want["per_box_sum"], rtol=1e-12, atol=0)
            per_stage = model.cost_per_stage(actx, trav, lto, dict(CALIBRATION))
            assert set(per_stage) == set(want["per_stage"])
            for k, v in want["per_stage"].items():
                assert np.isclose(per_stage[k], v, rtol=1e-12, atol=0), k
    unit = boxtree_b200.FMMCostModel.get_unit_calibration_params()
    est = boxtree_b200.FMMCostModel().estimate_calibration_params(
        [per_stage], [{k: {"wall_elapsed": 2.0 * v} for k, v in per_stage.items()}])
    assert set(est) == set(unit) and all(np.isclose(v, 2.0) or v == 0.0 for v in est.values())


def test_random_parity_sweep(actx, builders):
    """Ten seconds of tests/random_sweep.py (random shapes, options and seeds; ~600 cases): every
    array bit for bit.  A 140 s run over two master seeds (9210 cases) is recorded in
    profiles/r01_random_sweep.txt."""
    import time
    from tests.random_sweep import random_case
    tb, travs = builders
    rng = np.random.default_rng(7)
    t0, n = time.time(), 0
    while time.time() - t0 < 10.0:
        case = random_case(rng)
        bad = run_case(case, actx, tb, travs)
        n += 1
        assert not bad, ({k: v for k, v in case.items() if not k.startswith("_")}, bad[:6])
    assert n > 50


def test_error_behaviour(actx):
    """Invalid calls raise what the reference raises: the table of
    ``tests.parity_util.error_cases`` is checked against the reference's own code in
    ``tests/test_refexec.py``."""
    import boxtree_b200
    from boxtree_b200 import FMMTraversalBuilder, TreeBuilder
    from tests.parity_util import error_cases
    tb = TreeBuilder(actx)
    for name, particles, tkw, vkw, exc_name in error_cases():
        exc = getattr(boxtree_b200, exc_name, None) or getattr(__import__("builtins"), exc_name)
        dkw = {k: (actx.from_numpy(v) if isinstance(v, np.ndarray) else
                   [actx.from_numpy(x) for x in v] if k == "targets" else v) for k, v in tkw.items()}
        with pytest.raises(exc):
            tree, _ = tb(actx, [actx.from_numpy(p) for p in particles], **dkw)
            assert vkw is not None, f"{name}: the tree was built"
            FMMTraversalBuilder(actx, **vkw)(actx, tree)
    src = [actx.from_numpy(s) for s in normal_particles(100, 2, np.float64)]
    with pytest.warns(DeprecationWarning):
        tb(actx, src, max_particles_in_box=10, allocator=object())


def test_structured_bbox_and_wait_for(actx):
    """The bounding box in the reference's structured form (tree_build.py:479-488) and inputs
    produced on another stream handed over through ``wait_for``."""
    import torch
    from boxtree_b200 import TreeBuilder
    from tests.parity_util import tree_mismatches
    from oracle.tree_build import build_tree
    src = normal_particles(20000, 3, np.float64)
    lo, hi = -7.5, 8.5
    sb = np.empty(1, np.dtype([(f"{m}_{ax}", np.float64) for ax in "xyz" for m in ("min", "max")]))
    for ax in "xyz":
        sb[f"min_{ax}"], sb[f"max_{ax}"] = lo, hi
    side = torch.cuda.Stream(device=actx.device)
    with torch.cuda.stream(side):
        dsrc = [torch.from_numpy(s).to(actx.device, non_blocking=True) * 1.0 for s in src]
        ev = torch.cuda.Event()
        ev.record(side)
    tree, _ = TreeBuilder(actx)(actx, dsrc, max_particles_in_box=30, bbox=sb, wait_for=[ev])
    ref = build_tree(src, max_particles_in_box=30, bbox=np.array([[lo, hi]] * 3))
    assert not tree_mismatches(ref, actx.to_numpy(tree))
