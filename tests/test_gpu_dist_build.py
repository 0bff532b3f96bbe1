"""GPU parity of the DISTRIBUTED TREE BUILD (per-level all-reduced box counts, particles kept
on their ranks, one all-to-all to the local trees): several in-process ranks share the GPU
through ``ThreadComm``.  Every rank's global box arrays must equal the oracle's tree of the
concatenated input bit for bit; its local tree / local traversal / index arrays must equal the
oracle's restatement of the reference's distributed flow and the digests of the reference's
own distributed run (``tests/golden/refexec_distributed_digests.json``)."""
import json
import os

import numpy as np
import pytest

from tests.dist_build_util import check_rank, oracle_ranks, run_rank, run_threads
from tests.dist_cases import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
@pytest.mark.parametrize("name", sorted(CASES))
def test_distributed_build_matches_reference_flow(actx, name, nranks):
    src, tkw, vkw = CASES[name]()
    rtree, want = oracle_ranks(src, tkw, vkw, nranks)
    with open(os.path.join(os.path.dirname(__file__), "golden",
                           "refexec_distributed_digests.json")) as f:
        reference = json.load(f).get(f"{name}:{nranks}")
    outs = run_threads(nranks, lambda comm: run_rank(actx, comm, src, tkw, vkw))
    for r in range(nranks):
        bad = check_rank(actx, r, nranks, outs[r], rtree, want[r],
                         None if reference is None else reference[r])
        assert not bad, (r, bad[:10])
    # every particle reaches exactly one rank as a target
    assert sum(int(o[4].shape[0]) for o in outs) == rtree.ntargets


@pytest.mark.parametrize("name", sorted(CASES)[:3])
def test_distributed_build_without_deferred_extents(actx, name):
    """``run_rank`` defers the extents' all-reduce by default (as the bench does); the build that
    completes the extents itself must give the same arrays, also before any setup call."""
    src, tkw, vkw = CASES[name]()
    rtree, want = oracle_ranks(src, tkw, vkw, 2)
    outs = run_threads(2, lambda comm: run_rank(actx, comm, src, tkw, vkw, defer=False))
    for r in range(2):
        bad = check_rank(actx, r, 2, outs[r], rtree, want[r])
        assert not bad, (r, bad[:10])


@pytest.mark.parametrize("below", ["0", "1e30"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_distributed_build_extents_variants(actx, monkeypatch, name, below):
    """The own-particle pass of the box extents has a one-lane-per-box variant for ranks that hold
    only a particle or two of each box (chosen by density, ``BT_EXTENTS_SPARSE_BELOW``): both
    variants, forced, give the reference's arrays."""
    monkeypatch.setenv("BT_EXTENTS_SPARSE_BELOW", below)
    src, tkw, vkw = CASES[name]()
    rtree, want = oracle_ranks(src, tkw, vkw, 3)
    outs = run_threads(3, lambda comm: run_rank(actx, comm, src, tkw, vkw))
    for r in range(3):
        bad = check_rank(actx, r, 3, outs[r], rtree, want[r])
        assert not bad, (r, bad[:10])


@pytest.mark.parametrize("kind", ["adaptive", "non-adaptive", "adaptive-level-restricted"])
@pytest.mark.parametrize("dims", [1, 2, 3])
def test_distributed_build_kinds(actx, kind, dims):
    """Tree kinds / dimensions / an empty rank slice / skip_prune through the distributed build."""
    from tests.parity_util import normal_particles
    n = 3000 if kind == "non-adaptive" else 20000
    src = normal_particles(n, dims, np.float64, seed=3)
    tkw = dict(max_particles_in_box=25, kind=kind)
    if kind == "non-adaptive":
        tkw["skip_prune"] = False
    rtree, want = oracle_ranks(src, tkw, {}, 3)
    outs = run_threads(3, lambda comm: run_rank(actx, comm, src, tkw, {}))
    for r in range(3):
        bad = check_rank(actx, r, 3, outs[r], rtree, want[r])
        assert not bad, (r, bad[:10])
