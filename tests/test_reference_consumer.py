"""The oracle's Tree / FMMTraversalInfo consumed by the REFERENCE's own, unmodified
``drive_fmm`` + ``ConstantOneExpansionWrangler`` (``/root/reference/boxtree/fmm.py:342-532``,
``boxtree/constant_one.py:50-237``; the reference's ``test/test_fmm.py:166-391``): every target
must receive the sum of all source weights.  This pins the MEANING of every list of the
oracle (which box list indexes which CSR, per-level list 3 with its own target-box lists,
close lists, particle orderings) to reference code executed here -- the strongest pin
available without pyopencl.  Skipped where /root/reference is not mounted (the GPU box)."""
import numpy as np
import pytest

from oracle.fmm import constant_one_fmm
from oracle.traversal import build_traversal, merge_close_lists
from oracle.tree_build import build_tree
from tests import reference_consumer as rc
from tests.parity_util import normal_particles

pytestmark = pytest.mark.skipif(not rc.available(), reason="/root/reference is not mounted")


def _radii(n, seed=13):
    return 0.05 * 2 ** np.random.default_rng(seed).uniform(-10, 0, n)


CASES = {
    # test/test_fmm.py:166-215 option rows (dims, nsources, ntargets, tree and traversal options)
    "2d-same": (2, 3000, None, {}, {}),
    "3d-same": (3, 4000, None, {}, {}),
    "3d-lr": (3, 4000, None, {"kind": "adaptive-level-restricted"}, {}),
    "2d-src-tgt": (2, 3000, 2500, {}, {}),
    "3d-src-tgt-2away": (3, 3000, 2500, {}, {"well_sep_is_n_away": 2}),
    "3d-ext-linf-static": (3, 3000, 2500, {"radii": True, "extent_norm": "linf"},
                           {"from_sep_smaller_crit": "static_linf"}),
    "3d-ext-linf-precise": (3, 3000, 2500, {"radii": True, "extent_norm": "linf"},
                            {"from_sep_smaller_crit": "precise_linf"}),
    "2d-ext-l2-static": (2, 3000, 2500, {"radii": True, "extent_norm": "l2"},
                         {"from_sep_smaller_crit": "static_l2"}),
    "3d-ext-lr-2away": (3, 3000, 2500, {"radii": True, "extent_norm": "linf",
                                        "kind": "adaptive-level-restricted"},
                        {"well_sep_is_n_away": 2}),
    "3d-ext-minsrc": (3, 3000, 2500, {"radii": True, "extent_norm": "linf"},
                      {"_from_sep_smaller_min_nsources_cumul": 40}),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_drive_fmm_on_oracle_lists(name):
    drive_fmm, Wrangler, TreeIndep = rc.load()
    dims, ns, nt, topt, vopt = CASES[name]
    topt = dict(topt)
    src = normal_particles(ns, dims, np.float64, seed=12)
    tkw = dict(max_particles_in_box=30)
    if nt:
        tkw["targets"] = normal_particles(nt, dims, np.float64, seed=19)
    if topt.pop("radii", False):
        tkw.update(target_radii=_radii(nt), stick_out_factor=0.25)
    tkw.update(topt)
    tree = build_tree(src, **tkw)
    trav = build_traversal(tree, **vopt)
    weights = np.random.default_rng(3).integers(1, 6, ns).astype(np.float64)
    wrangler = Wrangler(TreeIndep(), trav)
    pot = drive_fmm(None, wrangler, (weights,))
    assert pot.shape == (tree.ntargets,)
    assert np.all(pot == weights.sum())
    # and the oracle's own restatement of that driver agrees with the reference's
    assert np.array_equal(pot, constant_one_fmm(tree, trav, weights))
    if trav.from_sep_close_smaller_starts is not None:
        merged = merge_close_lists(trav)
        pot2 = drive_fmm(None, Wrangler(TreeIndep(), merged), (weights,))
        assert np.all(pot2 == weights.sum())


class _ScatterComm:
    """mpi4py-style communicator for one emulated rank; the root's send buffer is shared."""

    def __init__(self, rank, size, shared):
        self.rank, self.size, self.shared = rank, size, shared

    def Get_rank(self):  # noqa: N802
        return self.rank

    def Get_size(self):  # noqa: N802
        return self.size

    def Scatter(self, sendbuf, recvbuf, root=0):  # noqa: N802
        if self.rank == root:
            self.shared["segments"] = np.array(sendbuf, copy=True)
        recvbuf[:] = self.shared["segments"][self.rank]


@pytest.mark.parametrize("nranks", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("dims", [2, 3])
def test_reference_partition_work_matches_oracle(dims, nranks):
    """Row c1: the reference's own ``get_box_ids_dfs_order`` / ``partition_work``
    (``boxtree/distributed/partition.py:38-121``, executed in place) against the oracle's
    restatement, with integer and with fractional costs."""
    from types import SimpleNamespace
    from oracle import distributed as od
    ref_dfs, ref_partition = rc.load_partition()
    src = normal_particles(6000, dims, np.float64, seed=12)
    tree = build_tree(src, max_particles_in_box=20)
    assert np.array_equal(ref_dfs(tree), od.get_box_ids_dfs_order(tree))
    nb = tree.nboxes
    rng = np.random.default_rng(4)
    for cost in ((1.0 + tree.box_source_counts_nonchild[:nb]).astype(np.float64),
                 rng.random(nb) * 3.0 + 0.01):
        want, _ = od.partition_work(cost, tree, nranks)
        shared = {}
        trav = SimpleNamespace(tree=tree)
        for r in range(nranks):
            got = ref_partition(cost if r == 0 else None, trav, _ScatterComm(r, nranks, shared))
            assert np.array_equal(got, want[r]), (r, nranks)
