"""Drivers: run the reference's ``FMMTraversalBuilder`` / ``TreeBuilder`` (unmodified host code and
kernel text, see ``tests/refexec/__init__.py``) and return plain-numpy results."""
from __future__ import annotations

import dataclasses
import types

import numpy as np

from . import reference_modules

TRAV_ARRAYS = (
    "source_boxes", "target_boxes", "source_parent_boxes", "target_or_target_parent_boxes",
    "level_start_source_box_nrs", "level_start_target_box_nrs",
    "level_start_source_parent_box_nrs", "level_start_target_or_target_parent_box_nrs",
    "same_level_non_well_sep_boxes_starts", "same_level_non_well_sep_boxes_lists",
    "neighbor_source_boxes_starts", "neighbor_source_boxes_lists",
    "from_sep_siblings_starts", "from_sep_siblings_lists",
    "from_sep_close_smaller_starts", "from_sep_close_smaller_lists",
    "from_sep_bigger_starts", "from_sep_bigger_lists",
    "from_sep_close_bigger_starts", "from_sep_close_bigger_lists",
)

_TREE_DEVICE_FIELDS = (
    "level_start_box_nrs", "box_source_starts", "box_source_counts_nonchild",
    "box_source_counts_cumul", "box_target_starts", "box_target_counts_nonchild",
    "box_target_counts_cumul", "box_parent_ids", "box_child_ids", "box_centers", "box_levels",
    "box_flags", "box_source_bounding_box_min", "box_source_bounding_box_max",
    "box_target_bounding_box_min", "box_target_bounding_box_max",
)


def tree_for_reference(fakecl, tree, actx):
    """A duck-typed stand-in for ``boxtree.Tree`` holding the arrays of *tree* (anything with the
    reference's field names, e.g. ``oracle.tree_build.OracleTree``) as device arrays."""
    ns = types.SimpleNamespace()
    for name in ("sources_are_targets", "sources_have_extent", "targets_have_extent",
                 "particle_id_dtype", "box_id_dtype", "coord_dtype", "box_level_dtype",
                 "root_extent", "stick_out_factor", "extent_norm", "_is_pruned", "dimensions",
                 "nboxes", "nlevels", "aligned_nboxes", "bounding_box"):
        if hasattr(tree, name):
            setattr(ns, name, getattr(tree, name))
    for name in _TREE_DEVICE_FIELDS:
        v = getattr(tree, name, None)
        setattr(ns, name, None if v is None else fakecl.Array(np.ascontiguousarray(v), actx.queue))
    ns.level_start_box_nrs = np.asarray(tree.level_start_box_nrs)      # host array in the reference
    return ns


def reference_traversal(tree, well_sep_is_n_away=1, from_sep_smaller_crit=None,
                        _from_sep_smaller_min_nsources_cumul=None,
                        source_boxes_mask=None, source_parent_boxes_mask=None,
                        merge_close_lists=False):
    """``FMMTraversalBuilder(actx, ...)(actx, tree, ...)`` of ``boxtree/traversal.py:1969-2345``.
    Returns a namespace of numpy arrays with ``FMMTraversalInfo``'s field names (compare with
    ``tests.parity_util.trav_mismatches``)."""
    with reference_modules() as fakecl:
        import boxtree.traversal as trav_mod
        return _reference_traversal(
            fakecl, trav_mod, tree, well_sep_is_n_away, from_sep_smaller_crit,
            _from_sep_smaller_min_nsources_cumul, source_boxes_mask, source_parent_boxes_mask,
            merge_close_lists)


def _reference_traversal(fakecl, trav_mod, tree, well_sep_is_n_away, from_sep_smaller_crit,
                         _from_sep_smaller_min_nsources_cumul, source_boxes_mask,
                         source_parent_boxes_mask, merge_close_lists=False):
    assert trav_mod.__file__.startswith("/root/reference/")
    actx = fakecl.PyOpenCLArrayContext()
    rtree = tree_for_reference(fakecl, tree, actx)
    builder = trav_mod.FMMTraversalBuilder(
        actx, well_sep_is_n_away=well_sep_is_n_away, from_sep_smaller_crit=from_sep_smaller_crit)
    kwargs = {}
    if source_boxes_mask is not None:
        kwargs["source_boxes_mask"] = fakecl.Array(np.asarray(source_boxes_mask, np.int8))
    if source_parent_boxes_mask is not None:
        kwargs["source_parent_boxes_mask"] = fakecl.Array(
            np.asarray(source_parent_boxes_mask, np.int8))
    info, _ = builder(
        actx, rtree,
        _from_sep_smaller_min_nsources_cumul=_from_sep_smaller_min_nsources_cumul, **kwargs)

    if merge_close_lists:
        # FMMTraversalInfo.merge_close_lists (traversal.py:1650-1693)
        info = info.merge_close_lists(actx)
    return _trav_namespace(fakecl, info)


def _trav_namespace(fakecl, info):
    def host(x):
        return None if x is None else np.asarray(fakecl._unwrap(x))

    out = types.SimpleNamespace(**{name: host(getattr(info, name)) for name in TRAV_ARRAYS})
    out.well_sep_is_n_away = info.well_sep_is_n_away
    out.from_sep_smaller_by_level = [
        types.SimpleNamespace(**{
            f.name: (getattr(bl, f.name) if f.name in ("count", "num_nonempty_lists")
                     else host(getattr(bl, f.name)))
            for f in dataclasses.fields(bl)})
        for bl in info.from_sep_smaller_by_level]
    out.target_boxes_sep_smaller_by_source_level = [
        host(a) for a in info.target_boxes_sep_smaller_by_source_level]
    return out


# {{{ tree build

_TREE_SCALARS = ("sources_are_targets", "sources_have_extent", "targets_have_extent",
                 "particle_id_dtype", "box_id_dtype", "coord_dtype", "box_level_dtype",
                 "root_extent", "stick_out_factor", "extent_norm", "_is_pruned")
_TREE_ARRAYS = _TREE_DEVICE_FIELDS + ("user_source_ids", "sorted_target_ids", "source_radii",
                                      "target_radii")


def reference_tree(particles, **kwargs):
    """``TreeBuilder(actx)(actx, particles, **kwargs)`` of ``boxtree/tree_build.py:145-1870``,
    unmodified host code and kernels.  Returns a namespace of numpy arrays with ``Tree``'s field
    names (plus ``nboxes``, ``nlevels``, ``dimensions``, ``aligned_nboxes``)."""
    with reference_modules() as fakecl:
        import boxtree.tree_build as tb_mod
        assert tb_mod.__file__.startswith("/root/reference/")
        actx = fakecl.PyOpenCLArrayContext()

        def dev(x):
            return None if x is None else fakecl.Array(np.array(x, copy=True), actx.queue)

        from pytools import obj_array
        kw = dict(kwargs)
        for name in ("source_radii", "target_radii", "refine_weights"):
            if kw.get(name) is not None:
                kw[name] = dev(kw[name])
        if kw.get("targets") is not None:
            kw["targets"] = obj_array.new_1d([dev(t) for t in kw["targets"]])
        parts = obj_array.new_1d([dev(p) for p in particles])
        tree, _ = tb_mod.TreeBuilder(actx)(actx, parts, **kw)

        return _tree_namespace(fakecl, tree)


def _tree_namespace(fakecl, tree):
    def host(x):
        if x is None:
            return None
        if isinstance(x, np.ndarray) and x.dtype == object:
            return [host(v) for v in x]
        return np.asarray(fakecl._unwrap(x))

    ns = types.SimpleNamespace()
    for name in _TREE_SCALARS:
        v = getattr(tree, name)
        setattr(ns, name, host(v) if isinstance(v, fakecl.Array) else v)
    for name in _TREE_ARRAYS:
        setattr(ns, name, host(getattr(tree, name)))
    ns.sources = host(tree.sources)
    ns.targets = host(tree.targets)
    ns.bounding_box = tuple(np.asarray(b) for b in tree.bounding_box)
    ns.nboxes = int(tree.nboxes)
    ns.nlevels = int(tree.nlevels)
    ns.dimensions = int(tree.dimensions)
    ns.aligned_nboxes = int(tree.aligned_nboxes)
    ns.nsources = int(tree.nsources)
    ns.ntargets = int(tree.ntargets)
    return ns

# }}}


# {{{ distributed setup

class _World:
    def __init__(self, size):
        import threading
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [None] * size
        self.box = None


class ThreadComm:
    """The handful of mpi4py calls the reference's setup makes, between threads of one process."""

    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def Get_rank(self):  # noqa: N802
        return self.rank

    def Get_size(self):  # noqa: N802
        return self.world.size

    def Scatter(self, sendbuf, recvbuf, root=0):  # noqa: N802
        if self.rank == root:
            self.world.box = np.asarray(sendbuf)
        self.world.barrier.wait()
        recvbuf[...] = self.world.box[self.rank]
        self.world.barrier.wait()

    def Gather(self, sendbuf, recvbuf, root=0):  # noqa: N802
        self.world.slots[self.rank] = np.array(sendbuf, copy=True)
        self.world.barrier.wait()
        if self.rank == root:
            recvbuf[...] = np.stack(self.world.slots)
        self.world.barrier.wait()

    def bcast(self, obj, root=0):
        if self.rank == root:
            self.world.box = obj
        self.world.barrier.wait()
        result = self.world.box
        self.world.barrier.wait()
        return result

    def gather(self, obj, root=0):
        self.world.slots[self.rank] = obj
        self.world.barrier.wait()
        result = list(self.world.slots) if self.rank == root else None
        self.world.barrier.wait()
        return result


_LOCAL_TREE_FIELDS = (
    "box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
    "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul", "box_flags",
    "box_parent_ids", "box_levels", "box_child_ids", "box_centers", "source_radii", "target_radii",
    "box_to_user_rank_starts", "box_to_user_rank_lists", "responsible_boxes_list",
    "responsible_boxes_mask", "ancestor_mask", "user_source_ids", "sorted_target_ids")


def reference_distributed_setup(particles, tree_kwargs, trav_kwargs, nranks, cost_per_box_fn):
    """The setup half of ``boxtree/distributed/__init__.py:150-265`` with the reference's own
    ``partition_work`` (partition.py:60-121), ``get_box_masks`` (:124-357), ``generate_local_tree``
    (local_tree.py:316-495) and ``generate_local_travs`` (local_traversal.py:34-62), one thread
    per rank.  *cost_per_box_fn(tree namespace)* gives the root rank's cost vector.  Returns
    ``(global tree, global trav, [per-rank dict])`` as numpy namespaces."""
    import threading
    with reference_modules() as fakecl:
        import boxtree.distributed.local_traversal as lt_mod
        import boxtree.distributed.local_tree as ltree_mod
        import boxtree.distributed.partition as part_mod
        import boxtree.traversal as trav_mod
        import boxtree.tree_build as tb_mod
        for m in (lt_mod, ltree_mod, part_mod):
            assert m.__file__.startswith("/root/reference/")
        actx = fakecl.PyOpenCLArrayContext()
        from pytools import obj_array

        def dev(x):
            return None if x is None else fakecl.Array(np.array(x, copy=True), actx.queue)

        kw = dict(tree_kwargs)
        for name in ("source_radii", "target_radii", "refine_weights"):
            if kw.get(name) is not None:
                kw[name] = dev(kw[name])
        if kw.get("targets") is not None:
            kw["targets"] = obj_array.new_1d([dev(t) for t in kw["targets"]])
        tree, _ = tb_mod.TreeBuilder(actx)(
            actx, obj_array.new_1d([dev(p) for p in particles]), **kw)
        ctor = {k: v for k, v in trav_kwargs.items()
                if k in ("well_sep_is_n_away", "from_sep_smaller_crit")}
        builder = trav_mod.FMMTraversalBuilder(actx, **ctor)
        global_trav, _ = builder(actx, tree)
        global_tree_np = _tree_namespace(fakecl, tree)
        cost_per_box = cost_per_box_fn(global_tree_np)

        world = _World(nranks)
        results = [None] * nranks
        errors = []

        def host(x):
            if x is None:
                return None
            if isinstance(x, np.ndarray) and x.dtype == object:
                return [host(v) for v in x]
            return np.asarray(fakecl._unwrap(x))

        def rank_main(rank):
            try:
                comm = ThreadComm(world, rank)
                rank_actx = fakecl.PyOpenCLArrayContext()
                resp = part_mod.partition_work(cost_per_box if rank == 0 else None,
                                               global_trav, comm)
                masks = part_mod.get_box_masks(rank_actx, global_trav,
                                               rank_actx.from_numpy(np.asarray(resp)))
                local_tree, src_idx, tgt_idx = ltree_mod.generate_local_tree(
                    rank_actx, global_trav, rank_actx.from_numpy(np.asarray(resp)), comm)
                local_trav = lt_mod.generate_local_travs(rank_actx, local_tree, builder)
                out = {"responsible_boxes_list": np.asarray(resp),
                       "src_idx": host(src_idx), "tgt_idx": host(tgt_idx),
                       "masks": {f: host(getattr(masks, f)) for f in (
                           "responsible_boxes", "ancestor_boxes", "point_src_boxes",
                           "multipole_src_boxes")},
                       "local_tree": {f: host(getattr(local_tree, f)) for f in _LOCAL_TREE_FIELDS},
                       "local_trav": _trav_namespace(fakecl, local_trav)}
                out["local_tree"]["sources"] = host(local_tree.sources)
                out["local_tree"]["targets"] = host(local_tree.targets)
                results[rank] = out
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
                world.barrier.abort()

        threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(nranks)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return global_tree_np, _trav_namespace(fakecl, global_trav), results

# }}}


# {{{ area queries

def reference_area_queries(tree, ball_centers, ball_radii):
    """``PeerListFinder``, ``AreaQueryBuilder``, ``LeavesToBallsLookupBuilder`` and
    ``SpaceInvaderQueryBuilder`` of ``boxtree/area_query.py`` on *tree* (any object with ``Tree``'s
    fields).  Returns a dict of numpy arrays."""
    with reference_modules() as fakecl:
        import boxtree.area_query as aq
        assert aq.__file__.startswith("/root/reference/")
        actx = fakecl.PyOpenCLArrayContext()
        rtree = tree_for_reference(fakecl, tree, actx)
        rtree.bounding_box = tuple(np.asarray(b) for b in tree.bounding_box)
        from pytools import obj_array
        centers = obj_array.new_1d([fakecl.Array(np.array(c, copy=True)) for c in ball_centers])
        radii = fakecl.Array(np.array(ball_radii, copy=True))

        def host(x):
            return np.asarray(fakecl._unwrap(x))

        peers, _ = aq.PeerListFinder(actx)(actx, rtree)
        res, _ = aq.AreaQueryBuilder(actx)(actx, rtree, centers, radii, peer_lists=peers)
        lbl, _ = aq.LeavesToBallsLookupBuilder(actx)(actx, rtree, centers, radii, peer_lists=peers)
        sq, _ = aq.SpaceInvaderQueryBuilder(actx)(actx, rtree, centers, radii, peer_lists=peers)
        return {"peer_list_starts": host(peers.peer_list_starts),
                "peer_lists": host(peers.peer_lists),
                "leaves_near_ball_starts": host(res.leaves_near_ball_starts),
                "leaves_near_ball_lists": host(res.leaves_near_ball_lists),
                "balls_near_box_starts": host(lbl.balls_near_box_starts),
                "balls_near_box_lists": host(lbl.balls_near_box_lists),
                "outer_space_invader_dists": host(sq)}

# }}}


# {{{ a session: real reference objects for the rows that need them

class Session:
    """``with Session() as s:`` -- the reference's modules are importable inside; ``s.tree(...)``
    and ``s.traversal(...)`` return the reference's REAL ``Tree`` / ``FMMTraversalInfo``."""

    def __enter__(self):
        self._cm = reference_modules()
        self.fakecl = self._cm.__enter__()
        self.actx = self.fakecl.PyOpenCLArrayContext()
        return self

    def __exit__(self, *exc):
        return self._cm.__exit__(*exc)

    def dev(self, x):
        return None if x is None else self.fakecl.Array(np.array(x, copy=True), self.actx.queue)

    def dev_vec(self, arrays):
        from pytools import obj_array
        return obj_array.new_1d([self.dev(a) for a in arrays])

    def host(self, x):
        if x is None:
            return None
        if isinstance(x, np.ndarray) and x.dtype == object:
            return [self.host(v) for v in x]
        return np.asarray(self.fakecl._unwrap(x))

    def tree(self, particles, **kwargs):
        import boxtree.tree_build as tb_mod
        kw = dict(kwargs)
        for name in ("source_radii", "target_radii", "refine_weights"):
            if kw.get(name) is not None:
                kw[name] = self.dev(kw[name])
        if kw.get("targets") is not None:
            kw["targets"] = self.dev_vec(kw["targets"])
        tree, _ = tb_mod.TreeBuilder(self.actx)(self.actx, self.dev_vec(particles), **kw)
        return tree

    def traversal(self, tree, **kwargs):
        import boxtree.traversal as trav_mod
        ctor = {k: kwargs.pop(k) for k in ("well_sep_is_n_away", "from_sep_smaller_crit")
                if k in kwargs}
        trav, _ = trav_mod.FMMTraversalBuilder(self.actx, **ctor)(self.actx, tree, **kwargs)
        return trav


def reference_particle_filter(particles, tree_kwargs, flags):
    """``ParticleListFilter`` (``boxtree/tree.py:1057-1239``) on the reference's own tree."""
    with Session() as s:
        import boxtree.tree as tree_mod
        tree = s.tree(particles, **tree_kwargs)
        plf = tree_mod.ParticleListFilter(s.actx)
        u = plf.filter_target_lists_in_user_order(s.actx, tree, s.dev(flags))
        t = plf.filter_target_lists_in_tree_order(s.actx, tree, s.dev(flags))
        return {"user.nfiltered_targets": int(u.nfiltered_targets),
                "user.target_starts": s.host(u.target_starts),
                "user.target_lists": s.host(u.target_lists),
                "tree.nfiltered_targets": int(t.nfiltered_targets),
                "tree.box_target_starts": s.host(t.box_target_starts),
                "tree.box_target_counts_nonchild": s.host(t.box_target_counts_nonchild),
                "tree.unfiltered_from_filtered_target_indices":
                    s.host(t.unfiltered_from_filtered_target_indices),
                "tree.targets": s.host(t.targets)}


def reference_link_point_sources(particles, tree_kwargs, point_source_starts, point_sources):
    """``link_point_sources`` (``boxtree/tree.py:772-955``) on the reference's own tree."""
    with Session() as s:
        import boxtree.tree as tree_mod
        tree = s.tree(particles, **tree_kwargs)
        linked = tree_mod.link_point_sources(
            s.actx, tree, s.dev(point_source_starts), s.dev_vec(point_sources))
        out = {f: s.host(getattr(linked, f)) for f in (
            "point_source_starts", "point_source_counts", "point_sources",
            "user_point_source_ids", "box_point_source_starts",
            "box_point_source_counts_nonchild", "box_point_source_counts_cumul")}
        out["npoint_sources"] = int(linked.npoint_sources)
        return out


def reference_translation_classes(particles, tree_kwargs, trav_kwargs, per_level=True):
    """``TranslationClassesBuilder`` (``boxtree/translation_classes.py:244-445``)."""
    with Session() as s:
        import boxtree.translation_classes as tc_mod
        tree = s.tree(particles, **tree_kwargs)
        trav = s.traversal(tree, **dict(trav_kwargs))
        info, _ = tc_mod.TranslationClassesBuilder(s.actx)(
            s.actx, trav, tree, is_translation_per_level=per_level)
        return {f: s.host(getattr(info, f)) for f in (
            "from_sep_siblings_translation_classes",
            "from_sep_siblings_translation_class_to_distance_vector",
            "from_sep_siblings_translation_classes_level_starts")}

# }}}


def reference_rotation_classes(particles, tree_kwargs, trav_kwargs):
    """``RotationClassesBuilder`` (``boxtree/rotation_classes.py:90-190``)."""
    with Session() as s:
        import boxtree.rotation_classes as rc_mod
        tree = s.tree(particles, **tree_kwargs)
        trav = s.traversal(tree, **dict(trav_kwargs))
        info, _ = rc_mod.RotationClassesBuilder(s.actx)(s.actx, trav, tree)
        return {f: s.host(getattr(info, f)) for f in (
            "from_sep_siblings_rotation_classes", "from_sep_siblings_rotation_class_to_angle")}


def reference_cost_model(particles, tree_kwargs, trav_kwargs, factory, level_to_order,
                         calibration_params):
    """The reference's DEVICE cost model, ``FMMCostModel`` (``boxtree/cost.py:715-1262``, OpenCL
    kernels), on the reference's own tree and traversal: ``(cost_per_box, cost_per_stage)``."""
    with Session() as s:
        import boxtree.cost as cost
        tree = s.tree(particles, **tree_kwargs)
        trav = s.traversal(tree, **dict(trav_kwargs))
        model = cost.FMMCostModel(getattr(cost, factory))
        per_box = model.cost_per_box(s.actx, trav, level_to_order, dict(calibration_params))
        per_stage = model.cost_per_stage(s.actx, trav, level_to_order, dict(calibration_params))
        return s.host(per_box), {k: float(s.fakecl._unwrap(v)) for k, v in per_stage.items()}
