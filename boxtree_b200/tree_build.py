"""``TreeBuilder``: drop-in for ``boxtree.tree_build.TreeBuilder`` on B200.

Same call signature, argument meaning, error behaviour and outputs as the
reference (``boxtree/tree_build.py:93-1878``); the work is done by hand-written
sm_100a kernels in ``libboxtree_b200.so`` (``include/boxtree_b200.h``) on
torch-owned device buffers.  There is no CPU fallback.

Algorithm: Morton keys for all levels at once -> one stable
one-sweep radix sort -> per-level split loop on box data only (child ranges by
binary search in the sorted keys) -> pruning / level-major renumbering ->
source/target split, permutation, flags, particle extents.  The host keeps the
reference's loop control (``tree_build.py:698-1276``) with one small readback
per level.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any

import numpy as np
import torch

from . import _cabi
from ._timing import mark
from ._cabi import (CTL_LR_FOUND, CTL_NBOXES, CTL_NBOXES_FINAL, CTL_NHUGE, CTL_NSPLIT,
                    CTL_NSPLIT_REGULAR, CTL_OVERFLOW, CTL_OVERSIZE, CTL_SIZE, STEP_ALL, STEP_COMMIT,
                    STEP_CREATE, STEP_DECIDE, bt_box_out, bt_particles, bt_pool, check, dptr)
from .array_context import TorchArrayContext, make_obj_array, numpy_dtype_of
from .tree import Tree, box_flags_enum


class MaxLevelsExceeded(RuntimeError):  # noqa: N818  (reference name, tree_build.py:79)
    pass


EXTENT_NORM_CODE = {None: 0, "linf": 1, "l2": 2}
_LEAF_SMEM_CAP = 4096


class _Pool:
    """Creation-order box pool (device), grown by doubling."""

    def __init__(self, actx, dim, coord_torch_dtype, capacity, distributed=False):
        self.actx, self.dim, self.cdt = actx, dim, coord_torch_dtype
        self.distributed = distributed
        self.capacity = 0
        self.arrays: dict[str, torch.Tensor] = {}
        self.center: list[torch.Tensor] = []
        self._alloc(capacity)

    def _alloc(self, capacity):
        dev = self.actx.device
        old = self.arrays
        oldc = self.center
        n_old = self.capacity
        spec = {"start": torch.int32, "count": torch.int32, "level": torch.uint8,
                "parent": torch.int32, "child0": torch.int32, "has_children": torch.uint8,
                "force_split": torch.uint8, "nonchild": torch.int32}
        if self.distributed:
            # sums over ranks of start/count/nonchild (include/boxtree_b200.h, bt_pool)
            spec.update(gstart=torch.int32, gcount=torch.int32, gnonchild=torch.int32)
        self.arrays = {}
        # (every field of a box is written when the box is created: no need to clear)
        for name, dt in spec.items():
            a = torch.empty(capacity, dtype=dt, device=dev)
            if n_old:
                a[:n_old].copy_(old[name])
            self.arrays[name] = a
        self.center = []
        for ax in range(self.dim):
            a = torch.empty(capacity, dtype=self.cdt, device=dev)
            if n_old:
                a[:n_old].copy_(oldc[ax])
            self.center.append(a)
        # the retry after CTL_OVERFLOW re-reads the split list of the decide scan: keep it
        old_split, old_flag = getattr(self, "split_list", None), getattr(self, "flag", None)
        self.split_list = torch.empty(capacity, dtype=torch.int32, device=dev)
        self.flag = torch.empty(capacity, dtype=torch.uint8, device=dev)
        if n_old:
            self.split_list[:n_old].copy_(old_split)
            self.flag[:n_old].copy_(old_flag)
        self.xch = torch.empty(2 * capacity, dtype=torch.int32, device=dev) \
            if self.distributed else None
        self.capacity = capacity

    def ensure(self, capacity):
        if capacity > self.capacity:
            new = self.capacity
            while new < capacity:
                new *= 2
            self._alloc(new)

    def struct(self) -> bt_pool:
        p = bt_pool()
        for name in ("start", "count", "level", "parent", "child0", "has_children",
                     "force_split", "nonchild"):
            setattr(p, name, dptr(self.arrays[name]))
        for ax in range(self.dim):
            p.center[ax] = dptr(self.center[ax])
        p.capacity = self.capacity
        if self.distributed:
            p.gstart, p.gcount = dptr(self.arrays["gstart"]), dptr(self.arrays["gcount"])
            p.gnonchild, p.xch = dptr(self.arrays["gnonchild"]), dptr(self.xch)
        return p


_LEVEL_START_WORDS = 128        # >= nlevels + 2 (trees end at level 31)


class _Pending:
    """Work that a deferred build still owes (``_defer_extents``): ``finish()`` runs it once."""

    def __init__(self, fn):
        self._fn = fn

    def finish(self):
        fn, self._fn = self._fn, None
        if fn is not None:
            fn()

    @property
    def done(self):
        return self._fn is None


class TreeBuilder:
    """Builds a :class:`boxtree_b200.Tree`; mirrors ``boxtree.TreeBuilder``."""

    morton_nr_dtype = np.dtype(np.int8)
    box_level_dtype = np.dtype(np.uint8)
    ROOT_EXTENT_STRETCH_FACTOR = 1e-4

    def __init__(self, array_context: TorchArrayContext) -> None:
        assert isinstance(array_context, TorchArrayContext)
        self._setup_actx = array_context
        self._lib = _cabi.load()
        self.last_stats: dict[str, Any] = {}

    # {{{ helpers

    @staticmethod
    def _as_device_1d(actx, a, name):
        if isinstance(a, np.ndarray):
            a = actx.from_numpy(a)
        if not isinstance(a, torch.Tensor):
            raise TypeError(f"'{name}' must be a torch tensor or numpy array")
        if not a.is_cuda:
            a = a.to(actx.device)
        return a.contiguous()

    # }}}

    def __call__(self, actx: TorchArrayContext, particles, kind="adaptive",
                 max_particles_in_box=None, allocator=None, debug=False, targets=None,
                 source_radii=None, target_radii=None, stick_out_factor=None,
                 refine_weights=None, max_leaf_refine_weight=None, wait_for=None,
                 extent_norm=None, bbox=None, **kwargs):
        """See ``boxtree/tree_build.py:145-214`` for the argument documentation.

        :returns: ``(tree, event)``; *event* is a :class:`torch.cuda.Event`
            recorded on the array context's stream.
        """
        assert isinstance(actx, TorchArrayContext)
        lib = self._lib

        if allocator is not None:
            from warnings import warn
            warn("Passing in 'allocator' is deprecated. The allocator of the "
                 "array context 'actx' is used throughout.", DeprecationWarning, stacklevel=2)

        # {{{ input processing (tree_build.py:225-295)

        if kind not in ["adaptive", "adaptive-level-restricted", "non-adaptive"]:
            raise ValueError(f"unknown tree kind: '{kind}'")

        dimensions = len(particles)
        if dimensions not in (1, 2, 3):
            raise ValueError("only 1, 2 and 3 dimensions are supported")

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

        particles = [self._as_device_1d(actx, p, "particles") for p in particles]
        coord_dtypes = {p.dtype for p in particles}
        if len(coord_dtypes) != 1:
            raise ValueError("all coordinate arrays must have the same dtype")
        coord_tdtype = particles[0].dtype
        coord_dtype = numpy_dtype_of(particles[0])
        if coord_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError(f"unsupported coordinate dtype: {coord_dtype}")
        dcode = _cabi.dtype_code(coord_dtype)

        nsources = int(particles[0].shape[0])
        if any(int(p.shape[0]) != nsources for p in particles):
            raise ValueError("coordinate arrays must have the same length")
        if targets is None:
            ntargets = 0
            nsrcntgts = nsources
        else:
            targets = [self._as_device_1d(actx, t, "targets") for t in targets]
            if len(targets) != dimensions:
                raise ValueError("sources and targets must have the same dimension")
            ntargets = int(targets[0].shape[0])
            if any(int(t.shape[0]) != ntargets for t in targets):
                raise ValueError("coordinate arrays must have the same length")
            nsrcntgts = nsources + ntargets

        if source_radii is not None:
            source_radii = self._as_device_1d(actx, source_radii, "source_radii")
            if tuple(source_radii.shape) != (nsources,):
                raise ValueError("'source_radii' has an invalid shape: "
                                 f"{tuple(source_radii.shape)} (expected ({nsources},))")
            if source_radii.dtype != coord_tdtype:
                raise TypeError("dtypes of coordinate array 'particles' and 'source_radii' "
                                f"must agree: got {coord_tdtype} and {source_radii.dtype}")
        if target_radii is not None:
            target_radii = self._as_device_1d(actx, target_radii, "target_radii")
            if tuple(target_radii.shape) != (ntargets,):
                raise ValueError("'target_radii' has an invalid shape: "
                                 f"{tuple(target_radii.shape)} (expected ({ntargets},))")
            if target_radii.dtype != coord_tdtype:
                raise TypeError("dtypes of coordinate array 'particles' and 'target_radii' "
                                f"must agree: got {coord_tdtype} and {target_radii.dtype}")

        if sources_have_extent or targets_have_extent:
            if stick_out_factor is None:
                raise ValueError("if sources or targets have extent, "
                                 "'stick_out_factor' must be explicitly specified")
        else:
            stick_out_factor = 0

        if targets is not None and targets[0].dtype != coord_tdtype:
            raise TypeError("sources and targets coordinates must have same dtype: "
                            f"got {coord_tdtype} and {targets[0].dtype}")

        if nsrcntgts >= 2**30:
            raise NotImplementedError("more than 2**30 particles per device are not supported")

        # distributed build (boxtree_b200/distributed/tree_build.py): the arguments are this
        # rank's slice of the global particle set; per-box counts are summed over `comm`
        comm = kwargs.get("comm")
        dist = comm is not None
        if dist and refine_weights is not None:
            raise NotImplementedError("refine weights are not supported by the distributed build")

        # }}}

        # {{{ refine weights (tree_build.py:405-452)

        specified_max = max_particles_in_box is not None
        specified_weights = refine_weights is not None and max_leaf_refine_weight is not None
        if specified_max and specified_weights:
            raise ValueError("may only specify one of 'max_particles_in_box' and "
                             "'refine_weights'/'max_leaf_refine_weight")
        elif not specified_max and not specified_weights:
            raise ValueError("must specify either 'max_particles_in_box' or "
                             "'refine_weights'/'max_leaf_refine_weight'")
        elif specified_max:
            refine_weights = None          # every particle weighs 1
            max_leaf_refine_weight = max_particles_in_box
        else:
            refine_weights = self._as_device_1d(actx, refine_weights, "refine_weights")
            if refine_weights.dtype != torch.int32:
                raise TypeError("'refine_weights' must have dtype 'int32' "
                                f"(got {refine_weights.dtype})")

        if max_leaf_refine_weight <= 0:
            raise ValueError(
                f"'max_leaf_refine_weight' must be positive: {max_leaf_refine_weight}")

        if refine_weights is not None and nsrcntgts:
            if max_leaf_refine_weight < int(refine_weights.max()):
                raise ValueError(
                    "entries of 'refine_weights' cannot exceed 'max_leaf_refine_weight'")
            if int(refine_weights.min()) < 0:
                raise ValueError("all entries of 'refine_weights' must be nonnegative")
            total_refine_weight = int(refine_weights.sum(dtype=torch.int64))
        else:
            total_refine_weight = nsrcntgts
        nsrcntgts_global = nsrcntgts         # distributed: known after the bounding box exchange
        max_leaf_refine_weight = int(max_leaf_refine_weight)

        # }}}

        stream = actx.stream
        sh = actx.stream_handle
        level_restrict = kind == "adaptive-level-restricted"
        adaptive = kind != "non-adaptive"
        have_ext = int(srcntgts_have_extent)
        nb = 2**dimensions

        # wait_for (tree_build.py:193, 229): events (or streams) that produced the inputs
        for dep in (wait_for or ()):
            if isinstance(dep, torch.cuda.Stream):
                stream.wait_stream(dep)
            else:
                stream.wait_event(dep)

        with torch.cuda.stream(stream), torch.cuda.device(actx.device):
            mark("start")
            # {{{ particle view (virtual concatenation, tree_build.py:328-388)

            P = bt_particles()
            for ax in range(dimensions):
                P.sources[ax] = dptr(particles[ax])
                P.targets[ax] = dptr(targets[ax]) if targets is not None else None
            P.source_radii = dptr(source_radii)
            P.target_radii = dptr(target_radii)
            P.nsources = nsources
            P.ntargets = ntargets

            # }}}

            # {{{ bounding box (tree_build.py:458-508)

            if nsrcntgts_global == 0 and not dist:
                raise ValueError("cannot build a tree without particles")

            bbox_dev = actx.empty(2 * dimensions, coord_dtype)
            check(lib.bt_bounding_box(dcode, dimensions, C.byref(P), dptr(bbox_dev), sh),
                  "bt_bounding_box")
            if dist:
                # ONE small all-gather + readback: every rank's particle counts and bounding box
                # (as float64, exact for either coordinate type); min / max / sums on the host
                # (counts through pinned memory, result back through pinned memory: no staging)
                cnt_host = torch.tensor([nsources, ntargets], dtype=torch.float64).pin_memory()
                mine = torch.empty(2 + 2 * dimensions, dtype=torch.float64, device=actx.device)
                mine[:2].copy_(cnt_host, non_blocking=True)
                mine[2:] = bbox_dev
                allv = actx.read_back(comm.allgather_tensor(mine))          # [size, 2 + 2 dim]
                nsources_global, ntargets_global = (int(x) for x in allv[:, :2].sum(axis=0))
                nsrcntgts_global = total_refine_weight = nsources_global + ntargets_global
                if nsrcntgts_global >= 2**31 - 1:
                    raise NotImplementedError("more than 2**31 - 2 particles in a global tree")
                if nsrcntgts_global == 0:
                    raise ValueError("cannot build a tree without particles")
                have = allv[:, :2].sum(axis=1) > 0          # ranks without particles: no box
                bbox_auto = np.empty(2 * dimensions, coord_dtype)
                bbox_auto[0::2] = allv[have][:, 2::2].min(axis=0).astype(coord_dtype)
                bbox_auto[1::2] = allv[have][:, 3::2].max(axis=0).astype(coord_dtype)
            else:
                # (pinned: a copy to pageable memory is staged and costs ~0.1 ms more)
                bbox_host = torch.empty(2 * dimensions, dtype=coord_tdtype, pin_memory=True)
                bbox_host.copy_(bbox_dev, non_blocking=True)
                stream.synchronize()
                bbox_auto = bbox_host.numpy().copy()
            auto_min = bbox_auto[0::2].copy()
            auto_max = bbox_auto[1::2].copy()

            if bbox is None:
                root_extent = max(auto_max[i] - auto_min[i] for i in range(dimensions)) \
                    * (1 + TreeBuilder.ROOT_EXTENT_STRETCH_FACTOR)
                bbox_min = auto_min.copy()
                bbox_max = bbox_min + root_extent
            else:
                if not isinstance(bbox, np.ndarray):
                    raise NotImplementedError(f"unsupported bounding box type: {type(bbox)}")
                # [dimensions][2] (min, max) rows, or the length-1 structured array with fields
                # min_x, max_x, ... that the reference's bounding box finder returns
                # (tree_build.py:479-488, bounding_box.py:35-52)
                structured = bbox.dtype.names is not None
                if not structured:
                    assert len(bbox) == dimensions
                else:
                    assert len(bbox) == 1
                bbox_min = np.empty(dimensions, coord_dtype)
                bbox_max = np.empty(dimensions, coord_dtype)
                for i, ax in enumerate("xyz"[:dimensions]):
                    bbox_min[i] = bbox[f"min_{ax}"][0] if structured else bbox[i][0]
                    bbox_max[i] = bbox[f"max_{ax}"][0] if structured else bbox[i][1]
                    assert bbox_min[i] < bbox_max[i]
                    assert bbox_min[i] <= auto_min[i]
                    assert bbox_max[i] >= auto_max[i]
                bbox_exts = bbox_max - bbox_min
                for ext in bbox_exts:
                    assert abs(ext - bbox_exts[0]) < 1e-15
                root_extent = bbox_exts[0]
            root_extent = coord_dtype.type(root_extent)
            root_center = [bbox_min[d] + (bbox_max[d] - bbox_min[d]) / 2 for d in range(dimensions)]

            # }}}

            mark("tb:bbox")
            # The sort key resolves `key_depth` levels.  The fewer, the fewer radix passes: a depth
            # estimated from the particle count is tried first, the full depth of the 64-bit key if
            # the tree turns out deeper (bt_make_keys; the result is identical either way).
            max_key_level = int(lib.bt_max_key_level(dimensions))
            est = 1
            while (nb ** est) * max_leaf_refine_weight < max(total_refine_weight, 1):
                est += 1
            first_depth = int(kwargs.get("_key_depth", 0)) or min(max_key_level, est + 6)
            # x, y, z, radius of every particle side by side for the gather into tree order
            records = actx.empty(4 * max(nsrcntgts, 1), coord_dtype)
            # ... and two-word keys beyond that, up to level 31, the deepest level the reference's
            # digit expression `1U << (1 + level)` resolves (deeper: MaxLevelsExceeded in both)
            max_tree_level = int(lib.bt_max_tree_level(dimensions))
            depth_tries = ([first_depth] if first_depth < max_key_level else []) + [max_key_level]
            if max_tree_level > max_key_level:
                depth_tries.append(max_tree_level)
            if kwargs.get("_key_depth") == -1:          # testing: two-word keys right away
                depth_tries = [max_tree_level]
            for key_depth in depth_tries:
                too_deep = False
                deep = key_depth > max_key_level
                # {{{ keys + sort

                key_bufs = [actx.empty(nsrcntgts, np.int64), actx.empty(nsrcntgts, np.int64)]
                id_bufs = [actx.empty(nsrcntgts, np.int32), actx.empty(nsrcntgts, np.int32)]
                lo_bufs = [actx.empty(nsrcntgts, np.int64), actx.empty(nsrcntgts, np.int64)] \
                    if deep else [None, None]
                check(lib.bt_make_keys(dcode, dimensions, C.byref(P), _cabi.darray(bbox_min),
                                       _cabi.darray(bbox_max), EXTENT_NORM_CODE[srcntgts_extent_norm],
                                       float(stick_out_factor), key_depth, dptr(key_bufs[0]),
                                       dptr(lo_bufs[0]), dptr(records), sh), "bt_make_keys")
                in_alt = C.c_int(0)
                if nsrcntgts and deep:
                    check(lib.bt_sort_particles_deep(
                        nsrcntgts, dimensions, have_ext, dptr(key_bufs[0]), dptr(key_bufs[1]),
                        dptr(lo_bufs[0]), dptr(lo_bufs[1]), dptr(id_bufs[0]), dptr(id_bufs[1]), sh),
                        "bt_sort_particles_deep")
                elif nsrcntgts:
                    check(lib.bt_sort_particles(nsrcntgts, dimensions, have_ext, key_depth,
                                                dptr(key_bufs[0]),
                                                dptr(key_bufs[1]), dptr(id_bufs[0]), dptr(id_bufs[1]),
                                                C.byref(in_alt), sh), "bt_sort_particles")
                keys = key_bufs[in_alt.value]
                ids = id_bufs[in_alt.value]
                keys_lo = lo_bufs[0]
                del key_bufs, id_bufs, lo_bufs

                wprefix = None
                if refine_weights is not None:
                    wprefix = actx.empty(nsrcntgts + 1, np.int64)
                    check(lib.bt_weight_prefix(nsrcntgts, dptr(ids), dptr(refine_weights),
                                               dptr(wprefix), sh), "bt_weight_prefix")

                # }}}

                mark("tb:keys+sort")
                # {{{ level loop (tree_build.py:653-1276)

                nboxes_guess = kwargs.get("nboxes_guess")
                if nboxes_guess is None:
                    nboxes_guess = int(nb * ((max_leaf_refine_weight + total_refine_weight - 1)
                                             // max_leaf_refine_weight)) + 1
                assert nboxes_guess > 0
                pool = _Pool(actx, dimensions, coord_tdtype, max(int(nboxes_guess), 2), dist)
                # control words, then the level starts of the final numbering (read back together)
                ctl = actx.zeros(CTL_SIZE + _LEVEL_START_WORDS, np.int32)
                ctl_host = torch.empty(CTL_SIZE + _LEVEL_START_WORDS, dtype=torch.int32,
                                       pin_memory=True)
                check(lib.bt_pool_init(dcode, dimensions, C.byref(pool.struct()), nsrcntgts, have_ext,
                                       dptr(keys), _cabi.darray(root_center), dptr(ctl), sh),
                      "bt_pool_init")
                if dist:        # the root holds every rank's particles
                    g = torch.stack([pool.arrays["count"][0], pool.arrays["nonchild"][0]])
                    comm.allreduce_(g, "sum")
                    pool.arrays["gcount"][0:1].copy_(g[0:1])
                    pool.arrays["gnonchild"][0:1].copy_(g[1:2])

                nlevels_max = 2 * (np.finfo(coord_dtype).nmant + 1)
                level = 1 if total_refine_weight > max_leaf_refine_weight else 0
                nboxes = 1
                level_block_start = 0            # first pool id of the boxes on level-1
                final_level_restrict_iteration = False
                nreallocs = 0
                niterations = 0

                def read_ctl():
                    ctl_host.copy_(ctl, non_blocking=True)
                    stream.synchronize()
                    return ctl_host.numpy()

                while level:
                    niterations += 1
                    if level > key_depth and key_depth < depth_tries[-1] \
                            and level + 1 < nlevels_max:
                        too_deep = True         # deeper than the key resolves: sort again
                        break
                    if level + 1 >= nlevels_max or level > key_depth:
                        raise MaxLevelsExceeded(
                            "Level count exceeded number of significant "
                            "bits in coordinate dtype. That means that a large number "
                            "of particles was indistinguishable up to floating point "
                            "precision (because they ended up in the same box).")

                    lo = 0 if level_restrict else level_block_start
                    ncand = nboxes - lo
                    if adaptive and not level_restrict:
                        bound = min(ncand, total_refine_weight // (max_leaf_refine_weight + 1) + 1)
                    else:
                        bound = min(ncand, nboxes - level_block_start
                                    + (int(kwargs.get("_lr_slack", 1024)) if level_restrict else 0))
                    # Room for the children: the worst case (every candidate splits) is not
                    # reserved up front -- at the deep levels it is several times the final tree
                    # and used to double the pool (13 array copies, 0.4 ms on config 3) right
                    # before the last, empty iteration.  A step that does not fit reports
                    # CTL_OVERFLOW with the number of boxes it wants; the pool grows then.
                    pool.ensure(min(nboxes + nb * bound, max(pool.capacity, nboxes + nb)))

                    skip_if_no_regular = int(bool(srcntgts_have_extent)
                                             and not final_level_restrict_iteration)

                    def run_step(phases):
                        check(lib.bt_level_step(
                            dcode, dimensions, C.byref(pool.struct()), dptr(keys), dptr(wprefix),
                            dptr(ctl), dptr(pool.split_list), dptr(pool.flag), lo, nboxes, level,
                            max_leaf_refine_weight, int(adaptive), int(level_restrict), have_ext,
                            skip_if_no_regular, float(root_extent), phases,
                            min(key_depth, max_key_level), dptr(keys_lo), sh), "bt_level_step")
                        if phases & STEP_COMMIT and level_restrict \
                                and not final_level_restrict_iteration:
                            check(lib.bt_level_restrict(dcode, dimensions, C.byref(pool.struct()),
                                                        dptr(ctl), level, pool.capacity,
                                                        float(root_extent), sh), "bt_level_restrict")

                    # one GPU: decide + children + commit in one go; distributed: the children's
                    # local (lower bound, count, nonchild) are summed over the ranks before the commit
                    first = STEP_DECIDE | STEP_CREATE if dist else STEP_ALL
                    if dist and ncand <= 4096:
                        # few candidates (the top levels): room for all of them to split and the
                        # counts of all their children summed, so that no readback is needed
                        # between the children and the commit
                        pool.ensure(nboxes + nb * ncand)
                        run_step(first)
                        comm.allreduce_(pool.xch[:2 * nb * ncand], "sum")
                        run_step(STEP_COMMIT)
                        h = read_ctl()
                        assert not h[CTL_OVERFLOW]
                    else:
                        run_step(first)
                        h = read_ctl()
                        while h[CTL_OVERFLOW]:
                            nreallocs += 1
                            pool.ensure(nboxes + nb * int(h[CTL_NSPLIT]))
                            run_step(first & ~STEP_DECIDE)
                            h = read_ctl()
                        if dist:
                            if h[CTL_NSPLIT] and not (skip_if_no_regular
                                                      and not h[CTL_NSPLIT_REGULAR]):
                                comm.allreduce_(pool.xch[:2 * nb * int(h[CTL_NSPLIT])], "sum")
                            run_step(STEP_COMMIT)
                            h = read_ctl()

                    nsplit_regular = int(h[CTL_NSPLIT_REGULAR])
                    have_oversize_split_box = int(h[CTL_OVERSIZE])

                    if nsplit_regular == 0:
                        # no new boxes on the new level (tree_build.py:1016-1025)
                        if srcntgts_have_extent and not final_level_restrict_iteration:
                            level -= 1
                            break
                        assert final_level_restrict_iteration

                    level_block_start = nboxes
                    nboxes = int(h[CTL_NBOXES])

                    if final_level_restrict_iteration:
                        level -= 1
                        break

                    if level_restrict:
                        did_upper_level_split = bool(level - 2 >= 1 and h[CTL_LR_FOUND + level - 2])
                        if have_oversize_split_box == 0 and did_upper_level_split:
                            final_level_restrict_iteration = True
                            level += 1
                            continue

                    if not have_oversize_split_box:
                        break
                    level += 1

                if not too_deep:
                    break
                del keys, ids, pool, keys_lo

            nlevels = level + 1

            # }}}

            mark("tb:level loop")
            # {{{ prune / renumber (tree_build.py:1330-1456)

            prune_empty_leaves = not kwargs.get("skip_prune")
            map_old2new = actx.empty(nboxes, np.int32)
            src_of_new = actx.empty(nboxes, np.int32)
            assert nlevels + 2 <= _LEVEL_START_WORDS
            level_start_dev = ctl[CTL_SIZE:CTL_SIZE + nlevels + 2]
            check(lib.bt_finalize_numbering(dcode, dimensions, C.byref(pool.struct()), nboxes,
                                            int(level_restrict), int(not prune_empty_leaves),
                                            dptr(ctl), dptr(map_old2new), dptr(src_of_new),
                                            dptr(level_start_dev), sh), "bt_finalize_numbering")
            h = read_ctl()
            nfinal = int(h[CTL_NBOXES_FINAL])
            level_start_box_nrs = h[CTL_SIZE:CTL_SIZE + nlevels + 1].copy()
            level_start_box_nrs[nlevels] = nfinal

            aligned_nboxes = ((nfinal + 31) // 32) * 32
            box_srcntgt_starts = actx.empty(nfinal, np.int32)
            box_srcntgt_counts_cumul = actx.empty(nfinal, np.int32)
            box_srcntgt_counts_nonchild = actx.empty(nfinal, np.int32)
            box_levels = actx.empty(nfinal, np.uint8)
            box_parent_ids = actx.empty(nfinal, np.int32)
            # (bt_gather_boxes writes every entry of the nfinal boxes: only the padding up to
            # aligned_nboxes is cleared, not 0.7 GB of box arrays of a 12 M-box global tree)
            box_child_ids = actx.empty((nb, aligned_nboxes), np.int32)
            box_centers = actx.empty((dimensions, aligned_nboxes), coord_dtype)
            box_child_ids[:, nfinal:].zero_()
            box_centers[:, nfinal:].zero_()
            box_has_children = actx.empty(nfinal, np.uint8)
            box_real_children = actx.empty(nfinal, np.uint8)
            out = bt_box_out()
            out.box_start = dptr(box_srcntgt_starts)
            out.box_count = dptr(box_srcntgt_counts_cumul)
            out.box_nonchild = dptr(box_srcntgt_counts_nonchild)
            out.box_levels = dptr(box_levels)
            out.box_parent_ids = dptr(box_parent_ids)
            out.box_child_ids = dptr(box_child_ids)
            out.box_centers = dptr(box_centers)
            out.has_children = dptr(box_has_children)
            out.real_children = dptr(box_real_children)
            if dist:        # ranges of the boxes in this rank's particles
                lbox_start = actx.empty(nfinal, np.int32)
                lbox_count = actx.empty(nfinal, np.int32)
                lbox_nonchild = actx.empty(nfinal, np.int32)
                out.local_start, out.local_count = dptr(lbox_start), dptr(lbox_count)
                out.local_nonchild = dptr(lbox_nonchild)
            else:
                lbox_start, lbox_count = box_srcntgt_starts, box_srcntgt_counts_cumul
                lbox_nonchild = box_srcntgt_counts_nonchild
            check(lib.bt_gather_boxes(dcode, dimensions, C.byref(pool.struct()), have_ext,
                                      dptr(src_of_new), dptr(map_old2new), nfinal, aligned_nboxes,
                                      C.byref(out), sh), "bt_gather_boxes")

            # }}}

            mark("tb:prune+gather")
            # {{{ particle order inside never-partitioned boxes

            big_list = actx.empty(max(nfinal, 1), np.int32)
            huge_list = actx.empty(max(nfinal, 1), np.int32)
            check(lib.bt_leaf_fixup(nfinal, dptr(lbox_start),
                                    dptr(lbox_count), dptr(box_real_children),
                                    dptr(ids), dptr(ctl), dptr(big_list), nfinal, dptr(huge_list),
                                    sh), "bt_leaf_fixup")
            leaves_bounded = (refine_weights is None and not srcntgts_have_extent and adaptive
                              and max_leaf_refine_weight <= _LEAF_SMEM_CAP)
            if not leaves_bounded:
                h = read_ctl()
                nhuge = int(h[CTL_NHUGE])
                if nhuge:
                    hl = huge_list[:nhuge].cpu().numpy()
                    st = lbox_start.cpu().numpy()
                    cn = lbox_count.cpu().numpy()
                    for b in hl:
                        seg = ids[int(st[b]):int(st[b]) + int(cn[b])]
                        check(lib.bt_sort_u32_segment(int(cn[b]), dptr(seg), sh),
                              "bt_sort_u32_segment")
            del big_list, huge_list

            # }}}

            # {{{ sources / targets (tree_build.py:1464-1620)

            if sources_are_targets:
                user_source_ids = ids
                sorted_target_ids = actx.empty(nsrcntgts, np.int32)
                check(lib.bt_reverse_index(nsrcntgts, dptr(ids), dptr(sorted_target_ids), sh),
                      "bt_reverse_index")
                source_numbers = None
                srcntgt_target_ids = None
            else:
                source_numbers = actx.empty(nsrcntgts + 1, np.int32)
                user_source_ids = actx.empty(nsources, np.int32)
                srcntgt_target_ids = actx.empty(ntargets, np.int32)
                sorted_target_ids = actx.empty(ntargets, np.int32)
                check(lib.bt_split_sources_targets(
                    nsrcntgts, nsources, dptr(ids), dptr(source_numbers), dptr(user_source_ids),
                    dptr(srcntgt_target_ids), dptr(sorted_target_ids), sh),
                    "bt_split_sources_targets")

            def permute(from_ids, n, want_radii):
                outs = [actx.empty(n, coord_dtype) for _ in range(dimensions)]
                radii = actx.empty(n, coord_dtype) if want_radii else None
                check(lib.bt_permute(dcode, dimensions, C.byref(P), dptr(records), dptr(from_ids), n,
                                     _cabi.ptr_array(outs), dptr(radii), sh), "bt_permute")
                return make_obj_array(outs), radii

            if sources_are_targets:
                sources, _ = permute(user_source_ids, nsrcntgts, False)
                tgts = sources
                out_source_radii = out_target_radii = None
            else:
                sources, out_source_radii = permute(user_source_ids, nsources,
                                                    srcntgts_have_extent)
                tgts, out_target_radii = permute(srcntgt_target_ids, ntargets,
                                                 srcntgts_have_extent)

            # }}}

            mark("tb:leaf fixup+split+permute")
            # {{{ per-box particle ranges and flags (tree_build.py:1666-1723)

            box_flags = actx.empty(nfinal, box_flags_enum.dtype)

            if sources_are_targets:
                box_source_starts = box_target_starts = box_srcntgt_starts
                box_source_counts_cumul = box_target_counts_cumul = box_srcntgt_counts_cumul
                box_source_counts_nonchild = box_target_counts_nonchild = \
                    actx.empty(nfinal, np.int32)
            else:
                box_source_starts = actx.empty(nfinal, np.int32)
                box_source_counts_cumul = actx.empty(nfinal, np.int32)
                box_source_counts_nonchild = actx.empty(nfinal, np.int32)
                box_target_starts = actx.empty(nfinal, np.int32)
                box_target_counts_cumul = actx.empty(nfinal, np.int32)
                box_target_counts_nonchild = actx.empty(nfinal, np.int32)
            local_ranges = None
            if not dist or sources_are_targets:
                check(lib.bt_box_info(
                    nfinal, int(sources_are_targets), have_ext, dptr(box_srcntgt_starts),
                    dptr(box_srcntgt_counts_cumul), dptr(box_srcntgt_counts_nonchild),
                    dptr(box_has_children), dptr(source_numbers),
                    dptr(box_source_starts), dptr(box_source_counts_nonchild),
                    dptr(box_source_counts_cumul), dptr(box_target_starts),
                    dptr(box_target_counts_nonchild), dptr(box_target_counts_cumul),
                    dptr(box_flags), sh), "bt_box_info")
                if dist:    # the same on the rank's own ranges: only leaves own particles
                    l_own = torch.where(box_has_children != 0, torch.zeros_like(lbox_count),
                                        lbox_count)
                    local_ranges = (lbox_start, l_own, lbox_count) * 2
            else:
                # rank-local source counts per box -> sums over ranks -> global ranges and flags
                src3 = actx.empty(3 * nfinal, np.int32)
                local_ranges = tuple(actx.empty(nfinal, np.int32) for _ in range(6))
                check(lib.bt_box_info_local(
                    nfinal, have_ext, dptr(lbox_start), dptr(lbox_count), dptr(lbox_nonchild),
                    dptr(box_has_children), dptr(source_numbers), dptr(src3),
                    *[dptr(a) for a in local_ranges], sh), "bt_box_info_local")
                comm.allreduce_(src3, "sum")
                check(lib.bt_box_info_global(
                    nfinal, have_ext, dptr(box_srcntgt_starts), dptr(box_srcntgt_counts_cumul),
                    dptr(box_srcntgt_counts_nonchild), dptr(box_has_children), dptr(src3),
                    dptr(box_source_starts), dptr(box_source_counts_nonchild),
                    dptr(box_source_counts_cumul), dptr(box_target_starts),
                    dptr(box_target_counts_nonchild), dptr(box_target_counts_cumul),
                    dptr(box_flags), sh), "bt_box_info_global")
                del src3

            # }}}

            mark("tb:box info")
            # {{{ box particle extents (tree_build.py:1730-1802)

            # source and target boxes side by side: the distributed build all-reduces all minima
            # (and all maxima) in one collective
            nsets = 1 if sources_are_targets else 2
            # (every box's entry is written by the own-particle pass, seeded with the centre)
            bb_min_all = actx.empty((nsets, dimensions, aligned_nboxes), coord_dtype)
            bb_max_all = actx.empty((nsets, dimensions, aligned_nboxes), coord_dtype)
            bb_min_all[:, :, nfinal:].zero_()
            bb_max_all[:, :, nfinal:].zero_()
            bb_src_min, bb_src_max = bb_min_all[0], bb_max_all[0]
            bb_tgt_min, bb_tgt_max = (bb_src_min, bb_src_max) if sources_are_targets else \
                (bb_min_all[1], bb_max_all[1])

            # own ranges of the boxes in the (rank's) tree-ordered particle arrays
            own_src = (local_ranges[0], local_ranges[1]) if dist else \
                (box_source_starts, box_source_counts_nonchild)
            own_tgt = (local_ranges[3], local_ranges[4]) if dist else \
                (box_target_starts, box_target_counts_nonchild)
            rounds = [(sources, out_source_radii if sources_have_extent else None,
                       own_src[0], own_src[1], bb_src_min, bb_src_max)]
            if not sources_are_targets:
                rounds.append((tgts, out_target_radii if targets_have_extent else None,
                               own_tgt[0], own_tgt[1], bb_tgt_min, bb_tgt_max))
            ls_host = (C.c_int32 * (nlevels + 1))(*[int(x) for x in level_start_box_nrs])
            # distributed: min/max over the rank's own particles, all-reduced (exact), then the
            # child merge on the global values
            # A rank's share of a box of a big global tree is a particle or two: then ONE lane per
            # box (flag 4) instead of 8 -- 0.68 -> 0.17 ms for 7.5e6 particles in 8.2e6 boxes, 0.17
            # -> 0.05 ms in 1.8e6 boxes (tests/extents_sparse_probe.py; same bits).  Decided per
            # particle kind by the rank's particles per box; BT_EXTENTS_SPARSE_BELOW overrides the
            # threshold (tests force either variant).
            sparse_below = float(os.environ.get("BT_EXTENTS_SPARSE_BELOW", 3.0))

            def extents_phase(phases):
                for parts, radii, pstarts, pcounts, bmin, bmax in rounds:
                    ph = phases
                    if dist and (phases & 1) and int(parts[0].shape[0]) < sparse_below * nfinal:
                        ph |= 4
                    check(lib.bt_box_extents_phase(
                        dcode, dimensions, nfinal, aligned_nboxes, nlevels, ls_host,
                        dptr(box_child_ids), dptr(box_centers), dptr(pstarts), dptr(pcounts),
                        _cabi.ptr_array(list(parts)), dptr(radii), dptr(bmin), dptr(bmax), ph,
                        sh), "bt_box_extents")

            pending_extents = None
            if not dist:
                extents_phase(3)
            elif not kwargs.get("_defer_extents"):
                extents_phase(1)
                comm.allreduce_(bb_min_all, "min")
                comm.allreduce_(bb_max_all, "max")
                extents_phase(2)
            else:
                # the two big reductions (2 x nsets x dim x nboxes coordinates) are only enqueued:
                # they run on the backend's stream while this stream goes on with work that does
                # not read the extents (work partition, colleagues, lists 2 and 4);
                # `pending_extents.finish()` waits for them and merges the children's boxes
                extents_phase(1)
                waits = [comm.allreduce_async_(bb_min_all, "min"),
                         comm.allreduce_async_(bb_max_all, "max")]

                def finish_extents():
                    with torch.cuda.stream(stream), torch.cuda.device(actx.device):
                        for w in waits:
                            w.wait()
                        extents_phase(2)
                pending_extents = _Pending(finish_extents)

            # }}}

            # (assembled on the device: a host -> device copy of pageable memory would synchronise)
            level_start_dev_out = ctl[CTL_SIZE:CTL_SIZE + nlevels + 1].clone()
            level_start_dev_out[nlevels:] = ctl[CTL_NBOXES_FINAL:CTL_NBOXES_FINAL + 1]
            evt = torch.cuda.Event()
            evt.record(stream)

            mark("tb:extents")
        self.last_stats = {"level_iterations": niterations, "nboxes_pre_prune": nboxes,
                           "reallocs": nreallocs, "key_depth": key_depth}

        extra = {}
        cls = Tree
        if dist:
            from .distributed.tree_build import DistributedTree
            cls = DistributedTree
            extra = dict(
                local_box_source_starts=local_ranges[0],
                local_box_source_counts_nonchild=local_ranges[1],
                local_box_source_counts_cumul=local_ranges[2],
                local_box_target_starts=local_ranges[3],
                local_box_target_counts_nonchild=local_ranges[4],
                local_box_target_counts_cumul=local_ranges[5],
                nsources_global=nsources_global, ntargets_global=ntargets_global,
                rank=comm.Get_rank(), nranks=comm.Get_size(), pending=pending_extents)
        tree = cls(
            **extra,
            sources_are_targets=sources_are_targets,
            sources_have_extent=sources_have_extent,
            targets_have_extent=targets_have_extent,
            particle_id_dtype=particle_id_dtype, box_id_dtype=box_id_dtype,
            coord_dtype=coord_dtype, box_level_dtype=self.box_level_dtype,
            bounding_box=(bbox_min, bbox_max), root_extent=root_extent,
            stick_out_factor=stick_out_factor, extent_norm=srcntgts_extent_norm,
            level_start_box_nrs=level_start_dev_out,
            sources=sources, targets=tgts,
            source_radii=out_source_radii if sources_have_extent else None,
            target_radii=out_target_radii if targets_have_extent else None,
            box_source_starts=box_source_starts,
            box_source_counts_nonchild=box_source_counts_nonchild,
            box_source_counts_cumul=box_source_counts_cumul,
            box_target_starts=box_target_starts,
            box_target_counts_nonchild=box_target_counts_nonchild,
            box_target_counts_cumul=box_target_counts_cumul,
            box_parent_ids=box_parent_ids, box_child_ids=box_child_ids,
            box_centers=box_centers, box_levels=box_levels, box_flags=box_flags,
            user_source_ids=user_source_ids, sorted_target_ids=sorted_target_ids,
            box_source_bounding_box_min=bb_src_min, box_source_bounding_box_max=bb_src_max,
            box_target_bounding_box_min=bb_tgt_min, box_target_bounding_box_max=bb_tgt_max,
            _is_pruned=prune_empty_leaves)
        return actx.freeze(tree), evt
