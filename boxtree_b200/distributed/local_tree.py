"""Per-rank local tree: ``boxtree/distributed/local_tree.py`` on B200."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any

import numpy as np
import torch

from .. import _cabi
from .._cabi import check, dptr
from ..array_context import make_obj_array
from ..tree import Tree
from .partition import _dev, get_box_masks


@dataclass(frozen=True)
class LocalTree(Tree):
    """``local_tree.py:287-313``: a :class:`Tree` whose particles are the rank's local ones
    (box geometry arrays are the global tree's) plus the distributed bookkeeping."""
    box_to_user_rank_starts: Any = None
    box_to_user_rank_lists: Any = None
    responsible_boxes_list: Any = None
    responsible_boxes_mask: Any = None
    ancestor_mask: Any = None


def _local_particles_and_lists(actx, lib, dimensions, nboxes, nparticles, coord_dtype, have_extent,
                               box_mask, particles, radii, starts, counts_nonchild, counts_cumul):
    """``construct_local_particles_and_lists``, ``local_tree.py:198-284``."""
    sh = actx.stream_handle
    dcode = _cabi.dtype_code(coord_dtype)
    starts, counts_nonchild, counts_cumul = (_dev(actx, a) for a in
                                             (starts, counts_nonchild, counts_cumul))
    particle_mask = actx.zeros(max(nparticles, 1), np.int32)
    check(lib.bt_dist_particle_mask(nboxes, dptr(box_mask), dptr(starts), dptr(counts_nonchild),
                                    dptr(particle_mask), sh), "bt_dist_particle_mask")
    g2l = actx.empty(nparticles + 1, np.int32)
    check(lib.bt_dist_mask_scan(nparticles, dptr(particle_mask), dptr(g2l), sh), "bt_dist_mask_scan")
    nlocal = int(g2l[-1].item())
    parts = [_dev(actx, p) for p in particles]
    local = [actx.empty(nlocal, coord_dtype) for _ in range(dimensions)]
    local_radii = actx.empty(nlocal, coord_dtype) if have_extent else None
    idx = actx.empty(nlocal, np.int64)
    check(lib.bt_dist_fetch_local_particles(
        dcode, dimensions, nparticles, dptr(particle_mask), dptr(g2l), _cabi.ptr_array(parts),
        dptr(_dev(actx, radii)) if have_extent else None, _cabi.ptr_array(local),
        dptr(local_radii), dptr(idx), sh), "bt_dist_fetch_local_particles")
    lstarts = actx.empty(nboxes, np.int32)
    lnonchild = actx.empty(nboxes, np.int32)
    lcumul = actx.empty(nboxes, np.int32)
    check(lib.bt_dist_local_lists(nboxes, dptr(box_mask), dptr(g2l), dptr(starts),
                                  dptr(counts_nonchild), dptr(counts_cumul), dptr(lstarts),
                                  dptr(lnonchild), dptr(lcumul), sh), "bt_dist_local_lists")
    return make_obj_array(local), local_radii, lstarts, lnonchild, lcumul, idx


def box_to_user_rank(actx, multipole_masks_all_ranks, bitsel=0xff):
    """MaskCompressorKernel on the gathered masks (``local_tree.py:376-406``,
    ``tools.py:647-740``): CSR box -> ranks that use the box's multipole expansion.  The mask
    entries are tested with ``& bitsel``."""
    lib = _cabi.load()
    masks = multipole_masks_all_ranks.contiguous()
    nranks, nboxes = int(masks.shape[0]), int(masks.shape[1])
    sh = actx.stream_handle
    starts = actx.empty(nboxes + 1, np.int32)
    total = actx.zeros(1, np.int64)
    check(lib.bt_dist_box_to_user_rank_bits(0, nboxes, nranks, bitsel, dptr(masks), dptr(starts),
                                            None, dptr(total), sh), "bt_dist_box_to_user_rank")
    lists = actx.empty(int(total.item()), np.int32)
    check(lib.bt_dist_box_to_user_rank_bits(1, nboxes, nranks, bitsel, dptr(masks), dptr(starts),
                                            dptr(lists), dptr(total), sh),
          "bt_dist_box_to_user_rank")
    return starts, lists


def generate_local_tree(actx, global_traversal, responsible_boxes_list, comm,
                        multipole_masks_all_ranks=None, box_masks=None):
    """``local_tree.py:316-495``.  Collective on *comm* (an all-gather of the int8 multipole
    masks replaces the reference's Gather to the root + bcast) unless
    *multipole_masks_all_ranks* ``[nranks, nboxes]`` is supplied.  *box_masks* may carry
    precomputed masks (sharded setup); then only ``global_traversal.tree`` is used.

    :returns: ``(local_tree, src_idx, tgt_idx)``; the index arrays (int64, device) give the
        position of every local source/target in the global tree's particle order."""
    lib = _cabi.load()
    gt = global_traversal.tree
    nb = int(gt.nboxes)
    dims = int(gt.dimensions)
    with torch.cuda.stream(actx.stream), torch.cuda.device(actx.device):
        masks = box_masks if box_masks is not None else \
            get_box_masks(actx, global_traversal, responsible_boxes_list)
        src = _local_particles_and_lists(
            actx, lib, dims, nb, int(gt.nsources), gt.coord_dtype, gt.sources_have_extent,
            masks.point_src_boxes, gt.sources, gt.source_radii, gt.box_source_starts,
            gt.box_source_counts_nonchild, gt.box_source_counts_cumul)
        tgt = _local_particles_and_lists(
            actx, lib, dims, nb, int(gt.ntargets), gt.coord_dtype, gt.targets_have_extent,
            masks.responsible_boxes, gt.targets, gt.target_radii, gt.box_target_starts,
            gt.box_target_counts_nonchild, gt.box_target_counts_cumul)
        if multipole_masks_all_ranks is None:
            multipole_masks_all_ranks = comm.allgather_tensor(masks.multipole_src_boxes)
        local_tree = assemble_local_tree(actx, gt, src, tgt, masks, multipole_masks_all_ranks,
                                         responsible_boxes_list)
    return actx.freeze(local_tree), src[5], tgt[5]


def assemble_local_tree(actx, gt, src, tgt, masks, multipole_masks_all_ranks,
                        responsible_boxes_list, bitsel=0xff):
    """The :class:`LocalTree` record of ``local_tree.py:430-495`` from the rank's local
    particles *src* / *tgt* (tuples as returned by ``_local_particles_and_lists``), its box
    masks and every rank's multipole mask."""
    lib = _cabi.load()
    nb = int(gt.nboxes)
    b2u_starts, b2u_lists = box_to_user_rank(actx, _dev(actx, multipole_masks_all_ranks), bitsel)
    local_flags = _dev(actx, gt.box_flags).clone()
    check(lib.bt_dist_modify_target_flags(nb, dptr(tgt[3]), dptr(tgt[4]), dptr(local_flags),
                                          actx.stream_handle), "bt_dist_modify_target_flags")
    base = {f.name: getattr(gt, f.name) for f in fields(Tree)}
    base.update(
        sources=src[0], targets=tgt[0],
        source_radii=src[1] if gt.sources_have_extent else None,
        target_radii=tgt[1] if gt.targets_have_extent else None,
        box_source_starts=src[2], box_source_counts_nonchild=src[3],
        box_source_counts_cumul=src[4], box_target_starts=tgt[2],
        box_target_counts_nonchild=tgt[3], box_target_counts_cumul=tgt[4],
        box_flags=local_flags, user_source_ids=None, sorted_target_ids=None)
    resp = responsible_boxes_list
    if not isinstance(resp, torch.Tensor):
        resp = actx.from_numpy(np.asarray(resp, np.int32))
    return LocalTree(
        **base, box_to_user_rank_starts=b2u_starts, box_to_user_rank_lists=b2u_lists,
        responsible_boxes_list=resp, responsible_boxes_mask=masks.responsible_boxes,
        ancestor_mask=masks.ancestor_boxes)
