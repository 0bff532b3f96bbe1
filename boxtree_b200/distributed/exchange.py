"""Particle redistribution of the distributed build: ONE all-to-all per particle kind.

Replaces the root-side ``fetch_local_particles`` + scatter of the reference
(``boxtree/distributed/local_tree.py:124-151, 198-284, 408-495``).  After
:func:`boxtree_b200.distributed.tree_build.build_distributed_tree` every rank holds the
global box arrays and its own input particles in tree order.  Rank *d*'s local tree needs
the sources of its ``point_src_boxes`` and the targets of its ``responsible_boxes``
(``local_tree.py:430-470``): the owners pack one record per needed particle and destination
(``bt_dist_pack_records``), the records travel in one variable all-to-all (NCCL grouped
send/recv over NVLink), and the receiver scatters them to their place in the global tree
order restricted to its boxes (``bt_dist_unpack_records``).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi
from .._cabi import check, dptr
from ..array_context import make_obj_array


def _record_bytes(coord_dtype, dims, have_radii):
    return np.dtype(coord_dtype).itemsize * (dims + (1 if have_radii else 0)) + 8


def preorder(actx, tree):
    """``(rank, boxes, subtree_size)`` of the boxes' pre-order with children in Morton order:
    the order of the global tree's particle arrays (own particles, then the children's)."""
    lib = _cabi.load()
    nb = int(tree.nboxes)
    sh = actx.stream_handle
    size = actx.empty(max(nb, 1), np.int32)
    rank = actx.empty(max(nb, 1), np.int32)
    boxes = actx.empty(max(nb, 1), np.int32)
    check(lib.bt_trav_dfs_rank(int(tree.dimensions), nb, int(tree.aligned_nboxes),
                               int(tree.nlevels), dptr(tree.level_start_box_nrs.to(torch.int32)),
                               dptr(tree.box_child_ids), dptr(size), dptr(rank), sh),
          "bt_trav_dfs_rank")
    check(lib.bt_reverse_index(nb, dptr(rank), dptr(boxes), sh), "bt_reverse_index")
    return rank, boxes, size


def _local_ranges(actx, lib, nb, mask, own_counts, pre):
    """``(local_starts, local_counts_nonchild, local_counts_cumul)`` of the rank's local particle
    arrays for the boxes of *mask* (``local_tree.py:249-284``), from the global own counts."""
    pre_rank, pre_boxes, subtree = pre
    prefix = actx.empty(nb + 1, np.int32)
    lstarts = actx.empty(nb, np.int32)
    lnonchild = actx.empty(nb, np.int32)
    lcumul = actx.empty(nb, np.int32)
    check(lib.bt_dist_local_ranges(nb, dptr(mask), dptr(own_counts), dptr(pre_rank), dptr(pre_boxes),
                                   dptr(subtree), dptr(prefix), dptr(lstarts), dptr(lnonchild),
                                   dptr(lcumul), actx.stream_handle), "bt_dist_local_ranges")
    return lstarts, lnonchild, lcumul


class _Exchange:
    """One particle kind's exchange in three steps -- count, pack + all-to-all, unpack -- so that
    several kinds can share the one small all-gather + readback of the chunk sizes."""

    def __init__(self, actx, comm, dtree, masks_all_ranks, bitsel, my_mask, kind, pre, ranges):
        self.actx, self.comm, self.dtree = actx, comm, dtree
        self.lib = _cabi.load()
        self.sh = actx.stream_handle
        self.nb = int(dtree.nboxes)
        self.dims = int(dtree.dimensions)
        self.nranks, self.rank = comm.Get_size(), comm.Get_rank()
        self.dcode = _cabi.dtype_code(dtree.coord_dtype)
        src = kind == "source"
        self.parts = list(dtree.sources if src else dtree.targets)
        self.have_radii = dtree.sources_have_extent if src else dtree.targets_have_extent
        self.radii = (dtree.source_radii if src else dtree.target_radii) if self.have_radii else None
        self.lstart = dtree.local_box_source_starts if src else dtree.local_box_target_starts
        self.lown = dtree.local_box_source_counts_nonchild if src else \
            dtree.local_box_target_counts_nonchild
        self.gstart = dtree.box_source_starts if src else dtree.box_target_starts
        self.gown = dtree.box_source_counts_nonchild if src else dtree.box_target_counts_nonchild
        self.recbytes = _record_bytes(dtree.coord_dtype, self.dims, self.have_radii)
        self.my_mask, self.pre, self.ranges = my_mask, pre, ranges
        self.masks_all_ranks, self.bitsel = masks_all_ranks, bitsel

    def count(self):
        """Records per destination (device): ``[nranks + 2]`` int64 = chunk offsets, number of
        boxes of my mask."""
        actx, lib, sh, nb, nranks = self.actx, self.lib, self.sh, self.nb, self.nranks
        self.dest_bits = actx.empty(max(nb, 1), np.int32)
        check(lib.bt_dist_mask_bits(nb, nranks, self.bitsel, dptr(self.masks_all_ranks),
                                    dptr(self.dest_bits), sh), "bt_dist_mask_bits")
        n = self.n = int(self.parts[0].shape[0])
        ntiles = max(int(lib.bt_dist_pack_ntiles(n)), 1)
        self.pbox = actx.empty(max(n, 1), np.int32)
        tile_counts = actx.empty(nranks * ntiles, np.int32)
        self.tile_offs = actx.empty(nranks * ntiles, np.int64)
        dest_offsets = actx.empty(nranks + 2, np.int64)
        check(lib.bt_dist_pack_count(nranks, nb, n, dptr(self.dest_bits), dptr(self.lstart),
                                     dptr(self.lown), dptr(self.pbox), dptr(tile_counts),
                                     dptr(self.tile_offs), dptr(dest_offsets), sh),
              "bt_dist_pack_count")
        self.compact = actx.empty(max(nb, 1), np.int32)
        nmasked_dev = actx.empty(1, np.int32)
        check(lib.bt_dist_compact_index(nb, dptr(self.my_mask), dptr(self.compact),
                                        dptr(nmasked_dev), sh), "bt_dist_compact_index")
        dest_offsets[nranks + 1:].copy_(nmasked_dev)
        return dest_offsets

    def send(self, gathered):
        """*gathered* ``[sender, nranks + 2]`` (host): pack and post the all-to-all."""
        import ctypes as C
        actx, lib, sh, nranks, rank = self.actx, self.lib, self.sh, self.nranks, self.rank
        counts = np.diff(gathered[:, :nranks + 1], axis=1)                  # [sender, dest]
        self.nmasked = int(gathered[rank, nranks + 1])
        send_counts, self.recv_counts = counts[rank], counts[:, rank]
        nsend = int(send_counts.sum())
        sendbuf = actx.empty(max(nsend, 1) * self.recbytes, np.uint8)
        check(lib.bt_dist_pack_records(self.dcode, nranks, self.dims, self.n, dptr(self.pbox),
                                       dptr(self.dest_bits), dptr(self.tile_offs),
                                       _cabi.ptr_array(self.parts), dptr(self.radii),
                                       dptr(self.lstart), dptr(sendbuf), sh), "bt_dist_pack_records")
        self.recvbuf = self.comm.all_to_all_bytes(
            sendbuf[:nsend * self.recbytes], [int(c) * self.recbytes for c in send_counts],
            [int(c) * self.recbytes for c in self.recv_counts])
        self.C = C

    def unpack(self):
        """The rank's local arrays (global tree order restricted to its boxes)."""
        actx, lib, sh, nb, nranks, dtree = self.actx, self.lib, self.sh, self.nb, self.nranks, self.dtree
        lstarts, lnonchild, lcumul = self.ranges if self.ranges is not None else \
            _local_ranges(actx, lib, nb, self.my_mask, self.gown, self.pre)
        nrecv = int(self.recv_counts.sum())
        count_tmp = actx.empty(max(nranks * self.nmasked, 1), np.int32)
        local = [actx.empty(nrecv, dtree.coord_dtype) for _ in range(self.dims)]
        local_radii = actx.empty(nrecv, dtree.coord_dtype) if self.have_radii else None
        idx = actx.empty(nrecv, np.int64)
        chunk = (self.C.c_int64 * (nranks + 1))(
            *np.concatenate([[0], np.cumsum(self.recv_counts)]).tolist())
        check(lib.bt_dist_unpack_records(self.dcode, nranks, self.dims, nrecv, int(self.have_radii),
                                         dptr(self.recvbuf), chunk, dptr(self.compact), self.nmasked,
                                         dptr(count_tmp), dptr(lstarts), dptr(self.gstart),
                                         _cabi.ptr_array(local), dptr(local_radii), dptr(idx), sh),
              "bt_dist_unpack_records")
        return make_obj_array(local), local_radii, lstarts, lnonchild, lcumul, idx


def exchange_particles(actx, comm, dtree, masks_all_ranks, bitsel, my_mask, kind, pre,
                       ranges=None):
    """Collective.  *masks_all_ranks* ``[nranks, nboxes]`` int8 bit fields: rank *d* needs the
    *kind* (``"source"`` / ``"target"``) particles of the boxes with
    ``masks_all_ranks[d][b] & bitsel``; *my_mask* (int8 0/1) is this rank's row.
    *pre* = :func:`preorder` of the tree; *ranges*: the result of ``_local_ranges`` for
    *my_mask* when the caller already has it.

    :returns: ``(particles, radii, local_starts, local_counts_nonchild, local_counts_cumul,
        idx)`` like ``construct_local_particles_and_lists`` (``local_tree.py:198-284``); *idx*
        (int64) is every local particle's position in the global tree order."""
    return exchange_particle_kinds(
        actx, comm, dtree, masks_all_ranks, [(bitsel, my_mask, kind, ranges)], pre)[0]


def exchange_particle_kinds(actx, comm, dtree, masks_all_ranks, kinds, pre):
    """Several exchanges (*kinds*: ``(bitsel, my_mask, kind, ranges)``) sharing ONE all-gather +
    readback of the chunk sizes; the all-to-alls are posted back to back."""
    xs = [_Exchange(actx, comm, dtree, masks_all_ranks, bitsel, my_mask, kind, pre, ranges)
          for bitsel, my_mask, kind, ranges in kinds]
    offs = torch.stack([x.count() for x in xs])                          # [kinds, nranks + 2]
    gathered = actx.read_back(comm.allgather_tensor(offs))               # [sender, kinds, ...]
    for k, x in enumerate(xs):
        x.send(gathered[:, k])
    return [x.unpack() for x in xs]
