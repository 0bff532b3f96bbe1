"""Communicators for the distributed setup: a thin mpi4py-like layer over
``torch.distributed`` (NCCL over NVLink on the B200 box, gloo on CPU for tests) and a
single-process stand-in.  The reference uses mpi4py (``boxtree/distributed/__init__.py``);
only the calls the tree/traversal setup needs are provided."""
from __future__ import annotations

import numpy as np
import torch


class SingleProcessComm:
    """One rank; collectives are identities."""

    def Get_rank(self):  # noqa: N802  (mpi4py spelling)
        return 0

    def Get_size(self):  # noqa: N802
        return 1

    def bcast_array(self, arr, root=0):
        return arr

    def scatter_rows(self, rows, root=0):
        return np.asarray(rows[0])

    def allgather_tensor(self, t):
        return t.unsqueeze(0)

    def gather_objects(self, obj, root=0):
        return [obj]

    def barrier(self):
        pass


class TorchDistComm:
    """mpi4py-flavoured subset over an initialised ``torch.distributed`` process group.

    With the ``nccl`` backend tensors travel device to device (NVLink / NVSwitch); with
    ``gloo`` they are staged through host memory."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist = dist
        self.group = group
        self.backend = dist.get_backend(group)
        if device is None:
            device = (torch.device("cuda", torch.cuda.current_device())
                      if self.backend == "nccl" else torch.device("cpu"))
        self.device = torch.device(device)

    def Get_rank(self):  # noqa: N802
        return self.dist.get_rank(self.group)

    def Get_size(self):  # noqa: N802
        return self.dist.get_world_size(self.group)

    def _stage(self, t):
        return t.to(self.device) if t.device != self.device else t

    def bcast_array(self, arr, root=0):
        """Broadcast a numpy array (shape/dtype known on the root only)."""
        meta = [None]
        if self.Get_rank() == root:
            meta = [(arr.shape, arr.dtype.str)]
        self.dist.broadcast_object_list(meta, src=root, group=self.group)
        shape, dtype = meta[0]
        if self.Get_rank() == root:
            t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1))
        else:
            t = torch.empty(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8)
        t = self._stage(t)
        self.dist.broadcast(t, src=root, group=self.group)
        return t.cpu().numpy().view(np.dtype(dtype)).reshape(shape)

    def scatter_rows(self, rows, root=0):
        """Row *r* of the root's int32 2-D array goes to rank *r* (MPI Scatter,
        ``distributed/partition.py:118``)."""
        size = self.Get_size()
        width = [None]
        if self.Get_rank() == root:
            rows = np.ascontiguousarray(rows, dtype=np.int32)
            assert rows.shape[0] == size
            width = [int(rows.shape[1])]
        self.dist.broadcast_object_list(width, src=root, group=self.group)
        out = self._stage(torch.empty(width[0], dtype=torch.int32))
        if self.Get_rank() == root:
            parts = [self._stage(torch.from_numpy(rows[r].copy())) for r in range(size)]
            self.dist.scatter(out, parts, src=root, group=self.group)
        else:
            self.dist.scatter(out, None, src=root, group=self.group)
        return out.cpu().numpy()

    def allgather_tensor(self, t):
        """Stack every rank's tensor: result[r] = rank r's *t* (replaces the Gather to the
        root + bcast of ``distributed/local_tree.py:376-406``)."""
        src_device = t.device
        t = self._stage(t.contiguous())
        out = torch.empty((self.Get_size(),) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(-1), t.view(-1), group=self.group)
        return out.to(src_device)

    def gather_objects(self, obj, root=0):
        out = [None] * self.Get_size() if self.Get_rank() == root else None
        self.dist.gather_object(obj, out, dst=root, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)
