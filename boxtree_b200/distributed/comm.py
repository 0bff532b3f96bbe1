"""Communicators for the distributed setup: a thin mpi4py-like layer over
``torch.distributed`` (NCCL over NVLink on the B200 box, gloo on CPU for tests) and a
single-process stand-in.  The reference uses mpi4py (``boxtree/distributed/__init__.py``);
only the calls the tree/traversal setup needs are provided."""
from __future__ import annotations

import numpy as np
import torch


class _Done:
    """Handle of a collective that has already been carried out."""

    def wait(self):
        pass


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

    def allreduce_(self, t, op="sum"):
        return t

    def allreduce_async_(self, t, op="sum"):
        return _Done()

    def all_to_all_bytes(self, send, send_splits, recv_splits):
        return send

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

    def allreduce_(self, t, op="sum"):
        """In-place all-reduce (sum / min / max) of a contiguous tensor."""
        ops = {"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN,
               "max": self.dist.ReduceOp.MAX}
        assert t.is_contiguous()
        if t.device == self.device:
            self.dist.all_reduce(t, op=ops[op], group=self.group)
        else:
            staged = t.to(self.device)
            self.dist.all_reduce(staged, op=ops[op], group=self.group)
            t.copy_(staged)
        return t

    def allreduce_async_(self, t, op="sum"):
        """In-place all-reduce that is only ENQUEUED: it runs on the backend's own stream after
        the work already queued on the current stream; kernels launched afterwards on the
        current stream overlap it.  ``handle.wait()`` makes the current stream wait for the
        result (no host synchronisation with NCCL)."""
        ops = {"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN,
               "max": self.dist.ReduceOp.MAX}
        assert t.is_contiguous()
        if t.device != self.device:
            self.allreduce_(t, op)
            return _Done()
        return self.dist.all_reduce(t, op=ops[op], group=self.group, async_op=True)

    def all_to_all_bytes(self, send, send_splits, recv_splits):
        """Variable all-to-all of a uint8 buffer: ``send_splits[d]`` bytes go to rank *d*,
        ``recv_splits[s]`` bytes arrive from rank *s* (NCCL: grouped send/recv over NVLink)."""
        src_device = send.device
        send = self._stage(send.contiguous())
        recv = torch.empty(int(sum(recv_splits)), dtype=torch.uint8, device=send.device)
        self.dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in recv_splits],
                                    input_split_sizes=[int(x) for x in send_splits],
                                    group=self.group)
        return recv.to(src_device)

    def barrier(self):
        self.dist.barrier(group=self.group)


class ThreadGroup:
    """Shared state of :class:`ThreadComm` ranks (one Python thread per rank)."""

    def __init__(self, size):
        import threading
        self.size = size
        self.slots = [None] * size
        self.barrier = threading.Barrier(size)


class ThreadComm:
    """In-process communicator: every rank is a thread of this process (all ranks may share one
    GPU).  Same interface as :class:`TorchDistComm`; used to exercise the collective code paths
    of the distributed build where only one device is available."""

    def __init__(self, group: ThreadGroup, rank: int):
        self.g, self.rank = group, rank

    def Get_rank(self):  # noqa: N802
        return self.rank

    def Get_size(self):  # noqa: N802
        return self.g.size

    def _exchange(self, obj):
        if isinstance(obj, torch.Tensor) and obj.is_cuda:
            torch.cuda.synchronize(obj.device)
        self.g.slots[self.rank] = obj
        self.g.barrier.wait()
        out = list(self.g.slots)
        self.g.barrier.wait()
        return out

    def bcast_array(self, arr, root=0):
        return self._exchange(arr)[root]

    def scatter_rows(self, rows, root=0):
        return np.asarray(self._exchange(rows)[root][self.rank])

    def allgather_tensor(self, t):
        return torch.stack([x.to(t.device) for x in self._exchange(t.contiguous().clone())])

    def gather_objects(self, obj, root=0):
        out = self._exchange(obj)
        return out if self.rank == root else None

    def allreduce_(self, t, op="sum"):
        allv = torch.stack([x.to(t.device) for x in self._exchange(t.clone())])
        if op == "sum":
            t.copy_(allv.sum(dim=0, dtype=t.dtype))
        elif op == "min":
            t.copy_(allv.amin(dim=0))
        else:
            t.copy_(allv.amax(dim=0))
        if t.is_cuda:
            torch.cuda.synchronize(t.device)
        self.g.barrier.wait()
        return t

    def allreduce_async_(self, t, op="sum"):
        self.allreduce_(t, op)
        return _Done()

    def all_to_all_bytes(self, send, send_splits, recv_splits):
        allv = self._exchange((send.clone(), [int(x) for x in send_splits]))
        parts = []
        for s, (buf, splits) in enumerate(allv):
            off = sum(splits[:self.rank])
            assert splits[self.rank] == int(recv_splits[s])
            parts.append(buf[off:off + splits[self.rank]].to(send.device))
        out = torch.cat(parts) if parts else send[:0]
        if out.is_cuda:
            torch.cuda.synchronize(out.device)
        self.g.barrier.wait()
        return out

    def barrier(self):
        self.g.barrier.wait()
