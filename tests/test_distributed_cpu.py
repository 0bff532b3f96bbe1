"""CPU tests of the distributed host logic: work-partition segments against the oracle's
restatement of the reference loop, and the torch.distributed communicator over gloo with
world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from boxtree_b200.distributed.partition import partition_segments
from oracle.distributed import get_box_ids_dfs_order, partition_work
from oracle.tree_build import build_tree
from tests.parity_util import normal_particles


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
@pytest.mark.parametrize("costs", ["ones", "counts", "random", "spiky"])
def test_partition_segments_match_reference_loop(nranks, costs):
    tree = build_tree(normal_particles(5000, 3, np.float64), max_particles_in_box=20)
    nb = tree.nboxes
    rng = np.random.default_rng(3)
    cost = {"ones": np.ones(nb), "counts": 1.0 + tree.box_source_counts_nonchild[:nb],
            "random": rng.random(nb) * 10, "spiky": np.where(rng.random(nb) < 0.01, 1e3, 0.0)}[costs]
    cost = np.asarray(cost, np.float64)
    want_lists, want_segments = partition_work(cost, tree, nranks)
    dfs = get_box_ids_dfs_order(tree)
    got = partition_segments(cost[dfs], nranks)
    assert np.array_equal(got, want_segments)
    assert sorted(np.concatenate(want_lists).tolist()) == list(range(nb))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boxtree_b200.distributed.comm import TorchDistComm
    comm = TorchDistComm()
    assert comm.Get_rank() == rank and comm.Get_size() == world
    rows = np.array([[0, 5], [5, 9]], np.int32) if rank == 0 else None
    mine = comm.scatter_rows(rows, root=0)
    mask = torch.full((7,), rank + 1, dtype=torch.int8)
    gathered = comm.allgather_tensor(mask)
    arr = comm.bcast_array(np.arange(6, dtype=np.float64).reshape(2, 3) if rank == 0 else None, 0)
    objs = comm.gather_objects({"rank": rank}, root=0)
    comm.barrier()
    out[rank] = (mine.tolist(), gathered.tolist(), arr.tolist(), objs)
    dist.destroy_process_group()


class _HostActx:
    """numpy <-> CPU torch stand-in so broadcast_tree can be exercised without a GPU."""

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def to_numpy(self, t):
        return t.numpy()


def _tree_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boxtree_b200 import Tree
    from boxtree_b200.array_context import make_obj_array
    from boxtree_b200.distributed import TorchDistComm, broadcast_tree
    import dataclasses
    actx = _HostActx()
    tree = None
    if rank == 0:
        ot = build_tree(normal_particles(500, 2, np.float64), max_particles_in_box=20)
        vals = {}
        for f in dataclasses.fields(Tree):
            v = getattr(ot, f.name)
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(np.ascontiguousarray(v))
            elif isinstance(v, list):
                v = make_obj_array([torch.from_numpy(np.ascontiguousarray(x)) for x in v])
            vals[f.name] = v
        tree = Tree(**vals)
    got = broadcast_tree(actx, tree, TorchDistComm(), root=0)
    out[rank] = (got.nboxes, got.nlevels, got.box_child_ids.sum().item(),
                 float(got.sources[1].sum()), got.targets is got.sources,
                 float(got.root_extent), got.bounding_box[0].tolist(), got._is_pruned)
    dist.destroy_process_group()


def test_broadcast_tree_gloo_world_size_2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_tree_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == out[1]
    assert out[1][4] is True and out[1][0] > 1


def test_torch_dist_comm_gloo_world_size_2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] == [0, 5] and out[1][0] == [5, 9]
    for r in (0, 1):
        assert out[r][1] == [[1] * 7, [2] * 7]
        assert out[r][2] == [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]]
    assert out[0][3] == [{"rank": 0}, {"rank": 1}] and out[1][3] is None


def _collective_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boxtree_b200.distributed.comm import TorchDistComm
    comm = TorchDistComm()
    # the collectives of the distributed tree build: per-level count sums, min/max of the box
    # extents, the variable all-to-all of particle records
    counts = torch.arange(6, dtype=torch.int32) * (rank + 1)
    comm.allreduce_(counts, "sum")
    lo = torch.tensor([1.0 + rank, -3.0 * rank], dtype=torch.float64)
    hi = lo.clone()
    # enqueued, both in flight, waited for in order: how the deferred build reduces the extents
    waits = [comm.allreduce_async_(lo, "min"), comm.allreduce_async_(hi, "max")]
    for w in waits:
        w.wait()
    # rank r sends (d + 1 + r) bytes of value 10 * r + d to rank d
    send_splits = [d + 1 + rank for d in range(world)]
    send = torch.cat([torch.full((n,), 10 * rank + d, dtype=torch.uint8)
                      for d, n in enumerate(send_splits)])
    recv_splits = [rank + 1 + s for s in range(world)]
    recv = comm.all_to_all_bytes(send, send_splits, recv_splits)
    out[rank] = (counts.tolist(), lo.tolist(), hi.tolist(), recv.tolist())
    dist.destroy_process_group()


def _check_collectives(out, world):
    for r in range(world):
        assert out[r][0] == [i * sum(range(1, world + 1)) for i in range(6)]
        assert out[r][1] == [1.0, -3.0 * (world - 1)] and out[r][2] == [float(world), 0.0]
        want = []
        for s in range(world):
            want += [10 * s + r] * (r + 1 + s)
        assert out[r][3] == want


def test_distributed_build_collectives_gloo_world_size_2():
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_collective_worker, args=(2, port, out), nprocs=2, join=True)
    _check_collectives(out, 2)


def test_thread_comm_collectives():
    """The in-process communicator of the GPU tests implements the same collectives."""
    import threading

    from boxtree_b200.distributed.comm import ThreadComm, ThreadGroup
    world = 3
    group = ThreadGroup(world)
    out = {}

    def work(rank):
        comm = ThreadComm(group, rank)
        counts = torch.arange(6, dtype=torch.int32) * (rank + 1)
        comm.allreduce_(counts, "sum")
        lo = torch.tensor([1.0 + rank, -3.0 * rank], dtype=torch.float64)
        hi = lo.clone()
        for w in [comm.allreduce_async_(lo, "min"), comm.allreduce_async_(hi, "max")]:
            w.wait()
        send_splits = [d + 1 + rank for d in range(world)]
        send = torch.cat([torch.full((n,), 10 * rank + d, dtype=torch.uint8)
                          for d, n in enumerate(send_splits)])
        recv = comm.all_to_all_bytes(send, send_splits, [rank + 1 + s for s in range(world)])
        gathered = comm.allgather_tensor(torch.tensor([rank], dtype=torch.int64))
        assert gathered.view(-1).tolist() == list(range(world))
        out[rank] = (counts.tolist(), lo.tolist(), hi.tolist(), recv.tolist())

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    _check_collectives(out, world)


@pytest.mark.parametrize("nranks", [2, 5])
def test_box_ranges_are_sums_of_rank_local_ranges(nranks):
    """The principle of the distributed tree build, on the host: with the particles dealt to
    ranks arbitrarily and every rank's share sorted by the tree order, a box's range in the
    global order has start = sum of the ranks' lower bounds and count = sum of the ranks' counts
    (what the build all-reduces per level), and a particle's global position is box start + own
    particles of the box on lower ranks + index in the rank's own range (what the receiver of
    the all-to-all reconstructs)."""
    tree = build_tree(normal_particles(4000, 3, np.float64), max_particles_in_box=15)
    n, nb = tree.nsources, tree.nboxes
    user_ids = np.asarray(tree.user_source_ids)              # tree order -> user id
    owner = np.arange(n) * nranks // n                        # user id -> rank (contiguous slices)
    tree_pos_of_user = np.empty(n, np.int64)
    tree_pos_of_user[user_ids] = np.arange(n)
    starts = np.asarray(tree.box_source_starts)[:nb].astype(np.int64)
    cumul = np.asarray(tree.box_source_counts_cumul)[:nb].astype(np.int64)
    # every rank's particles in tree order (global tree positions, ascending)
    local = [np.sort(tree_pos_of_user[np.nonzero(owner == r)[0]]) for r in range(nranks)]
    lower = np.stack([np.searchsorted(p, starts) for p in local])
    upper = np.stack([np.searchsorted(p, starts + cumul) for p in local])
    assert np.array_equal(lower.sum(0), starts)
    assert np.array_equal((upper - lower).sum(0), cumul)
    # global position = box start + counts of lower ranks in the box's own range + local index
    own = np.asarray(tree.box_source_counts_nonchild)[:nb].astype(np.int64)
    own_upper = np.stack([np.searchsorted(p, starts + own) for p in local])
    own_cnt = own_upper - lower
    excl = np.cumsum(own_cnt, axis=0) - own_cnt
    for b in np.nonzero(own)[0][:300]:
        for r in range(nranks):
            for k in range(int(own_cnt[r, b])):
                # rank r's k-th own particle of box b really sits at that global position
                assert local[r][lower[r, b] + k] == starts[b] + excl[r, b] + k


def test_pending_work_runs_once():
    """``build_distributed_tree(defer_extents=True)`` hands its unfinished work over as an object
    whose ``finish()`` may be called from several places (the traversal's hook, the setup's
    catch-all): it must run exactly once."""
    from boxtree_b200.tree_build import _Pending
    ran = []
    p = _Pending(lambda: ran.append(1))
    assert not p.done
    p.finish()
    p.finish()
    assert ran == [1] and p.done


def test_single_process_comm_async_handle():
    from boxtree_b200.distributed.comm import SingleProcessComm
    t = torch.arange(4, dtype=torch.float64)
    h = SingleProcessComm().allreduce_async_(t, "max")
    h.wait()
    assert t.tolist() == [0.0, 1.0, 2.0, 3.0]
