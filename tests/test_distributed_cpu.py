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
