#!/usr/bin/env python
"""Benchmark of the hot path: TreeBuilder.__call__ + FMMTraversalBuilder.__call__.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one tree build + one full traversal build over one batch of
synthetic particles.  Prints ONE JSON line (see the keys below).  Workloads are
the configurations of BASELINE.json; the default is config3 (3-D, 1e7 points:
5e6 sources + 5e6 targets with target radii, level-restricted), the
configuration the metric is quoted on.

* ``value``: Mpoints/s with the inputs already resident in HBM, CUDA-event timed
  over exactly K steps, max over ranks.
* ``e2e``: the same through the public API starting from pinned HOST buffers: every step
  copies its own coordinates/radii host->device and reads its result summary device->host
  inside the timed region.  ``e2e.value`` is the double-buffered loop (the H2D copy of step
  k+1 runs on a copy stream into a second, preallocated set of device buffers while step k
  builds); ``e2e.serial`` is upload, build, read back one after the other.
* ``roofline``: the dominant kernel scope, timed live with CUDA events on the
  launching stream (library instrumentation ``bt_prof_*``), achieved algorithmic
  GB/s against the measured HBM peak of MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle (a restatement of the
  reference algorithm; the reference itself cannot run here) timed
  on the host cores on a bounded sample of the same recipe.

With N > 1 (torchrun) the default is ``--parallelism distributed``: ONE global problem of
N x the workload's points (weak scaling: rank r contributes the workload generated with seed
shift r), built by the distributed tree build (all-reduced bounding box and per-level box
counts over NCCL, particles stay on their ranks), partitioned by the reference's DFS-order
cost partition, particles sent to the local trees in one NCCL all-to-all, local traversal per
rank (``boxtree.distributed`` semantics).  The same run also times the STRONG-scaling arm (the
workload's N=1 problem cut into N slices; its global box arrays are checked against the
committed digests of the reference's own run) and reports it in ``distributed_strong``.
``--parallelism replicas`` (independent problems, no collective) and ``--parallelism sharded``
(round 1: all-gathered particles, replicated tree build) remain selectable.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

WORKLOADS = {
    # name: (recipe, n, dtype, description)
    "config2": ("uniform", 1_000_000, "f64",
                "3D 1e6 uniform fp64, sources are targets, adaptive, max 30"),
    "config3": ("config3", 10_000_000, "f64",
                "3D 1e7: 5e6 sources + 5e6 targets fp64, target radii, stick_out 0.25, linf, "
                "adaptive-level-restricted, max 30"),
    "config4": ("plummer", 100_000_000, "f32", "3D 1e8 Plummer fp32, adaptive, max 30"),
    "uniform1e7": ("uniform", 10_000_000, "f64",
                   "3D 1e7 uniform fp64, sources are targets, adaptive, max 30"),
    # per GPU; with --gpus 8 the global problem is BASELINE config 5 (5e8 uniform points)
    "config5": ("uniform", 62_500_000, "f64",
                "3D 6.25e7 uniform fp64 per GPU (5e8 over 8 GPUs), sources are targets, adaptive, "
                "max 30"),
}
# CPU arm (oracle port, all host cores): the in-run `cpu_baseline` leg does ONE step on the full
# workload when it has at most CPU_BASELINE_POINTS points (config3: all 1e7, 10-30 s); the
# `--impl reference` arm, which the driver runs for K + W steps, takes CPU_SAMPLE_POINTS per step
CPU_BASELINE_POINTS = 10_000_000
CPU_SAMPLE_POINTS = 4_000_000


def make_inputs(recipe, n, dtype, seed_shift=0):
    from tests.parity_util import plummer_particles
    dt = np.float32 if dtype == "f32" else np.float64
    kw = {"max_particles_in_box": 30}
    if recipe == "uniform":
        pts = np.random.default_rng(15 + seed_shift).random((3, n))
        src = [np.ascontiguousarray(pts[i]).astype(dt) for i in range(3)]
    elif recipe == "plummer":
        src = plummer_particles(n, dt, seed=15 + seed_shift)
    elif recipe == "config3":
        ns, nt = n // 2, n - n // 2
        s = np.random.default_rng(12 + seed_shift).random((3, ns))
        t = np.random.default_rng(19 + seed_shift).random((3, nt))
        radii = 2 ** np.random.default_rng(13 + seed_shift).uniform(-14, -4, nt)
        src = [np.ascontiguousarray(s[i]).astype(dt) for i in range(3)]
        kw.update(targets=[np.ascontiguousarray(t[i]).astype(dt) for i in range(3)],
                  target_radii=radii.astype(dt), stick_out_factor=0.25, extent_norm="linf",
                  kind="adaptive-level-restricted")
    else:
        raise ValueError(recipe)
    return src, kw


def host_threads():
    """Threads the CPU arm uses: all host cores (torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would silently serialise the OpenMP traversal kernels)."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)      # read when the oracle's C library is loaded
    return n


def time_oracle(recipe, n, dtype, steps, warmup):
    """The CPU restatement of the reference on the host cores (reported baseline)."""
    host_threads()
    from oracle.traversal import build_traversal
    from oracle.tree_build import build_tree
    src, kw = make_inputs(recipe, n, dtype)
    for _ in range(warmup):
        build_traversal(build_tree(src, **kw))
    t0 = time.perf_counter()
    for _ in range(steps):
        build_traversal(build_tree(src, **kw))
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n / dt / 1e6, dt


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = f"/tmp/bt_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(scope, tree, trav, n, dims, s, heavy_entries=0):
    """Compulsory bytes of one launch of the scope's kernel ('Kernels')."""
    nb = 2 ** dims
    B, aB = tree.nboxes, tree.aligned_nboxes
    tree_read = aB * (4 * nb + dims * s) + B * 6
    T = int(trav.target_boxes.shape[0])
    TP = int(trav.target_or_target_parent_boxes.shape[0])
    nl = tree.nlevels

    def csr(starts, lists):
        return 4 * (int(starts.shape[0]) + int(lists.shape[0]))

    coll = csr(trav.same_level_non_well_sep_boxes_starts, trav.same_level_non_well_sep_boxes_lists)
    l3_lists = sum(int(b.lists.shape[0]) for b in trav.from_sep_smaller_by_level)
    l3_close = 0 if trav.from_sep_close_smaller_lists is None else \
        int(trav.from_sep_close_smaller_lists.shape[0])
    l1 = csr(trav.neighbor_source_boxes_starts, trav.neighbor_source_boxes_lists)
    g_counts = 4 * (nl + 2) * (T + 1)                    # per-(slot, row) counts / offsets
    table = {
        # top-down colleagues: box geometry + child table in, staged rows + list-2 counts/masks out
        "trav_colleagues_count": tree_read + coll + 4 * B + 32 * B,
        "trav_colleagues_fill": 2 * coll,
        "trav_list2_count": 4 * TP + 4 * B + 4 * (TP + 1),
        "trav_list2_fill": coll + 32 * B + 4 * TP + aB * 4 * nb
        + csr(trav.from_sep_siblings_starts, trav.from_sep_siblings_lists),
        # fused list-1+3 walk (count pass also stages the entries the fill pass copies)
        "l13_walk_count": tree_read + coll + 4 * T + g_counts + 4 * (l3_lists + l3_close) + l1,
        "l13_walk_fill": tree_read + coll + 4 * T + g_counts + 4 * (l3_lists + l3_close) + l1,
        "l13_unstage": 2 * (4 * (l3_lists + l3_close) + l1) + g_counts,
        "trav_list4_count": tree_read + coll + 4 * TP + 4 * (TP + 1),
        "trav_list4_fill": tree_read + coll + 4 * TP + csr(trav.from_sep_bigger_starts,
                                                           trav.from_sep_bigger_lists),
        # one radix pass: read + write (8-byte key, 4-byte id)
        "rs_onesweep_pass": 2 * 12 * n,
        "rs_histogram": 8 * n,
        "bt_make_keys": n * (dims * s + 8),
        "bt_permute": n * (4 + 2 * dims * s),
        "bt_bounding_box": n * dims * s,
        "bt_box_extents": n * dims * s + 2 * aB * dims * s,
    }
    if heavy_entries:
        # position-map mode: one byte per appended box out, then map in + list entries out
        table["l13_heavy_expand"] = tree_read + heavy_entries
        table["l13_heavy_extract"] = 5 * heavy_entries
        table["l13_heavy_sort_pass"] = 2 * 8 * heavy_entries
        table["l13_heavy_scatter"] = 12 * heavy_entries
        table["l13_heavy_steps_fill"] = 8 * heavy_entries
    return table.get(scope)


def whole_path_bytes(tree, trav, nsources, ntargets, dims, s, sources_are_targets, have_radii):
    """B_tree + B_trav, compulsory bytes of one step, each byte once."""
    nb = 2 ** dims
    B, aB = int(tree.nboxes), int(tree.aligned_nboxes)
    n_in = nsources + ntargets
    k = 1 if sources_are_targets else 2
    b_tree = n_in * dims * s + (nsources + (0 if sources_are_targets else ntargets)) * dims * s
    if have_radii:
        b_tree += 2 * ntargets * s
    b_tree += 4 * nsources + 4 * (nsources if sources_are_targets else ntargets)
    b_tree += B * (k * 12 + 4 + 1 + 1) + aB * (4 * nb + dims * s + 2 * k * dims * s)
    b_trav = aB * (4 * nb + dims * s) + B * 6
    import dataclasses

    import torch

    def nbytes(v):
        if isinstance(v, torch.Tensor):
            return v.numel() * v.element_size()
        if isinstance(v, np.ndarray) and v.dtype == object:
            return sum(nbytes(x) for x in v)
        if dataclasses.is_dataclass(v):
            return sum(nbytes(getattr(v, f.name)) for f in dataclasses.fields(v))
        return 0
    for f in dataclasses.fields(trav):
        if f.name != "tree":
            b_trav += nbytes(getattr(trav, f.name))
    return b_tree, b_trav


def run_ours(args):
    import torch

    from boxtree_b200 import FMMTraversalBuilder, TorchArrayContext, TreeBuilder, _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")

    recipe, n, dtype, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    mode = args.parallelism if world > 1 else "single"
    dims = 3
    s_bytes = 4 if dtype == "f32" else 8

    actx = TorchArrayContext(device)
    tb = TreeBuilder(actx)
    tg = FMMTraversalBuilder(actx)
    lib = _cabi.load()
    comm = bd = None
    if world > 1:
        from boxtree_b200 import distributed as bd
        comm = bd.TorchDistComm()

    def to_dev(v):
        return actx.from_numpy(v)

    def dev_inputs(src, kw):
        return [to_dev(x) for x in src], {
            k: (to_dev(v) if isinstance(v, np.ndarray) else
                [to_dev(x) for x in v] if k == "targets" else v) for k, v in kw.items()}

    def my_slice(a):
        m = int(a.shape[0])
        return a[rank * m // world:(rank + 1) * m // world]

    def slice_inputs(src, kw):
        return [np.ascontiguousarray(my_slice(x)) for x in src], {
            k: (np.ascontiguousarray(my_slice(v)) if isinstance(v, np.ndarray) else
                [np.ascontiguousarray(my_slice(x)) for x in v] if k == "targets" else v)
            for k, v in kw.items()}

    # {{{ this rank's particles
    # single / replicas / distributed (weak): the workload generated with seed shift `rank`;
    # for "distributed" the global problem is the concatenation over ranks (N x n points).
    # distributed-strong / sharded: the r-th slice of the N=1 problem (n points in total).
    if mode in ("distributed-strong", "sharded"):
        src, kw = slice_inputs(*make_inputs(recipe, n, dtype, seed_shift=0))
        npoints_job = n
    else:
        src, kw = make_inputs(recipe, n, dtype, seed_shift=rank)
        npoints_job = world * n
    dsrc, dkw = dev_inputs(src, kw)
    # }}}

    def make_step(dsrc, dkw, mode):
        if mode in ("single", "replicas"):
            def step():
                tree, _ = tb(actx, dsrc, **dkw)
                if args.chunks != 1:
                    # large-list mode: the traversal in row pieces with int32 CSR arrays each
                    pieces = tg.build_in_chunks(actx, tree, nchunks=args.chunks or None)
                    nonlocal_state["nchunks"] = len(pieces)
                    return tree, pieces[-1][1]
                trav, _ = tg(actx, tree)
                return tree, trav
        elif mode in ("distributed", "distributed-strong"):
            def step():
                # all-reduced bounding box + per-level box counts (NCCL), particles stay put;
                # DFS-order cost partition; ONE all-to-all of particles; local traversal
                # (the all-reduce of the boxes' particle extents overlaps the partition and the
                # colleague pass of the setup, which completes the extents: defer_extents)
                dtree = bd.build_distributed_tree(actx, tb, comm, dsrc, defer_extents=True, **dkw)
                lt, ltrav, _, _ = bd.distributed_tree_setup(actx, dtree, tg, comm,
                                                            traversal_pieces=args.pieces or None)
                if isinstance(ltrav, list):         # row pieces (int32 CSR range)
                    nonlocal_state["nchunks"] = len(ltrav)
                    ltrav = ltrav[-1]
                return lt, ltrav
        else:
            def step():
                # round 1: NCCL all-gather of the particle slices -> replicated tree build ->
                # only this rank's share of the traversal
                g = bd.allgather_particles(actx, comm, dsrc)
                gk = dict(dkw)
                if "targets" in dkw:
                    gk["targets"] = bd.allgather_particles(actx, comm, dkw["targets"])
                if "target_radii" in dkw:
                    gk["target_radii"] = bd.allgather_particles(actx, comm, [dkw["target_radii"]])[0]
                tree, _ = tb(actx, g, **gk)
                lt, ltrav, _, _ = bd.sharded_setup(actx, tree, tg, comm)
                return lt, ltrav
        return step

    nonlocal_state = {"nchunks": 1}
    step_resident = make_step(dsrc, dkw, mode)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = None          # drop the previous step's outputs before building the next
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(args.warmup):
        out = step_resident()
        del out
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.launch_count()
    ms, out = timed(step_resident, args.steps)
    launches = _cabi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    tree, trav = out
    ms_per_step = ms / args.steps
    value = npoints_job / (ms_per_step * 1e-3) / 1e6
    nboxes, nlevels = int(tree.nboxes), int(tree.nlevels)

    # {{{ N > 1, weak arm: also the strong-scaling arm on the N=1 problem, with a parity check

    distributed_strong = None
    if mode == "distributed" and not args.no_strong:
        ssrc, skw = dev_inputs(*slice_inputs(*make_inputs(recipe, n, dtype, seed_shift=0)))
        step_strong = make_step(ssrc, skw, "distributed-strong")
        for _ in range(max(args.warmup, 3)):
            o = step_strong()
            del o
        ssteps = max(args.steps, 10)
        ms_s, o = timed(step_strong, ssteps)
        del o
        parity = strong_parity(actx, bd, tb, comm, ssrc, skw, args.workload, n)
        distributed_strong = {
            "value": n / (ms_s / ssteps * 1e-3) / 1e6, "unit": "Mpoints/s",
            "ms_per_step": ms_s / ssteps, "steps": ssteps, "scaling": "strong", "points_total": n,
            "parity": parity}
        del ssrc, skw, step_strong

    # }}}

    # {{{ end to end: pinned host buffers in, result summary out, every step

    def pin(a):
        return torch.from_numpy(a).pin_memory()

    hsrc = [pin(x) for x in src]
    hkw = {k: (pin(v) if isinstance(v, np.ndarray) else [pin(x) for x in v] if k == "targets"
               else v) for k, v in kw.items()}
    h2d_bytes = sum(x.numel() * x.element_size() for x in hsrc)
    for k, v in hkw.items():
        if isinstance(v, torch.Tensor):
            h2d_bytes += v.numel() * v.element_size()
        elif k == "targets":
            h2d_bytes += sum(x.numel() * x.element_size() for x in v)

    def upload():
        g = [x.to(device, non_blocking=True) for x in hsrc]
        gk = {k: (v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else
                  [x.to(device, non_blocking=True) for x in v] if k == "targets" else v)
              for k, v in hkw.items()}
        return g, gk

    def summary_of(t, tr):
        return torch.cat([t.level_start_box_nrs.to(torch.int64),
                          tr.from_sep_siblings_starts[-1:].to(torch.int64),
                          tr.neighbor_source_boxes_starts[-1:].to(torch.int64)]).cpu()

    def step_e2e():
        g, gk = upload()
        t, tr = make_step(g, gk, mode)()
        return t, tr, summary_of(t, tr)

    e2e_steps = max(1, min(args.steps, 10))
    e2e_value, d2h_bytes = None, 0
    if not args.no_e2e:
        o = step_e2e()
        del o
        ms_e2e, o = timed(step_e2e, e2e_steps)
        d2h_bytes = int(o[2].numel() * 8)
        e2e_value = npoints_job / (ms_e2e / e2e_steps * 1e-3) / 1e6
        del o

    # the same, double buffered: two sets of device input buffers allocated once; the H2D copy of
    # step k+1 (copy stream, from the same pinned buffers) overlaps the build of step k; every step
    # still copies its own inputs and reads its own summary back inside the timed region
    e2e_pipelined = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(device=device)
        main_stream = actx.stream

        def flat(g, gk):
            out = list(g)
            for k in sorted(gk):
                v = gk[k]
                out += [v] if isinstance(v, torch.Tensor) else (list(v) if k == "targets" else [])
            return out

        host_flat = flat(hsrc, hkw)
        bufs = []
        for _ in range(2):
            g = [torch.empty_like(x, device=device) for x in hsrc]
            gk = {k: (torch.empty_like(v, device=device) if isinstance(v, torch.Tensor) else
                      [torch.empty_like(x, device=device) for x in v] if k == "targets" else v)
                  for k, v in hkw.items()}
            bufs.append((g, gk))
        consumed = [None, None]         # main-stream event: the step that read buffer i is done

        def upload_into(i):
            with torch.cuda.stream(copy_stream):
                if consumed[i] is not None:
                    copy_stream.wait_event(consumed[i])
                for d, h in zip(flat(*bufs[i]), host_flat):
                    d.copy_(h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ev

        def run_pipelined(steps):
            ev = upload_into(0)
            last = None
            for k in range(steps):
                i = k & 1
                main_stream.wait_event(ev)
                if k + 1 < steps:
                    ev = upload_into((k + 1) & 1)
                t, tr = make_step(*bufs[i], mode)()
                last = summary_of(t, tr)           # D2H + host sync: the step's result
                consumed[i] = torch.cuda.Event()
                consumed[i].record(main_stream)
                del t, tr
            return last

        run_pipelined(3)
        psteps = max(2, min(args.steps, 10))
        ms_p, _ = timed(lambda: run_pipelined(psteps), 1)
        e2e_pipelined = {"value": npoints_job / (ms_p / psteps * 1e-3) / 1e6, "unit": "Mpoints/s",
                         "steps": psteps, "ms_per_step": ms_p / psteps,
                         "how": "double buffered: H2D of step k+1 on a copy stream into a second, "
                                "preallocated set of device buffers during the build of step k; "
                                "every step copies its own inputs from pinned host memory and "
                                "reads its own summary back"}
        del bufs

    # }}}

    # {{{ roofline: per-scope CUDA-event times on the launching stream (rank 0)

    roofline = None
    prof_steps = 2
    collective = mode not in ("single", "replicas")
    if rank == 0:
        lib.bt_prof_reset()
        lib.bt_prof_enable(1)
    if rank == 0 or collective:          # collective steps: every rank takes part
        for _ in range(prof_steps):
            tree, trav = step_resident()
        torch.cuda.synchronize()
    if rank == 0:
        lib.bt_prof_enable(0)
        rep = _cabi.profile_report()
        # scopes that only wrap other scopes
        nested_parents = {"bt_sort_particles", "trav_list13_count", "trav_list13_fill"}
        leaf = {k: v for k, v in rep.items() if k not in nested_parents
                and not k.startswith("l13h_")}          # (l13h_*: parts of l13_heavy_expand)
        total_ms = sum(v[1] for v in leaf.values())
        top = max(leaf.items(), key=lambda kv: kv[1][1])
        scope, (calls, tot) = top
        avg_ms = tot / calls
        nsrc_l = int(tree.sources[0].shape[0])
        ntgt_l = 0 if tree.sources_are_targets else int(tree.targets[0].shape[0])
        abytes = algorithmic_bytes(scope, tree, trav, n, dims, s_bytes,
                                   heavy_entries=int(tg.last_stats.get("heavy_entries_list3", 0)))
        peaks_path = os.path.join(HERE, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = (abytes / (avg_ms * 1e-3) / 1e9) if abytes else None
        # DRAM bytes per launch of every scope, from the committed ncu --set full capture
        traffic = None
        per_scope_traffic = {}
        tkey = {"config3": "config3_10000000", "uniform1e7": "uniform_10000000_f64"}.get(args.workload)
        for tname in ("r01_traffic.json", "r02_traffic.json"):      # later captures override
            tpath = os.path.join(HERE, "profiles", tname)
            if tkey and not args.n and os.path.exists(tpath):
                per_scope_traffic.update(
                    json.load(open(tpath)).get(tkey, {}).get("bytes_per_launch", {}))
        traffic = per_scope_traffic.get(scope)
        # walks chase pointers: they are bound by issue slots / latency, not by DRAM bandwidth
        walk_scopes = ("l13_walk", "l13_heavy_expand", "trav_colleagues", "trav_list4", "bt_level")
        bound = "issue" if scope.startswith(walk_scopes) else "hbm"
        # the whole path: (B_tree + B_trav) / step time / peak
        b_tree, b_trav = whole_path_bytes(tree, trav, nsrc_l, ntgt_l, dims, s_bytes,
                                          bool(tree.sources_are_targets),
                                          bool(tree.targets_have_extent))
        whole = (b_tree + b_trav) / (ms_per_step * 1e-3) / 1e9
        # time-weighted DRAM fraction of the kernels: sum(ncu dram bytes) / sum(live time) / peak
        tw = None
        if per_scope_traffic:
            tb_, tt_ = 0.0, 0.0
            for k_, (c_, ms_) in leaf.items():
                if k_ in per_scope_traffic and per_scope_traffic[k_]:
                    tb_ += per_scope_traffic[k_] * c_
                    tt_ += ms_
            if tt_ > 0:
                tw = {"dram_gbs": tb_ / (tt_ * 1e-3) / 1e9, "frac": tb_ / (tt_ * 1e-3) / 1e9 / peak,
                      "covered_share_of_step": tt_ / total_ms}
        roofline = {"bound": bound, "kernel": scope, "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                    "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": avg_ms, "launches_per_step": calls / prof_steps,
                    "share_of_step": tot / total_ms if total_ms else None,
                    "algorithmic_bytes_per_launch": abytes,
                    "whole_path": {"b_tree": b_tree, "b_trav": b_trav, "achieved": whole,
                                   "frac": whole / peak, "unit": "GB/s",
                                   "how": "(B_tree + B_trav of SURVEY 8d) / ms_per_step / peak"},
                    "time_weighted_dram": tw,
                    "top_scopes_ms_per_step": {k: round(v[1] / prof_steps, 3) for k, v in
                                               sorted(leaf.items(), key=lambda kv: -kv[1][1])[:8]}}

    # }}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_n = min(n, CPU_BASELINE_POINTS)
        v, dt = time_oracle(recipe, sample_n, dtype, steps=1, warmup=0)
        cpu_baseline = {"value": v, "unit": "Mpoints/s", "cores": host_threads(), "kind": "port",
                        "sample": ("the full workload" if sample_n == n else
                                   f"same recipe at {sample_n} points") + f", 1 step, {dt:.1f} s "
                                  "(oracle port: OpenMP over all host cores in the Morton scan, "
                                  "renumbering, level restriction, extents and every list builder)"}

    if rank == 0:
        par = {"single": "single",
               "replicas": "replicas: independent problems, no collective",
               "distributed": "distributed tree build of ONE global problem of n_gpus x points_per_gpu "
                              "points: all-reduced bbox and per-level box counts (NCCL), DFS-order "
                              "cost partition, one all-to-all of particles to the local trees, "
                              "local traversal per rank (boxtree.distributed semantics)",
               "distributed-strong": "distributed tree build of the N=1 problem cut into n_gpus slices",
               "sharded": "round 1: NCCL all-gather of particle slices, replicated tree build, "
                          "traversal rows sharded by the DFS-order work partition"}[mode]
        line = {
            "metric": "Mpoints/s TreeBuilder+FMMTraversalBuilder", "value": value,
            "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if mode in ("distributed-strong", "sharded") else "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}" + (f" (n={n})" if args.n else ""),
                       "points_per_gpu": npoints_job // world, "points_total": npoints_job,
                       "nboxes": nboxes, "nlevels": nlevels,
                       "traversal_row_pieces": nonlocal_state["nchunks"],
                       "l2_policy": "inputs larger than L2"
                       if (npoints_job // world) * dims * s_bytes > 126e6
                       else "inputs smaller than L2 (no flush)",
                       "parallelism": par},
            # headline: the double-buffered loop (a user's steady state); `serial` = upload, build
            # and read back one after the other
            "e2e": {"value": e2e_pipelined["value"] if e2e_pipelined else e2e_value,
                    "unit": "Mpoints/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_pipelined["steps"] if e2e_pipelined else e2e_steps,
                    "ms_per_step": e2e_pipelined["ms_per_step"] if e2e_pipelined else None,
                    "how": e2e_pipelined["how"] if e2e_pipelined else "serial",
                    "serial": {"value": e2e_value, "unit": "Mpoints/s", "steps": e2e_steps}},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        if distributed_strong is not None:
            line["distributed_strong"] = distributed_strong
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def strong_parity(actx, bd, tb, comm, ssrc, skw, workload, n):
    """Distributed build of the N=1 problem: the global box arrays every rank ends with are
    hashed and compared with the digests of the REFERENCE's own run on that input
    (tests/golden/full_size_digests.json); all ranks must agree."""
    import torch
    import torch.distributed as dist

    from tests.golden.make_golden import digest
    key = {"config3": "config3_3d_1e7", "uniform1e7": "uniform_3d_1e7_f64",
           "config2": "config2_3d_1e6"}.get(workload)
    full = {"config3": 10_000_000, "uniform1e7": 10_000_000, "config2": 1_000_000}.get(workload)
    dtree = bd.build_distributed_tree(actx, tb, comm, ssrc, **skw)
    nb = int(dtree.nboxes)
    got = {}
    for f in ("box_source_starts", "box_source_counts_nonchild", "box_source_counts_cumul",
              "box_target_starts", "box_target_counts_nonchild", "box_target_counts_cumul",
              "box_parent_ids", "box_levels", "box_flags"):
        got["tree." + f] = digest(getattr(dtree, f).cpu().numpy()[:nb])
    for f in ("box_child_ids", "box_centers", "box_source_bounding_box_min",
              "box_source_bounding_box_max", "box_target_bounding_box_min",
              "box_target_bounding_box_max", "level_start_box_nrs"):
        got["tree." + f] = digest(getattr(dtree, f).cpu().numpy())
    result = {"checked_fields": len(got), "nboxes": nb}
    ok = 1
    path = os.path.join(HERE, "tests", "golden", "full_size_digests.json")
    if key and n == full and os.path.exists(path):
        want = json.load(open(path))[key]
        bad = [k for k, v in got.items() if k in want and want[k] != v]
        missing = [k for k in got if k not in want]
        ok = int(not bad and nb == want["_nboxes"])
        result.update(against="reference run digests (tests/golden/full_size_digests.json: "
                              f"{key})", mismatches=bad, not_in_golden=missing)
    else:
        result.update(against="rank agreement only (no committed digest for this size)")
    # every rank holds the same global box arrays
    h = int(digest(np.frombuffer("".join(sorted(got.values())).encode(), np.uint8))[:15], 16)
    t = torch.tensor([h, -h, ok], device=actx.device, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    agree = int(t[0].item()) == h and int(t[1].item()) == -h
    result["ranks_agree"] = bool(agree)
    result["ok"] = bool(agree and int(t[2].item()) == 1)
    return result


def run_reference(args):
    """The reference arm: the CPU restatement of the reference algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    recipe, n, dtype, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    # the FULL workload per step when the whole run then stays within ~3 minutes (measured:
    # 0.53 s per 1e6 points of config 3 on 16 cores -- 20 steps of 1e7 points: under 2 minutes),
    # otherwise the largest sample of the same recipe that does (at least CPU_SAMPLE_POINTS)
    nsteps = args.steps + min(args.warmup, 1)
    per_mpoint = 0.53 * 16.0 / max(os.cpu_count() or 1, 1)
    fit = int(180.0 / (max(nsteps, 1) * per_mpoint) * 1e6)
    sample_n = n if (n <= CPU_BASELINE_POINTS and n <= fit) else min(n, max(CPU_SAMPLE_POINTS, min(fit, n)))
    v, dt = time_oracle(recipe, sample_n, dtype, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "Mpoints/s TreeBuilder+FMMTraversalBuilder", "value": v,
        "unit": "Mpoints/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "points_per_step": sample_n,
                   "points_per_gpu": sample_n, "points_total": sample_n},
        "cpu_baseline": {"value": v, "unit": "Mpoints/s", "cores": host_threads(), "kind": "port",
                         "sample": ("the full workload per step" if sample_n == n else
                                    f"same recipe at {sample_n} points per step")
                                   + " (the reference needs pyopencl/PoCL, absent here: the oracle "
                                     "port is timed instead, OpenMP over all host cores)"},
        "e2e": {"value": v, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--points", dest="n", type=int, default=0,
                    help="override the number of points (per GPU for weak workloads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=1,
                    help="single GPU: build the traversal in this many row pieces (0 = as many as "
                         "the int32 CSR range needs; config4 needs > 1)")
    ap.add_argument("--pieces", type=int, default=0,
                    help="N > 1: build every rank's local traversal in this many row pieces "
                         "(0 = as many as the int32 CSR range needs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm")
    ap.add_argument("--no-strong", action="store_true",
                    help="N > 1: skip the strong-scaling arm of the distributed run")
    ap.add_argument("--parallelism", default="distributed",
                    choices=["distributed", "distributed-strong", "replicas", "sharded"],
                    help="N > 1: distributed tree build of one global problem of N x the workload "
                         "(weak scaling, default) or of the workload itself (strong); independent "
                         "replicas; round 1's all-gather + replicated tree build")
    args = ap.parse_args()
    if args.workload == "config4" and args.chunks == 1 and not args.n:
        args.chunks = 0
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
