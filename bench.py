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
* ``e2e``: the same through the public API starting from pinned HOST buffers
  (H2D copies of coordinates/radii and a D2H read of the result summary inside
  the timed region).
* ``roofline``: the dominant kernel scope, timed live with CUDA events on the
  launching stream (library instrumentation ``bt_prof_*``), achieved algorithmic
  GB/s against the measured HBM peak of MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle (a restatement of the
  reference algorithm; the reference itself cannot run here) timed
  on the host cores on a bounded sample of the same recipe.

With N > 1 (torchrun) the default is ``--parallelism replicas``: every rank builds its
own replica of the workload with a rank-specific seed, no data-path collective, the value
is the aggregate (weak scaling).  ``--parallelism sharded`` runs ONE global problem: NCCL
all-gather of the ranks' particle slices, replicated tree build, and only the rank's share
of the traversal (box masks, local tree, local traversal of the reference's distributed
setup) -- strong scaling.
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
}
CPU_SAMPLE_POINTS = 2_000_000


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
    sharded = world > 1 and args.parallelism == "sharded"
    # replicas: every rank owns an independent problem (rank-specific seed);
    # sharded: ONE global problem, rank r starts with the r-th slice of every particle array
    src, kw = make_inputs(recipe, n, dtype, seed_shift=0 if sharded else rank)
    dims = len(src)
    s_bytes = 4 if dtype == "f32" else 8

    actx = TorchArrayContext(device)
    tb = TreeBuilder(actx)
    tg = FMMTraversalBuilder(actx)
    lib = _cabi.load()

    def to_dev(v):
        return actx.from_numpy(v)

    dsrc = [to_dev(x) for x in src]
    dkw = {k: (to_dev(v) if isinstance(v, np.ndarray) else
               [to_dev(x) for x in v] if k == "targets" else v) for k, v in kw.items()}

    def step_resident():
        tree, _ = tb(actx, dsrc, **dkw)
        trav, _ = tg(actx, tree)
        return tree, trav

    step_sharded = None
    if world > 1:
        from boxtree_b200 import distributed as bd
        comm = bd.TorchDistComm()
        # ONE global problem = rank 0's particle set; rank r starts with its r-th slice
        gsrc, gkw_np = (src, kw) if sharded else make_inputs(recipe, n, dtype, seed_shift=0)
        if sharded:
            gd, gdk = dsrc, dkw
        else:
            gd = [to_dev(x) for x in gsrc]
            gdk = {k: (to_dev(v) if isinstance(v, np.ndarray) else
                       [to_dev(x) for x in v] if k == "targets" else v) for k, v in gkw_np.items()}

        def my_slice(t):
            m = int(t.shape[0])
            return t[rank * m // world:(rank + 1) * m // world].contiguous()

        ssrc = [my_slice(x) for x in gd]
        skw = {k: (my_slice(v) if isinstance(v, torch.Tensor) else
                   [my_slice(x) for x in v] if k == "targets" else v) for k, v in gdk.items()}
        del gd, gdk

        def step_sharded():
            # NCCL all-gather of the particle slices -> replicated tree build -> only this
            # rank's share of the traversal (masks, local tree, local traversal)
            g = bd.allgather_particles(actx, comm, ssrc)
            gk = dict(skw)
            if "targets" in skw:
                gk["targets"] = bd.allgather_particles(actx, comm, skw["targets"])
            if "target_radii" in skw:
                gk["target_radii"] = bd.allgather_particles(actx, comm, [skw["target_radii"]])[0]
            tree, _ = tb(actx, g, **gk)
            local_tree, local_trav, _, _ = bd.sharded_setup(actx, tree, tg, comm)
            return local_tree, local_trav

        if sharded:
            step_resident = step_sharded  # noqa: F811

    # pinned host copies for the end-to-end arm
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

    def step_e2e():
        g = [x.to(device, non_blocking=True) for x in hsrc]
        gk = {k: (v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else
                  [x.to(device, non_blocking=True) for x in v] if k == "targets" else v)
              for k, v in hkw.items()}
        tree, _ = tb(actx, g, **gk)
        trav, _ = tg(actx, tree)
        summary = torch.cat([tree.level_start_box_nrs.to(torch.int64),
                             trav.from_sep_siblings_starts[-1:].to(torch.int64),
                             trav.neighbor_source_boxes_starts[-1:].to(torch.int64)]).cpu()
        return tree, trav, summary

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
    npoints_job = n if sharded else world * n
    value = npoints_job / (ms_per_step * 1e-3) / 1e6

    # N > 1, replicas as the main arm: also time the distributed path (ONE global problem of
    # the same workload, strong scaling) so that both numbers come from the same run
    distributed = None
    if world > 1 and not sharded:
        for _ in range(max(args.warmup, 1)):
            o = step_sharded()
            del o
        dsteps = max(1, min(args.steps, 3))
        ms_d, o = timed(step_sharded, dsteps)
        del o
        distributed = {"value": n / (ms_d / dsteps * 1e-3) / 1e6, "unit": "Mpoints/s",
                       "ms_per_step": ms_d / dsteps, "steps": dsteps, "scaling": "strong",
                       "points_total": n,
                       "parallelism": "sharded: NCCL all-gather of particle slices, replicated tree "
                                      "build, traversal rows sharded by the reference's DFS-order "
                                      "work partition (boxtree.distributed setup)"}

    # end to end (host buffers in, result summary out)
    e2e_steps = max(1, min(args.steps, 3))
    if sharded:
        def step_e2e():  # noqa: F811
            nonlocal ssrc, skw
            ssrc = [my_slice(x.to(device, non_blocking=True)) for x in hsrc]
            skw = {k: (my_slice(v.to(device, non_blocking=True)) if isinstance(v, torch.Tensor) else
                       [my_slice(x.to(device, non_blocking=True)) for x in v] if k == "targets"
                       else v) for k, v in hkw.items()}
            lt, ltrav = step_resident()
            summary = torch.cat([lt.level_start_box_nrs.to(torch.int64),
                                 ltrav.from_sep_siblings_starts[-1:].to(torch.int64),
                                 ltrav.neighbor_source_boxes_starts[-1:].to(torch.int64)]).cpu()
            return lt, ltrav, summary
    o = step_e2e()
    del o
    ms_e2e, o = timed(step_e2e, e2e_steps)
    d2h_bytes = int(o[2].numel() * 8)
    e2e_value = npoints_job / (ms_e2e / e2e_steps * 1e-3) / 1e6
    del o

    # the same, double buffered: the H2D copy of step k+1 (copy stream) overlaps the build of step
    # k; every step still copies its own inputs from pinned memory and reads its summary back
    e2e_pipelined = None
    if not sharded:
        copy_stream = torch.cuda.Stream(device=device)
        main_stream = actx.stream

        def upload():
            with torch.cuda.stream(copy_stream):
                g = [x.to(device, non_blocking=True) for x in hsrc]
                gk = {k: (v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else
                          [x.to(device, non_blocking=True) for x in v] if k == "targets" else v)
                      for k, v in hkw.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return g, gk, ev

        def run_pipelined(steps):
            nxt = upload()
            last = None
            for k in range(steps):
                g, gk, ev = nxt
                main_stream.wait_event(ev)
                if k + 1 < steps:
                    nxt = upload()
                tree, _ = tb(actx, g, **gk)
                trav, _ = tg(actx, tree)
                last = torch.cat([tree.level_start_box_nrs.to(torch.int64),
                                  trav.from_sep_siblings_starts[-1:].to(torch.int64),
                                  trav.neighbor_source_boxes_starts[-1:].to(torch.int64)]).cpu()
                used = list(g)                     # the buffers were allocated on the copy stream
                for v in gk.values():
                    used += [v] if isinstance(v, torch.Tensor) else (v if isinstance(v, list) else [])
                for t_ in used:
                    t_.record_stream(main_stream)
                del tree, trav
            return last

        run_pipelined(2)
        psteps = max(2, min(args.steps, 5))
        ms_p, _ = timed(lambda: run_pipelined(psteps), 1)
        e2e_pipelined = {"value": npoints_job / (ms_p / psteps * 1e-3) / 1e6, "unit": "Mpoints/s",
                         "steps": psteps, "ms_per_step": ms_p / psteps,
                         "how": "double buffered: H2D of step k+1 on a copy stream during the build "
                                "of step k; every step copies its own inputs and reads its summary"}

    # dominant kernel, timed live with CUDA events on the launching stream
    roofline = None
    prof_steps = 2
    if rank == 0:
        lib.bt_prof_reset()
        lib.bt_prof_enable(1)
    if rank == 0 or sharded:          # sharded steps are collective: every rank takes part
        for _ in range(prof_steps):
            tree, trav = step_resident()
        torch.cuda.synchronize()
    if rank == 0:
        lib.bt_prof_enable(0)
        rep = _cabi.profile_report()
        # scopes that only wrap other scopes
        nested_parents = {"bt_sort_particles", "trav_list13_count", "trav_list13_fill"}
        leaf = {k: v for k, v in rep.items() if k not in nested_parents}
        total_ms = sum(v[1] for v in leaf.values())
        top = max(leaf.items(), key=lambda kv: kv[1][1])
        scope, (calls, tot) = top
        avg_ms = tot / calls
        abytes = algorithmic_bytes(scope, tree, trav, n, dims, s_bytes,
                                   heavy_entries=int(tg.last_stats.get("heavy_entries_list3", 0)))
        peaks_path = os.path.join(HERE, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = (abytes / (avg_ms * 1e-3) / 1e9) if abytes else None
        # DRAM bytes of the same scope from the committed ncu --set full capture of this build
        traffic = None
        tkey = {"config3": "config3_10000000", "uniform1e7": "uniform_10000000_f64"}.get(args.workload)
        tpath = os.path.join(HERE, "profiles", "r01_traffic.json")
        if tkey and not args.n and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(tkey, {}).get("bytes_per_launch", {}).get(scope)
        roofline = {"bound": "hbm", "kernel": scope, "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                    "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": avg_ms, "launches_per_step": calls / prof_steps,
                    "share_of_step": tot / total_ms if total_ms else None,
                    "algorithmic_bytes_per_launch": abytes,
                    "top_scopes_ms_per_step": {k: round(v[1] / prof_steps, 3) for k, v in
                                               sorted(leaf.items(), key=lambda kv: -kv[1][1])[:6]}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_n = min(n, CPU_SAMPLE_POINTS)
        v, dt = time_oracle(recipe, sample_n, dtype, steps=1, warmup=0)
        cpu_baseline = {"value": v, "unit": "Mpoints/s", "cores": host_threads(), "kind": "port",
                        "sample": f"same recipe at {sample_n} points, 1 step, {dt:.1f} s "
                                  "(tree-build kernels single-threaded, traversal OpenMP)"}

    if rank == 0:
        line = {
            "metric": "Mpoints/s TreeBuilder+FMMTraversalBuilder", "value": value,
            "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}" + (f" (n={n})" if args.n else ""),
                       "points_per_gpu": n, "nboxes": tree.nboxes, "nlevels": tree.nlevels,
                       "l2_policy": "inputs larger than L2" if n * dims * s_bytes > 126e6
                       else "inputs smaller than L2 (no flush)",
                       "parallelism": ("sharded: NCCL all-gather of particle slices, replicated "
                                       "tree build, traversal rows sharded by the reference's "
                                       "DFS-order work partition" if sharded else
                                       "replicas" if world > 1 else "single")},
            "e2e": {"value": e2e_value, "unit": "Mpoints/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        if distributed is not None:
            line["distributed"] = distributed
        if e2e_pipelined is not None:
            line["e2e"]["pipelined"] = e2e_pipelined
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_reference(args):
    """The reference arm: the CPU restatement of the reference algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    recipe, n, dtype, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    sample_n = min(n, CPU_SAMPLE_POINTS)
    v, dt = time_oracle(recipe, sample_n, dtype, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "Mpoints/s TreeBuilder+FMMTraversalBuilder", "value": v,
        "unit": "Mpoints/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "points_per_step": sample_n},
        "cpu_baseline": {"value": v, "unit": "Mpoints/s", "cores": host_threads(), "kind": "port",
                         "sample": f"same recipe at {sample_n} points per step (the reference "
                                   "needs pyopencl/PoCL, absent here: oracle port timed instead)"},
        "e2e": {"value": v, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the number of points")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default="replicas", choices=["replicas", "sharded"],
                    help="N > 1: independent replicas per rank (weak scaling, default) or one "
                         "global problem with a sharded traversal (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
