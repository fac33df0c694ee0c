#!/usr/bin/env python
"""Benchmark of the hot path on B200: BASELINE.json's metric (particle-pushes/s, one push = interpolate +
Boris push + deposit of one macro-particle) on config 5 (3-D uniform plasma, 128^3 cells and 64 ppc PER GPU,
interp order 1), weak scaling over the GPUs of one node.

    python bench.py --gpus N --steps K --warmup W              # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path on the host cores
    python bench.py --config k ...                             # the other BASELINE configs (1-4, phare_b200/configs.py):
                                                               # their patches are dealt to the N GPUs (strong scaling)

At N > 1 a small global problem is first advanced 3 steps on the N ranks AND as one process holding every patch on
rank 0; the two must agree (fields / moments <= 1e-10 per node, particle counts per patch equal, ghost nodes equal to
the owner's values): "parity_check" in the JSON line, non-zero exit status on failure.

A "step" = one SolverPPC::advanceLevel: 3 x (Faraday, Ampere, Ohm), 2 x average, the domain_only and the all
particle sweeps (2 pushes per particle) and every same-level exchange phase.  For N > 1 the driver launches
this file under torch.distributed.run, one rank per GPU; patches are sharded one per GPU (Cartesian grid) and
halos / migrating particles go through NVLink peer memory (no collective on the data path).  Rank 0 prints ONE JSON line.

Who enqueues the step (--host): by default (auto) `value` / `ms_per_step` are timed with the C++ level driver
(include/phare_b200/solver_ppc.hpp: the host code in the reference's own language; W warm-up + K steps in one call per rank),
then the same kernels are driven once more from Python (phare_b200/solver.py, same C ABI), which brackets every launch with
CUDA events: `roofline`, `roofline_other`, `e2e` and `python_host` come from that pass.  If the C++ driver cannot run, the line
carries the Python-driven figures and `cpp_host_error`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CELLS_PER_GPU = (128, 128, 128)
PPC = 64
INTERP = 1
DX = (0.2, 0.2, 0.2)
DT = 1e-3
GPU_GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
# cells per patch of the multi-rank parity problem, by dimension
PARITY_CELLS = {1: (256,), 2: (32, 32), 3: (16, 16, 16)}
PARITY_PPC = 16
BYTES_PUSH = {1: 80, 2: 104, 3: 128}      # K1 algorithmic bytes per particle
BYTES_DEPOSIT = {1: 52, 2: 64, 3: 76}     # K3
METRIC = "particle-pushes/sec (interp+push+deposit), whole job"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region.  nvidia-smi needs a second or more to come up (longer
    on an 8-GPU box) and the timed region can be a fraction of a second, so the sampler is started BEFORE the warm-up,
    samples every 50 ms with a timestamp, and window(t0, t1) keeps the samples that fall inside the timed region
    (host clock taken right after the synchronisations that bracket it)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        rows = []
        for seen, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                rows.append((seen, float(f[2]), float(f[3]), f[6:10]))
            except ValueError:
                continue
        # a line is read a few ms after nvidia-smi took the sample: the window is widened by one sampling period
        inside = [r for r in rows if t0 is None or (t0 <= r[0] <= t1 + 0.06)]
        note = None
        if not inside and rows:
            mid = 0.5 * (t0 + t1)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            note = "timed region shorter than the sampling period: the 3 samples nearest to it"
        sm, mx, reasons = [], [], set()
        for _, a, b, flags in inside:
            sm.append(a)
            mx.append(b)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        out = dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                   reasons=sorted(reasons), samples=len(sm))
        if note:
            out["note"] = note
        return out


# ---------------------------------------------------------------------------------------------------------
def bench_config(k, n_gpus):
    """config 5: one 128^3 patch per GPU on a periodic Cartesian GPU grid (weak scaling); configs 1-4: the config's own
    domain and patches, dealt to the GPUs in order (strong scaling)"""
    from phare_b200 import configs
    if int(k) == 5:
        return configs.config5(GPU_GRIDS[n_gpus]), "weak"
    return configs.get(k), "strong"


def build_gpu_solver(cfg, n_gpus, device, comm=None):
    """the config on the CUDA back end, particles created on the device (phare_b200/configs.py)"""
    from phare_b200 import configs
    from phare_b200.messenger import LocalComm, TorchComm
    from phare_b200.solver import GpuOps
    comm = comm or (TorchComm(device) if n_gpus > 1 else LocalComm())
    ops = GpuOps(cfg.dim, cfg.interp, device)
    # capacity: below the 85 % fill at which the stores would grow
    solver = configs.build_device_loaded(ops, comm, cfg, capacity_factor=1.2)
    return solver, ops


def _node_err(got, want):
    """largest per-node |got - want| / (|want| + 1e-3 max|want|)"""
    import numpy as np
    scale = float(np.max(np.abs(want)))
    if scale == 0.0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - want) / (np.abs(want) + 1e-3 * scale)))


def _snapshot(solver):
    """every local patch: E, B (whole arrays, ghosts included), Ni, Vi, per-population moments and particle counts"""
    ops, out = solver.ops, {}
    for p in solver.patches:
        rec = dict(lo=[int(p.layout.amr_lower[d]) for d in range(p.layout.dim)],
                   n=[int(p.layout.ncells[d]) for d in range(p.layout.dim)])
        for name in ("E", "B", "Vi"):
            rec[name] = [ops.get_field(getattr(p, name)[c]) for c in range(3)]
        rec["Ni"] = ops.get_field(p.Ne)
        rec["pops"] = [dict(count=int(ops.count(pop.domain)), rho=ops.get_field(pop.rho_n),
                            flux=[ops.get_field(pop.flux[c]) for c in range(3)]) for pop in p.pops]
        out[p.geom.id] = rec
    return out


def multi_rank_parity(cfg_key, n_gpus, rank, device, steps=3):
    """N-rank correctness, visible to the driver: a small global problem of the benchmarked config (same profiles, dl, dt;
    PARITY_CELLS per patch, PARITY_PPC particles per cell, one patch per rank at least) advanced `steps` steps (a) on the
    N ranks through the peer-memory halo and migration path that the timed region uses and (b) by rank 0 alone holding
    every patch.  Compared on rank 0: every field and moment per node, particle counts per patch and population, and the
    reference's own multi-rank criterion (tests/simulator/test_advance.py:180, overlap coherence): the ghost nodes of
    every patch equal the values of the patch that owns them."""
    import gc
    import numpy as np
    import torch
    import torch.distributed as dist
    from phare_b200 import abi, configs
    from phare_b200.messenger import LocalComm, centering
    base, _ = bench_config(cfg_key, n_gpus)
    grid = base.patch_grid
    if int(np.prod(grid)) < n_gpus:
        raise RuntimeError("fewer patches than GPUs")
    per = PARITY_CELLS[base.dim]
    small = base.with_cells(tuple(per[d] * grid[d] for d in range(base.dim)), grid)
    small.pops = [dict(p, ppc=PARITY_PPC) for p in small.pops]
    solver, ops = build_gpu_solver(small, n_gpus, device)
    solver.updater.predict_min_cells = 0  # the sweeps of the timed region (predicted re-binning), whatever the patch size
    for _ in range(steps):
        solver.advance_level(small.dt)
    mine = _snapshot(solver)
    everyone = [None] * n_gpus
    dist.all_gather_object(everyone, mine)
    del solver, ops
    gc.collect()
    verdict = None
    if rank == 0:
        multi = {}
        for part in everyone:
            multi.update(part)
        one, ops1 = build_gpu_solver(small, 1, device, comm=LocalComm())
        one.updater.predict_min_cells = 0
        for _ in range(steps):
            one.advance_level(small.dt)
        single = _snapshot(one)
        del one, ops1
        worst, worst_name, counts_equal = 0.0, "", True
        g = 2 if small.interp == 1 else 4
        dim = small.dim
        for pid, a in multi.items():
            b = single[pid]
            for name in ("E", "B", "Vi"):
                for c in range(3):
                    whole = name != "Vi"
                    qty = dict(E=abi.EX, B=abi.BX, Vi=abi.VX)[name] + c
                    sl = tuple(slice(None) if whole else slice(g, g + a["n"][d] + (1 if centering(qty, d) == 0 else 0))
                               for d in range(dim))
                    e = _node_err(a[name][c][sl], b[name][c][sl])
                    if e > worst:
                        worst, worst_name = e, f"{name}{'xyz'[c]}@patch{pid}"
            phys = tuple(slice(g, g + a["n"][d] + 1) for d in range(dim))
            e = _node_err(a["Ni"][phys], b["Ni"][phys])
            if e > worst:
                worst, worst_name = e, f"Ni@patch{pid}"
            for i, (pa, pb) in enumerate(zip(a["pops"], b["pops"])):
                counts_equal = counts_equal and pa["count"] == pb["count"]
                for got, want in [(pa["rho"], pb["rho"])] + list(zip(pa["flux"], pb["flux"])):
                    e = _node_err(got[phys], want[phys])
                    if e > worst:
                        worst, worst_name = e, f"pop{i} moment@patch{pid}"
        # overlap coherence of the N-rank run: assemble the periodic global array of every E, B component from the
        # patch interiors, then every node of every patch (ghosts included) must carry the global value
        overlap = 0.0
        for name, q0 in (("E", abi.EX), ("B", abi.BX)):
            for c in range(3):
                prim = [centering(q0 + c, d) == 0 for d in range(dim)]
                G = np.full([small.cells[d] for d in range(dim)], np.nan)
                for a in multi.values():  # dual: n nodes; primal: the first n of n + 1 (the last is the neighbour's first)
                    src = tuple(slice(g, g + a["n"][d]) for d in range(dim))
                    dst = tuple(slice(a["lo"][d], a["lo"][d] + a["n"][d]) for d in range(dim))
                    G[dst] = a[name][c][src]
                scale = float(np.nanmax(np.abs(G))) or 1.0
                for a in multi.values():
                    arr = a[name][c]
                    idx = np.ix_(*[(np.arange(arr.shape[d]) - g + a["lo"][d]) % small.cells[d] for d in range(dim)])
                    overlap = max(overlap, float(np.max(np.abs(arr - G[idx]))) / scale)
        ok = bool(worst <= 1e-10 and counts_equal and overlap <= 1e-12)
        verdict = dict(ok=ok, max_rel=worst, where=worst_name, counts_equal=bool(counts_equal), overlap_max_rel=overlap,
                       steps=steps, problem=f"{small.name.split(':')[0]} profiles, {list(small.cells)} cells in "
                                            f"{list(grid)} patches, {PARITY_PPC} ppc, N ranks vs one process",
                       bar="fields and moments <= 1e-10 per node (|d| / (|ref| + 1e-3 max|ref|)), particle counts per patch "
                           "equal, every ghost node of E and B within 1e-12 max|.| of its owner's value")
    box = [verdict]
    dist.broadcast_object_list(box, src=0)
    torch.cuda.synchronize()
    return box[0]


def _cpp_snapshot(level):
    from phare_b200 import abi, host_cpp
    out, counts = {}, level.counts()
    for ip, pid in enumerate(level.patch_ids):
        rec = dict(counts=list(counts[ip]))
        for which, name, q0 in ((host_cpp.B, "B", abi.BX), (host_cpp.E, "E", abi.EX), (host_cpp.VI, "Vi", abi.VX)):
            rec[name] = [level.get_field(ip, which, c, q0 + c) for c in range(3)]
        rec["Ni"] = level.get_field(ip, host_cpp.NI, 0, abi.RHO)
        out[pid] = rec
    return out


def cpp_multi_rank_parity(cfg_key, n_gpus, rank, device, comm, steps=3):
    """multi_rank_parity for the C++ driver: the small problem of the benchmarked config advanced by the C++ level driver on the
    N ranks (peer-memory exchange between the C++ messengers) and by ONE C++ driver holding every patch on rank 0"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from phare_b200.host_cpp import CppLevel
    base, _ = bench_config(cfg_key, n_gpus)
    grid = base.patch_grid
    per = PARITY_CELLS[base.dim]
    small = base.with_cells(tuple(per[d] * grid[d] for d in range(base.dim)), grid)
    small.pops = [dict(p, ppc=PARITY_PPC) for p in small.pops]
    os.environ["PHB_PREDICT_MIN_CELLS"] = "0"  # the sweeps of the timed region (predicted re-binning), whatever the patch size
    level = CppLevel(small, device, comm=comm)
    level.initialize()
    level.advance(small.dt, steps)
    everyone = [None] * n_gpus
    dist.all_gather_object(everyone, _cpp_snapshot(level))
    level.close()
    verdict = None
    if rank == 0:
        multi = {}
        for part in everyone:
            multi.update(part)
        one = CppLevel(small, device)
        one.initialize()
        one.advance(small.dt, steps)
        single = _cpp_snapshot(one)
        one.close()
        worst, where, counts_equal = 0.0, "", True
        for pid, a in multi.items():
            b = single[pid]
            counts_equal = counts_equal and a["counts"] == b["counts"]
            for name in ("B", "E", "Vi"):
                for c in range(3):
                    ok = np.isfinite(b[name][c])
                    e = _node_err(a[name][c][ok], b[name][c][ok])
                    if e > worst:
                        worst, where = e, f"{name}{'xyz'[c]}@patch{pid}"
            e = _node_err(a["Ni"], b["Ni"])
            if e > worst:
                worst, where = e, f"Ni@patch{pid}"
        verdict = dict(ok=bool(worst <= 1e-10 and counts_equal), max_rel=worst, where=where, counts_equal=bool(counts_equal),
                       steps=steps, problem=f"{small.name.split(':')[0]} profiles, {list(small.cells)} cells in {list(grid)} "
                                            f"patches, {PARITY_PPC} ppc, C++ driver on N ranks vs one C++ driver",
                       bar="fields and moments <= 1e-10 per node, particle counts per patch equal")
    os.environ.pop("PHB_PREDICT_MIN_CELLS", None)
    box = [verdict]
    dist.broadcast_object_list(box, src=0)
    torch.cuda.synchronize()
    return box[0]


def cpp_timed_region(args, cfg, device, comm, world, local):
    """W warm-up steps, then EXACTLY K steps of the C++ level driver (ONE call into C++ per rank), bracketed by barrier +
    synchronize, CUDA events, max over ranks.  Returns dict(ms, n_local, n_total, launches, clocks, patches)."""
    import torch
    import torch.distributed as dist
    from phare_b200.host_cpp import CppLevel
    level = CppLevel(cfg, device, comm=comm)
    try:
        level.initialize()

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        sampler = ClockSampler(local)
        sampler.start()
        level.advance(cfg.dt, max(args.warmup, 3))
        barrier()
        launches0 = level.ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        e0.record()
        level.advance(cfg.dt, args.steps)
        e1.record()
        barrier()
        t1 = time.time()
        clocks = sampler.stop(t0, t1)
        ms = e0.elapsed_time(e1)
        n_local = sum(sum(c) for c in level.counts())
        n_total = n_local
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            cnt = torch.tensor([n_local], device=device, dtype=torch.int64)
            dist.all_reduce(cnt)
            n_total = int(cnt.item())
        return dict(ms=ms, n_local=n_local, n_total=n_total, launches=level.ctx.launches - launches0, clocks=clocks,
                    patches=len(level.layouts))
    finally:
        level.close()


def cpp_host_arm(args):
    """the same step driven by the C++ level driver (include/phare_b200/solver_ppc.hpp through libphare_b200_host.so):
    Python only builds the problem; the K timed steps are ONE call into C++ per rank.  At N > 1 the level is dealt to the
    ranks (one process per GPU) and the C++ messengers exchange through NVLink peer memory; torch.distributed serves the
    set-up (the IPC handles of the arenas) and this harness's own barrier / max-over-ranks.  Kernel-level roofline figures
    come from the default arm, whose Python-driven pass brackets the launches one by one."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    assert world == args.gpus, "launch with torch.distributed.run --nproc-per-node N"
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    comm, parity = None, None
    if world > 1:
        from phare_b200.messenger import TorchComm
        dist.init_process_group("nccl", device_id=device)
        comm = TorchComm(device)
        if not args.no_parity:
            parity = cpp_multi_rank_parity(args.config, world, rank, device, comm)
    cfg, scaling = bench_config(args.config, world)
    r = cpp_timed_region(args, cfg, device, comm, world, local)
    ms, n_total, n_local = r["ms"], r["n_total"], r["n_local"]
    value = 2 * n_total * args.steps / (ms * 1e-3)
    peak, _ = measured_peaks()
    dim = cfg.dim
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="particle-pushes/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling=scaling,
                    vs_baseline=None, dtype="f64", data="synthetic", host="cpp",
                    config=dict(workload=cfg.name, cells=list(cfg.cells), patch_grid=list(cfg.patch_grid),
                                ppc=[p["ppc"] for p in cfg.pops], interp_order=cfg.interp, particles_total=n_total,
                                pushes_per_step=2 * n_total,
                                parallelism=f"{r['patches']} patch(es) per GPU, {world} GPU(s), C++ level driver"),
                    per_gpu=value / world, clocks=r["clocks"], gpu_launches=r["launches"], e2e=None, roofline=None,
                    roofline_other=dict(whole_step_frac_of_hbm=round(
                        2 * n_local * (BYTES_PUSH[dim] + BYTES_DEPOSIT[dim]) / (ms / args.steps * 1e-3) / 1e9 / peak, 4)),
                    cpu_baseline=None)
        if parity is not None:
            line["parity_check"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.exit(3)


def cpp_primary(args, cfg, device, world, local, rank):
    """The default arm's PRIMARY timed region: the step as the product drives it, i.e. by the C++ level driver (one call into
    C++ per rank for the K steps; at N > 1 its own parity check first).  Any failure, on any rank, makes every rank fall back
    to the Python-driven region, and the JSON line says so."""
    import torch
    import torch.distributed as dist
    ok, res, err = 1, None, None
    try:
        comm = None
        if world > 1:
            from phare_b200.messenger import TorchComm
            comm = TorchComm(device)
        par = cpp_multi_rank_parity(args.config, world, rank, device, comm) if world > 1 and not args.no_parity else None
        res = cpp_timed_region(args, cfg, device, comm, world, local)
        res["parity"] = par
    except Exception as e:  # noqa: BLE001 - whatever it is, the Python-driven region is the fallback
        ok, err = 0, f"{type(e).__name__}: {e}"
    if world > 1:
        t = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = int(t.item())
    return res if ok else dict(error=err or "the C++ driver failed on another rank")


def our_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    n_gpus = args.gpus
    assert world == n_gpus or world == 1 and n_gpus == 1, "launch with torch.distributed.run --nproc-per-node N"
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    # host placement next to this rank's GPU, before any pinned buffer exists (phare_b200/numa.py; PHB_NUMA=0: leave it)
    from phare_b200 import numa
    placement = numa.bind_near_gpu(local) if os.environ.get("PHB_NUMA", "1") != "0" else dict(bound=False, why="PHB_NUMA=0")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    parity = multi_rank_parity(args.config, n_gpus, rank, device) if world > 1 and not args.no_parity else None
    cfg, scaling = bench_config(args.config, n_gpus)
    DT, dim = cfg.dt, cfg.dim
    # --host auto (default): `value` is timed with the C++ level driver enqueueing the step (the product's host code, the
    # reference's own language); the Python-driven pass that follows runs the same kernels through the same C ABI and brackets
    # them one by one (roofline, per-kernel ms, e2e).  Both numbers are in the line.
    cpp = cpp_primary(args, cfg, device, world, local, rank) if args.host == "auto" else None
    solver, ops = build_gpu_solver(cfg, n_gpus, device)
    n_local = sum(ops.count(pop.domain) for p in solver.patches for pop in p.pops)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        solver.advance_level(DT)
    barrier()

    # ---- timed region: EXACTLY K steps, CUDA events on the launching stream, max over ranks
    ops.kernel_timing = True
    ops.timed = {}
    launches0 = ops.ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        solver.advance_level(DT)
    e1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ops.kernel_timing = False
    ms = e0.elapsed_time(e1)
    launches = ops.ctx.launches - launches0
    kernel_ms = {k: [a.elapsed_time(b) for a, b in v] for k, v in ops.timed.items()}
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        cnt = torch.tensor([n_local], device=device, dtype=torch.int64)
        dist.all_reduce(cnt)
        n_total = int(cnt.item())
    else:
        n_total = n_local
    pushes = 2 * n_total * args.steps  # two particle sweeps per PPC step
    value = pushes / (ms * 1e-3)
    python_host = dict(value=value, ms_per_step=ms / args.steps, gpu_launches=launches, clocks=clocks)
    host = "python"
    if cpp is not None and "error" not in cpp and cpp["n_total"] == n_total:
        host, ms, launches, clocks = "cpp", cpp["ms"], cpp["launches"], cpp["clocks"]
        value = pushes / (ms * 1e-3)

    # ---- roofline of the dominant kernel (K1 fused interpolate + push), measured live in the timed region
    peak, peak_src = measured_peaks()
    push_ms = kernel_ms.get("push", [])
    dep_ms = kernel_ms.get("deposit", [])
    bin_ms = kernel_ms.get("bin", [])
    avg = lambda v: sum(v) / len(v) if v else None
    roofline = None
    extra = {}
    if push_ms or kernel_ms.get("move_domain_only"):
        # one K1 launch = one population of one patch: the average launch moves n_local / launches-per-sweep particles
        n_launch = n_local / max(1, sum(len(p.pops) for p in solver.patches))
        prof = os.path.join(ROOT, "profiles", "traffic_r2.json")
        prof = json.load(open(prof)).get(f"c{cfg.key}", {}) if os.path.exists(prof) else {}

        def kernel_roofline(label, name, ms_list, alg_bpp, traffic_key, note):
            """SURVEY 8(d) algorithmic bytes per particle x particles of one launch / the launch's CUDA-event time; traffic =
            dram__bytes_read + dram__bytes_write per particle of one `ncu --set full` capture of EXACTLY this kernel
            instantiation at this config (profiles/traffic_r2.json), or null when that capture does not exist"""
            alg = n_launch * alg_bpp
            ach = alg / (avg(ms_list) * 1e-3) / 1e9
            rec = prof.get(traffic_key) if traffic_key else None
            return dict(bound="hbm", kernel=label, kernel_name=name, achieved=round(ach, 1), peak=peak, unit="GB/s",
                        frac=round(ach / peak, 4), traffic=(rec["dram_bytes_per_particle"] * n_launch) if rec else None,
                        peak_source=peak_src, algorithmic_bytes_per_particle=alg_bpp, algorithmic_bytes_per_launch=alg,
                        avg_launch_ms=round(avg(ms_list), 4), launches_timed=len(ms_list), note=note)

        planned = not kernel_ms.get("bin_plan")  # phb_push_plan: the bin count rides on the in-place push
        k1 = None if not push_ms else kernel_roofline("K1 fused interpolate + Boris push, in place" + (" + bin count (phb_push_plan)" if planned else ""),
                             "push_tma_kernel", push_ms, BYTES_PUSH[dim] + (4 if planned else 0),
                             None if planned else "push_in_place",
                             "reads iCell, delta, v, charge; writes iCell, delta, v" + (" and the 4-byte slot" if planned else ""))
        mv_all = kernel_ms.get("move_domain_only", [])
        rb_ms = kernel_ms.get("move_all_rebin", [])
        unit = BYTES_PUSH[dim] + BYTES_DEPOSIT[dim]  # one particle-push = K1 + K3 bytes
        if rb_ms:
            # predicted re-binning (csrc/predict.cu): both sweeps are ONE pass each over the store.  The dominant kernel of the
            # step is the all sweep: interpolate + push + deposit + re-binning of a particle (K1+K3+K2 fused).  Unit of work =
            # one particle-push (K1 + K3 bytes; K2 is overhead, not algorithmic bytes); it MOVES 2 x K3 + 4 bytes.
            roofline = kernel_roofline("K1+K3+K2 fused, the all sweep in one pass: interpolate + push + deposit + write to the slot "
                                       "planned by the domain_only sweep (tile kernel, E,B block in shared memory, PLAN_REBIN)",
                                       "tile_kernel", rb_ms, unit, "move_all_rebin",
                                       "algorithmic = one particle-push of SURVEY 8(d) (K1 + K3 bytes); the kernel moves 2 x K3 + 4 "
                                       "bytes (reads the store and the 4-byte plan word, writes the re-binned store) and is bound "
                                       "by dependent FP64 issue (exact mode: no FMA contraction), ncu: profiles/r2p_kernels.md")
            roofline["frac_of_bytes_moved"] = round(n_launch * (2 * BYTES_DEPOSIT[dim] + 4) / (avg(rb_ms) * 1e-3) / 1e9 / peak, 4)
            roofline["plans_that_did_not_hold"] = solver.updater.misfiled
            roofline["rebin_fallbacks"] = solver.updater.rebin_fallbacks
            if mv_all:
                extra["move_domain_only"] = kernel_roofline(
                    "K1+K3 fused, the domain_only sweep + the plan of the re-binning (tile kernel, PLAN_PREDICT)", "tile_kernel",
                    mv_all, unit, "move_domain_only", "moves K3 bytes + the 4-byte plan word; nothing written back")
                extra["move_domain_only"]["frac_of_bytes_moved"] = round(
                    n_launch * (BYTES_DEPOSIT[dim] + 4) / (avg(mv_all) * 1e-3) / 1e9 / peak, 4)
            if k1:
                extra["push"] = k1
        elif mv_all and sum(mv_all) > sum(push_ms):  # (1-D configs run both sweeps fused: there is no separate K1 launch)
            # the dominant kernel of the step: interpolate + push + deposit of a particle in ONE pass (the domain_only sweep).
            # Its unit of work is SURVEY 8(d)'s whole "particle-push" (K1 + K3 = 80/104/128 + 52/64/76 B); it MOVES only the
            # K3 bytes (the pushed copy is never written nor re-read), which is why traffic < algorithmic bytes
            roofline = kernel_roofline("K1+K3 fused: interpolate + push + deposit in one pass (tile kernel, E,B block in shared memory)",
                                       "tile_kernel", mv_all, unit, "move_domain_only_plain",
                                       "algorithmic = one particle-push of SURVEY 8(d) (K1 + K3 bytes); the kernel moves only the "
                                       "K3 bytes and is bound by the FP64 pipe (exact mode: no FMA contraction)")
            roofline["frac_of_bytes_moved"] = round(n_launch * BYTES_DEPOSIT[dim] / (avg(mv_all) * 1e-3) / 1e9 / peak, 4)
            if k1:
                extra["push"] = k1
        else:
            roofline = k1
        if dep_ms:
            a = n_launch * BYTES_DEPOSIT[dim] / (avg(dep_ms) * 1e-3) / 1e9
            extra["deposit"] = dict(achieved=round(a, 1), frac=round(a / peak, 4), avg_launch_ms=round(avg(dep_ms), 4))
        if bin_ms:
            extra["bin"] = dict(avg_ms=round(avg(bin_ms), 4))
        ds_ms, plan_ms = kernel_ms.get("deposit_scatter", []), kernel_ms.get("bin_plan", [])
        if ds_ms:
            # K3+K2 fused (the `all` sweep): reads the store + 4 B slot, writes the re-binned store
            a = n_launch * (2 * BYTES_DEPOSIT[dim] + 4) / (avg(ds_ms) * 1e-3) / 1e9
            extra["deposit_scatter"] = dict(achieved=round(a, 1), frac=round(a / peak, 4),
                                            avg_launch_ms=round(avg(ds_ms), 4),
                                            what="deposit carried by the re-binning scatter pass (phb_deposit_scatter)")
            if plan_ms:
                extra["bin_plan"] = dict(avg_ms=round(avg(plan_ms), 4))
            else:
                extra["bin_plan"] = "folded into the in-place push (phb_push_plan)"
        NOT_KERNELS = ("exchange_phases", "finish_particles", "fp_alltoall_counts", "fp_send_recv", "fp_maintain_arrays")  # brackets of the exchange / host-synchronising parts
        extra["particle_kernels_share_of_step"] = round(
            sum(sum(v) for k, v in kernel_ms.items() if k not in NOT_KERNELS) / (python_host["ms_per_step"] * args.steps), 4)
        extra["kernel_ms_per_step"] = {k: round(sum(v) / args.steps, 3) for k, v in sorted(kernel_ms.items())}
        # whole-step fraction: 2 sweeps x (K1 + K3 algorithmic bytes) / step time
        extra["whole_step_frac_of_hbm"] = round(
            2 * n_local * (BYTES_PUSH[dim] + BYTES_DEPOSIT[dim]) / (ms / args.steps * 1e-3) / 1e9 / peak, 4)
        # bytes the particle kernels of one step actually move per particle (their own column reads + writes)
        moved = {"push": BYTES_PUSH[dim], "deposit": BYTES_DEPOSIT[dim], "move_domain_only": BYTES_DEPOSIT[dim],
                 "move_all": BYTES_PUSH[dim] + 8, "move_all_rebin": 2 * BYTES_DEPOSIT[dim] + 4,
                 "bin_plan": 4 * dim + 4, "deposit_scatter": 2 * BYTES_DEPOSIT[dim] + 4, "bin": 4 * dim + 8 + 2 * BYTES_DEPOSIT[dim]}
        if rb_ms:
            moved["move_domain_only"] += 4  # the predicting sweep writes the 4-byte plan word
        sweeps = max(1, sum(len(p.pops) for p in solver.patches))
        extra["bytes_moved_per_particle_per_step"] = round(
            sum(moved.get(k, 0) * len(v) / (args.steps * sweeps) for k, v in kernel_ms.items()), 1)
        extra["algorithmic_bytes_per_particle_per_step"] = 2 * (BYTES_PUSH[dim] + BYTES_DEPOSIT[dim])

    # ---- e2e: the same step driven with HOST buffers: the step's field inputs come from pinned host memory
    # and the step's results (moments and new fields) are read back, inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(solver, ops, args, world, device, n_total, DT)
        places = [placement]
        if world > 1:
            places = [None] * world
            dist.all_gather_object(places, placement)
        e2e["host_placement"] = places

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu_baseline = cpu_reference_rate(args.config)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="particle-pushes/s", n_gpus=n_gpus, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling=scaling,
                    vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=cfg.name, cells=list(cfg.cells), patch_grid=list(cfg.patch_grid),
                                ppc=[p["ppc"] for p in cfg.pops], interp_order=cfg.interp,
                                particles_total=n_total, pushes_per_step=2 * n_total,
                                l2=f"inputs ({n_local * BYTES_DEPOSIT[dim] / 1e9:.1f} GB of particle columns per GPU) "
                                   "against the 126 MB L2",
                                parallelism=f"{len(solver.patches)} patch(es) per GPU, {n_gpus} GPU(s)"),
                    per_gpu=value / n_gpus, clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline,
                    roofline_other=extra, cpu_baseline=cpu_baseline, host=host,
                    host_note=("value / ms_per_step / gpu_launches / clocks: K steps enqueued by the C++ level driver "
                               "(include/phare_b200/solver_ppc.hpp), one call per rank; roofline, per-kernel ms and e2e: the "
                               "Python-driven pass of the same kernels (python_host), which brackets every launch"
                               if host == "cpp" else "the step enqueued by phare_b200/solver.py"),
                    python_host=python_host)
        if cpp is not None and "error" in cpp:
            line["cpp_host_error"] = cpp["error"]
        if parity is not None:
            line["parity_check"] = parity
        if cpp is not None and cpp.get("parity") is not None:
            line["parity_check_cpp_host"] = cpp["parity"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.exit(3)
    if cpp is not None and cpp.get("parity") is not None and not cpp["parity"]["ok"]:
        sys.exit(3)


def measure_e2e(solver, ops, args, world, device, n_total, DT):
    import torch
    import torch.distributed as dist
    from phare_b200.solver import HostStaging
    staging = HostStaging(ops, solver.patches)
    bi, bo = staging.h2d_bytes, staging.d2h_bytes
    steps = max(2, min(args.steps, 5))

    def one():
        # the public host-buffer call: E,B pinned host -> device, advanceLevel, moments + E,B -> pinned host
        solver.advance_level(DT, staging=staging)
        staging.results_become_inputs()

    one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    staging.timing = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    staging.copy_stream.synchronize()  # the last step's moments are still travelling: they belong to the timed region
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # per-rank diagnosis: this rank's step time and the duration of each host transfer (PCIe / host-memory contention
    # between the ranks of a node shows up here, not in the kernels)
    mine = dict(ms_per_step=round(ms / steps, 3), **{k: round(v, 3) for k, v in staging.timings_ms().items()})
    per_rank = [mine]
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    out = dict(value=2 * n_total * steps / (ms * 1e-3), unit="particle-pushes/s", h2d_bytes_per_step=bi,
               d2h_bytes_per_step=bo, steps=steps, staging_ms_per_rank=per_rank,
               what="SolverPPC.advance_level(dt, staging=HostStaging): per step E,B from pinned host memory -> device, "
                    "advanceLevel, the new E,B -> pinned host as soon as the corrector has produced them (the next step's "
                    "inputs: on the critical path), per-population + total moments -> pinned host from a device-side "
                    "snapshot taken when the all sweep has produced them, behind the fields and underneath the next step; "
                    "the timed region ends when every result of every step is on the host; the particle store stays "
                    "device-resident (it is solver state, like the reference's ParticlesData)")
    # the naive drop-in for comparison: AoS Particle<3> records cross PCIe both ways around one sweep
    if solver.patches[0].layout.dim == 3:
        try:
            out["particles_roundtrip"] = particle_roundtrip(solver, ops, DT)
        except Exception as e:  # pragma: no cover
            out["particles_roundtrip"] = dict(error=str(e))
    return out


def particle_roundtrip(solver, ops, DT):
    """host AoS records -> device SoA, one `all` sweep kernels (push + deposit), -> host AoS, on a 16.8 M particle
    sub-sample (PCIe-bound, so the per-particle rate does not depend on the sample size)"""
    import ctypes as C
    import torch
    from phare_b200 import abi
    patch, pop = solver.patches[0], solver.patches[0].pops[0]
    n = min(ops.count(pop.domain), 64 ** 3 * 64)
    stride = 80
    host = torch.empty(n * stride, dtype=torch.uint8).pin_memory()
    view = abi.Particles()
    C.memmove(C.byref(view), C.byref(pop.domain.c), C.sizeof(abi.Particles))
    view.n = n
    lib, h = ops.ctx.lib, ops.ctx.h
    ops.ctx._check(lib.phb_particles_to_aos(h, C.byref(view), host.data_ptr()))
    tmp = ops.particles(n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    ops.ctx._check(lib.phb_particles_from_aos(h, host.data_ptr(), n, C.byref(tmp.c)))
    tmp.n = n
    ops.push(patch.layout, patch.Eavg, patch.Bavg, tmp, tmp, 1.0, DT)
    ops.deposit(patch.layout, tmp, pop.scratch[0], pop.scratch[1], patch.Ve, 1.0, 0, n, patch.non_level_ghost)
    ops.ctx._check(lib.phb_particles_to_aos(h, C.byref(tmp.c), host.data_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(value=n / (ms * 1e-3), unit="particle-pushes/s", particles=n, h2d_bytes=n * stride, d2h_bytes=n * stride,
                what="naive drop-in: Particle<3> AoS records (80 B) H2D, push+deposit, D2H, every sweep")


# ---------------------------------------------------------------------------------------------------------
CPU_SLAB = {1: (20000,), 2: (128, 128), 3: (32, 32, 32)}  # cells of one reference worker (one MPI rank of the reference)


def _cpu_worker(args):
    """one reference worker = one MPI rank of the reference: one patch of the config's (dim, interp, populations, ppc) with
    a uniform plasma, updatePopulations(domain_only) + (all) + updateIons"""
    seed, cfg_key, reps = args
    import numpy as np
    sys.path.insert(0, ROOT)
    import oracle
    from phare_b200 import abi, configs
    cfg = configs.get(cfg_key)
    dim, interp = cfg.dim, cfg.interp
    nc = CPU_SLAB[dim]
    impl = "ref" if oracle.have_ref() else "oracle"
    cpu = oracle.Cpu(impl)
    L = abi.make_layout(dim, interp, list(nc), list(cfg.dl))
    rng = np.random.default_rng(seed)
    cells = np.stack(np.meshgrid(*[np.arange(c) for c in nc], indexing="ij"), -1).reshape(-1, dim)
    soas = []
    for p in cfg.pops:
        ppc = p["ppc"]
        n = len(cells) * ppc
        soas.append((np.repeat(cells, ppc, axis=0).astype(np.int32), rng.random((n, dim)), np.full(n, 1.0 / ppc),
                     np.ones(n), rng.standard_normal((n, 3)) * 0.3))
    ntot = sum(len(s[2]) for s in soas)
    shape = lambda q: cpu.field_shape(L, q)
    E = [np.zeros(shape(abi.EX + c)) for c in range(3)]
    B = [np.zeros(shape(abi.BX + c)) for c in range(3)]
    B[0][...] = 1.0
    pg = 1 if interp == 1 else 2
    ghost = abi.make_box([-pg] * dim, [c - 1 + pg for c in nc])
    masses = [p["mass"] for p in cfg.pops]
    best = None
    for _ in range(reps):
        P = [oracle.HostParticles.from_soa(*s, capacity=int(len(s[2]) * 1.1)) for s in soas]
        if impl == "ref":
            pgs = [oracle.HostParticles(dim, len(s[2]) // 4 + 16) for s in soas]
            lgs = [oracle.HostParticles(dim, 16) for _ in soas]
            # time of the reference calls only (the wrapper's AoS staging is excluded)
            r1 = cpu.ion_update(L, E, B, masses, P, pgs, lgs, [ghost], cfg.dt, 1, update_ions=False)
            r2 = cpu.ion_update(L, E, B, masses, P, pgs, lgs, [ghost], cfg.dt, 2, update_ions=True)
            dt = r1["seconds"] + r2["seconds"]
        else:
            t0 = time.perf_counter()
            for mode in (1, 2):
                for Pi, m in zip(P, masses):
                    rc, Q = cpu.push(L, E, B, Pi, m, cfg.dt)
                    cpu.deposit(L, Q, sel=[ghost])
            dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 2 * ntot / best, impl, ntot


def cpu_reference_rate(cfg_key=5, workers=None):
    """The reference's own CPU implementation of the two particle sweeps of one PPC step
    (IonUpdater::updatePopulations(domain_only) + (all) + updateIons, compiled from the reference headers into
    oracle/_ref), one independent patch per host core (the reference is one MPI rank per core; no MPI in this
    image, so no halo cost: an upper bound of the MPI path).  Bounded sample: CPU_SLAB cells per worker."""
    import multiprocessing as mp
    from phare_b200 import configs
    cfg = configs.get(cfg_key)
    cores = workers or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(100 + i, cfg_key, 2) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    kind = "reference" if res[0][1] == "ref" else "port"
    return dict(value=total, unit="particle-pushes/s", cores=cores, kind=kind, per_core=total / cores,
                particles_per_worker=res[0][2],
                sample=f"{cores} independent patches of {'x'.join(str(c) for c in CPU_SLAB[cfg.dim])} cells, "
                       f"{len(cfg.pops)} population(s) x {cfg.pops[0]['ppc']} ppc ({cfg.dim}-D, interp {cfg.interp}), "
                       f"updatePopulations(domain_only)+(all)+updateIons, best of 2, {wall:.1f} s wall",
                flags="g++ -O3 -DNDEBUG, x86-64 baseline ISA, -ffp-contract=off (the reference's default build)")


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    t0 = time.perf_counter()
    rates = []
    from phare_b200 import configs
    cfg = configs.get(args.config)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate(args.config)
    for _ in range(max(1, min(args.steps, 3))):
        rates.append(cpu_reference_rate(args.config))
    best = max(rates, key=lambda r: r["value"])
    n_sample = best["cores"] * best["particles_per_worker"]
    line = dict(metric=METRIC, value=best["value"], unit="particle-pushes/s", n_gpus=args.gpus, steps=len(rates),
                warmup=1 if args.warmup else 0, ms_per_step=2 * n_sample / best["value"] * 1e3, higher_is_better=True,
                scaling="weak" if args.config == 5 else "strong", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference",
                config=dict(workload=cfg.name + f"; CPU arm: per-core slabs of {list(CPU_SLAB[cfg.dim])} cells, uniform "
                                                "plasma; rates are per particle, so they compare directly with the GPU arm",
                            parallelism=f"{best['cores']} host cores, one patch per core"),
                cpu_baseline=best, e2e=dict(value=best["value"], unit="particle-pushes/s", h2d_bytes_per_step=0,
                                            d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=round(time.perf_counter() - t0, 1))
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank parity check that precedes the timed region")
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--host", default="auto", choices=["auto", "python", "cpp"],
                    help="who enqueues the timed steps: auto (default) = the C++ level driver for `value`, then the "
                         "Python-driven pass of the same kernels for the per-kernel figures and e2e; python / cpp = only that one")
    a = ap.parse_args()
    os.environ.setdefault("PHB_PEER_TIMEOUT_S", "60")  # a lost neighbour ends the bench instead of holding the box
    if a.impl == "reference":
        reference_arm(a)
    elif a.host == "cpp":
        cpp_host_arm(a)
    else:
        our_arm(a)
