#!/usr/bin/env python
"""Benchmark of the hot path on B200: BASELINE.json's metric (particle-pushes/s, one push = interpolate +
Boris push + deposit of one macro-particle) on config 5 (3-D uniform plasma, 128^3 cells and 64 ppc PER GPU,
interp order 1), weak scaling over the GPUs of one node.

    python bench.py --gpus N --steps K --warmup W              # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path on the host cores

A "step" = one SolverPPC::advanceLevel: 3 x (Faraday, Ampere, Ohm), 2 x average, the domain_only and the all
particle sweeps (2 pushes per particle) and every same-level exchange phase.  For N > 1 the driver launches
this file under torch.distributed.run, one rank per GPU; patches are sharded one per GPU (Cartesian grid) and
halos / migrating particles go through NCCL point-to-point.  Rank 0 prints ONE JSON line.
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
BYTES_PUSH = {1: 80, 2: 104, 3: 128}      # K1 algorithmic bytes per particle
BYTES_DEPOSIT = {1: 52, 2: 64, 3: 76}     # K3
METRIC = "particle-pushes/sec (interp+push+deposit), whole job"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------
def build_gpu_solver(n_gpus, rank, device):
    """one 128^3 patch per GPU on a periodic Cartesian GPU grid; particles created on the device"""
    import torch
    from phare_b200 import abi
    from phare_b200.messenger import LocalComm, TorchComm
    from phare_b200.solver import GpuOps, Patch, SolverPPC, make_level
    from phare_b200.torch_interop import uniform_sorted_particles
    grid = GPU_GRIDS[n_gpus]
    domain = tuple(CELLS_PER_GPU[d] * grid[d] for d in range(3))
    comm = TorchComm(device) if n_gpus > 1 else LocalComm()
    ops = GpuOps(3, INTERP, device)
    geom, layouts = make_level(domain, grid, INTERP, DX, nranks=n_gpus)
    patches = []
    for pg, L in zip(geom.patches, layouts):
        if pg.owner != rank:
            continue
        n = int(L.ncells[0]) * int(L.ncells[1]) * int(L.ncells[2]) * PPC
        patch = Patch(ops, pg, L, [dict(name="protons", mass=1.0, n=n)], capacity_factor=1.2)  # below the 85 % fill at which the stores would grow
        pop = patch.pops[0]
        # uniform Maxwellian: n = 1, V = 0, vth = 0.3, charge 1 (functional uniform bench)
        P = uniform_sorted_particles(ops.ctx, L, PPC, 0.3, device, seed=1337 + pg.id, store=pop.domain)
        patch.B[0].t.fill_(1.0)  # B = (1, 0, 0)
        patches.append(patch)
    solver = SolverPPC(ops, patches, geom, comm, resistivity=0.0, hyper_resistivity=1e-4, Te=0.12)
    solver.initialize()
    return solver, ops


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
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    solver, ops = build_gpu_solver(n_gpus, rank, device)
    n_local = sum(ops.count(p.pops[0].domain) for p in solver.patches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        solver.advance_level(DT)
    barrier()

    # ---- timed region: EXACTLY K steps, CUDA events on the launching stream, max over ranks
    ops.kernel_timing = True
    ops.timed = {}
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ops.ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        solver.advance_level(DT)
    e1.record()
    barrier()
    clocks = sampler.stop()
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

    # ---- roofline of the dominant kernel (K1 fused interpolate + push), measured live in the timed region
    peak, peak_src = measured_peaks()
    push_ms = kernel_ms.get("push", [])
    dep_ms = kernel_ms.get("deposit", [])
    bin_ms = kernel_ms.get("bin", [])
    avg = lambda v: sum(v) / len(v) if v else None
    roofline = None
    extra = {}
    if push_ms:
        alg = n_local * BYTES_PUSH[3]
        ach = alg / (avg(push_ms) * 1e-3) / 1e9
        # DRAM traffic per launch: dram__bytes_read+write of one `ncu --set full` capture of this kernel
        # (profiles/traffic_c5s.json, taken on the same kernel at 16.8 M particles), scaled per particle
        traffic = None
        prof = os.path.join(ROOT, "profiles", "traffic_c5s.json")
        if os.path.exists(prof):
            for k, v in json.load(open(prof)).items():
                if k.startswith("push_tma_kernel<3, 1"):
                    traffic = v["dram_bytes_per_particle"] * n_local
        roofline = dict(bound="hbm", kernel="push (K1 fused interpolate+Boris)", achieved=round(ach, 1), peak=peak,
                        unit="GB/s", frac=round(ach / peak, 4), traffic=traffic, peak_source=peak_src,
                        algorithmic_bytes_per_launch=alg, avg_launch_ms=round(avg(push_ms), 4),
                        launches_timed=len(push_ms))
        if dep_ms:
            a = n_local * BYTES_DEPOSIT[3] / (avg(dep_ms) * 1e-3) / 1e9
            extra["deposit"] = dict(achieved=round(a, 1), frac=round(a / peak, 4), avg_launch_ms=round(avg(dep_ms), 4))
        if bin_ms:
            extra["bin"] = dict(avg_ms=round(avg(bin_ms), 4))
        ds_ms, plan_ms = kernel_ms.get("deposit_scatter", []), kernel_ms.get("bin_plan", [])
        if ds_ms:
            # K3+K2 fused (the `all` sweep): reads the store + 4 B slot, writes the re-binned store
            a = n_local * (2 * BYTES_DEPOSIT[3] + 4) / (avg(ds_ms) * 1e-3) / 1e9
            extra["deposit_scatter"] = dict(achieved=round(a, 1), frac=round(a / peak, 4),
                                            avg_launch_ms=round(avg(ds_ms), 4),
                                            what="deposit carried by the re-binning scatter pass (phb_deposit_scatter)")
            extra["bin_plan"] = dict(avg_ms=round(avg(plan_ms), 4))
        extra["particle_kernels_share_of_step"] = round(sum(sum(v) for v in kernel_ms.values()) / ms, 4)
        extra["kernel_ms_per_step"] = {k: round(sum(v) / args.steps, 3) for k, v in sorted(kernel_ms.items())}
        # whole-step fraction: 2 sweeps x (K1 + K3 algorithmic bytes) / step time
        extra["whole_step_frac_of_hbm"] = round(
            2 * n_local * (BYTES_PUSH[3] + BYTES_DEPOSIT[3]) / (ms / args.steps * 1e-3) / 1e9 / peak, 4)

    # ---- e2e: the same step driven with HOST buffers: the step's field inputs come from pinned host memory
    # and the step's results (moments and new fields) are read back, inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(solver, ops, args, world, device, n_total)

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        cpu_baseline = cpu_reference_rate(sample_seconds=12.0)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="particle-pushes/s", n_gpus=n_gpus, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload="config 5: 3-D uniform Maxwellian plasma, periodic, 128^3 cells and 64 ppc "
                                         "per GPU, 1 population, interp order 1, dt=1e-3, dl=0.2",
                                cells_per_gpu=list(CELLS_PER_GPU), ppc=PPC, interp_order=INTERP,
                                particles_total=n_total, gpu_grid=list(GPU_GRIDS[n_gpus]),
                                pushes_per_step=2 * n_total,
                                l2="inputs (10.2 GB of particle columns per GPU) are far larger than the 126 MB L2",
                                parallelism=f"one patch per GPU, {n_gpus} GPU(s)"),
                    per_gpu=value / n_gpus, clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline,
                    roofline_other=extra, cpu_baseline=cpu_baseline)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(solver, ops, args, world, device, n_total):
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = dict(value=2 * n_total * steps / (ms * 1e-3), unit="particle-pushes/s", h2d_bytes_per_step=bi,
               d2h_bytes_per_step=bo, steps=steps,
               what="SolverPPC.advance_level(dt, staging=HostStaging): per step E,B from pinned host memory -> device, "
                    "advanceLevel, per-population + total moments and the new E,B -> pinned host (read back on a copy "
                    "stream as soon as each is final, overlapping the particle re-binning that ends the step); the "
                    "particle store stays device-resident (it is solver state, like the reference's ParticlesData)")
    # the naive drop-in for comparison: AoS Particle<3> records cross PCIe both ways around one sweep
    try:
        out["particles_roundtrip"] = particle_roundtrip(solver, ops)
    except Exception as e:  # pragma: no cover
        out["particles_roundtrip"] = dict(error=str(e))
    return out


def particle_roundtrip(solver, ops):
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
def _cpu_worker(args):
    """one reference worker = one MPI rank of the reference: a 3-D patch, domain_only + all + updateIons"""
    seed, ncell, reps = args
    import numpy as np
    sys.path.insert(0, ROOT)
    import oracle
    from phare_b200 import abi
    impl = "ref" if oracle.have_ref() else "oracle"
    cpu = oracle.Cpu(impl)
    L = abi.make_layout(3, INTERP, [ncell] * 3, DX)
    rng = np.random.default_rng(seed)
    n = ncell ** 3 * PPC
    cells = np.stack(np.meshgrid(*[np.arange(ncell)] * 3, indexing="ij"), -1).reshape(-1, 3)
    icell = np.repeat(cells, PPC, axis=0).astype(np.int32)
    soa = (icell, rng.random((n, 3)), np.full(n, 1.0 / PPC), np.ones(n), rng.standard_normal((n, 3)) * 0.3)
    shape = lambda q: cpu.field_shape(L, q)
    E = [np.zeros(shape(abi.EX + c)) for c in range(3)]
    B = [np.zeros(shape(abi.BX + c)) for c in range(3)]
    B[0][...] = 1.0
    dom = abi.make_box([0] * 3, [ncell - 1] * 3)
    ghost = abi.make_box([-1] * 3, [ncell] * 3)
    best = None
    for _ in range(reps):
        if impl == "ref":
            P = oracle.HostParticles.from_soa(*soa, capacity=int(n * 1.1))
            pg, lg = oracle.HostParticles(3, n // 4 + 16), oracle.HostParticles(3, 16)
            # time of the reference calls only (the wrapper's AoS staging is excluded)
            r1 = cpu.ion_update(L, E, B, [1.0], [P], [pg], [lg], [ghost], DT, 1, update_ions=False)
            r2 = cpu.ion_update(L, E, B, [1.0], [P], [pg], [lg], [ghost], DT, 2, update_ions=True)
            dt = r1["seconds"] + r2["seconds"]
        else:
            P = oracle.HostParticles.from_soa(*soa, capacity=int(n * 1.1))
            t0 = time.perf_counter()
            for mode in (1, 2):
                rc, Q = cpu.push(L, E, B, P, 1.0, DT)
                cpu.deposit(L, Q, sel=[ghost])
            dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 2 * n / best, impl


def cpu_reference_rate(sample_seconds=12.0, workers=None):
    """The reference's own CPU implementation of the two particle sweeps of one PPC step
    (IonUpdater::updatePopulations(domain_only) + (all) + updateIons, compiled from the reference headers into
    oracle/_ref), one independent patch per host core (the reference is one MPI rank per core; no MPI in this
    image, so no halo cost: an upper bound of the MPI path).  Bounded sample: 32^3 cells x 64 ppc per worker."""
    import multiprocessing as mp
    cores = workers or os.cpu_count() or 1
    ncell = 32
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(100 + i, ncell, 2) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    kind = "reference" if res[0][1] == "ref" else "port"
    return dict(value=total, unit="particle-pushes/s", cores=cores, kind=kind, per_core=total / cores,
                sample=f"{cores} independent patches of {ncell}^3 cells x {PPC} ppc (3-D, interp 1), "
                       f"updatePopulations(domain_only)+(all)+updateIons, best of 2, {wall:.1f} s wall",
                flags="g++ -O3 -DNDEBUG, x86-64 baseline ISA, -ffp-contract=off (the reference's default build)")


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    t0 = time.perf_counter()
    rates = []
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate()
    for _ in range(max(1, min(args.steps, 3))):
        rates.append(cpu_reference_rate())
    best = max(rates, key=lambda r: r["value"])
    n_sample = best["cores"] * 32 ** 3 * PPC
    line = dict(metric=METRIC, value=best["value"], unit="particle-pushes/s", n_gpus=args.gpus, steps=len(rates),
                warmup=1 if args.warmup else 0, ms_per_step=2 * n_sample / best["value"] * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload="config 5 (3-D uniform plasma, 64 ppc, interp 1) per-core slabs of 32^3 cells; "
                                     "CPU arm: rates are per particle, so they compare directly with 128^3 per GPU",
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
    a = ap.parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        our_arm(a)
