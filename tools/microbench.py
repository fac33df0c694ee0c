"""Kernel micro-benchmarks on one B200 (CUDA events on the stream the kernels are launched on).
Not the contract bench (that is bench.py); used to steer optimisation.  usage: python tools/microbench.py [cfg]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phare_b200 import abi
from phare_b200.device import Context
from phare_b200.torch_interop import (TorchParticles, TorchArray, TorchVec, current_stream_ptr,
                                      uniform_sorted_particles)

HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0

CFGS = {
    "c1": dict(dim=1, interp=1, ncells=[16384 * 64], ppc=100, dx=[0.2]),
    "c2": dict(dim=1, interp=2, ncells=[500 * 2048], ppc=100, dx=[1.0]),
    "c3": dict(dim=2, interp=1, ncells=[1024, 1024], ppc=100, dx=[0.4, 0.4]),
    "c4": dict(dim=2, interp=3, ncells=[1024, 512], ppc=100, dx=[0.2, 0.2]),
    "c5": dict(dim=3, interp=1, ncells=[128, 128, 128], ppc=64, dx=[0.2, 0.2, 0.2]),
    "c5s": dict(dim=3, interp=1, ncells=[64, 64, 64], ppc=64, dx=[0.2, 0.2, 0.2]),
}


def timeit(fn, iters=5, warm=2, prep=None):
    for _ in range(warm):
        if prep:
            prep()
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if prep:
            prep()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def run(name, cfg):
    dim, interp = cfg["dim"], cfg["interp"]
    dev = torch.device("cuda:0")
    ctx = Context(dim, interp, device=0, stream=current_stream_ptr())
    L = abi.make_layout(dim, interp, cfg["ncells"], cfg["dx"])
    P = uniform_sorted_particles(ctx, L, cfg["ppc"], 0.3, dev)
    n = P.n
    Q = TorchParticles(dim, P.capacity, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    E = TorchVec(ctx, L, abi.EX, dev)
    B = TorchVec(ctx, L, abi.BX, dev)
    for c in range(3):
        E[c].t.normal_(0, 0.01, generator=g)
        B[c].t.normal_(0, 0.01, generator=g)
    B[0].t += 1.0
    rn, rq = TorchArray(ctx.field_shape(L, abi.RHO), dev), TorchArray(ctx.field_shape(L, abi.RHO), dev)
    F = TorchVec(ctx, L, abi.VX, dev)
    lo = [0] * dim
    hi = [cfg["ncells"][d] - 1 for d in range(dim)]
    dom = abi.make_box(lo, hi)
    pg = 1 if interp == 1 else 2
    keep = [abi.make_box([l - pg for l in lo], [h + pg for h in hi])]
    cs = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)
    bpp_push = {1: 80, 2: 104, 3: 128}[dim]
    bpp_dep = {1: 52, 2: 64, 3: 76}[dim]
    dt = 1e-3 * min(cfg["dx"]) / 0.2
    out = {}

    def report(label, ms, bytes_pp):
        gbs = n * bytes_pp / ms / 1e6
        out[label] = dict(ms=round(ms, 3), gparts_s=round(n / ms / 1e6, 2), gbs=round(gbs, 1), frac=round(gbs / HBM, 3))
        print(f"{name:4s} {label:22s} {ms:9.3f} ms  {n / ms / 1e6:8.2f} Gp/s  {gbs:8.1f} GB/s  {gbs / HBM:6.3f} of HBM peak",
              flush=True)

    if os.environ.get('PHB_MB_ONLY') == 'predict':
        for d in range(dim):
            P.icell[d][:n] = torch.randint(0, cfg["ncells"][d], (n,), generator=g, device=dev, dtype=torch.int32)
        out = run_predict(name, cfg, ctx, L, P, Q, E, B, rn, rq, F, dom, keep, cs, dt, report)
        ctx.poll_error()
        ctx.close()
        return out
    if os.environ.get('PHB_MB_ONLY') == 'tile':
        if os.environ.get('PHB_MB_RANDOM', '1') != '0':
            # cells drawn at random (Poisson counts per cell, like a store a few steps into a run) instead of exactly
            # ppc per cell: kernels whose lanes own cells must not depend on every cell holding the same count
            for d in range(dim):
                P.icell[d][:n] = torch.randint(0, cfg["ncells"][d], (n,), generator=g, device=dev, dtype=torch.int32)
        out = run_tile(name, cfg, ctx, L, P, Q, E, B, rn, rq, F, dom, keep, cs, dt, report)
        ctx.poll_error()
        ctx.close()
        return out
    if os.environ.get('PHB_MB_ONLY') != 'fused':
        for exact in (True, False):
            ctx.set_exact(exact)
            ms, _ = timeit(lambda: ctx.push(L, E, B, P, Q, 1.0, dt))
            report(f"push[{'exact' if exact else 'fma'}] oop", ms, bpp_push + 16)
        ctx.set_exact(True)
        if os.environ.get('PHB_MB_ONLY') == 'push':
            ms, _ = timeit(lambda: ctx.push(L, E, B, P, P, 1.0, 0.0))
            report("push[exact] in place", ms, bpp_push)
            ctx.close()
            return out
        # bin first so that cell_start is valid (P is sorted already, Q pushed)
        counts = ctx.bin(L, Q, P, dom, keep, cs)
        ms, _ = timeit(lambda: ctx.bin(L, Q, P, dom, keep, cs))
        report("bin (count+scan+scatter)", ms, 2 * bpp_dep)
        print("   counts", counts, flush=True)

        def dep_cells():
            ctx.deposit(L, P, rn, rq, F, sel=keep, domain=dom, cell_start=cs, last=counts[0])
        ms, _ = timeit(dep_cells)
        report("deposit[cells]", ms, bpp_dep)
        if n <= 40_000_000:
            ms, _ = timeit(lambda: ctx.deposit(L, P, rn, rq, F, sel=keep), iters=2, warm=1)
            report("deposit[atomic]", ms, bpp_dep)
        ms, _ = timeit(lambda: ctx.push(L, E, B, P, P, 1.0, 0.0))  # dt = 0: in place, store stays sorted
        report("push[exact] in place", ms, bpp_push)
        # the way the step uses it: deposit right after a real push, on the stale order (movers -> list kernel)
        ctx.push(L, E, B, P, P, 1.0, dt)
        ms, _ = timeit(dep_cells)
        report("deposit[cells, stale order]", ms, bpp_dep)
    # K1+K3 fused (phb_push_deposit) on the freshly binned store, the way the step uses it
    state = {"P": P, "Q": Q}

    def rebin():
        c = ctx.bin(L, state["P"], state["Q"], dom, keep, cs)
        state["P"], state["Q"] = state["Q"], state["P"]
        state["n"] = c[0]
    rebin()

    def fused(wb):
        S = state["P"]
        ctx.push_deposit(L, E, B, S, 1.0, dt, rn, rq, F, 1.0, 0, state["n"], keep, dom, cs, write_back=wb)
    bpp_ro = bpp_dep
    ms, _ = timeit(lambda: fused(False))
    report("move[domain_only] fused", ms, bpp_ro)
    ms, _ = timeit(lambda: fused(True), prep=rebin)
    report("move[all] fused", ms, bpp_ro + 4 * dim + 8 * dim + 24)
    # the two-pass equivalents on the same state
    rebin()
    S, T = state["P"], state["Q"]

    def two_pass_domain_only():
        ctx.push(L, E, B, S, T, 1.0, dt)
        ctx.deposit(L, T, rn, rq, F, sel=keep, domain=dom, cell_start=cs, last=state["n"])
    ms, _ = timeit(two_pass_domain_only)
    report("push+deposit [domain_only]", ms, bpp_push + 16 + bpp_dep)
    # K3+K2 fused: the `all` sweep's deposit carried by the re-binning scatter, against deposit then bin
    rebin()
    S, T = state["P"], state["Q"]
    ctx.push(L, E, B, S, S, 1.0, dt)
    cs2 = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)

    def plan_and_scatter():
        ctx.bin_plan(L, S, dom, keep, cs2)
        ctx.deposit_scatter(L, S, state["n"], rn, rq, F, 1.0, keep, dom, cs, keep, T, cs2)

    def deposit_then_bin():
        ctx.deposit(L, S, rn, rq, F, sel=keep, domain=dom, cell_start=cs, last=state["n"])
        ctx.bin(L, S, T, dom, keep, cs2)
    ms, _ = timeit(plan_and_scatter)
    report("plan + deposit_scatter [all]", ms, 3 * bpp_dep)
    ms, _ = timeit(deposit_then_bin)
    report("deposit + bin [all]", ms, 3 * bpp_dep)
    ctx.poll_error()
    ctx.close()
    return out


def run_tile(name, cfg, ctx, L, P, Q, E, B, rn, rq, F, dom, keep, cs, dt, report):
    """the tile kernels (E,B block in shared memory) the way the step uses them: sweep 1 = push+deposit without
    write-back, sweep 2 = push in place + deposit + plan, then the scatter; plus K1 alone (phb_push_cells)"""
    dim = cfg["dim"]
    dev = torch.device("cuda:0")
    n = P.n
    bpp_push = {1: 80, 2: 104, 3: 128}[dim]
    bpp_dep = {1: 52, 2: 64, 3: 76}[dim]
    state = {"P": P, "Q": Q}

    def rebin():
        c = ctx.bin(L, state["P"], state["Q"], dom, keep, cs)
        state["P"], state["Q"] = state["Q"], state["P"]
        state["n"] = c[0]
    rebin()
    gss = [int(g) for g in os.environ.get("PHB_MB_GS", "16").split(",")]
    cs2 = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)
    for gs in gss:
        os.environ["PHB_TILE_GS"] = str(gs)
        S = state["P"]
        ms, _ = timeit(lambda: ctx.push_deposit(L, E, B, S, 1.0, dt, rn, rq, F, 1.0, 0, state["n"], keep, dom, cs,
                                                write_back=False))
        report(f"tile sweep1 push+dep gs{gs}", ms, bpp_dep)
        ms, _ = timeit(lambda: ctx.push_cells(L, E, B, S, S, state["n"], 1.0, 0.0, dom, cs))
        report(f"tile K1 in place gs{gs}", ms, bpp_push)
        T = state["Q"]
        view = type("V", (), {})()
        view.c = abi.Particles()
        import ctypes as C
        C.memmove(C.byref(view.c), C.byref(T.c), C.sizeof(abi.Particles))
        view.c.weight, view.c.charge = S.c.weight, S.c.charge
        ms, _ = timeit(lambda: ctx.push_cells(L, E, B, S, view, state["n"], 1.0, dt, dom, cs))
        report(f"tile K1 out of place gs{gs}", ms, bpp_push)

        def sweep2():
            ctx.push_deposit_plan(L, E, B, state["P"], state["n"], 1.0, dt, rn, rq, F, 1.0, keep, dom, cs, keep, cs2)

        def scatter():
            ctx.scatter_planned(L, state["P"], state["n"], dom, cs, keep, state["Q"], cs2)
        # time the two passes separately: plan on a fresh order each time (prep = the scatter of the previous one + swap)
        sweep2()
        ts_plan, ts_sc = [], []
        for it in range(5):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            scatter()
            b.record()
            counts = ctx.bin_counts(L, dom, cs2, state["Q"])
            state["P"], state["Q"] = state["Q"], state["P"]
            state["n"] = counts[0]
            cs.t.copy_(cs2.t)
            a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a2.record()
            sweep2()
            b2.record()
            torch.cuda.synchronize()
            ts_sc.append(a.elapsed_time(b))
            ts_plan.append(a2.elapsed_time(b2))
        report(f"tile sweep2 push+dep+plan gs{gs}", min(ts_plan), bpp_dep + bpp_dep - 16)
        report(f"scatter_planned gs{gs}", min(ts_sc), 2 * bpp_dep)
        scatter()
        counts = ctx.bin_counts(L, dom, cs2, state["Q"])
        state["P"], state["Q"] = state["Q"], state["P"]
        state["n"] = counts[0]
        cs.t.copy_(cs2.t)
    return {}


def run_predict(name, cfg, ctx, L, P, Q, E, B, rn, rq, F, dom, keep, cs, dt, report):
    """the two sweeps of a step as the predicted re-binning runs them (csrc/predict.cu), beside the plain sweep 1"""
    dim = cfg["dim"]
    dev = torch.device("cuda:0")
    bpp_dep = {1: 52, 2: 64, 3: 76}[dim]
    c = ctx.bin(L, P, Q, dom, keep, cs)
    S, T, ns = Q, P, c[0]
    S.n = ns
    cs2 = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)
    nbytes = ctx.predict_plan_bytes(L, dom, S.capacity)
    plan = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=dev)
    ms, _ = timeit(lambda: ctx.push_deposit(L, E, B, S, 1.0, dt, rn, rq, F, 1.0, 0, ns, keep, dom, cs, write_back=False))
    report("sweep1 plain (push+dep)", ms, bpp_dep)
    predict = lambda: ctx.push_deposit_predict(L, E, B, S, ns, 1.0, dt, rn, rq, F, 1.0, keep, dom, cs, keep,
                                               plan.data_ptr(), nbytes)
    rebin = lambda: ctx.push_deposit_rebin(L, E, B, S, ns, 1.0, dt, rn, rq, F, 1.0, keep, dom, cs, keep, T, cs2,
                                           plan.data_ptr(), nbytes)
    t1, t2 = [], []
    for it in range(6):
        a, b, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        predict()
        b.record()
        rebin()
        c2.record()
        torch.cuda.synchronize()
        t1.append(a.elapsed_time(b))
        t2.append(b.elapsed_time(c2))
    report("sweep1 predict (push+dep+plan)", min(t1[1:]), bpp_dep + 4)
    report("sweep2 rebin (push+dep+scatter)", min(t2[1:]), 2 * bpp_dep + 4)
    print("   counts+misfiled", ctx.predict_counts(L, dom, cs2, plan.data_ptr(), T), flush=True)
    return {}


if __name__ == "__main__":
    names = sys.argv[1:] or ["c5s", "c5", "c3", "c4", "c1", "c2"]
    res = {}
    for nm in names:
        res[nm] = run(nm, CFGS[nm])
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(os.environ.get("PHB_MB_OUT", "gpurun_out/microbench.json"), "w"), indent=1)
