"""Micro-benchmark of the coarse <-> fine level operators (K9, csrc/level.cu) and of phb_split on one B200: a 128^3
fine patch (config-5 size) refined from a 64^3 coarse patch.  The field kernels run for a few microseconds, less than
one ctypes call takes to issue, so they are timed as a CUDA graph of 20 back-to-back calls (captured on the stream the
library launches on, replayed 5 times, best replay / 20; the NaN re-fill the conditional refiners need is captured too
and its own time subtracted); phb_split returns a count to the host and is timed with CUDA events.  Algorithmic bytes = nodes written x 8 B + coarse nodes read x 8 B (refine), fine read + coarse written
(coarsen), 76 B read + nref x 76 B written per split coarse particle.
usage: python tools/microbench_level.py  -> gpurun_out/microbench_level.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phare_b200 import abi
from phare_b200.device import Context
from phare_b200.split import pattern
from phare_b200.torch_interop import TorchVec, TorchParticles, current_stream_ptr, uniform_sorted_particles
import ctypes as C

from microbench import timeit, HBM

REPS = 20


def timeit_graph(ctx, fn, prep=None):
    """ms per call of fn (GPU time only): a graph of REPS x (prep; fn) minus a graph of REPS x prep"""
    s = torch.cuda.Stream()
    ctx._check(ctx.lib.phb_set_stream(ctx.h, C.c_void_p(s.cuda_stream)))

    def one(body):
        with torch.cuda.stream(s):
            body()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(REPS):
                    body()
            best = 1e30
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(s)
                g.replay()
                b.record(s)
                s.synchronize()
                best = min(best, a.elapsed_time(b))
        return best / REPS

    def both():
        if prep:
            prep()
        fn()
    t = one(both) - (one(prep) if prep else 0.0)
    ctx._check(ctx.lib.phb_set_stream(ctx.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return t


def main():
    dim, interp, g = 3, 1, 2
    dev = torch.device("cuda:0")
    ctx = Context(dim, interp, device=0, stream=current_stream_ptr())
    nc = np.array([64, 64, 64])
    clo = np.array([32, 32, 32])
    flo, fhi = 2 * clo, 2 * (clo + nc - 1) + 1
    Lf = abi.make_layout(dim, interp, list(2 * nc), [0.1] * 3, amr_lower=list(flo), level=1)
    Lc = abi.make_layout(dim, interp, list(nc), [0.2] * 3, amr_lower=list(clo))
    out = {}

    def report(label, ms, nbytes):
        gbs = nbytes / ms / 1e6
        out[label] = dict(ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / HBM, 3), mbytes=round(nbytes / 1e6, 2))
        print(f"{label:34s} {ms:8.4f} ms {nbytes / 1e6:9.2f} MB {gbs:8.1f} GB/s {gbs / HBM:6.3f} of HBM peak", flush=True)

    prim = lambda qty: np.array([1 if (qty <= abi.BZ and qty - abi.BX == d) or (abi.EX <= qty <= abi.JZ and (qty - abi.EX) % 3 != d)
                                 or qty >= abi.RHO else 0 for d in range(3)])
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    B, E = TorchVec(ctx, Lf, abi.BX, dev), TorchVec(ctx, Lf, abi.EX, dev)
    cB, cE = TorchVec(ctx, Lc, abi.BX, dev), TorchVec(ctx, Lc, abi.EX, dev)
    for v in (cB, cE):
        for c in range(3):
            v[c].t.normal_(0, 1, generator=gen)
    nan = float("nan")
    for name, op, fine, coarse, q0 in (("refine magnetic_init (3 comps)", abi.REFINE_MAGNETIC_INIT, B, cB, abi.BX),
                                       ("refine magnetic NaN-only (3 comps)", abi.REFINE_MAGNETIC, B, cB, abi.BX),
                                       ("refine electric (3 comps)", abi.REFINE_ELECTRIC, E, cE, abi.EX),
                                       ("refine default (3 comps)", abi.REFINE_DEFAULT, E, cE, abi.EX)):
        def prep():
            for c in range(3):
                fine[c].t.fill_(nan)

        def fn():
            for c in range(3):
                p = prim(q0 + c)
                ctx.field_refine(op, q0 + c, coarse[c], clo - g, fine[c], flo - g, flo - g + 1, fhi + g + p - 1)
        ms = timeit_graph(ctx, fn, prep=prep)
        nb = sum(fine[c].size * 8 + coarse[c].size * 8 for c in range(3))
        report(name, ms, nb)
    ms = timeit_graph(ctx, lambda: ctx.magnetic_postprocess(Lf, B, flo - g, fhi + g))
    # every new fine face (half of the faces of each component) is written, the coarse faces around it are read
    report("magnetic post-process (Toth-Roe)", ms, sum(B[c].size * 8 for c in range(3)))
    for name, op, fine, coarse, q0 in (("coarsen electric (3 comps)", abi.COARSEN_ELECTRIC, E, cE, abi.EX),):
        def fn():
            for c in range(3):
                ctx.field_coarsen(op, q0 + c, fine[c], flo - g, coarse[c], clo - g, clo, clo + nc - 1 + prim(q0 + c))
        ms = timeit_graph(ctx, fn)
        report(name, ms, sum(coarse[c].size * 8 * 3 for c in range(3)))  # 1 write + <= 2 fine reads per coarse node
    a, b = E[0], E[1]
    ms = timeit_graph(ctx, lambda: ctx.axpy(a, b, 0.25))
    report("fluxSum axpy (1 comp)", ms, a.size * 24)
    ms = timeit_graph(ctx, lambda: ctx.box_fill(a, [0, 0, 0], a.shape, nan))
    report("NaN fill (1 comp, whole array)", ms, a.size * 8)
    # particle splitting: a 64^3 coarse patch with 64 ppc -> nref = 6 children each, all inside the fine patch
    P = uniform_sorted_particles(ctx, Lc, 64, 0.3, dev)
    nref = 6
    d, w, m = pattern(dim, interp, nref)
    F = TorchParticles(dim, P.n * nref + 1024, dev)
    box = [abi.make_box(list(flo), list(fhi))]

    def split():
        F.n = 0
        return ctx.split(P, 0, P.n, d, w, m, box, F)
    ms, _ = timeit(split, iters=3, warm=1)
    kept = F.n
    report(f"split {P.n / 1e6:.1f} M coarse -> {kept / 1e6:.1f} M fine", ms, P.n * 76 + kept * 76)
    out["split_particles"] = dict(coarse=P.n, fine=kept, gcoarse_s=round(P.n / ms / 1e6, 2))
    ctx.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/microbench_level.json", "w"), indent=1)


if __name__ == "__main__":
    main()
