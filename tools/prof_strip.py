"""One launch of the strip K1 and of the streaming K1 on the same binned store (for ncu / quick timing):
python tools/prof_strip.py [cfg]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phare_b200 import abi
from phare_b200.device import Context
from phare_b200.torch_interop import TorchParticles, TorchArray, TorchVec, current_stream_ptr, uniform_sorted_particles
from microbench import CFGS

name = sys.argv[1] if len(sys.argv) > 1 else "c5s"
cfg = CFGS[name]
dim, interp = cfg["dim"], cfg["interp"]
dev = torch.device("cuda:0")
ctx = Context(dim, interp, device=0, stream=current_stream_ptr())
L = abi.make_layout(dim, interp, cfg["ncells"], cfg["dx"])
P = uniform_sorted_particles(ctx, L, cfg["ppc"], 0.3, dev)
n = P.n
g = torch.Generator(device=dev)
g.manual_seed(1)
for d in range(dim):  # Poisson counts per cell
    P.icell[d][:n] = torch.randint(0, cfg["ncells"][d], (n,), generator=g, device=dev, dtype=torch.int32)
Q = TorchParticles(dim, P.capacity, dev)
E, B = TorchVec(ctx, L, abi.EX, dev), TorchVec(ctx, L, abi.BX, dev)
for c in range(3):
    E[c].t.normal_(0, 0.01, generator=g)
    B[c].t.normal_(0, 0.01, generator=g)
B[0].t += 1.0
lo = [0] * dim
hi = [cfg["ncells"][d] - 1 for d in range(dim)]
dom = abi.make_box(lo, hi)
pg = 1 if interp == 1 else 2
keep = [abi.make_box([l - pg for l in lo], [h + pg for h in hi])]
cs = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)
cs2 = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)
counts = ctx.bin(L, P, Q, dom, keep, cs)
P, Q = Q, P
n = counts[0]
dt = 1e-3 * min(cfg["dx"]) / 0.2


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


bpp = {1: 80, 2: 104, 3: 128}[dim]
for label, fn in (("strip K1 in place (dt=0)", lambda: ctx.push_cells(L, E, B, P, P, n, 1.0, 0.0, dom, cs)),
                  ("streaming K1 in place (dt=0)", lambda: ctx.push(L, E, B, P, P, 1.0, 0.0)),
                  ("strip K1 + plan (dt=0)", lambda: ctx.push_plan(L, E, B, P, 1.0, 0.0, dom, keep, cs2, n, cs)),
                  ("streaming K1 + plan (dt=0)", lambda: ctx.push_plan(L, E, B, P, 1.0, 0.0, dom, keep, cs2))):
    ms = timeit(fn)
    print(f"{name} {label:32s} {ms:8.3f} ms  {P.n * bpp / ms / 1e6 / 6536.4:6.3f} of HBM peak", flush=True)
ctx.poll_error()
