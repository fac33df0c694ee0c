"""One launch of each tile-path kernel (for ncu): python tools/prof_tile.py [cfg] [gs...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phare_b200 import abi
from phare_b200.device import Context
from phare_b200.torch_interop import TorchParticles, TorchArray, TorchVec, current_stream_ptr, uniform_sorted_particles
from microbench import CFGS

name = sys.argv[1] if len(sys.argv) > 1 else "c5s"
gss = [int(g) for g in sys.argv[2:]] or [8, 16]
cfg = CFGS[name]
dim, interp = cfg["dim"], cfg["interp"]
dev = torch.device("cuda:0")
ctx = Context(dim, interp, device=0, stream=current_stream_ptr())
L = abi.make_layout(dim, interp, cfg["ncells"], cfg["dx"])
P = uniform_sorted_particles(ctx, L, cfg["ppc"], 0.3, dev)
Q = TorchParticles(dim, P.capacity, dev)
E, B = TorchVec(ctx, L, abi.EX, dev), TorchVec(ctx, L, abi.BX, dev)
g = torch.Generator(device=dev)
g.manual_seed(1)
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
cs2 = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)
dt = 1e-3 * min(cfg["dx"]) / 0.2
counts = ctx.bin(L, P, Q, dom, keep, cs)
P, Q = Q, P
n = counts[0]
for gs in gss:
    os.environ["PHB_TILE_GS"] = str(gs)
    os.environ["PHB_SCATTER_GS"] = str(gs)
    ctx.push_deposit(L, E, B, P, 1.0, dt, rn, rq, F, 1.0, 0, n, keep, dom, cs, write_back=False)
    ctx.push_cells(L, E, B, P, P, n, 1.0, 0.0, dom, cs)
    ctx.push_deposit_plan(L, E, B, P, n, 1.0, dt, rn, rq, F, 1.0, keep, dom, cs, keep, cs2)
    ctx.scatter_planned(L, P, n, dom, cs, keep, Q, cs2)
    counts = ctx.bin_counts(L, dom, cs2, Q)
    P, Q = Q, P
    n = counts[0]
    cs.t.copy_(cs2.t)
ctx.poll_error()
torch.cuda.synchronize()
print("done", counts)
