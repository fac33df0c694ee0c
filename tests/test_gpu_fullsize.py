"""Size-independent properties at BASELINE.json's FULL sizes (the oracle cannot finish these in seconds):
  C1 1-D 16384 cells x 100 ppc (order 1)          C2 1-D order 2, 1000 cells x 100 ppc
  C3 2-D 512x256 cells x 100 ppc x 2 pops (o 1)   C4 2-D order 3, one 512x512 GPU shard x 100 ppc
  C5 3-D 128^3 cells x 64 ppc (order 1)
Checked through the C ABI: particle-number conservation through push+bin, sortedness and offset consistency of
the binned store, idempotence of bin, B-spline partition of unity of the deposit (sum over nodes of rho_n ==
sum of weights, likewise charge and flux), linearity of the deposit in the weights / coef, agreement of the
cell-ordered kernel with the any-order kernel, and |v| conservation of the Boris rotation when E = 0."""
import numpy as np
import pytest
import torch

from phare_b200 import abi
from phare_b200.device import Context
from phare_b200.torch_interop import TorchParticles, TorchArray, TorchVec, current_stream_ptr, uniform_sorted_particles

pytestmark = pytest.mark.gpu

CONFIGS = {
    "C1": dict(dim=1, interp=1, ncells=[16384], ppc=100, dx=[0.2]),
    "C2": dict(dim=1, interp=2, ncells=[1000], ppc=100, dx=[1.0]),
    "C3": dict(dim=2, interp=1, ncells=[512, 256], ppc=100, dx=[0.4, 0.4]),
    "C4": dict(dim=2, interp=3, ncells=[512, 512], ppc=100, dx=[0.2, 0.2]),
    "C5": dict(dim=3, interp=1, ncells=[128, 128, 128], ppc=64, dx=[0.2, 0.2, 0.2]),
}


def cell_keys(P, n, lo, ext, dim):
    k = torch.zeros(n, dtype=torch.int64, device=P.weight.device)
    for d in range(dim):
        k = k * ext[d] + (P.icell[d][:n].to(torch.int64) - lo[d])
    return k


@pytest.mark.parametrize("name", list(CONFIGS))
def test_fullsize_properties(name):
    cfg = CONFIGS[name]
    dim, interp = cfg["dim"], cfg["interp"]
    dev = torch.device("cuda:0")
    ctx = Context(dim, interp, device=0, stream=current_stream_ptr())
    L = abi.make_layout(dim, interp, cfg["ncells"], cfg["dx"])
    pg = 1 if interp == 1 else 2
    lo, hi = [0] * dim, [c - 1 for c in cfg["ncells"]]
    dom = abi.make_box(lo, hi)
    keep = [abi.make_box([l - pg for l in lo], [h + pg for h in hi])]
    P = uniform_sorted_particles(ctx, L, cfg["ppc"], 0.3, dev, seed=5)
    n = P.n
    Q = TorchParticles(dim, P.capacity, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    E, B = TorchVec(ctx, L, abi.EX, dev), TorchVec(ctx, L, abi.BX, dev)
    for c in range(3):
        B[c].t.normal_(0, 0.3, generator=g)  # E = 0: the Boris rotation must conserve |v| to rounding
    speed0 = torch.sqrt(sum(P.v[k][:n] ** 2 for k in range(3)))
    wsum = float(P.weight[:n].sum())
    dt = 0.2 * min(cfg["dx"])

    # ---- push (out of place), then bin
    ctx.push(L, E, B, P, Q, 1.0, dt)
    ctx.poll_error()
    speed1 = torch.sqrt(sum(Q.v[k][:n] ** 2 for k in range(3)))
    assert float((speed1 - speed0).abs().max() / speed0.max()) < 1e-13
    assert all(bool(((Q.delta[d][:n] >= 0) & (Q.delta[d][:n] <= 1)).all()) for d in range(dim))
    cs = TorchArray((ctx.bin_nkeys(L, dom) + 1,), dev, dtype=torch.int32)
    S = TorchParticles(dim, P.capacity, dev)
    counts = ctx.bin(L, Q, S, dom, keep, cs)
    assert sum(counts) == n and counts[2] == 0  # nobody can leave the ghost box with |v| dt / dx << 1
    nd = int(np.prod(cfg["ncells"]))
    starts = cs.t.to(torch.int64)
    assert int(starts[0]) == 0 and int(starts[nd]) == counts[0] and int(starts[-1]) == n
    assert bool((starts[1:] >= starts[:-1]).all())
    # sortedness: keys of the domain range are non-decreasing and agree with cell_start
    keys = cell_keys(S, counts[0], lo, cfg["ncells"], dim)
    assert bool((keys[1:] >= keys[:-1]).all())
    hist = torch.bincount(keys, minlength=nd)
    assert torch.equal(hist, starts[1:nd + 1] - starts[:nd])
    # conservation of the multiset: sums of every column are unchanged by the sort (exact for the integer column)
    assert int(S.icell[0][:n].to(torch.int64).sum()) == int(Q.icell[0][:n].to(torch.int64).sum())
    assert abs(float(S.weight[:n].sum()) - wsum) < 1e-9 * wsum
    # idempotence
    S2 = TorchParticles(dim, P.capacity, dev)
    cs2 = TorchArray(cs.shape, dev, dtype=torch.int32)
    counts2 = ctx.bin(L, S, S2, dom, keep, cs2)
    assert counts2 == (counts[0], counts[1], 0) and torch.equal(cs.t[:-1], cs2.t[:-1])

    # ---- deposit: partition of unity, linearity, cell-ordered == any-order
    def moments():
        return [TorchArray(ctx.field_shape(L, abi.RHO), dev) for _ in range(2)], TorchVec(ctx, L, abi.VX, dev)

    (rn, rq), F = moments()
    ctx.deposit(L, S, rn, rq, F, 1.0, 0, counts[0], keep, dom, cs)
    if counts[1]:
        ctx.deposit(L, S, rn, rq, F, 1.0, counts[0], n, keep)
    for arr, col in ((rn, None), (rq, S.charge), (F[0], S.v[0]), (F[1], S.v[1]), (F[2], S.v[2])):
        want = float((S.weight[:n] * (col[:n] if col is not None else 1.0)).sum())
        scale = float((S.weight[:n] * (col[:n].abs() if col is not None else 1.0)).sum())
        assert abs(float(arr.t.sum()) - want) <= 1e-10 * scale
    (rn2, rq2), F2 = moments()
    ctx.deposit(L, S, rn2, rq2, F2, 2.0, 0, n, keep)  # any-order kernel, coef = 2
    for a, b in ((rn, rn2), (rq, rq2), (F[0], F2[0]), (F[1], F2[1]), (F[2], F2[2])):
        scale = float(a.t.abs().max())
        assert float((2.0 * a.t - b.t).abs().max()) <= 1e-10 * scale
    ctx.close()
