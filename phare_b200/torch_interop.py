"""Plumbing only: wrap torch CUDA tensors as the pointer structs of the C ABI (torch is used for device
memory, streams and torch.distributed; no torch op is on the compute path)."""
import torch

from . import abi


class TorchParticles:
    """SoA particle store whose columns are torch tensors (one per component)."""

    def __init__(self, dim, capacity, device):
        self.dim = dim
        cap = max(int(capacity), 1)
        self.icell = [torch.zeros(cap, dtype=torch.int32, device=device) for _ in range(dim)]
        self.delta = [torch.zeros(cap, dtype=torch.float64, device=device) for _ in range(dim)]
        self.v = [torch.zeros(cap, dtype=torch.float64, device=device) for _ in range(3)]
        self.weight = torch.zeros(cap, dtype=torch.float64, device=device)
        self.charge = torch.zeros(cap, dtype=torch.float64, device=device)
        self.c = abi.Particles()
        for d in range(dim):
            self.c.icell[d] = self.icell[d].data_ptr()
            self.c.delta[d] = self.delta[d].data_ptr()
        for k in range(3):
            self.c.v[k] = self.v[k].data_ptr()
        self.c.weight = self.weight.data_ptr()
        self.c.charge = self.charge.data_ptr()
        self.c.n = 0
        self.c.capacity = cap

    @property
    def n(self):
        return int(self.c.n)

    @n.setter
    def n(self, v):
        self.c.n = int(v)

    @property
    def capacity(self):
        return int(self.c.capacity)


class TorchArray:
    def __init__(self, shape, device, dtype=torch.float64, tensor=None):
        self.t = tensor if tensor is not None else torch.zeros(tuple(shape), dtype=dtype, device=device)
        self.ptr = self.t.data_ptr()
        self.shape = tuple(self.t.shape)
        self.size = self.t.numel()

    def zero(self):
        self.t.zero_()


class TorchVec:
    def __init__(self, ctx, layout, qty0, device, tensors=None):
        self.comps = [TorchArray(ctx.field_shape(layout, qty0 + c), device,
                                 tensor=None if tensors is None else tensors[c]) for c in range(3)]
        self.c = abi.VecField()
        for c in range(3):
            self.c.comp[c] = self.comps[c].ptr

    def __getitem__(self, i):
        return self.comps[i]


def current_stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def uniform_sorted_particles(ctx, layout, ppc, vth, device, seed=0, capacity_factor=1.0, store=None):
    """ppc particles per domain cell in row-major cell order: synthetic input created on the device."""
    dim = layout.dim
    nc = [int(layout.ncells[d]) for d in range(dim)]
    ncell = 1
    for n_ in nc:
        ncell *= n_
    n = ncell * ppc
    P = store if store is not None else TorchParticles(dim, int(n * capacity_factor) + 1024, device)
    assert P.capacity >= n
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cell = torch.arange(ncell, device=device, dtype=torch.int64).repeat_interleave(ppc)
    rem = cell
    for d in reversed(range(dim)):
        P.icell[d][:n] = (rem % nc[d]).to(torch.int32) + int(layout.amr_lower[d])
        rem = rem // nc[d]
    del cell, rem
    for d in range(dim):
        P.delta[d][:n] = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    for k in range(3):
        P.v[k][:n] = torch.randn(n, generator=g, device=device, dtype=torch.float64) * vth
    P.weight[:n] = 1.0 / ppc
    P.charge[:n] = 1.0
    P.n = n
    return P
