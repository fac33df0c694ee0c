"""The C++ level driver (include/phare_b200/solver_ppc.hpp behind phare_b200/lib/libphare_b200_host.so) seen from Python:
Python builds the problem (fields uploaded, particles created on the device in the stores the C++ side owns) and then
C++ runs whole steps — no interpreter between the kernels of a step.  bench.py --host cpp, tests/test_host_api_gpu.py."""
import ctypes as C
import os

import numpy as np

from . import abi
from .device import Context

HOST_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libphare_b200_host.so")
B, E, J, VI, NI, RHO_M = range(6)  # `which` of field()


class _Store:
    """view on a phb_particles struct that lives in the C++ solver"""

    def __init__(self, ptr):
        self.c = abi.Particles.from_address(ptr)

    @property
    def n(self):
        return int(self.c.n)

    @n.setter
    def n(self, v):
        self.c.n = int(v)

    @property
    def capacity(self):
        return int(self.c.capacity)


class CppLevel:
    """comm (messenger.TorchComm) with size > 1: the level is dealt to the ranks like make_level deals it (one process per
    GPU); the C++ driver of every rank holds its own patches and reaches its neighbours' memory over NVLink.  torch.distributed
    is used ONCE here, to all-gather the 64-byte CUDA IPC handles of the ranks' arenas."""

    def __init__(self, cfg, device="cuda:0", capacity_factor=1.2, comm=None):
        import torch
        from . import torch_interop as ti
        from .solver import make_level
        if not os.path.exists(HOST_LIB):
            raise RuntimeError(f"{HOST_LIB} not built: run `make host` (or __graft_entry__.build())")
        abi.load()  # libphare_b200.so first (the host library resolves its symbols against it)
        lib = self.lib = C.CDLL(HOST_LIB)
        lib.phh_create.restype = C.c_void_p
        lib.phh_create_distributed.restype = C.c_void_p
        lib.phh_ctx.restype = C.c_void_p
        lib.phh_layout.restype = C.POINTER(abi.Layout)
        lib.phh_field.restype = C.c_void_p
        lib.phh_particles.restype = C.c_void_p
        lib.phh_last_error.restype = C.c_char_p
        for fn in (lib.phh_ctx, lib.phh_npatch, lib.phh_layout, lib.phh_field, lib.phh_particles, lib.phh_initialize,
                   lib.phh_advance, lib.phh_add_population, lib.phh_destroy, lib.phh_patch_id, lib.phh_arena_export,
                   lib.phh_arena_open, lib.phh_release_peers):
            fn.argtypes = None
        self.cfg, self.torch, self.comm = cfg, torch, comm
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        dim = cfg.dim
        world = comm.size if comm is not None else 1
        rank = comm.rank if comm is not None else 0
        geom, layouts = make_level(cfg.cells, cfg.patch_grid, cfg.interp, cfg.dl, nranks=world)
        boxes = (abi.Box * len(layouts))()
        for i, L in enumerate(layouts):
            for d in range(dim):
                boxes[i].lower[d] = L.amr_lower[d]
                boxes[i].upper[d] = L.amr_lower[d] + L.ncells[d] - 1
        dbl = lambda xs: (C.c_double * len(xs))(*[float(x) for x in xs])
        cells = (C.c_int * dim)(*[int(c) for c in cfg.cells])
        if world > 1:
            owners = (C.c_int * len(layouts))(*[int(pg.owner) for pg in geom.patches])
            h = lib.phh_create_distributed(self.device.index or 0, dim, cfg.interp, len(layouts), boxes, owners, rank, world,
                                           dbl(cfg.dl), cells, C.c_double(cfg.eta), C.c_double(cfg.nu), 0, C.c_double(cfg.Te))
        else:
            h = lib.phh_create(self.device.index or 0, dim, cfg.interp, len(layouts), boxes, dbl(cfg.dl), cells,
                               C.c_double(cfg.eta), C.c_double(cfg.nu), 0, C.c_double(cfg.Te))
        if not h:
            raise RuntimeError("phh_create: " + lib.phh_last_error().decode())
        self.h = C.c_void_p(h)
        self.patch_ids = [pg.id for pg in geom.patches if pg.owner == rank]
        if world > 1:
            # the one exchange between the ranks that does not go through device memory: the handles of the arenas
            mine = (C.c_ubyte * 64)()
            self._check(lib.phh_arena_export(self.h, mine), "phh_arena_export")
            everyone = [None] * world
            comm.dist.all_gather_object(everyone, bytes(mine))
            self._check(lib.phh_arena_open(self.h, (C.c_ubyte * (64 * world))(*b"".join(everyone))), "phh_arena_open")
            assert [lib.phh_patch_id(self.h, C.c_int(i)) for i in range(len(self.patch_ids))] == self.patch_ids
            layouts = [layouts[i] for i in self.patch_ids]
        # everything on torch's current stream, like GpuOps
        self.ctx = Context.adopt(lib.phh_ctx(self.h), dim, cfg.interp, stream=ti.current_stream_ptr())
        self.layouts = layouts
        self._load(capacity_factor)

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: status {rc}: {self.lib.phh_last_error().decode()}")

    def field_ptr(self, patch, which, comp=0):
        return self.lib.phh_field(self.h, C.c_int(patch), C.c_int(which), C.c_int(comp))

    def get_field(self, patch, which, comp=0, qty=None):
        qty = qty if qty is not None else {B: abi.BX, E: abi.EX, J: abi.JX, VI: abi.VX}.get(which, abi.RHO - comp) + comp
        shape = self.ctx.field_shape(self.layouts[patch], qty)
        out = np.empty(shape)
        self.ctx._check(self.ctx.lib.phb_d2h(self.ctx.h, out.ctypes.data, C.c_void_p(self.field_ptr(patch, which, comp)),
                                             out.nbytes))
        return out

    def store(self, patch, pop):
        return _Store(self.lib.phh_particles(self.h, C.c_int(patch), C.c_int(pop)))

    def _load(self, capacity_factor):
        """phare_b200.configs.build_device_loaded for the C++-owned buffers: B at the node coordinates, particles by the
        device loader (same Philox streams: the same particles as the Python-driven level)"""
        from .messenger import centering
        from .setup import node_coords
        from .solver import GpuOps
        cfg, dim, Ld = self.cfg, self.cfg.dim, self.cfg.lengths
        ops = GpuOps.__new__(GpuOps)  # only its loader helper is used, on the adopted context
        import torch
        from . import torch_interop as ti
        ops.torch, ops.ti, ops.device, ops.ctx, ops.dim, ops.interp = torch, ti, self.device, self.ctx, dim, cfg.interp
        for ip, L in enumerate(self.layouts):
            for c in range(3):
                mesh = np.meshgrid(*node_coords(L, abi.BX + c, centering, cfg.cells), indexing="ij")
                host = np.ascontiguousarray(np.broadcast_to(cfg.B[c](mesh, Ld), mesh[0].shape), dtype=np.float64)
                self.ctx._check(self.ctx.lib.phb_h2d(self.ctx.h, C.c_void_p(self.field_ptr(ip, B, c)), host.ctypes.data,
                                                     host.nbytes))
            self.ctx.sync()
        ncells = [int(np.prod([L.ncells[d] for d in range(dim)])) for L in self.layouts]
        for ipop, p in enumerate(cfg.pops):
            caps = (C.c_size_t * len(ncells))(*[int(n * p["ppc"] * capacity_factor) + 4096 for n in ncells])
            idx = self.lib.phh_add_population(self.h, p["name"].encode(), C.c_double(p["mass"]), caps)
            if idx != ipop:
                raise RuntimeError("phh_add_population: " + self.lib.phh_last_error().decode())
            for ip, L in enumerate(self.layouts):
                ncell = ncells[ip]
                axes = [(np.arange(L.ncells[d]) + L.amr_lower[d] + 0.5) * L.dx[d] for d in range(dim)]
                centres = [m.reshape(-1) for m in np.meshgrid(*axes, indexing="ij")]
                n = np.ascontiguousarray(np.broadcast_to(p["density"](centres, Ld), (ncell,)), dtype=np.float64)
                vth = np.ascontiguousarray(np.broadcast_to(p["vth"](centres, Ld), (ncell,)), dtype=np.float64)
                V = [np.full(ncell, float(p["bulk"][k])) for k in range(3)]
                first = np.arange(ncell + 1, dtype=np.uint32) * np.uint32(p["ppc"])
                st = self.store(ip, ipop)
                ops.maxwellian_load(L, n, V, [vth, vth, vth], first, ncell * p["ppc"], p["charge"], p["ppc"], p["seed"],
                                    cfg.cells, st)
                st.n = ncell * p["ppc"]

    def initialize(self):
        self._check(self.lib.phh_initialize(self.h), "phh_initialize")

    def advance(self, dt, nsteps=1):
        self._check(self.lib.phh_advance(self.h, C.c_double(dt), C.c_int(nsteps)), "phh_advance")

    def counts(self):
        return [[self.store(ip, i).n for i in range(len(self.cfg.pops))] for ip in range(len(self.layouts))]

    def close(self):
        if self.h:
            self.torch.cuda.synchronize()
            if self.comm is not None and self.comm.size > 1:
                # nobody frees an arena that a neighbour still has mapped
                self.lib.phh_release_peers(self.h)
                self.comm.dist.barrier()
            self.lib.phh_destroy(self.h)
            self.h = None
