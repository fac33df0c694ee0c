"""The five BASELINE.json configurations restated as synthetic inputs; used by bench.py --config k and
by the parity tests against the step oracle (tests/test_configs_gpu.py), so both run THE SAME definitions.

Profiles are the reference's own input scripts restated (paths relative to the PHARE tree):
  C1  tools/bench/functional/simulation_setup.py:10-67           1-D uniform Maxwellian, 1 population, 100 ppc, interp 1
  C2  tests/functional/td/td1d.py:17-87                          1-D tangential discontinuity, interp 2
  C3  tests/functional/harris/harris_2d.py:49-122                2-D Harris double current sheet, 2 populations, interp 1
  C4  tests/functional/ionIonBeam/ion_ion_beam1d.py:19-92        ion-ion beam lifted to 2-D, main + beam, interp 3
  C5  tools/bench/functional/test_cases/uniform.py:13-26         3-D uniform plasma, 64 ppc, interp 1

Every function of space takes the node / cell-centre coordinates (one array per direction) and the domain lengths."""
import numpy as np


def _S(x, x0, l):
    return 0.5 * (1.0 + np.tanh((x - x0) / l))


def _const(val):
    return lambda X, Ld: np.full_like(X[0], val, dtype=np.float64)


class Config:
    """one restated configuration; `cells` may be overridden (scaled-down parity runs keep profiles, dl, dt, ppc)"""

    def __init__(self, key, name, interp, cells, dl, dt, pops, B, Te, eta, nu, patch_grid, steps):
        self.key, self.name, self.interp = key, name, interp
        self.cells, self.dl, self.dt = tuple(cells), tuple(dl), dt
        self.pops, self.B = pops, B
        self.Te, self.eta, self.nu = Te, eta, nu
        self.patch_grid, self.steps = tuple(patch_grid), steps

    @property
    def dim(self):
        return len(self.cells)

    @property
    def lengths(self):
        return tuple(c * d for c, d in zip(self.cells, self.dl))

    def with_cells(self, cells, patch_grid=None):
        return Config(self.key, self.name, self.interp, cells, self.dl, self.dt, self.pops, self.B, self.Te, self.eta,
                      self.nu, patch_grid or self.patch_grid, self.steps)

    def B_fn(self):
        """B_fn(component, *mesh) for phare_b200.setup.build"""
        Ld = self.lengths
        return lambda c, *mesh: self.B[c](mesh, Ld)

    def solver_kw(self):
        return dict(resistivity=self.eta, hyper_resistivity=self.nu, Te=self.Te)

    def particles_total(self):
        return int(np.prod(self.cells)) * sum(p["ppc"] for p in self.pops)


def _pop(name, density, vth, bulk=(0., 0., 0.), ppc=100, mass=1.0, charge=1.0, seed=1337):
    return dict(name=name, density=density, vth=vth, bulk=bulk, ppc=ppc, mass=mass, charge=charge, seed=seed)


def config1():
    return Config(1, "config 1: 1-D uniform Maxwellian plasma, 1 population, 100 ppc, interp order 1, 16384 cells, dl=0.2, "
                     "dt=1e-3", 1, (16384,), (0.2,), 1e-3,
                  [_pop("protons", _const(1.0), _const(0.3), ppc=100, seed=1337)],
                  [_const(1.0), _const(0.0), _const(0.0)], Te=0.12, eta=0.0, nu=1e-4, patch_grid=(8,), steps=20)


def config2():
    def by(X, Ld):
        L = Ld[0]
        return -1.0 + 2.0 * (_S(X[0], L * 0.25, 1.0) - _S(X[0], L * 0.75, 1.0))

    def T(X, Ld):  # td1d.py: vth = T (sic), T = (K - B^2/2) / n with K = 1, n = 1, bx = 0, bz = 0.5
        return 1.0 - (by(X, Ld) ** 2 + 0.25) * 0.5

    return Config(2, "config 2: 1-D tangential discontinuity (td1d), interp order 2, 500 cells, dl=1.0, 100 ppc, root level",
                  2, (500,), (1.0,), 1e-2, [_pop("protons", _const(1.0), T, ppc=100, seed=2001)],
                  [_const(0.0), by, _const(0.5)], Te=0.12, eta=0.0, nu=1e-4, patch_grid=(5,), steps=10)


def config3():
    Lsheet = 0.5

    def density(X, Ld):
        y, Ly = X[1], Ld[1]
        return 0.4 + 1.0 / np.cosh((y - Ly * 0.3) / Lsheet) ** 2 + 1.0 / np.cosh((y - Ly * 0.7) / Lsheet) ** 2

    def bx(X, Ld):
        x, y = X
        Lx, Ly = Ld
        x0, y1, y2 = x - 0.5 * Lx, y - 0.3 * Ly, y - 0.7 * Ly
        dBx1 = -2 * 0.1 * y1 * np.exp(-(x0 ** 2 + y1 ** 2))
        dBx2 = 2 * 0.1 * y2 * np.exp(-(x0 ** 2 + y2 ** 2))
        return -1.0 + 2.0 * (_S(y, Ly * 0.3, Lsheet) - _S(y, Ly * 0.7, Lsheet)) + dBx1 + dBx2

    def by(X, Ld):
        x, y = X
        Lx, Ly = Ld
        x0, y1, y2 = x - 0.5 * Lx, y - 0.3 * Ly, y - 0.7 * Ly
        return 2 * 0.1 * x0 * np.exp(-(x0 ** 2 + y1 ** 2)) - 2 * 0.1 * x0 * np.exp(-(x0 ** 2 + y2 ** 2))

    def vth(X, Ld):
        T = (0.7 - (bx(X, Ld) ** 2 + by(X, Ld) ** 2) * 0.5) / density(X, Ld)
        return np.sqrt(T)

    half = lambda X, Ld: 0.5 * density(X, Ld)  # the sheet's density carried half by each of the two populations
    return Config(3, "config 3: 2-D Harris double current sheet, 2 populations x 100 ppc, interp order 1, 512x256 cells, "
                     "dl=0.4, dt=5e-3", 1, (512, 256), (0.4, 0.4), 5e-3,
                  [_pop("protons", half, vth, ppc=100, seed=12334), _pop("background", half, vth, ppc=100, seed=12335)],
                  [bx, by, _const(0.0)], Te=0.0, eta=1e-3, nu=2e-3, patch_grid=(4, 2), steps=10)


def config4():
    vth = _const(float(np.sqrt(0.1)))
    return Config(4, "config 4: 2-D ion-ion beam, main (n=1) + beam (n=0.01, Vx=5), 100 ppc each, interp order 3, "
                     "2048x1024 cells, dl=0.2, dt=1e-3", 3, (2048, 1024), (0.2, 0.2), 1e-3,
                  [_pop("main", _const(1.0), vth, ppc=100, seed=4001),
                   _pop("beam", _const(0.01), vth, bulk=(5.0, 0., 0.), ppc=100, seed=4002)],
                  [_const(1.0), _const(0.0), _const(0.0)], Te=0.0, eta=0.0, nu=0.01, patch_grid=(4, 2), steps=10)


def config5(gpu_grid=(1, 1, 1)):
    cells = tuple(128 * g for g in gpu_grid)
    return Config(5, "config 5: 3-D uniform Maxwellian plasma, periodic, 128^3 cells and 64 ppc per GPU, 1 population, "
                     "interp order 1, dt=1e-3, dl=0.2", 1, cells, (0.2, 0.2, 0.2), 1e-3,
                  [_pop("protons", _const(1.0), _const(0.3), ppc=64, seed=1337)],
                  [_const(1.0), _const(0.0), _const(0.0)], Te=0.12, eta=0.0, nu=1e-4, patch_grid=gpu_grid, steps=10)


CONFIGS = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5}


def get(k):
    return CONFIGS[int(k)]()


# ---- host-side particle loading for the parity runs (numpy; both sides of a comparison get THE SAME arrays) ----------
def host_particles(cfg, ipop):
    """ppc particles per cell over the whole domain: weight = n(cell centre)/ppc, uniform delta, Maxwellian velocities
    with the population's local thermal speed and bulk velocity (the distribution MaxwellianParticleInitializer samples,
    maxwellian_particle_initializer.hpp:138-199; the reference's mt19937_64 stream is reproduced by
    phb_maxwellian_load_host and checked in tests/test_frontend.py — here both sides only need identical inputs)."""
    pop = cfg.pops[ipop]
    dim, ppc = cfg.dim, pop["ppc"]
    rng = np.random.default_rng(pop["seed"])
    grids = np.meshgrid(*[np.arange(c) for c in cfg.cells], indexing="ij")
    cells = np.stack([g.ravel() for g in grids], 1)
    centres = [(cells[:, d] + 0.5) * cfg.dl[d] for d in range(dim)]
    dens = np.broadcast_to(pop["density"](centres, cfg.lengths), (len(cells),))
    vth = np.broadcast_to(pop["vth"](centres, cfg.lengths), (len(cells),))
    icell = np.repeat(cells, ppc, axis=0).astype(np.int32)
    n = len(icell)
    delta = rng.random((n, dim))
    v = rng.standard_normal((n, 3)) * np.repeat(vth, ppc)[:, None] + np.asarray(pop["bulk"])[None, :]
    weight = np.repeat(dens / ppc, ppc)
    return icell, delta, weight, np.full(n, pop["charge"]), v


def build_host_loaded(ops, comm, cfg):
    """the config on the given back end with host-loaded particles (parity runs); returns (solver, gparts)"""
    from .setup import build
    gparts = [host_particles(cfg, i) for i in range(len(cfg.pops))]

    def particles_fn(i, L, pid):
        icell, delta, w, q, v = gparts[i]
        inside = np.ones(len(w), bool)
        for d in range(L.dim):
            inside &= (icell[:, d] >= L.amr_lower[d]) & (icell[:, d] < L.amr_lower[d] + L.ncells[d])
        return icell[inside], delta[inside], w[inside], q[inside], v[inside]

    pops = [dict(name=p["name"], mass=p["mass"]) for p in cfg.pops]
    solver = build(ops, comm, cfg.cells, cfg.patch_grid, cfg.interp, cfg.dl, pops, cfg.B_fn(), particles_fn, cfg.solver_kw())
    return solver, gparts


def build_device_loaded(ops, comm, cfg, capacity_factor=1.2):
    """the config on the CUDA back end with particles created on the device (phb_maxwellian_load: Philox keyed by the
    population seed, counted by global cell — the same particles whatever the patch decomposition and rank count);
    used by bench.py and by its multi-rank parity check"""
    from . import abi
    from .messenger import centering
    from .setup import node_coords
    from .solver import Patch, SolverPPC, make_level
    geom, layouts = make_level(cfg.cells, cfg.patch_grid, cfg.interp, cfg.dl, nranks=comm.size)
    dim, Ld = cfg.dim, cfg.lengths
    patches = []
    for pg, L in zip(geom.patches, layouts):
        if pg.owner != comm.rank:
            continue
        ncell = int(np.prod([L.ncells[d] for d in range(dim)]))
        spec = [dict(name=p["name"], mass=p["mass"], n=ncell * p["ppc"]) for p in cfg.pops]
        patch = Patch(ops, pg, L, spec, capacity_factor=capacity_factor)
        for c in range(3):
            coords = node_coords(L, abi.BX + c, centering, cfg.cells)
            mesh = np.meshgrid(*coords, indexing="ij")
            ops.set_field(patch.B[c], np.broadcast_to(cfg.B[c](mesh, Ld), mesh[0].shape))
        axes = [(np.arange(L.ncells[d]) + L.amr_lower[d] + 0.5) * L.dx[d] for d in range(dim)]
        centres = [m.reshape(-1) for m in np.meshgrid(*axes, indexing="ij")]
        for pop, p in zip(patch.pops, cfg.pops):
            n = np.ascontiguousarray(np.broadcast_to(p["density"](centres, Ld), (ncell,)), dtype=np.float64)
            vth = np.ascontiguousarray(np.broadcast_to(p["vth"](centres, Ld), (ncell,)), dtype=np.float64)
            V = [np.full(ncell, float(p["bulk"][k])) for k in range(3)]
            first = np.arange(ncell + 1, dtype=np.uint32) * np.uint32(p["ppc"])
            ops.maxwellian_load(L, n, V, [vth, vth, vth], first, ncell * p["ppc"], p["charge"], p["ppc"], p["seed"],
                                cfg.cells, pop.domain)
            ops.set_count(pop.domain, ncell * p["ppc"])
        patches.append(patch)
    solver = SolverPPC(ops, patches, geom, comm, **cfg.solver_kw())
    solver.initialize()
    return solver
