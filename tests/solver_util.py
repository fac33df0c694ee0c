"""Helpers shared by the solver tests: build the same periodic problem on any back end / decomposition and
gather patch data into global arrays for comparison."""
import numpy as np

from phare_b200 import abi
from phare_b200.messenger import LocalComm, centering
from phare_b200.setup import build, maxwellian_particles


def global_particles(domain_cells, interp, dx, ppc, seed, vth=0.3, pops=1):
    """one global particle set (so that every decomposition starts from identical particles)"""
    dim = len(domain_cells)
    L = abi.make_layout(dim, interp, domain_cells, dx)
    out = []
    for i in range(pops):
        dens = (lambda *x: 1.0 + 0.2 * np.sin(2 * np.pi * x[0] / (domain_cells[0] * dx[0]))) if i == 0 \
            else (lambda *x: 0.3 + 0 * x[0])
        out.append(maxwellian_particles(L, ppc, dens, vth, 1.0, seed + 17 * i, bulk=(0.1 * i, 0., 0.)))
    return out


def B_init(domain_cells, dx):
    Lx = domain_cells[0] * dx[0]

    def fn(c, *mesh):
        x = mesh[0]
        if c == 0:
            return np.ones_like(x)
        if c == 1:
            return 0.1 * np.cos(2 * np.pi * x / Lx)
        return 0.05 * np.sin(2 * np.pi * x / Lx)
    return fn


def make_solver(ops, domain_cells, patch_grid, interp, dx, gparts, comm=None, masses=(1.0, 2.0), solver_kw=None):
    comm = comm or LocalComm()
    pops = [dict(name=f"pop{i}", mass=masses[i]) for i in range(len(gparts))]

    def particles_fn(i, L, pid):
        icell, delta, w, q, v = gparts[i]
        inside = np.ones(len(w), bool)
        for d in range(L.dim):
            inside &= (icell[:, d] >= L.amr_lower[d]) & (icell[:, d] < L.amr_lower[d] + L.ncells[d])
        return icell[inside], delta[inside], w[inside], q[inside], v[inside]

    kw = dict(resistivity=1e-3, hyper_resistivity=1e-3, Te=0.12)
    kw.update(solver_kw or {})
    return build(ops, comm, domain_cells, patch_grid, interp, dx, pops, B_init(domain_cells, dx), particles_fn, kw)


def gather_field(solver, attr, comp, qty, domain_cells):
    """physical nodes of every local patch assembled into the global periodic array (shared primal
    border nodes are checked for bit-equality between patches)"""
    ops = solver.ops
    dim = len(domain_cells)
    shape = [domain_cells[d] + (1 if centering(qty, d) == 0 else 0) for d in range(dim)]
    out = np.full(shape, np.nan)
    for p in solver.patches:
        h = getattr(p, attr)
        a = ops.get_field(h[comp] if comp is not None else h)
        g = 2 if p.layout.interp == 1 else 4
        sl_src, sl_dst = [], []
        for d in range(dim):
            n = p.layout.ncells[d] + (1 if centering(qty, d) == 0 else 0)
            sl_src.append(slice(g, g + n))
            lo = p.layout.amr_lower[d]
            sl_dst.append(slice(lo, lo + n))
        block = a[tuple(sl_src)]
        cur = out[tuple(sl_dst)]
        mask = ~np.isnan(cur)
        assert np.array_equal(cur[mask], block[mask]), f"{attr}[{comp}]: shared border nodes differ between patches"
        out[tuple(sl_dst)] = block
    return out


def all_particles(solver, ipop):
    ops = solver.ops
    parts = [ops.get_particles(p.pops[ipop].domain) for p in solver.patches]
    return tuple(np.concatenate([x[k] for x in parts]) for k in range(5))


FIELDS = [("B", 0, abi.BX), ("B", 1, abi.BY), ("B", 2, abi.BZ), ("E", 0, abi.EX), ("E", 1, abi.EY), ("E", 2, abi.EZ),
          ("Ne", None, abi.RHO), ("Vi", 0, abi.VX), ("Vi", 1, abi.VY), ("Vi", 2, abi.VZ)]
