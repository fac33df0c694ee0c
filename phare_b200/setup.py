"""Builds a single-level periodic simulation (patches, fields, particles) on a given back end from
host-side initial conditions: the restated equivalent of HybridModel::initialize
(src/amr/physical_models/hybrid_model.hpp:125-144) for the synthetic configs of SURVEY.md §8(d)."""
import numpy as np

from . import abi
from .solver import Patch, SolverPPC, make_level


def node_coords(layout, qty, centering_fn, domain_cells=None):
    """physical coordinates of every node of the allocation of `qty` (ghosts included), per direction.
    Computed from the GLOBAL node index, wrapped into the periodic domain, so that every patch (and every
    periodic image) evaluates the initial profile at bit-identical coordinates."""
    g = 2 if layout.interp == 1 else 4
    out = []
    for d in range(layout.dim):
        primal = centering_fn(qty, d) == 0
        n = layout.ncells[d] + (1 if primal else 0) + 2 * g
        idx = np.arange(n) - g + layout.amr_lower[d]
        if domain_cells is not None:
            idx = idx % domain_cells[d]
        x = (idx + (0.0 if primal else 0.5)) * layout.dx[d]
        out.append(x)
    return out


def build(ops, comm, domain_cells, patch_grid, interp, dx, pops, B_fn, particles_fn, solver_kw=None):
    """pops: [{"name","mass"}]; B_fn(component, coords...) -> array on the node mesh;
    particles_fn(pop index, patch layout, patch id) -> (icell, delta, weight, charge, v) host arrays."""
    from .messenger import centering
    geom, layouts = make_level(domain_cells, patch_grid, interp, dx, nranks=comm.size)
    patches = []
    for pg, L in zip(geom.patches, layouts):
        if pg.owner != comm.rank:
            continue
        loaded = [particles_fn(i, L, pg.id) for i in range(len(pops))]
        spec = [dict(name=p["name"], mass=p["mass"], n=len(ld[2])) for p, ld in zip(pops, loaded)]
        patch = Patch(ops, pg, L, spec)
        for c in range(3):
            coords = node_coords(L, abi.BX + c, centering, domain_cells)
            mesh = np.meshgrid(*coords, indexing="ij")
            ops.set_field(patch.B[c], np.broadcast_to(B_fn(c, *mesh), mesh[0].shape))
        for pop, ld in zip(patch.pops, loaded):
            ops.set_particles(pop.domain, *ld)
        patches.append(patch)
    solver = SolverPPC(ops, patches, geom, comm, **(solver_kw or {}))
    solver.initialize()
    return solver


def maxwellian_particles(layout, ppc, density_fn, vth, charge, seed, bulk=(0., 0., 0.)):
    """numpy stand-in for MaxwellianParticleInitializer (particle_initializers/maxwellian_particle_initializer.hpp
    :138-199): ppc particles per cell, weight = n(cell centre)/ppc, uniform delta, Maxwellian velocity.
    (The reference's std::mt19937_64 streams cannot be reproduced in numpy; parity tests feed BOTH sides
    the same arrays, and test_oracle_vs_reference.py checks the reference initializer separately.)"""
    dim = layout.dim
    rng = np.random.default_rng(seed)
    nc = [layout.ncells[d] for d in range(dim)]
    grids = np.meshgrid(*[np.arange(nc[d]) + layout.amr_lower[d] for d in range(dim)], indexing="ij")
    cells = np.stack([g.ravel() for g in grids], 1)
    centres = [layout.origin[d] + (cells[:, d] - layout.amr_lower[d] + 0.5) * layout.dx[d] for d in range(dim)]
    dens = np.broadcast_to(density_fn(*centres), (len(cells),))
    icell = np.repeat(cells, ppc, axis=0).astype(np.int32)
    n = len(icell)
    delta = rng.random((n, dim))
    v = rng.standard_normal((n, 3)) * vth + np.asarray(bulk)[None, :]
    weight = np.repeat(dens / ppc, ppc)
    return icell, delta, weight, np.full(n, charge), v
