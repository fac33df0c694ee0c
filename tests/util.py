"""Shared seeded input generators for the parity tests (used on both sides of every comparison)."""
import numpy as np

from phare_b200 import abi

ALL_DIM_INTERP = [(d, o) for d in (1, 2, 3) for o in (1, 2, 3)]
# the (dim, interp) pairs of BASELINE.json's five configs
CONFIG_DIM_INTERP = [(1, 1), (1, 2), (2, 1), (2, 3), (3, 1)]


def bits(x):
    x = np.ascontiguousarray(x)
    return x.view(np.int64) if x.dtype == np.float64 else x


def bit_equal(a, b):
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def small_layout(dim, interp, ncells=None, level=0):
    nc = (ncells or [12, 9, 7])[:dim]
    dx = [0.2, 0.3, 0.25][:dim]
    return abi.make_layout(dim, interp, nc, dx, amr_lower=[5, -3, 10][:dim], level=level)


def domain_box(L):
    lo = [L.amr_lower[d] for d in range(L.dim)]
    hi = [L.amr_lower[d] + L.ncells[d] - 1 for d in range(L.dim)]
    return abi.make_box(lo, hi)


def grown(box, dim, w):
    return abi.make_box([box.lower[d] - w for d in range(dim)], [box.upper[d] + w for d in range(dim)])


def particle_ghosts(interp):
    return 1 if interp == 1 else 2


def random_vec(rng, shape_of, L, q0, scale=1.0):
    return [scale * rng.standard_normal(shape_of(L, q0 + c)) for c in range(3)]


def random_particles(rng, L, n, spread=None, vth=0.5):
    """particles uniformly spread over the patch grown by `spread` cells (default: particle ghost width)"""
    dim = L.dim
    pg = particle_ghosts(L.interp) if spread is None else spread
    icell = np.stack([rng.integers(L.amr_lower[d] - pg, L.amr_lower[d] + L.ncells[d] + pg, n)
                      for d in range(dim)], 1).astype(np.int32)
    delta = rng.random((n, dim))
    v = rng.standard_normal((n, 3)) * vth
    w = rng.random(n) + 0.1
    q = np.where(rng.random(n) < 0.5, 1.0, 2.0)
    return icell, delta, w, q, v


def sorted_particles(rng, L, ppc, vth=0.5):
    """ppc particles in every domain cell, in row-major cell order (what phb_bin produces)"""
    dim = L.dim
    nc = [L.ncells[d] for d in range(dim)]
    grids = np.meshgrid(*[np.arange(nc[d]) + L.amr_lower[d] for d in range(dim)], indexing="ij")
    cells = np.stack([g.ravel() for g in grids], 1)
    icell = np.repeat(cells, ppc, axis=0).astype(np.int32)
    n = len(icell)
    delta = rng.random((n, dim))
    v = rng.standard_normal((n, 3)) * vth
    w = rng.random(n) + 0.1
    q = np.ones(n)
    return icell, delta, w, q, v
