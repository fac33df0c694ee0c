"""Small refined hierarchies shared by the CPU and GPU tests of phare_b200.amr (same seeded particles and fields on
every back end)."""
import numpy as np

from phare_b200.amr import build_hierarchy
from solver_util import global_particles, B_init

SOLVER_KW = dict(resistivity=1e-3, hyper_resistivity=1e-3, Te=0.12)

# name: (domain cells, dx, root patch grid, interp, ppc, refinement boxes per level (index space of the level below))
CASES = {
    "1d_o1": ([64], [0.2], [2], 1, 20, [[([20], [39])]]),
    "1d_o2_td": ([80], [0.25], [1], 2, 16, [[([24], [55])]]),           # config 2: 1-D, order 2, one refined level
    "1d_o3_two_patches": ([64], [0.2], [1], 3, 12, [[([12], [27]), ([28], [45])]]),
    "1d_o1_three_levels": ([64], [0.2], [1], 1, 16, [[([16], [47])], [([48], [79])]]),
    "1d_o2_three_levels_two_root_patches": ([64], [0.2], [2], 2, 12, [[([14], [47])], [([40], [71])]]),
    "1d_o1_at_the_periodic_boundary": ([64], [0.2], [2], 1, 20, [[([0], [19])]]),
    "2d_o1": ([32, 24], [0.2, 0.25], [2, 1], 1, 8, [[([8, 6], [15, 13]), ([16, 6], [21, 13])]]),
    "2d_o2_L": ([28, 28], [0.25, 0.25], [1, 1], 2, 6, [[([8, 8], [19, 13]), ([8, 14], [13, 19])]]),
    "3d_o1": ([16, 14, 12], [0.2, 0.25, 0.3], [2, 1, 1], 1, 4, [[([4, 4, 4], [9, 8, 7])]]),
}


def B_init_nd(cells, dx):
    if len(cells) == 1:
        return B_init(cells, dx)
    Lx, Ly = cells[0] * dx[0], cells[1] * dx[1]
    kx, ky = 2 * np.pi / Lx, 2 * np.pi / Ly

    def fn(c, x, y, *z):
        zero = 0 * x + 0 * y + (0 * z[0] if z else 0)
        if c == 0:
            return 1.0 - 0.1 * ky * np.cos(kx * x) * np.sin(ky * y) + zero
        if c == 1:
            return 0.1 * kx * np.sin(kx * x) * np.cos(ky * y) + zero
        return 0.05 * np.cos(kx * x) + zero
    return fn


def make_hierarchy(ops, name, pops=1):
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    dim = len(cells)
    gparts = global_particles(cells, interp, dx, ppc, 7, pops=pops)
    spec = [dict(name=f"pop{i}", mass=(1.0, 2.0)[i]) for i in range(pops)]

    def particles_fn(i, L, pid):
        icell, delta, w, q, v = gparts[i]
        inside = np.ones(len(w), bool)
        for d in range(dim):
            inside &= (icell[:, d] >= L.amr_lower[d]) & (icell[:, d] < L.amr_lower[d] + L.ncells[d])
        return icell[inside], delta[inside], w[inside], q[inside], v[inside]

    return build_hierarchy(ops, cells, grid, interp, dx, spec, B_init_nd(cells, dx), particles_fn,
                           refinement_boxes=boxes, solver_kw=SOLVER_KW)


def level_fields(h):
    """{(level, patch id, name): host array} of the state every level ends a step with"""
    ops, out = h.ops, {}
    for lvl in h.levels:
        for p in lvl.solver.patches:
            for name, vec in (("B", p.B), ("E", p.E), ("Vi", p.Vi), ("J", p.J)):
                for c in range(3):
                    out[(lvl.number, p.geom.id, f"{name}{c}")] = ops.get_field(vec[c])
            out[(lvl.number, p.geom.id, "Ne")] = ops.get_field(p.Ne)
            for i, pop in enumerate(p.pops):
                out[(lvl.number, p.geom.id, f"rho{i}")] = ops.get_field(pop.rho_q)
    return out
