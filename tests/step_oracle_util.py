"""Helpers of the step-oracle tests: build the STEP ORACLE (oracle/ref_step.py: the reference's own functors, its own C++
level loop and box algebra) on the same inputs as a phare_b200 SolverPPC and compare the two states node by node."""
import numpy as np

from phare_b200 import abi
from oracle import canonical_rows
from oracle import ref_step as rs

PER_NODE_RTOL = 1e-10   # north_star: moments and fields <= 1e-10 relative after N steps (FP64 atomic reordering)
FLOOR = 1e-13           # absolute floor, in units of max|field|, for nodes where the field crosses zero


def oracle_for(solver, gparts, masses, Te, eta, nu, hyper_mode=0, B_arrays=None):
    """RefStep with the patches, particles and B of `solver` as they were BEFORE its initialize() changed anything
    (initialize() derives moments, J and E and leaves B and the particles alone)"""
    ops = solver.ops
    p0 = solver.patches[0]
    dim, interp = p0.layout.dim, p0.layout.interp
    boxes = [([p.layout.amr_lower[d] for d in range(dim)],
              [p.layout.amr_lower[d] + p.layout.ncells[d] - 1 for d in range(dim)]) for p in solver.patches]
    dx = [p0.layout.dx[d] for d in range(dim)]
    domain_cells = list(solver.geom.domain_cells) if hasattr(solver.geom, "domain_cells") else None
    if domain_cells is None:
        hi = np.max([b[1] for b in boxes], axis=0)
        domain_cells = [int(h) + 1 for h in hi]
    ref = rs.RefStep(dim, interp, boxes, dx, [0.0] * dim, domain_cells, masses, Te=Te, eta=eta, nu=nu, hyper_mode=hyper_mode)
    for ip, p in enumerate(solver.patches):
        ref.set_vec(ip, rs.B, B_arrays[ip] if B_arrays is not None else [ops.get_field(p.B[c]) for c in range(3)])
        for i, (icell, delta, w, q, v) in enumerate(gparts):
            inside = np.ones(len(w), bool)
            for d in range(dim):
                inside &= (icell[:, d] >= boxes[ip][0][d]) & (icell[:, d] <= boxes[ip][1][d])
            ref.set_particles(ip, i, icell[inside], delta[inside], w[inside], q[inside], v[inside])
    ref.initialize()
    return ref


def physical(a, layout, qty, centering_fn):
    """the physical nodes (no ghosts) of a field array"""
    g = 2 if layout.interp == 1 else 4
    sl = tuple(slice(g, g + layout.ncells[d] + (1 if centering_fn(qty, d) == 0 else 0)) for d in range(layout.dim))
    return a[sl]


def node_errors(got, want):
    """largest |got - want| / (|want| PER NODE + FLOOR * max|want|)"""
    scale = np.max(np.abs(want))
    if scale == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - want) / (np.abs(want) + (FLOOR / PER_NODE_RTOL) * scale)))


def compare_fields(solver, ref, ghosts=True):
    """every field and moment of every patch, physical nodes per node; with ghosts=True also the ghost nodes of E and B
    (what the next push gathers from).  Returns {name: worst per-node error in units of PER_NODE_RTOL-relative}"""
    from phare_b200.messenger import centering
    ops = solver.ops
    worst = {}

    def check(name, got, want, layout, qty, whole):
        a = got if whole else physical(got, layout, qty, centering)
        b = want if whole else physical(want, layout, qty, centering)
        assert not np.isnan(b).any(), name
        worst[name] = max(worst.get(name, 0.0), node_errors(a, b))

    for ip, p in enumerate(solver.patches):
        L = p.layout
        for which, attr, q0 in ((rs.B, "B", abi.BX), (rs.E, "E", abi.EX), (rs.J, "J", abi.JX), (rs.VI, "Vi", abi.VX)):
            want = ref.vec(ip, which)
            for c in range(3):
                whole = ghosts and attr in ("B", "E")
                check(f"{attr}{'xyz'[c]}", ops.get_field(getattr(p, attr)[c]), want[c], L, q0 + c, whole)
        check("Ni", ops.get_field(p.Ne), ref.scalar(ip, rs.NI), L, abi.RHO, False)
        check("rho_m", ops.get_field(p.rho_m), ref.scalar(ip, rs.RHO_M), L, abi.RHO, False)
        for i, pop in enumerate(p.pops):
            check(f"pop{i}.rho_n", ops.get_field(pop.rho_n), ref.scalar(ip, rs.POP_RHO_N + 2 * i), L, abi.RHO, False)
            check(f"pop{i}.rho_q", ops.get_field(pop.rho_q), ref.scalar(ip, rs.POP_RHO_Q + 2 * i), L, abi.RHO, False)
            want = ref.vec(ip, rs.POP_FLUX + i)
            for c in range(3):
                check(f"pop{i}.F{'xyz'[c]}", ops.get_field(pop.flux[c]), want[c], L, abi.VX + c, False)
    return worst


def compare_particles(solver, ref, rtol=1e-12):
    """per patch and population: counts exact, cells exact, delta / v <= rtol (north_star: positions and velocities
    <= 1e-12 relative per step; free-running they drift with the fields, so callers pass what N steps allow).
    Particles are matched through their (weight, charge, initial-order-independent) canonical sort by cell then by
    position.  Returns the worst relative differences (delta, v)."""
    ops = solver.ops
    worst_d = worst_v = 0.0
    for ip, p in enumerate(solver.patches):
        for i, pop in enumerate(p.pops):
            a = ops.get_particles(pop.domain)
            b = ref.particles(ip, i, 0)
            assert len(a[2]) == len(b[2]) == ref.count(ip, i, 0), (ip, i, len(a[2]), len(b[2]))
            assert ref.count(ip, i, 1) == 0  # patch ghosts are cleared by fillIonGhostParticles

            def order(s):
                icell, delta, w, q, v = s
                keys = [w, v[:, 2], v[:, 1], v[:, 0]] + [delta[:, d] for d in range(delta.shape[1])][::-1] \
                    + [icell[:, d] for d in range(icell.shape[1])][::-1]
                return np.lexsort(keys)
            oa, ob = order(a), order(b)
            assert np.array_equal(a[0][oa], b[0][ob]), f"patch {ip} pop {i}: cell indices differ"
            assert np.array_equal(a[2][oa], b[2][ob]) and np.array_equal(a[3][oa], b[3][ob])
            worst_d = max(worst_d, float(np.max(np.abs(a[1][oa] - b[1][ob]), initial=0.0)))
            vs = np.max(np.abs(b[4])) if len(b[2]) else 1.0
            worst_v = max(worst_v, float(np.max(np.abs(a[4][oa] - b[4][ob]), initial=0.0)) / vs)
    return worst_d, worst_v
