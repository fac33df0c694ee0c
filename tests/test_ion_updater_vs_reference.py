"""IonUpdater (phare_b200/solver.py) against the REFERENCE IonUpdater::updatePopulations + updateIons
(src/core/numerics/ion_updater/ion_updater.hpp:90-295, compiled in place into oracle/_ref) on a patch that has,
like the reference's own fixture (tests/core/numerics/ion_updater/test_updater.cpp:402-526), level-ghost particles on
its lower side and a same-level neighbour (patch-ghost layer) on its upper side, two populations.
Particle arrays are compared as multisets, bit for bit (the reference has no particle order); moments to 1e-12
on the CPU back end (different summation order) and 1e-10 on the GPU."""
import numpy as np
import pytest

from phare_b200 import abi
from phare_b200.boxes import Box
from phare_b200.messenger import PatchGeom
from phare_b200.solver import Patch, IonUpdater, DOMAIN_ONLY, ALL
from oracle import HostParticles, canonical_rows
from util import small_layout, random_vec, particle_ghosts

CASES = [(1, 1), (1, 2), (1, 3), (2, 1), (2, 3), (3, 1)]


def make_inputs(dim, interp, seed):
    rng = np.random.default_rng(seed)
    L = small_layout(dim, interp, ncells=[14, 10, 8])
    pg = particle_ghosts(interp)
    lo = np.array([L.amr_lower[d] for d in range(dim)])
    hi = lo + np.array([L.ncells[d] for d in range(dim)]) - 1
    # nonLevelGhostBox: domain + the ghost strip towards the upper-x neighbour (amr_utils.hpp:233-253)
    strip_lo, strip_hi = lo - pg, hi + pg
    strip_lo[0] = hi[0] + 1
    boxes = [abi.make_box(lo, hi), abi.make_box(strip_lo, strip_hi)]
    pops = []
    for ip, mass in enumerate((1.0, 4.0)):
        ppc = 12
        grids = np.meshgrid(*[np.arange(lo[d], hi[d] + 1) for d in range(dim)], indexing="ij")
        cells = np.repeat(np.stack([g.ravel() for g in grids], 1), ppc, axis=0).astype(np.int32)
        n = len(cells)
        dom = (cells, rng.random((n, dim)), rng.random(n) + .1, np.full(n, 1.0 + ip), rng.standard_normal((n, 3)))
        # level ghosts: the pg cells below the domain in x, all y/z of the ghost box
        lg_lo, lg_hi = lo - pg, hi + pg
        lg_hi[0] = lo[0] - 1
        g2 = np.meshgrid(*[np.arange(lg_lo[d], lg_hi[d] + 1) for d in range(dim)], indexing="ij")
        cl = np.repeat(np.stack([g.ravel() for g in g2], 1), ppc, axis=0).astype(np.int32)
        m = len(cl)
        vlg = rng.standard_normal((m, 3))
        vlg[:, 0] = np.abs(vlg[:, 0]) * 2  # many of them enter the domain
        lgp = (cl, rng.random((m, dim)), rng.random(m) + .1, np.full(m, 1.0 + ip), vlg)
        pops.append(dict(mass=mass, domain=dom, level_ghost=lgp))
    E = random_vec(rng, lambda L_, q: tuple(int(L_.ncells[d]) + (1 if _primal(q, d) else 0) + 2 * (2 if interp == 1 else 4)
                                            for d in range(dim)), L, abi.EX, 0.3)
    B = random_vec(rng, lambda L_, q: tuple(int(L_.ncells[d]) + (1 if _primal(q, d) else 0) + 2 * (2 if interp == 1 else 4)
                                            for d in range(dim)), L, abi.BX, 0.5)
    return L, boxes, pops, E, B


def _primal(q, d):
    from phare_b200.messenger import centering
    return centering(q, d) == 0


def run_reference(cpu_ref, L, boxes, pops, E, B, dt, mode):
    cap = lambda p: len(p["domain"][2]) + len(p["level_ghost"][2]) + 64
    dom = [HostParticles.from_soa(*p["domain"], capacity=cap(p)) for p in pops]
    pgh = [HostParticles(L.dim, cap(p)) for p in pops]
    lgh = [HostParticles.from_soa(*p["level_ghost"], capacity=cap(p)) for p in pops]
    res = cpu_ref.ion_update(L, E, B, [p["mass"] for p in pops], dom, pgh, lgh, boxes, dt, mode, update_ions=True)
    return res, dom, pgh, lgh


def run_ours(ops, L, boxes, pops, E, B, dt, mode):
    lo = [L.amr_lower[d] for d in range(L.dim)]
    hi = [L.amr_lower[d] + L.ncells[d] - 1 for d in range(L.dim)]
    spec = [dict(name=f"pop{i}", mass=p["mass"], n=len(p["domain"][2]) + len(p["level_ghost"][2])) for i, p in enumerate(pops)]
    patch = Patch(ops, PatchGeom(0, Box(lo, hi), 0), L, spec)
    patch.non_level_ghost = boxes
    for c in range(3):
        ops.set_field(patch.E[c], E[c])
        ops.set_field(patch.B[c], B[c])
    for pop, p in zip(patch.pops, pops):
        ops.set_particles(pop.spare, *p["domain"])
        counts = ops.bin(L, pop.spare, pop.domain, patch.domain_box, boxes, pop.cell_start)
        pop.n_sorted = counts[0]
        ops.set_count(pop.domain, counts[0])
        pop.set_level_ghosts(ops, len(p["level_ghost"][2]) + 16)
        ops.set_particles(pop.level_ghost, *p["level_ghost"])
    upd = IonUpdater(ops)
    upd.update_populations(patch, patch.E, patch.B, dt, mode)
    assert ops.poll_error() == 0
    upd.update_ions(patch)
    return patch


def compare(ops, patch, ref, tol):
    res, dom, pgh, lgh = ref
    for i, pop in enumerate(patch.pops):
        got = [ops.get_field(m) for m in pop.moments()]
        want = [res["rho_n"][i], res["rho_q"][i]] + res["flux"][i]
        for g, w in zip(got, want):
            assert np.max(np.abs(g - w)) <= tol * (np.max(np.abs(w)) + 1e-300)
        for mine, theirs in ((pop.domain, dom[i]), (pop.patch_ghost, pgh[i]), (pop.level_ghost, lgh[i])):
            a, b = ops.get_particles(mine), theirs.soa()
            assert len(a[2]) == len(b[2])
            assert np.array_equal(canonical_rows(*a), canonical_rows(*b))
    for g, w in ((patch.Ne, res["rho_q_tot"]), (patch.rho_m, res["rho_m_tot"])):
        assert np.max(np.abs(ops.get_field(g) - w)) <= tol * np.max(np.abs(w))
    for c in range(3):
        g, w = ops.get_field(patch.Vi[c]), res["V"][c]
        ok = np.isfinite(w)
        assert np.array_equal(np.isnan(g), np.isnan(w))  # 0/0 on untouched ghost nodes, exactly like the reference
        assert np.max(np.abs(g[ok] - w[ok])) <= tol * 10 * (np.max(np.abs(w[ok])) + 1e-300)


@pytest.mark.parametrize("dim,interp", CASES)
@pytest.mark.parametrize("mode", [DOMAIN_ONLY, ALL])
def test_cpu_backend_matches_reference_updater(cpu_ref, dim, interp, mode):
    from oracle.cpu_ops import CpuOps
    L, boxes, pops, E, B = make_inputs(dim, interp, 40 + dim * 3 + interp)
    ref = run_reference(cpu_ref, L, boxes, pops, E, B, 0.05, mode)
    ops = CpuOps(dim, interp)
    patch = run_ours(ops, L, boxes, pops, E, B, 0.05, mode)
    compare(ops, patch, ref, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,interp", CASES)
@pytest.mark.parametrize("mode", [DOMAIN_ONLY, ALL])
def test_gpu_matches_reference_updater(cpu_ref, dim, interp, mode):
    from phare_b200.solver import GpuOps
    L, boxes, pops, E, B = make_inputs(dim, interp, 40 + dim * 3 + interp)
    ref = run_reference(cpu_ref, L, boxes, pops, E, B, 0.05, mode)
    ops = GpuOps(dim, interp, "cuda:0")
    patch = run_ours(ops, L, boxes, pops, E, B, 0.05, mode)
    compare(ops, patch, ref, 1e-10)
