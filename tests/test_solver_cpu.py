"""Step orchestration (SolverPPC restatement + same-level messenger) on the CPU back end: the result of N
steps must not depend on how the periodic domain is cut into patches (up to FP reordering of the moment
sums), neighbouring patches must agree bit-for-bit on shared nodes (the reference's overlap-coherence tests,
tests/simulator/test_advance.py:27-190), and the L0 particle number is conserved
(tests/simulator/advance/test_advance_hybrid.py:263-289)."""
import numpy as np
import pytest

from phare_b200 import abi
from oracle.cpu_ops import CpuOps
from solver_util import global_particles, make_solver, gather_field, all_particles, FIELDS

CASES = [
    # domain, patch grid, interp, dx, ppc, pops, steps
    ((48,), (4,), 1, (0.2,), 20, 1, 5),
    ((48,), (3,), 2, (0.25,), 20, 2, 4),
    ((40,), (2,), 3, (0.2,), 20, 1, 4),
    ((16, 12), (2, 2), 1, (0.4, 0.4), 8, 2, 3),
    ((16, 16), (2, 1), 3, (0.2, 0.2), 6, 1, 2),
    ((8, 8, 8), (2, 1, 2), 1, (0.2, 0.2, 0.2), 4, 1, 2),
]


@pytest.mark.parametrize("domain,grid,interp,dx,ppc,npop,steps", CASES)
def test_result_independent_of_patch_decomposition(domain, grid, interp, dx, ppc, npop, steps):
    dim = len(domain)
    gparts = global_particles(domain, interp, dx, ppc, seed=3, pops=npop)
    ntot = [len(g[2]) for g in gparts]
    one = make_solver(CpuOps(dim, interp), domain, (1,) * dim, interp, dx, gparts)
    many = make_solver(CpuOps(dim, interp), domain, grid, interp, dx, gparts)
    dt = 0.005
    for s in range(steps):
        one.advance_level(dt)
        many.advance_level(dt)
    for attr, comp, qty in FIELDS:
        a = gather_field(one, attr, comp, qty, domain)
        b = gather_field(many, attr, comp, qty, domain)
        assert not np.isnan(a).any() and not np.isnan(b).any()
        scale = np.max(np.abs(a)) + 1e-30
        assert np.max(np.abs(a - b)) <= 1e-10 * scale + 1e-13, (attr, comp, np.max(np.abs(a - b)), scale)
    for i in range(npop):
        pa, pb = all_particles(one, i), all_particles(many, i)
        assert len(pa[2]) == len(pb[2]) == ntot[i]  # L0 particle number conservation
        # same particles up to rounding: compare per-particle after sorting by (weight, charge, v0 at t=0 is lost)
        # -> use global position x = icell + delta, sorted
        xa = np.sort((pa[0][:, 0] % domain[0]) + pa[1][:, 0])
        xb = np.sort((pb[0][:, 0] % domain[0]) + pb[1][:, 0])
        assert np.max(np.abs(xa - xb)) < 1e-9


def test_domain_only_does_not_touch_particles():
    """IonUpdaterTest particlesUntouchedInMomentOnlyMode (tests/core/numerics/ion_updater/test_updater.cpp:744)"""
    from phare_b200.solver import DOMAIN_ONLY
    domain, interp, dx = (32,), 1, (0.2,)
    gparts = global_particles(domain, interp, dx, 10, seed=5)
    s = make_solver(CpuOps(1, interp), domain, (2,), interp, dx, gparts)
    before = [s.ops.get_particles(p.pops[0].domain) for p in s.patches]
    rho_before = [s.ops.get_field(p.pops[0].rho_n).copy() for p in s.patches]
    for p in s.patches:
        s.updater.update_populations(p, p.E, p.B, 0.01, DOMAIN_ONLY)
    after = [s.ops.get_particles(p.pops[0].domain) for p in s.patches]
    for b, a in zip(before, after):
        for x, y in zip(b, a):
            assert np.array_equal(x, y)
    # ... but the moments did change (momentsAreChangedInMomentsOnlyMode, :815)
    assert any(not np.array_equal(r, s.ops.get_field(p.pops[0].rho_n)) for r, p in zip(rho_before, s.patches))


def test_no_nan_on_physical_nodes_and_density_close_to_prescribed():
    """thatNoNaNsExistOnPhysicalNodesMoments (test_updater.cpp:855) + density within 0.07 of the profile (:835)"""
    domain, interp, dx = (64,), 1, (0.2,)
    gparts = global_particles(domain, interp, dx, 400, seed=11)
    s = make_solver(CpuOps(1, interp), domain, (2,), interp, dx, gparts)
    for _ in range(3):
        s.advance_level(0.005)
    ne = gather_field(s, "Ne", None, abi.RHO, domain)
    x = np.arange(domain[0] + 1) * dx[0]
    want = 1.0 + 0.2 * np.sin(2 * np.pi * x / (domain[0] * dx[0]))
    assert not np.isnan(ne).any()
    assert np.max(np.abs(ne - want)) < 0.12  # 400 ppc: noise ~ 1/sqrt(ppc*2)


def test_fused_and_two_pass_sweeps_are_the_same_orchestration():
    """IonUpdater(fused=True) issues one push_deposit per array and sweep, fused=False the reference's push then
    deposit; on the CPU checker both are the same arithmetic in the same order -> bit-identical state"""
    domain, grid, interp, dx = (16, 12), (2, 2), 1, (0.4, 0.4)
    gparts = global_particles(domain, interp, dx, 8, seed=21, pops=2)
    a = make_solver(CpuOps(2, interp), domain, grid, interp, dx, gparts, solver_kw=dict(fused=True))
    b = make_solver(CpuOps(2, interp), domain, grid, interp, dx, gparts, solver_kw=dict(fused=False))
    for _ in range(3):
        a.advance_level(0.01)
        b.advance_level(0.01)
    for attr, comp, qty in FIELDS:
        assert np.array_equal(gather_field(a, attr, comp, qty, domain), gather_field(b, attr, comp, qty, domain))
    for i in range(2):
        for x, y in zip(all_particles(a, i), all_particles(b, i)):
            assert np.array_equal(x, y)


def test_particle_stores_grow_before_they_overflow():
    """a population that fills its stores gets larger ones at the end of the step (std::vector growth in the reference)"""
    domain, interp, dx = (32,), 1, (0.2,)
    gparts = global_particles(domain, interp, dx, 10, seed=5)
    s = make_solver(CpuOps(1, interp), domain, (2,), interp, dx, gparts)
    pop = s.patches[0].pops[0]
    n, cap0 = s.ops.count(pop.domain), s.ops.capacity(pop.domain)
    s.GROW_AT = (n - 1) / cap0  # pretend the store is nearly full
    before = all_particles(s, 0)
    s.advance_level(0.005)
    assert s.ops.capacity(pop.domain) == s.ops.capacity(pop.spare) > cap0
    s.GROW_AT = 0.85
    s.advance_level(0.005)  # and the step goes on with the new stores
    assert len(all_particles(s, 0)[2]) == len(before[2])


def test_patch_without_particles_and_empty_population():
    """a population absent from one patch (which receives its first particles by migration) and a population that is
    empty everywhere, next to a normal one (the plasma density itself is never zero: Ohm divides by it): every sweep,
    binning and exchange phase must cope with zero-length ranges"""
    domain, interp, dx = (32,), 1, (0.2,)
    full, second = global_particles(domain, interp, dx, 20, seed=8, pops=2)
    left = second[0][:, 0] < 16
    gparts = [full, tuple(a[left] for a in second), tuple(a[:0] for a in full)]
    s = make_solver(CpuOps(1, interp), domain, (2,), interp, dx, gparts, masses=(1.0, 2.0, 1.0))
    assert s.ops.count(s.patches[1].pops[1].domain) == 0
    for _ in range(6):
        s.advance_level(0.02)
    n = [s.ops.count(p.pops[1].domain) for p in s.patches]
    assert sum(n) == int(left.sum()) and n[1] > 0          # conserved, and some have crossed into the patch without any
    assert sum(s.ops.count(p.pops[0].domain) for p in s.patches) == len(full[2])
    assert all(s.ops.count(p.pops[2].domain) == 0 for p in s.patches)
    for attr, comp, qty in FIELDS:
        assert not np.isnan(gather_field(s, attr, comp, qty, domain)).any()
