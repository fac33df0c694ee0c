"""Pins the C restatement (oracle/phare_oracle.c) against the reference's own code compiled in place
(oracle/_ref/libphare_ref.so <- /root/reference/src/core headers): bit-for-bit, every operator of the
hot path, all nine (dim, interp) pairs.  CPU only."""
import numpy as np
import pytest

from phare_b200 import abi
from oracle import HostParticles
from util import ALL_DIM_INTERP, bit_equal, small_layout, random_vec, random_particles, domain_box, grown, \
    particle_ghosts


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_push_gather_deposit_bitexact(cpu_oracle, cpu_ref, dim, interp):
    rng = np.random.default_rng(100 * dim + interp)
    L = small_layout(dim, interp)
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX)
    P = HostParticles.from_soa(*random_particles(rng, L, 3000))
    rc_o, po = cpu_oracle.push(L, E, B, P, 1.3, 0.05)
    rc_r, pr = cpu_ref.push(L, E, B, P, 1.3, 0.05)
    assert rc_o == 0 and rc_r == 0
    for a, b in zip(po.soa(), pr.soa()):
        assert bit_equal(a, b)
    assert bit_equal(cpu_oracle.gather(L, E, B, P), cpu_ref.gather(L, E, B, P))
    for a, b in zip(cpu_oracle.deposit(L, P, coef=0.7), cpu_ref.deposit(L, P, coef=0.7)):
        assert bit_equal(a, b)


@pytest.mark.parametrize("dim,interp", [(1, 1), (2, 2), (3, 3)])
def test_push_first_selector(cpu_oracle, cpu_ref, dim, interp):
    """first selector = inGhostBox (ion_updater.hpp:137-140): rejected particles keep the pre-pushed
    position and the old velocity.  The reference partitions (reorders) -> compare canonical rows."""
    from oracle import canonical_rows
    rng = np.random.default_rng(7)
    L = small_layout(dim, interp)
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX)
    P = HostParticles.from_soa(*random_particles(rng, L, 2000, vth=0.8))
    ghost = grown(domain_box(L), dim, particle_ghosts(interp))
    rc_o, po = cpu_oracle.push(L, E, B, P, 1.0, 0.1, first_selector=ghost)
    rc_r, pr = cpu_ref.push(L, E, B, P, 1.0, 0.1, first_selector=ghost)
    assert rc_o == 0 and rc_r == 0
    assert np.array_equal(canonical_rows(*po.soa()), canonical_rows(*pr.soa()))


def test_move_two_cells_is_an_error(cpu_oracle, cpu_ref):
    L = small_layout(1, 1)
    E = [np.zeros(cpu_oracle.field_shape(L, abi.EX + c)) for c in range(3)]
    B = [np.zeros(cpu_oracle.field_shape(L, abi.BX + c)) for c in range(3)]
    P = HostParticles.from_soa(np.array([[8]], np.int32), np.array([[0.5]]), np.ones(1), np.ones(1),
                               np.array([[100., 0, 0]]))
    rc_o, _ = cpu_oracle.push(L, E, B, P, 1.0, 0.1)
    rc_r, _ = cpu_ref.push(L, E, B, P, 1.0, 0.1)
    assert rc_o == abi.PHB_ERR_MOVE_TWO_CELL and rc_r == abi.PHB_ERR_MOVE_TWO_CELL


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_field_operators_bitexact(cpu_oracle, cpu_ref, dim, interp):
    o, r = cpu_oracle, cpu_ref
    rng = np.random.default_rng(200 * dim + interp)
    L = small_layout(dim, interp, level=2)
    E, B = random_vec(rng, o.field_shape, L, abi.EX), random_vec(rng, o.field_shape, L, abi.BX)
    J, Ve = random_vec(rng, o.field_shape, L, abi.JX), random_vec(rng, o.field_shape, L, abi.VX)
    n = rng.random(o.field_shape(L, abi.RHO)) + 0.5
    Pe = rng.random(o.field_shape(L, abi.P))
    for a, b in zip(o.faraday(L, B, E, 0.01), r.faraday(L, B, E, 0.01)):
        assert bit_equal(a, b)
    for a, b in zip(o.ampere(L, B), r.ampere(L, B)):
        assert bit_equal(a, b)
    for hyper_mode in (0, 1):
        for a, b in zip(o.ohm(L, n, Ve, Pe, B, J, 0.3, 0.02, hyper_mode), r.ohm(L, n, Ve, Pe, B, J, 0.3, 0.02, hyper_mode)):
            assert bit_equal(a, b)
    (ve_o, pe_o), (ve_r, pe_r) = o.electrons_update(L, n, Ve, J, 0.12), r.electrons_update(L, n, Ve, J, 0.12)
    assert bit_equal(pe_o, pe_r)
    for a, b in zip(ve_o, ve_r):
        assert bit_equal(a, b)
    rn = [rng.random(n.shape) + .1 for _ in range(2)]
    rq = [rng.random(n.shape) for _ in range(2)]
    fl = [random_vec(rng, o.field_shape, L, abi.VX) for _ in range(2)]
    to, tr = o.ions_totals(L, rn, rq, fl, [1.0, 4.0]), r.ions_totals(L, rn, rq, fl, [1.0, 4.0])
    assert bit_equal(to[0], tr[0]) and bit_equal(to[1], tr[1])
    for a, b in zip(to[2], tr[2]):
        assert bit_equal(a, b)
    assert bit_equal(o.average(n, Pe), r.average(n, Pe))


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_reference_maxwellian_initializer_feeds_both_sides(cpu_oracle, cpu_ref, dim):
    """MaxwellianParticleInitializer::loadParticles (maxwellian_particle_initializer.hpp:138-199), run unmodified
    through oracle/_ref with a fixed seed: identical initial particles for the oracle and the reference."""
    L = small_layout(dim, 1)
    ncell = int(np.prod([L.ncells[d] for d in range(dim)]))
    n = 1.0 + 0.5 * np.arange(ncell) / ncell
    zeros, vth = np.zeros(ncell), np.full(ncell, 0.3)
    ppc = 7
    P1 = cpu_ref.maxwellian(L, n, [zeros] * 3, [vth] * 3, 1.0, ppc, 1337)
    P2 = cpu_ref.maxwellian(L, n, [zeros] * 3, [vth] * 3, 1.0, ppc, 1337)
    assert P1.n == ncell * ppc
    for a, b in zip(P1.soa(), P2.soa()):
        assert bit_equal(a, b)  # seeded std::mt19937_64 stream is reproducible
    ic, de, w, q, v = P1.soa()
    assert np.all((de >= 0) & (de < 1)) and np.all(q == 1.0)
    assert np.allclose(w.reshape(ncell, ppc), (n / ppc)[:, None], rtol=0, atol=0)
    # row-major cell order over the patch box
    lo = np.array([L.amr_lower[d] for d in range(dim)])
    idx = np.stack(np.unravel_index(np.arange(ncell), [L.ncells[d] for d in range(dim)]), 1) + lo
    assert np.array_equal(ic.reshape(ncell, ppc, dim)[:, 0, :], idx)
    rng = np.random.default_rng(0)
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.1)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.1)
    _, po = cpu_oracle.push(L, E, B, P1, 1.0, 0.02)
    _, pr = cpu_ref.push(L, E, B, P1, 1.0, 0.02)
    for a, b in zip(po.soa(), pr.soa()):
        assert bit_equal(a, b)
