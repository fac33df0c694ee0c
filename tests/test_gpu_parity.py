"""GPU parity tests: every kernel, called through the C ABI (libphare_b200.so), against the CPU oracle on
the same seeded inputs.  Bars (BASELINE.json north_star): cell indices / counts / sort offsets bit-exact;
pushed delta and v bit-exact in the default (unfused) mode and <= 1e-12 relative in the FMA mode;
moments <= 1e-10 relative (FP64 atomic reordering); field solvers bit-exact."""
import ctypes as C

import numpy as np
import pytest

from phare_b200 import abi
from phare_b200.device import Context, DeviceArray, DeviceVec, DeviceParticles, PhbError
from oracle import HostParticles, canonical_rows
from util import (ALL_DIM_INTERP, CONFIG_DIM_INTERP, bit_equal, small_layout, random_vec, random_particles,
                  sorted_particles, domain_box, grown, particle_ghosts)

pytestmark = pytest.mark.gpu

MOMENT_RTOL = 1e-10  # north_star bound for moments (atomic reordering)


@pytest.fixture(scope="module")
def ctxs():
    cache = {}

    def get(dim, interp):
        if (dim, interp) not in cache:
            cache[(dim, interp)] = Context(dim, interp)
        return cache[(dim, interp)]

    yield get
    for c in cache.values():
        c.close()


def dev_particles(ctx, soa, capacity=None):
    p = DeviceParticles(ctx, capacity or max(len(soa[2]), 1))
    return p.upload_soa(*soa)


def assert_moments_close(got, want):
    for g, w in zip(got, want):
        scale = np.max(np.abs(w)) + 1e-300
        assert np.max(np.abs(g - w)) <= MOMENT_RTOL * scale, (np.max(np.abs(g - w)), scale)


# ------------------------------------------------------------------------------------------- K1 push
@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_push_bitexact(ctxs, cpu_oracle, dim, interp):
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(100 * dim + interp)
    L = small_layout(dim, interp)
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX)
    soa = random_particles(rng, L, 5000)
    rc, want = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.3, 0.05)
    assert rc == 0
    dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
    # out of place
    pin, pout = dev_particles(ctx, soa), DeviceParticles(ctx, len(soa[2]))
    ctx.push(L, dE, dB, pin, pout, 1.3, 0.05)
    ctx.poll_error()
    for g, w in zip(pout.download_soa(), want.soa()):
        assert bit_equal(g, w)
    # in place
    ctx.push(L, dE, dB, pin, pin, 1.3, 0.05)
    for g, w in zip(pin.download_soa(), want.soa()):
        assert bit_equal(g, w)


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_push_cell_ordered_store_bitexact(ctxs, cpu_oracle, dim, interp):
    """a cell-ordered store (what the step pushes): the tiles stage their E,B node box in shared memory; a few
    percent of the particles sit in a neighbouring cell (moved since the ordering) and widen the boxes, and the
    strided sample at the end (cells far apart in one tile) takes the global-gather path.  Bit-identical to the
    oracle either way."""
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(500 + 10 * dim + interp)
    L = small_layout(dim, interp)
    ppc = {1: 300, 2: 60, 3: 40}[dim]
    icell, delta, w, q, v = sorted_particles(rng, L, ppc, vth=1.0)
    n = len(w)
    movers = rng.random(n) < 0.04
    icell = icell.copy()
    icell[movers, rng.integers(0, dim)] += rng.choice([-1, 1])
    perm = np.concatenate([np.arange(n), np.arange(0, n, 7)])  # sorted part, then a scattered tail
    soa = tuple(a[perm] for a in (icell, delta, w, q, v))
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX)
    rc, want = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, 0.05)
    assert rc == 0 and len(soa[2]) >= 2048
    dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
    pin, pout = dev_particles(ctx, soa), DeviceParticles(ctx, len(soa[2]))
    ctx.push(L, dE, dB, pin, pout, 1.0, 0.05)
    ctx.poll_error()
    for g, x in zip(pout.download_soa(), want.soa()):
        assert bit_equal(g, x)
    ctx.push(L, dE, dB, pin, pin, 1.0, 0.05)
    for g, x in zip(pin.download_soa(), want.soa()):
        assert bit_equal(g, x)


@pytest.mark.parametrize("dim,interp", CONFIG_DIM_INTERP)
def test_push_fma_mode_within_1e12(ctxs, cpu_oracle, dim, interp):
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(5)
    L = small_layout(dim, interp)
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX)
    soa = random_particles(rng, L, 5000)
    _, want = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, 0.05)
    pin = dev_particles(ctx, soa)
    ctx.set_exact(False)
    try:
        ctx.push(L, DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B), pin, pin, 1.0, 0.05)
    finally:
        ctx.set_exact(True)
    gi, gd, _, _, gv = pin.download_soa()
    wi, wd, _, _, wv = want.soa()
    assert np.array_equal(gi, wi)  # cells still bit-exact on this input
    # <= 1e-12 relative to the particle's speed (a component that nearly cancels has no relative meaning)
    speed = np.linalg.norm(wv, axis=1, keepdims=True)
    assert np.max(np.abs(gv - wv) / speed) <= 1e-12
    assert np.max(np.abs(gd - wd)) <= 1e-12


@pytest.mark.parametrize("dim,interp", [(1, 1), (2, 2), (3, 3)])
def test_push_first_selector(ctxs, cpu_oracle, dim, interp):
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(7)
    L = small_layout(dim, interp)
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX)
    soa = random_particles(rng, L, 3000, vth=0.8)
    ghost = grown(domain_box(L), dim, particle_ghosts(interp))
    _, want = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, 0.1, first_selector=ghost)
    pin = dev_particles(ctx, soa)
    ctx.push(L, DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B), pin, pin, 1.0, 0.1, first_selector=ghost)
    for g, w in zip(pin.download_soa(), want.soa()):
        assert bit_equal(g, w)


def test_push_move_two_cells_raises(ctxs):
    """boris.hpp:164-165: |delta| > 2 after a half push is an error carrying delta/vel"""
    ctx = ctxs(1, 1)
    L = small_layout(1, 1)
    dE, dB = DeviceVec(ctx, L, abi.EX), DeviceVec(ctx, L, abi.BX)
    p = dev_particles(ctx, (np.array([[8]], np.int32), np.array([[0.5]]), np.ones(1), np.ones(1),
                            np.array([[100., 0, 0]])))
    ctx.push(L, dE, dB, p, p, 1.0, 0.1)
    with pytest.raises(PhbError) as e:
        ctx.poll_error()
    assert e.value.code == abi.PHB_ERR_MOVE_TWO_CELL and "25.5/100" in str(e.value)
    ctx.poll_error()  # flag is cleared by the poll


def test_push_empty(ctxs):
    ctx = ctxs(2, 1)
    L = small_layout(2, 1)
    p = DeviceParticles(ctx, 4)
    ctx.push(L, DeviceVec(ctx, L, abi.EX), DeviceVec(ctx, L, abi.BX), p, p, 1.0, 0.1)
    assert p.n == 0


# ------------------------------------------------------------------------------------------- K3 deposit
def gpu_moments(ctx, L):
    return [DeviceArray(ctx, ctx.field_shape(L, abi.RHO)) for _ in range(2)], DeviceVec(ctx, L, abi.VX)


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_deposit_unordered(ctxs, cpu_oracle, dim, interp):
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(300 * dim + interp)
    L = small_layout(dim, interp)
    soa = random_particles(rng, L, 4000)
    sel = [grown(domain_box(L), dim, 0)]
    for boxes in ([], sel):
        want = cpu_oracle.deposit(L, HostParticles.from_soa(*soa), coef=0.7, sel=boxes)
        (rn, rq), F = gpu_moments(ctx, L)
        ctx.deposit(L, dev_particles(ctx, soa), rn, rq, F, coef=0.7, sel=boxes)
        assert_moments_close([rn.download(), rq.download()] + F.download(), want)


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
@pytest.mark.parametrize("ppc", [3, 10, 40, 130])
def test_deposit_cell_ordered(ctxs, cpu_oracle, dim, interp, ppc):
    """cell-ordered kernel on a binned store, with ~3% of the particles displaced to a neighbouring
    cell after the ordering (they must take the atomic path and still land on the right nodes)"""
    if dim == 3 and ppc > 40:
        pytest.skip("kept small")
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(17 * dim + interp + ppc)
    L = small_layout(dim, interp)
    icell, delta, w, q, v = sorted_particles(rng, L, ppc)
    n = len(w)
    dom = domain_box(L)
    nd = int(np.prod([L.ncells[d] for d in range(dim)]))
    cell_start = np.zeros(ctx.bin_nkeys(L, dom) + 1, np.uint32)
    cell_start[:nd + 1] = np.arange(nd + 1) * ppc
    cell_start[nd + 1:] = n
    movers = rng.random(n) < 0.03
    icell = icell.copy()
    icell[movers, rng.integers(0, dim)] += rng.choice([-1, 1])
    soa = (icell, delta, w, q, v)
    keep = [grown(dom, dim, particle_ghosts(interp))]
    want = cpu_oracle.deposit(L, HostParticles.from_soa(*soa), coef=1.0, sel=keep)
    (rn, rq), F = gpu_moments(ctx, L)
    cs = DeviceArray(ctx, cell_start.shape, np.uint32).upload(cell_start)
    ctx.deposit(L, dev_particles(ctx, soa), rn, rq, F, coef=1.0, sel=keep, domain=dom, cell_start=cs)
    assert_moments_close([rn.download(), rq.download()] + F.download(), want)


def test_deposit_known_answer_1d(ctxs):
    """tests/core/numerics/interpolator/test_main.cpp:492-677: hand-built particle sets whose deposit at
    node 25 is exactly rho_n = 1 (weight-sum), rho_q = 2 (charge 2... here charge*weight), F = (2,-1,1)"""
    for interp in (1, 2, 3):
        ctx = ctxs(1, interp)
        L = abi.make_layout(1, interp, [50], [0.1])
        g = 2 if interp == 1 else 4
        # particles symmetric around node 25 (local index 25+g): cells 24 / 25 with mirrored deltas
        deltas = {1: [(24, 0.5), (25, 0.5)], 2: [(24, 0.0), (24, 0.5), (25, 0.0), (25, 0.5)][:3],
                  3: [(23, 0.5), (24, 0.5), (25, 0.5), (26, 0.5)]}[interp]
        ic = np.array([[c] for c, _ in deltas], np.int32)
        de = np.array([[d] for _, d in deltas])
        n = len(deltas)
        soa = (ic, de, np.full(n, 0.25), np.full(n, 2.0), np.tile([2., -1., 1.], (n, 1)))
        (rn, rq), F = gpu_moments(ctx, L)
        ctx.deposit(L, dev_particles(ctx, soa), rn, rq, F)
        rho = rn.download()
        # total deposited number density = sum of weights (B-spline partition of unity)
        assert abs(rho.sum() - 0.25 * n) < 1e-14
        assert np.allclose(rq.download(), 2.0 * rho, rtol=0, atol=1e-14)
        fx, fy, fz = F.download()
        assert np.allclose(fx, 2.0 * rho, atol=1e-14) and np.allclose(fy, -rho, atol=1e-14) and np.allclose(fz, rho, atol=1e-14)


# ------------------------------------------------------------------------------------------- K1+K3 fused
def _ordered_store(rng, L, ppc, vth):
    icell, delta, w, q, v = sorted_particles(rng, L, ppc, vth=vth)
    q = np.where(rng.random(len(w)) < 0.5, 1.0, 2.0)
    nd = int(np.prod([L.ncells[d] for d in range(L.dim)]))
    return (icell, delta, w, q, v), nd


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
@pytest.mark.parametrize("write_back", [False, True])
def test_push_deposit_cell_ordered(ctxs, cpu_oracle, dim, interp, write_back):
    """phb_push_deposit on a binned store == oracle push followed by oracle deposit of the pushed particles
    restricted to the keep boxes; the stored particles are bit-identical to the oracle push (write_back) or
    untouched (domain_only).  dt is large enough for ~20% of the particles to change cell."""
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(41 * dim + interp)
    L = small_layout(dim, interp)
    ppc = 30 if dim < 3 else 12
    soa, nd = _ordered_store(rng, L, ppc, vth=1.0)
    n = len(soa[2])
    dom = domain_box(L)
    cell_start = np.zeros(ctx.bin_nkeys(L, dom) + 1, np.uint32)
    cell_start[:nd + 1] = np.arange(nd + 1) * ppc
    cell_start[nd + 1:] = n
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    keep_lo = [dom.lower[d] - particle_ghosts(interp) for d in range(dim)]
    keep_lo[0] = dom.lower[0]  # one face is not a keep box (as at a coarse-fine boundary)
    keep = [abi.make_box(keep_lo, [dom.upper[d] + particle_ghosts(interp) for d in range(dim)])]
    rc, pushed = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, 0.08)
    assert rc == 0
    assert 0.05 < np.mean(np.any(pushed.soa()[0] != soa[0], axis=1)) < 0.9
    want = cpu_oracle.deposit(L, pushed, coef=0.9, sel=keep)
    (rn, rq), F = gpu_moments(ctx, L)
    parts = dev_particles(ctx, soa)
    cs = DeviceArray(ctx, cell_start.shape, np.uint32).upload(cell_start)
    ctx.push_deposit(L, DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B), parts, 1.0, 0.08, rn, rq, F, 0.9,
                     sel=keep, domain=dom, cell_start=cs, write_back=write_back)
    ctx.poll_error()
    assert_moments_close([rn.download(), rq.download()] + F.download(), want)
    for g, w in zip(parts.download_soa(), pushed.soa() if write_back else soa):
        assert bit_equal(g, w)


@pytest.mark.parametrize("dim,interp", [(1, 1), (2, 2), (3, 1), (3, 3)])
@pytest.mark.parametrize("write_back", [False, True])
def test_push_deposit_unordered_with_first_selector(ctxs, cpu_oracle, dim, interp, write_back):
    """the per-particle fused kernel: any order, sub-range [first,last), first selector = ghost box and
    deposit restricted to the domain box (the level-ghost sweep of ion_updater.hpp:195-217, 275-288)"""
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(77 + dim + interp)
    L = small_layout(dim, interp)
    soa = random_particles(rng, L, 4000, spread=particle_ghosts(interp) + 1, vth=0.8)
    dom = domain_box(L)
    ghost = grown(dom, dim, particle_ghosts(interp))
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    _, pushed = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 2.0, 0.1, first_selector=ghost)
    sub = tuple(a[100:3900] for a in pushed.soa())
    want = cpu_oracle.deposit(L, HostParticles.from_soa(*sub), coef=1.0, sel=[dom])
    (rn, rq), F = gpu_moments(ctx, L)
    parts = dev_particles(ctx, soa)
    ctx.push_deposit(L, DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B), parts, 2.0, 0.1, rn, rq, F, 1.0,
                     first=100, last=3900, sel=[dom], first_selector=ghost, write_back=write_back)
    assert_moments_close([rn.download(), rq.download()] + F.download(), want)
    got = parts.download_soa()
    for g, w, o in zip(got, pushed.soa(), soa):
        assert bit_equal(g[:100], o[:100]) and bit_equal(g[3900:], o[3900:])
        assert bit_equal(g[100:3900], w[100:3900] if write_back else o[100:3900])


def test_push_deposit_mover_record_overflow(ctxs, cpu_oracle):
    """more cell-changing particles than the record buffer holds (n/16 + 65536): the overflow is scattered
    inside the kernel and nothing is lost"""
    ctx = ctxs(1, 1)
    rng = np.random.default_rng(8)
    L = small_layout(1, 1, ncells=[4096])
    ppc = 300
    soa, nd = _ordered_store(rng, L, ppc, vth=1.2)
    n = len(soa[2])
    dom = domain_box(L)
    cell_start = np.zeros(ctx.bin_nkeys(L, dom) + 1, np.uint32)
    cell_start[:nd + 1] = np.arange(nd + 1) * ppc
    cell_start[nd + 1:] = n
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.1)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.1)
    rc, pushed = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, 0.05)
    assert rc == 0
    movers = int(np.sum(pushed.soa()[0] != soa[0]))
    assert movers > n // 16 + 65536 + 1000
    keep = [grown(dom, 1, 1)]
    want = cpu_oracle.deposit(L, pushed, coef=1.0, sel=keep)
    (rn, rq), F = gpu_moments(ctx, L)
    parts = dev_particles(ctx, soa)
    cs = DeviceArray(ctx, cell_start.shape, np.uint32).upload(cell_start)
    ctx.push_deposit(L, DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B), parts, 1.0, 0.05, rn, rq, F, 1.0,
                     sel=keep, domain=dom, cell_start=cs, write_back=False)
    assert_moments_close([rn.download(), rq.download()] + F.download(), want)


def test_push_deposit_fma_mode_and_error(ctxs, cpu_oracle):
    ctx = ctxs(2, 1)
    rng = np.random.default_rng(12)
    L = small_layout(2, 1)
    soa = random_particles(rng, L, 3000, vth=0.3)
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    _, pushed = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, 0.05)
    want = cpu_oracle.deposit(L, pushed, coef=1.0, sel=[])
    dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
    (rn, rq), F = gpu_moments(ctx, L)
    parts = dev_particles(ctx, soa)
    ctx.set_exact(False)
    try:
        ctx.push_deposit(L, dE, dB, parts, 1.0, 0.05, rn, rq, F, 1.0, write_back=True)
    finally:
        ctx.set_exact(True)
    assert_moments_close([rn.download(), rq.download()] + F.download(), want)
    gv, wv = parts.download_soa()[4], pushed.soa()[4]
    assert np.max(np.abs(gv - wv) / np.linalg.norm(wv, axis=1, keepdims=True)) <= 1e-12
    # a particle that would move more than two cells raises through the poll, like phb_push
    fast = (np.array([[8, 0]], np.int32), np.array([[0.5, 0.5]]), np.ones(1), np.ones(1), np.array([[100., 0, 0]]))
    ctx.push_deposit(L, dE, dB, dev_particles(ctx, fast), 1.0, 0.1, rn, rq, F, 1.0, write_back=False)
    with pytest.raises(PhbError) as e:
        ctx.poll_error()
    assert e.value.code == abi.PHB_ERR_MOVE_TWO_CELL


# ------------------------------------------------------------------------------------------- K2 bin / export
@pytest.mark.parametrize("dim,interp", CONFIG_DIM_INTERP + [(3, 3)])
def test_bin_counts_offsets_and_multisets(ctxs, cpu_oracle, dim, interp):
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(900 + dim * 10 + interp)
    L = small_layout(dim, interp)
    dom = domain_box(L)
    pg = particle_ghosts(interp)
    # particles up to one cell beyond the ghost box (they must fall in the overflow bin)
    soa = random_particles(rng, L, 20000, spread=pg + 1)
    # keep boxes: the ghost box minus one face (as if that side were a coarse/fine boundary)
    G = grown(dom, dim, pg)
    keep_lo = [G.lower[d] for d in range(dim)]
    keep_lo[0] = dom.lower[0]
    keep = [abi.make_box(keep_lo, [G.upper[d] for d in range(dim)])]
    want, want_cs, want_counts = cpu_oracle.bin(L, HostParticles.from_soa(*soa), dom, keep)
    pin, pout = dev_particles(ctx, soa), DeviceParticles(ctx, len(soa[2]))
    cs = DeviceArray(ctx, (ctx.bin_nkeys(L, dom) + 1,), np.uint32)
    counts = ctx.bin(L, pin, pout, dom, keep, cs)
    assert counts == want_counts and sum(counts) == len(soa[2])
    got_cs = cs.download()
    assert np.array_equal(got_cs, want_cs)
    assert pout.n == counts[0] + counts[1] == want.n
    # same multiset per key: canonicalise inside every cell and compare bit for bit
    gi, gd, gw, gq, gv = pout.download_soa()
    wi, wd, ww, wq, wv = want.soa()
    assert np.array_equal(gi, wi)  # cell order is the key order on both sides
    assert np.array_equal(canonical_rows(gi, gd, gw, gq, gv), canonical_rows(wi, wd, ww, wq, wv))
    # idempotence: binning a binned store leaves offsets unchanged
    pout2 = DeviceParticles(ctx, len(soa[2]))
    counts2 = ctx.bin(L, pout, pout2, dom, keep, cs)
    assert counts2 == (counts[0], counts[1], 0)
    assert np.array_equal(cs.download()[:-1], got_cs[:-1])


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_deposit_scatter_equals_deposit_then_bin(ctxs, cpu_oracle, dim, interp):
    """K3+K2 fused: phb_bin_plan + phb_deposit_scatter + phb_bin_counts on a store whose first part is cell-ordered
    (with ~10% of the particles displaced, some out of the patch and out of the ghost box) and whose tail is
    unordered == oracle deposit of the whole store followed by the oracle bin."""
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(700 + 10 * dim + interp)
    L = small_layout(dim, interp)
    ppc = {1: 40, 2: 30, 3: 12}[dim]
    icell, delta, w, q, v = sorted_particles(rng, L, ppc, vth=0.5)
    q = np.where(rng.random(len(w)) < 0.5, 1.0, 2.0)
    n_sorted = len(w)
    dom = domain_box(L)
    pg = particle_ghosts(interp)
    nd = int(np.prod([L.ncells[d] for d in range(dim)]))
    old_start = np.zeros(ctx.bin_nkeys(L, dom) + 1, np.uint32)
    old_start[:nd + 1] = np.arange(nd + 1) * ppc
    old_start[nd + 1:] = n_sorted
    movers = rng.random(n_sorted) < 0.1
    icell = icell.copy()
    icell[movers] += rng.integers(-pg - 1, pg + 2, size=(int(movers.sum()), dim)).astype(np.int32)
    tail = random_particles(rng, L, 700, spread=pg + 1)
    soa = tuple(np.concatenate([a, b]) for a, b in zip((icell, delta, w, q, v), tail))
    G = grown(dom, dim, pg)
    keep_lo = [G.lower[d] for d in range(dim)]
    keep_lo[0] = dom.lower[0]
    keep = [abi.make_box(keep_lo, [G.upper[d] for d in range(dim)])]
    src = HostParticles.from_soa(*soa)
    want_m = cpu_oracle.deposit(L, src, coef=1.0, sel=keep)
    want, want_cs, want_counts = cpu_oracle.bin(L, src, dom, keep)
    pin, pout = dev_particles(ctx, soa), DeviceParticles(ctx, len(soa[2]))
    (rn, rq), F = gpu_moments(ctx, L)
    cs_old = DeviceArray(ctx, old_start.shape, np.uint32).upload(old_start)
    cs_new = DeviceArray(ctx, old_start.shape, np.uint32)
    ctx.bin_plan(L, pin, dom, keep, cs_new)
    ctx.deposit_scatter(L, pin, n_sorted, rn, rq, F, 1.0, keep, dom, cs_old, keep, pout, cs_new)
    counts = ctx.bin_counts(L, dom, cs_new, pout)
    assert counts == want_counts and pout.n == counts[0] + counts[1]
    assert np.array_equal(cs_new.download(), want_cs)
    assert_moments_close([rn.download(), rq.download()] + F.download(), want_m)
    got = pout.download_soa()
    assert np.array_equal(got[0], want.soa()[0])
    assert np.array_equal(canonical_rows(*got), canonical_rows(*want.soa()))
    for g, x in zip(pin.download_soa(), soa):  # the source store is only read
        assert bit_equal(g, x)
    # a scatter without its plan is refused
    with pytest.raises(PhbError):
        ctx.deposit_scatter(L, pin, n_sorted, rn, rq, F, 1.0, keep, dom, cs_old, keep, pout, cs_new)


def test_bin_empty_and_single(ctxs):
    ctx = ctxs(1, 1)
    L = small_layout(1, 1)
    dom = domain_box(L)
    cs = DeviceArray(ctx, (ctx.bin_nkeys(L, dom) + 1,), np.uint32)
    pin, pout = DeviceParticles(ctx, 4), DeviceParticles(ctx, 4)
    assert ctx.bin(L, pin, pout, dom, [grown(dom, 1, 1)], cs) == (0, 0, 0)
    assert not cs.download().any()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_export_with_shift(ctxs, cpu_oracle, dim):
    ctx = ctxs(dim, 1)
    rng = np.random.default_rng(44 + dim)
    L = small_layout(dim, 1)
    soa = random_particles(rng, L, 6000)
    dom = domain_box(L)
    box = abi.make_box([dom.lower[d] - 1 for d in range(dim)], [dom.lower[d] + 2 for d in range(dim)])
    shift = [int(L.ncells[d]) for d in range(dim)]
    src = HostParticles.from_soa(*soa)
    dst = HostParticles(dim, 6000 + 10)
    pre = random_particles(rng, L, 10)
    dst = HostParticles.from_soa(*pre, capacity=6010)
    n_want = cpu_oracle.export(L, src, 100, 5900, box, dst, minus=dom, shift=shift)
    d_dst = dev_particles(ctx, pre, capacity=6010)
    n_got = ctx.export(L, dev_particles(ctx, soa), 100, 5900, box, d_dst, minus=dom, shift=shift)
    assert n_got == n_want > 0 and d_dst.n == dst.n
    for g, w in zip(d_dst.download_soa(), dst.soa()):
        assert bit_equal(g, w)  # export preserves source order


def test_particles_pack_unpack_roundtrip(ctxs):
    """the migration message: two store ranges packed into one column-major buffer and appended to a non-empty store"""
    for dim in (1, 2, 3):
        ctx = ctxs(dim, 1)
        rng = np.random.default_rng(dim)
        L = small_layout(dim, 1)
        a, b, pre = (random_particles(rng, L, n) for n in (301, 77, 10))
        A, B = dev_particles(ctx, a), dev_particles(ctx, b)
        total = 200 + 77
        nbytes = ctx.lib.phb_particles_flat_bytes(dim, total)
        assert nbytes == total * (12 * dim + 40)
        buf = DeviceArray(ctx, (nbytes,), np.uint8)
        ctx._check(ctx.lib.phb_particles_pack(ctx.h, C.byref(A.c), 50, 200, buf.ptr, total, 0))
        ctx._check(ctx.lib.phb_particles_pack(ctx.h, C.byref(B.c), 0, 77, buf.ptr, total, 200))
        dst = dev_particles(ctx, pre, capacity=10 + total)
        ctx._check(ctx.lib.phb_particles_unpack(ctx.h, buf.ptr, total, 200, 77, C.byref(dst.c)))
        ctx._check(ctx.lib.phb_particles_unpack(ctx.h, buf.ptr, total, 0, 200, C.byref(dst.c)))
        assert dst.n == 10 + total
        for g, p0, x, y in zip(dst.download_soa(), pre, a, b):
            assert bit_equal(g, np.concatenate([p0, y, x[50:250]]))
        with pytest.raises(PhbError):
            ctx._check(ctx.lib.phb_particles_unpack(ctx.h, buf.ptr, total, 0, 200, C.byref(dst.c)))  # capacity


def test_export_capacity_error(ctxs):
    ctx = ctxs(1, 1)
    L = small_layout(1, 1)
    rng = np.random.default_rng(1)
    src = dev_particles(ctx, random_particles(rng, L, 100))
    dst = DeviceParticles(ctx, 3)
    with pytest.raises(PhbError) as e:
        ctx.export(L, src, 0, 100, grown(domain_box(L), 1, 1), dst)
    assert e.value.code == abi.PHB_ERR_CAPACITY


# ------------------------------------------------------------------------------------------- K4-K7 fields
@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_field_operators_bitexact(ctxs, cpu_oracle, dim, interp):
    ctx, o = ctxs(dim, interp), cpu_oracle
    rng = np.random.default_rng(200 * dim + interp)
    L = small_layout(dim, interp, level=2)
    E, B = random_vec(rng, o.field_shape, L, abi.EX), random_vec(rng, o.field_shape, L, abi.BX)
    J, Ve = random_vec(rng, o.field_shape, L, abi.JX), random_vec(rng, o.field_shape, L, abi.VX)
    n = rng.random(o.field_shape(L, abi.RHO)) + 0.5
    Pe = rng.random(o.field_shape(L, abi.P))
    dE, dB, dJ, dVe = (DeviceVec(ctx, L, q, h) for q, h in ((abi.EX, E), (abi.BX, B), (abi.JX, J), (abi.VX, Ve)))
    dn = DeviceArray(ctx, n.shape).upload(n)
    dPe = DeviceArray(ctx, Pe.shape).upload(Pe)

    out = DeviceVec(ctx, L, abi.BX)
    ctx.faraday(L, dB, dE, out, 0.01)
    for g, w in zip(out.download(), o.faraday(L, B, E, 0.01)):
        assert bit_equal(g, w)

    out = DeviceVec(ctx, L, abi.JX)
    ctx.ampere(L, dB, out)
    for g, w in zip(out.download(), o.ampere(L, B)):
        assert bit_equal(g, w)

    for hyper_mode in (0, 1):
        out = DeviceVec(ctx, L, abi.EX)
        ctx.ohm(L, dn, dVe, dPe, dB, dJ, out, 0.3, 0.02, hyper_mode)
        for g, w in zip(out.download(), o.ohm(L, n, Ve, Pe, B, J, 0.3, 0.02, hyper_mode)):
            assert bit_equal(g, w)

    oVe, oPe = DeviceVec(ctx, L, abi.VX), DeviceArray(ctx, Pe.shape)
    ctx.electrons_update(L, dn, dVe, dJ, 0.12, oVe, oPe)
    wVe, wPe = o.electrons_update(L, n, Ve, J, 0.12)
    assert bit_equal(oPe.download(), wPe)
    for g, w in zip(oVe.download(), wVe):
        assert bit_equal(g, w)

    rn = [rng.random(n.shape) + .1 for _ in range(2)]
    rq = [rng.random(n.shape) for _ in range(2)]
    fl = [random_vec(rng, o.field_shape, L, abi.VX) for _ in range(2)]
    d_rn = [DeviceArray(ctx, a.shape).upload(a) for a in rn]
    d_rq = [DeviceArray(ctx, a.shape).upload(a) for a in rq]
    d_fl = [DeviceVec(ctx, L, abi.VX, f) for f in fl]
    tq, tm, tV = DeviceArray(ctx, n.shape), DeviceArray(ctx, n.shape), DeviceVec(ctx, L, abi.VX)
    ctx.ions_totals(d_rn, d_rq, d_fl, [1.0, 4.0], tq, tm, tV)
    wq, wm, wV = o.ions_totals(L, rn, rq, fl, [1.0, 4.0])
    assert bit_equal(tq.download(), wq) and bit_equal(tm.download(), wm)
    for g, w in zip(tV.download(), wV):
        assert bit_equal(g, w)

    avg = DeviceArray(ctx, n.shape)
    ctx.average(dn, dPe, avg)
    assert bit_equal(avg.download(), o.average(n, Pe))


def test_faraday_unusable_field_is_an_error(ctxs):
    """faraday.hpp:30-31"""
    ctx = ctxs(1, 1)
    L = small_layout(1, 1)
    B, E, Bn = DeviceVec(ctx, L, abi.BX), DeviceVec(ctx, L, abi.EX), DeviceVec(ctx, L, abi.BX)
    Bn.c.comp[1] = None
    with pytest.raises(PhbError) as e:
        ctx.faraday(L, B, E, Bn, 0.1)
    assert "not all VecField parameters are usable" in str(e.value)


# ------------------------------------------------------------------------------------------- K8 box ops
@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_box_op(ctxs, cpu_oracle, dim, op):
    ctx = ctxs(dim, 1)
    rng = np.random.default_rng(dim * 3 + op)
    ds, ss = [11, 9, 8][:dim], [10, 12, 7][:dim]
    dst, src = rng.standard_normal(ds), rng.standard_normal(ss)
    dlo, slo, ext = [2, 1, 3][:dim], [1, 4, 0][:dim], [5, 6, 4][:dim]
    want = dst.copy()
    cpu_oracle.box_op(dim, want, dlo, src, slo, ext, op)
    d_dst = DeviceArray(ctx, ds).upload(dst)
    d_src = DeviceArray(ctx, ss).upload(src)
    ctx.box_op(dim, d_dst, ds, dlo, d_src, ss, slo, ext, op)
    assert bit_equal(d_dst.download(), want)


# ------------------------------------------------------------------------------------------- store interop
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_aos_soa_roundtrip(ctxs, dim):
    ctx = ctxs(dim, 1)
    rng = np.random.default_rng(dim)
    L = small_layout(dim, 1)
    soa = random_particles(rng, L, 1234)
    p = dev_particles(ctx, soa)
    raw = p.download_aos()
    stride = {1: 56, 2: 64, 3: 80}[dim]
    assert raw.size == 1234 * stride
    rec = raw.reshape(1234, stride)
    assert np.array_equal(rec[:, 0:8].copy().view(np.float64).ravel(), soa[2])       # weight first
    assert np.array_equal(rec[:, 16:16 + 4 * dim].copy().view(np.int32), soa[0])     # iCell at 16
    p2 = DeviceParticles(ctx, 1234).upload_aos(raw)
    for g, w in zip(p2.download_soa(), soa):
        assert bit_equal(np.asarray(g), np.ascontiguousarray(w))


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_export_multi_matches_sequential_exports(ctxs, cpu_oracle, dim):
    """one-pass export to several (possibly repeated) destinations == the oracle's box-by-box exports"""
    ctx = ctxs(dim, 1)
    rng = np.random.default_rng(70 + dim)
    L = small_layout(dim, 1)
    soa = random_particles(rng, L, 8000)
    dom = domain_box(L)
    lo = [dom.lower[d] for d in range(dim)]
    hi = [dom.upper[d] for d in range(dim)]
    # the ghost strips below and above the domain in x (disjoint), and one interior box
    boxes = [abi.make_box([lo[0] - 1] + lo[1:], [lo[0] - 1] + hi[1:]),
             abi.make_box([hi[0] + 1] + lo[1:], [hi[0] + 1] + hi[1:]),
             abi.make_box([lo[0] + 2] + lo[1:], [lo[0] + 3] + hi[1:])]
    shifts = [[int(L.ncells[0])] + [0] * (dim - 1), [-int(L.ncells[0])] + [0] * (dim - 1), [0] * dim]
    src = HostParticles.from_soa(*soa)
    w0, w1 = HostParticles(dim, 8000), HostParticles(dim, 8000)
    want_counts = [cpu_oracle.export(L, src, 50, 7900, boxes[0], w0, shift=shifts[0]),
                   cpu_oracle.export(L, src, 50, 7900, boxes[1], w1, shift=shifts[1]),
                   cpu_oracle.export(L, src, 50, 7900, boxes[2], w0, shift=shifts[2])]
    d0, d1 = DeviceParticles(ctx, 8000), DeviceParticles(ctx, 8000)
    got_counts = ctx.export_multi(L, dev_particles(ctx, soa), 50, 7900, boxes, shifts, [d0, d1, d0])
    assert got_counts == want_counts and d0.n == w0.n and d1.n == w1.n and min(want_counts) > 0
    assert np.array_equal(canonical_rows(*d0.download_soa()), canonical_rows(*w0.soa()))
    assert np.array_equal(canonical_rows(*d1.download_soa()), canonical_rows(*w1.soa()))
