"""GPU parity of the tile kernels (csrc/tile.cuh: E,B block of a CTA staged in shared memory by bulk copies) through
the C ABI against the CPU oracle: phb_push_cells bit-exact, phb_push_deposit_plan + phb_scatter_planned ==
oracle push -> deposit -> bin (offsets, counts and per-cell multisets bit-exact, moments <= 1e-10)."""
import os

import numpy as np
import pytest

from phare_b200 import abi
from phare_b200.device import Context, DeviceArray, DeviceVec, DeviceParticles, PhbError
from oracle import HostParticles, canonical_rows
from util import (ALL_DIM_INTERP, bit_equal, random_vec, random_particles, domain_box, grown, particle_ghosts)

pytestmark = pytest.mark.gpu
MOMENT_RTOL = 1e-10

# patches that hold whole blocks (3-D block = 4 x 4 x 8 cells, 2-D 8 x 16, 1-D 128) AND clipped ones
SHAPES = {1: [300], 2: [21, 37], 3: [9, 6, 19]}


@pytest.fixture(scope="module")
def ctxs():
    cache = {}

    def get(dim, interp, no_tile=False, strip=False):
        key = (dim, interp, no_tile, strip)
        if key not in cache:
            old = {k: os.environ.get(k) for k in ("PHB_NO_TILE", "PHB_STRIP")}
            os.environ["PHB_NO_TILE"] = "1" if no_tile else "0"
            os.environ["PHB_STRIP"] = "1" if strip else "0"
            try:
                cache[key] = Context(dim, interp)
            finally:
                for k, v in old.items():
                    if v is None:
                        del os.environ[k]
                    else:
                        os.environ[k] = v
        return cache[key]

    yield get
    for c in cache.values():
        c.close()


def layout_for(dim, interp):
    return abi.make_layout(dim, interp, SHAPES[dim], [0.2, 0.3, 0.25][:dim], amr_lower=[5, -3, 10][:dim])


def binned_store(cpu, rng, L, n, vth, tail=0):
    """a store as the step keeps it: [0, n_sorted) ordered by the oracle's bin (cells hold 0 .. many particles),
    then `tail` unordered particles (received since the last binning); returns soa, n_sorted, cell_start"""
    dim = L.dim
    dom = domain_box(L)
    icell = np.stack([rng.integers(dom.lower[d], dom.upper[d] + 1, n) for d in range(dim)], 1).astype(np.int32)
    # leave some cells empty and make others crowded
    icell[: n // 5] = icell[n // 5: 2 * (n // 5)]
    delta = rng.random((n, dim))
    v = rng.standard_normal((n, 3)) * vth
    w = rng.random(n) + 0.1
    q = np.where(rng.random(n) < 0.5, 1.0, 2.0)
    keep = [grown(dom, dim, particle_ghosts(L.interp))]
    sorted_, cs, counts = cpu.bin(L, HostParticles.from_soa(icell, delta, w, q, v), dom, keep)
    assert counts == (n, 0, 0)
    soa = sorted_.soa()
    if tail:
        t = random_particles(rng, L, tail, spread=0, vth=vth)
        soa = tuple(np.concatenate([a, b]) for a, b in zip(soa, t))
    return soa, n, cs


def moments(ctx, L, cpu):
    shape = cpu.field_shape(L, abi.RHO)
    return (DeviceArray(ctx, shape), DeviceArray(ctx, shape)), DeviceVec(ctx, L, abi.VX)


def close(got, want):
    for g, w in zip(got, want):
        scale = np.max(np.abs(w)) + 1e-300
        assert np.max(np.abs(g - w)) <= MOMENT_RTOL * scale, (np.max(np.abs(g - w)), scale)


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
@pytest.mark.parametrize("in_place", [True, False])
def test_push_cells_bitexact(ctxs, cpu_oracle, dim, interp, in_place):
    """K1 with the E,B of each strip of cells in shared memory (csrc/strip.cuh) == the oracle push, bit for bit; vth is
    large enough for many particles to be pre-pushed out of their strip's rows (they gather from global memory instead)"""
    ctx = ctxs(dim, interp, strip=True)
    rng = np.random.default_rng(3000 + 10 * dim + interp)
    L = layout_for(dim, interp)
    ncell = int(np.prod(SHAPES[dim]))
    soa, n_sorted, cs = binned_store(cpu_oracle, rng, L, ncell * (12 if dim < 3 else 6), vth=1.5, tail=333)
    # |v| <= 3.4: at most one cell per half step (h = 0.25) ...
    np.clip(soa[4], -3.4, 3.4, out=soa[4])
    # ... except a few planted particles, well inside the patch, that are pre-pushed exactly two cells (legal, boris.hpp:164)
    dom0 = domain_box(L)
    inner = np.flatnonzero(np.all([(soa[0][:n_sorted, d] >= dom0.lower[d] + 3) & (soa[0][:n_sorted, d] <= dom0.upper[d] - 3)
                                   for d in range(dim)], axis=0))
    fast = inner[:: max(1, len(inner) // 40)]
    soa[1][fast, 0] = 0.01
    soa[4][fast] = [-4.5, 0.1, -0.2]
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    dt = 0.1
    rc, want = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, dt)
    assert rc == 0
    moved = np.abs(want.soa()[0] - soa[0]).max(axis=1)
    assert (moved >= 1).mean() > 0.2 and (moved >= 2).sum() >= 10
    pin = DeviceParticles(ctx, len(soa[2])).upload_soa(*soa)
    dcs = DeviceArray(ctx, cs.shape, np.uint32).upload(cs)
    dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
    if in_place:
        pout = pin
    else:
        # the step's out-of-place target: its own position / velocity columns, weight and charge aliased
        pout = DeviceParticles(ctx, len(soa[2]))
        own_w, own_q = pout.c.weight, pout.c.charge
        pout.c.weight, pout.c.charge = pin.c.weight, pin.c.charge
    ctx.push_cells(L, dE, dB, pin, pout, n_sorted, 1.0, dt, domain_box(L), dcs)
    ctx.poll_error()
    pout.n = pin.n
    for g, w in zip(pout.download_soa(), want.soa()):
        assert bit_equal(g, w)
    if not in_place:
        for g, w in zip(pin.download_soa(), soa):
            assert bit_equal(g, w)
        pout.c.weight, pout.c.charge = own_w, own_q


@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_push_deposit_plan_then_scatter(ctxs, cpu_oracle, dim, interp):
    """the `all` sweep in two passes: (push in place + deposit + plan) then (scatter) == oracle push, deposit of the
    pushed particles inside the keep boxes, bin.  ~25 % of the particles change cell, some leave the patch through the
    face that is not a keep box (dropped) and some through the others (new patch ghosts)."""
    ctx = ctxs(dim, interp)
    rng = np.random.default_rng(4000 + 10 * dim + interp)
    L = layout_for(dim, interp)
    ncell = int(np.prod(SHAPES[dim]))
    soa, n_sorted, cs = binned_store(cpu_oracle, rng, L, ncell * (12 if dim < 3 else 6), vth=1.0, tail=257)
    n = len(soa[2])
    dom = domain_box(L)
    pg = particle_ghosts(interp)
    G = grown(dom, dim, pg)
    keep_lo = [G.lower[d] for d in range(dim)]
    keep_lo[0] = dom.lower[0]
    keep = [abi.make_box(keep_lo, [G.upper[d] for d in range(dim)])]
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    dt = 0.1
    rc, pushed = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, dt)
    assert rc == 0
    want_m = cpu_oracle.deposit(L, pushed, coef=0.9, sel=keep)
    want, want_cs, want_counts = cpu_oracle.bin(L, pushed, dom, keep)
    assert want_counts[1] > 0 and want_counts[2] > 0
    pin, pout = DeviceParticles(ctx, n).upload_soa(*soa), DeviceParticles(ctx, n)
    (rn, rq), F = moments(ctx, L, cpu_oracle)
    cs_old = DeviceArray(ctx, cs.shape, np.uint32).upload(cs)
    cs_new = DeviceArray(ctx, cs.shape, np.uint32)
    dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
    ctx.push_deposit_plan(L, dE, dB, pin, n_sorted, 1.0, dt, rn, rq, F, 0.9, keep, dom, cs_old, keep, cs_new)
    ctx.poll_error()
    tiled = (interp + 1 if interp != 2 else 4) ** dim <= 16  # tile_supported<dim, interp>()
    if tiled:  # (a download stages through the context scratch, where the plan of the three-pass fallback lives)
        for g, w in zip(pin.download_soa(), pushed.soa()):  # pushed in place, order untouched
            assert bit_equal(g, w)
    close([rn.download(), rq.download()] + F.download(), want_m)
    assert np.array_equal(cs_new.download(), want_cs)
    ctx.scatter_planned(L, pin, n_sorted, dom, cs_old, keep, pout, cs_new)
    counts = ctx.bin_counts(L, dom, cs_new, pout)
    assert counts == want_counts and pout.n == counts[0] + counts[1]
    got = pout.download_soa()
    assert np.array_equal(got[0], want.soa()[0])
    assert np.array_equal(canonical_rows(*got), canonical_rows(*want.soa()))
    with pytest.raises(PhbError):  # a scatter without its plan is refused
        ctx.scatter_planned(L, pin, n_sorted, dom, cs_old, keep, pout, cs_new)


@pytest.mark.parametrize("dim,interp", [(1, 1), (2, 1), (2, 3), (3, 1)])
@pytest.mark.parametrize("write_back", [False, True])
def test_push_deposit_tile_equals_round1_kernel(ctxs, cpu_oracle, dim, interp, write_back):
    """phb_push_deposit on the cell-ordered store: the tile kernel and the round-1 kernel (PHB_NO_TILE=1: gathers
    through L1) store the same bits and agree on the moments; both against the oracle"""
    rng = np.random.default_rng(5000 + 10 * dim + interp)
    L = layout_for(dim, interp)
    ncell = int(np.prod(SHAPES[dim]))
    soa, n_sorted, cs = binned_store(cpu_oracle, rng, L, ncell * (20 if dim < 3 else 8), vth=0.7)
    dom = domain_box(L)
    keep = [grown(dom, dim, particle_ghosts(interp))]
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    dt = 0.06
    rc, pushed = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, dt)
    assert rc == 0
    want_m = cpu_oracle.deposit(L, pushed, coef=1.0, sel=keep)
    stored = []
    for no_tile in (False, True):
        ctx = ctxs(dim, interp, no_tile)
        parts = DeviceParticles(ctx, len(soa[2])).upload_soa(*soa)
        (rn, rq), F = moments(ctx, L, cpu_oracle)
        dcs = DeviceArray(ctx, cs.shape, np.uint32).upload(cs)
        ctx.push_deposit(L, DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B), parts, 1.0, dt, rn, rq, F, 1.0,
                         0, n_sorted, keep, dom, dcs, write_back=write_back)
        ctx.poll_error()
        close([rn.download(), rq.download()] + F.download(), want_m)
        stored.append(parts.download_soa())
    for a, b, w, s in zip(stored[0], stored[1], pushed.soa(), soa):
        assert bit_equal(a, b) and bit_equal(a, w if write_back else s)


def test_move_two_cells_is_reported_by_the_tile_kernels(ctxs, cpu_oracle):
    ctx = ctxs(2, 1)
    rng = np.random.default_rng(77)
    L = layout_for(2, 1)
    soa, n_sorted, cs = binned_store(cpu_oracle, rng, L, 5000, vth=0.1)
    soa[4][1234, 0] = 1000.0  # one runaway
    dom = domain_box(L)
    keep = [grown(dom, 2, 1)]
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.1)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.1)
    parts, out = DeviceParticles(ctx, 5000).upload_soa(*soa), DeviceParticles(ctx, 5000)
    (rn, rq), F = moments(ctx, L, cpu_oracle)
    cs_old = DeviceArray(ctx, cs.shape, np.uint32).upload(cs)
    cs_new = DeviceArray(ctx, cs.shape, np.uint32)
    ctx.push_deposit_plan(L, DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B), parts, n_sorted, 1.0, 0.1, rn, rq,
                          F, 1.0, keep, dom, cs_old, keep, cs_new)
    with pytest.raises(PhbError) as e:
        ctx.poll_error()
    assert e.value.code == abi.PHB_ERR_MOVE_TWO_CELL
    # the plan stays consistent: the scatter runs and keeps every particle
    ctx.scatter_planned(L, parts, n_sorted, dom, cs_old, keep, out, cs_new)
    counts = ctx.bin_counts(L, dom, cs_new, out)
    assert sum(counts) == 5000


@pytest.mark.parametrize("strips", [True, False], ids=["strip_kernel", "streaming_kernel"])
@pytest.mark.parametrize("dim,interp", ALL_DIM_INTERP)
def test_push_plan_equals_push_then_bin_plan(ctxs, cpu_oracle, dim, interp, strips):
    """phb_push_plan (K1 in place with the count of the re-binning folded in) == phb_push in place then phb_bin_plan:
    particles bit-exact in place, the same cell_start, and phb_deposit_scatter consumes the plan to the same re-binned
    store (per-cell multisets) and moments as the oracle's push -> deposit -> bin"""
    ctx = ctxs(dim, interp, strip=strips)
    rng = np.random.default_rng(5000 + 10 * dim + interp)
    L = layout_for(dim, interp)
    ncell = int(np.prod(SHAPES[dim]))
    soa, n_sorted, cs = binned_store(cpu_oracle, rng, L, ncell * (12 if dim < 3 else 6) + 77, vth=1.0, tail=131)
    n = len(soa[2])
    dom = domain_box(L)
    keep = [grown(dom, dim, particle_ghosts(interp))]
    E = random_vec(rng, cpu_oracle.field_shape, L, abi.EX, 0.3)
    B = random_vec(rng, cpu_oracle.field_shape, L, abi.BX, 0.3)
    dt = 0.1
    rc, pushed = cpu_oracle.push(L, E, B, HostParticles.from_soa(*soa), 1.0, dt)
    assert rc == 0
    want_m = cpu_oracle.deposit(L, pushed, coef=1.0, sel=keep)
    want, want_cs, want_counts = cpu_oracle.bin(L, pushed, dom, keep)
    pin, pout = DeviceParticles(ctx, n).upload_soa(*soa), DeviceParticles(ctx, n)
    (rn, rq), F = moments(ctx, L, cpu_oracle)
    cs_old = DeviceArray(ctx, cs.shape, np.uint32).upload(cs)
    cs_new = DeviceArray(ctx, cs.shape, np.uint32)
    dE, dB = DeviceVec(ctx, L, abi.EX, E), DeviceVec(ctx, L, abi.BX, B)
    if strips:
        ctx.push_plan(L, dE, dB, pin, 1.0, dt, dom, keep, cs_new, n_sorted, cs_old)
    else:
        ctx.push_plan(L, dE, dB, pin, 1.0, dt, dom, keep, cs_new)
    ctx.deposit_scatter(L, pin, n_sorted, rn, rq, F, 1.0, keep, dom, cs_old, keep, pout, cs_new)
    counts = ctx.bin_counts(L, dom, cs_new, pout)
    ctx.poll_error()
    for g, w in zip(pin.download_soa(), pushed.soa()):  # pushed in place, order untouched
        assert bit_equal(g, w)
    assert np.array_equal(cs_new.download(), want_cs)
    assert counts == want_counts
    got = pout.download_soa()
    assert np.array_equal(got[0], want.soa()[0])
    assert np.array_equal(canonical_rows(*got), canonical_rows(*want.soa()))
    close([rn.download(), rq.download()] + F.download(), want_m)
