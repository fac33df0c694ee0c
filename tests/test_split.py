"""Particle splitting: the committed pattern table against the reference's Splitter classes, and
phb_split (GPU) against the reference's toFineGrid + Splitter::operator() compiled in place (oracle/_ref), bit for bit."""
import numpy as np
import pytest

from phare_b200 import abi
from phare_b200.split import pattern, permutations
from oracle import HostParticles, canonical_rows
from util import bit_equal


def test_table_covers_the_reference_permutations():
    perms = permutations()
    assert len(perms) == 32 and (1, 1, 2) in perms and (2, 3, 25) in perms and (3, 1, 6) in perms
    for dim, interp, nref in perms:
        d, w, m = pattern(dim, interp, nref)
        assert d.shape == (nref, dim) and w.shape == (nref,) and d.dtype == np.float32
        assert m == int(np.ceil((interp + 1) * 0.5))          # ASplitter::maxCellDistanceFromSplit
        assert np.allclose(d.sum(axis=0), 0, atol=1e-6)       # patterns are symmetric around the coarse particle
        # refined weights sum to 1 (the 2^dim factor of dispatch() compensates the finer cell volume) ... except where the
        # reference's own tables do not: Splitter<2,1,8> is built from the literals {1,2},{2,1}, <2,1,9> sums to 0.815 and
        # the 27-particle 3-D tables are marked TODO upstream; the table mirrors the reference, quirks included
        if (dim, interp, nref) not in ((2, 1, 8), (2, 1, 9)) and nref != 27:
            assert abs(float(w.sum()) - 1.0) < 2e-3
    with pytest.raises(KeyError):
        pattern(1, 1, 7)


def test_table_equals_the_reference_splitters(cpu_ref):
    for dim, interp, nref in permutations():
        d, w, m = pattern(dim, interp, nref)
        rd, rw, rm = cpu_ref.split_pattern(dim, interp, nref)
        assert bit_equal(d, rd) and bit_equal(w, rw) and m == rm, (dim, interp, nref)


def _coarse(rng, dim, n):
    icell = rng.integers(-4, 12, size=(n, dim)).astype(np.int32)
    delta = rng.random((n, dim))
    delta[::17] = 0.0          # edge values of delta
    delta[5::19] = 0.5
    return icell, delta, rng.random(n) + 0.1, np.where(rng.random(n) < 0.5, 1.0, 2.0), rng.standard_normal((n, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("dim,interp,nref", [(1, 1, 2), (1, 2, 3), (1, 3, 5), (2, 1, 4), (2, 2, 16), (2, 3, 25), (2, 1, 9),
                                             (3, 1, 6), (3, 2, 12), (3, 3, 27)])
def test_split_matches_reference_bit_for_bit(cpu_ref, dim, interp, nref):
    from phare_b200.device import Context, DeviceParticles
    ctx = Context(dim, interp)
    try:
        rng = np.random.default_rng(1000 * dim + 10 * interp + nref)
        soa = _coarse(rng, dim, 3000)
        want = cpu_ref.split(dim, interp, nref, HostParticles.from_soa(*soa)).soa()
        d, w, m = pattern(dim, interp, nref)
        src = DeviceParticles(ctx, 3000).upload_soa(*soa)
        # 1. no restriction: every refined particle, in the reference's order (source order, pattern order)
        big = 2 ** 30
        dst = DeviceParticles(ctx, 3000 * nref)
        n = ctx.split(src, 0, 3000, d, w, m, [abi.make_box([-big] * dim, [big] * dim)], dst)
        assert n == 3000 * nref == dst.n
        for g, x in zip(dst.download_soa(), want):
            assert bit_equal(g, x)
        # 2. two disjoint destination boxes on the fine level, a sub-range of the source, appended to a non-empty store
        boxes = [abi.make_box([0] * dim, [7] * dim), abi.make_box([8] + [0] * (dim - 1), [15] + [9] * (dim - 1))]
        pre = tuple(a[:5] for a in soa)
        dst = DeviceParticles(ctx, 5 + 2000 * nref).upload_soa(*pre)
        n = ctx.split(src, 500, 2500, d, w, m, boxes, dst)
        wi = want[0][500 * nref:2500 * nref]
        keep = np.zeros(len(wi), bool)
        for b in boxes:
            inb = np.ones(len(wi), bool)
            for k in range(dim):
                inb &= (wi[:, k] >= b.lower[k]) & (wi[:, k] <= b.upper[k])
            keep |= inb
        assert n == int(keep.sum()) > 0 and dst.n == 5 + n
        got = dst.download_soa()
        for g, p0 in zip(got, pre):
            assert bit_equal(g[:5], p0)
        assert np.array_equal(canonical_rows(*[g[5:] for g in got]),
                              canonical_rows(*[x[500 * nref:2500 * nref][keep] for x in want]))
        # 3. capacity is checked, nothing is appended
        small = DeviceParticles(ctx, 3)
        from phare_b200.device import PhbError
        with pytest.raises(PhbError) as e:
            ctx.split(src, 0, 3000, d, w, m, [abi.make_box([-big] * dim, [big] * dim)], small)
        assert e.value.code == abi.PHB_ERR_CAPACITY and small.n == 0
    finally:
        ctx.close()


@pytest.mark.gpu
def test_pybindlibs_split_pyarray_particles(cpu_ref):
    import importlib
    m = importlib.import_module("pybindlibs.cpp_2_1_4")
    rng = np.random.default_rng(3)
    soa = _coarse(rng, 2, 200)
    want = cpu_ref.split(2, 1, 4, HostParticles.from_soa(*soa)).soa()
    got = m.split_pyarray_particles((soa[0].reshape(-1), soa[1].reshape(-1), soa[2], soa[3], soa[4].reshape(-1)))
    assert bit_equal(got[0].reshape(-1, 2), want[0]) and bit_equal(got[1].reshape(-1, 2), want[1])
    assert bit_equal(got[2], want[2]) and bit_equal(got[3], want[3]) and bit_equal(got[4].reshape(-1, 3), want[4])
    sp = m.Splitter()
    assert sp.nbRefinedPart == 4 and sp.max_cell_distance == 1
