"""Coarse <-> fine level operators: refiners (default / magnetic + Toth-Roe post-process / electric),
coarseners (electric / moments), NaN fill and fluxSum accumulation.

CPU: the C oracle against golden vectors produced by the REFERENCE's own Python restatements of these operators
(tests/golden/make_amr_golden.py -> amr_operators.npz) and against analytic properties (div B, linear fields).
GPU: the CUDA kernels through the C ABI against the oracle, bit for bit."""
import os

import numpy as np
import pytest

from phare_b200 import abi
from phare_b200.messenger import centering, PRIMAL
from util import bit_equal

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "amr_operators.npz"))
QTY = dict(Bx=abi.BX, By=abi.BY, Bz=abi.BZ, Ex=abi.EX, Ey=abi.EY, Ez=abi.EZ, rho=abi.RHO)
CASES = sorted({k.rsplit("_", 1)[0] for k in GOLD.files if k.endswith("_coarse") and "_cz_" not in k})


def _geom(key):
    ndim, interp, name = int(key[0]), int(key.split("_")[1][1:]), key.split("_")[2]
    g = 2 if interp == 1 else 4
    box = GOLD[key + "_box"]
    return ndim, interp, name, g, box[0], box[1]


def _prim(qty, ndim):
    return np.array([1 if centering(qty, d) == PRIMAL else 0 for d in range(ndim)])


def close(a, b):
    # the reference Python sums the same products in another order: a few ulp of the largest operand
    return np.abs(a - b).max() <= 8 * np.finfo(float).eps * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("key", CASES)
def test_oracle_refiners_match_the_reference_python(cpu_oracle, key):
    cpu = cpu_oracle
    ndim, interp, name, g, lo, hi = _geom(key)
    qty, coarse = QTY[name], GOLD[key + "_coarse"]
    prim = _prim(qty, ndim)
    flo, fhi = 2 * lo, 2 * hi + 1
    layout = abi.make_layout(ndim, interp, list(fhi - flo + 1), [0.2] * ndim, amr_lower=list(flo))
    # --- DefaultFieldRefiner on the physical nodes of the refined patch
    want = GOLD[key + "_default_fine"]
    fine = np.full(want.shape, np.nan)
    cpu.field_refine(ndim, abi.REFINE_DEFAULT, qty, coarse, lo - g, fine, flo - g, flo, fhi + prim)
    inner = tuple(slice(g, s - g) for s in fine.shape)
    assert not np.isnan(fine[inner]).any() and close(fine[inner], want[inner]), key
    assert np.isnan(fine).sum() == fine.size - fine[inner].size          # nothing written outside the box
    keep = fine.copy()
    keep[inner] = 7.0                                                    # only NaN nodes are written
    cpu.field_refine(ndim, abi.REFINE_DEFAULT, qty, coarse, lo - g, keep, flo - g, flo, fhi + prim)
    assert (keep[inner] == 7.0).all()
    # --- MagneticFieldRefiner + postprocessRefine over the whole fine ghost box
    if key + "_magnetic_fine" in GOLD.files:
        want = GOLD[key + "_magnetic_fine"]
        B = [np.full(tuple(fhi - flo + 1 + 2 * g + _prim(abi.BX + c, ndim)) if ndim > 0 else (), np.nan) for c in range(3)]
        comp = qty - abi.BX
        for op in (abi.REFINE_MAGNETIC, abi.REFINE_MAGNETIC_INIT):
            fine = np.full(want.shape, np.nan if op == abi.REFINE_MAGNETIC else 3.0)
            cpu.field_refine(ndim, op, qty, coarse, lo - g, fine, flo - g, flo - g, fhi + g + prim)
            if ndim == 1:
                B[comp] = fine
                if comp == 0:
                    assert np.isnan(fine[1::2]).all() or op != abi.REFINE_MAGNETIC
                    cpu.magnetic_postprocess(layout, B, flo - g, fhi + g)
                assert bit_equal(B[comp], want), key
            else:
                # the reference's Python omits the transverse Toth-Roe terms: compare the coarse faces (all of Bz)
                sl = [slice(None)] * ndim
                if comp < ndim:
                    sl[comp] = slice(0, None, 2)
                assert bit_equal(fine[tuple(sl)], want[tuple(sl)]), key
    # --- ElectricFieldRefiner
    if key + "_electric_fine" in GOLD.files:
        want = GOLD[key + "_electric_fine"]
        fine = np.full(want.shape, np.nan)
        cpu.field_refine(ndim, abi.REFINE_ELECTRIC, qty, coarse, lo - g, fine, flo - g, flo - g + 1, fhi + g + prim - 1)
        inner1 = tuple(slice(1, s - 1) for s in fine.shape)
        assert bit_equal(fine[inner1], want[inner1]), key


@pytest.mark.parametrize("key", [k for k in CASES if k.split("_")[2][0] in "Er"])
def test_oracle_coarseners_match_the_reference_python(cpu_oracle, key):
    cpu = cpu_oracle
    ndim, interp, name, g, lo, hi = _geom(key)
    qty = QTY[name]
    fine, want = GOLD[key + "_cz_fine"], GOLD[key + "_cz_coarse"]
    coarse = np.zeros(want.shape)
    op = abi.COARSEN_MOMENTS if name == "rho" else abi.COARSEN_ELECTRIC
    cpu.field_coarsen(ndim, op, qty, fine, 2 * lo - g, coarse, lo - g, lo, hi)
    assert bit_equal(coarse, want), key


def _random_divfree_2d(rng, n, g):
    """Bx, By on a coarse patch [0,n)^2 with ghosts from a vector potential Az on the nodes: div B = 0 exactly-ish"""
    A = rng.standard_normal((n + 1 + 2 * g, n + 1 + 2 * g))
    Bx = A[:, 1:] - A[:, :-1]                # primal x, dual y:  dAz/dy
    By = -(A[1:, :] - A[:-1, :])             # dual x, primal y: -dAz/dx
    return Bx, By


def test_magnetic_refinement_conserves_divergence_2d(cpu_oracle):
    cpu = cpu_oracle
    """refined + post-processed B keeps div B of every fine cell at the rounding level (Toth & Roe 2002): the point of
    MagneticRefinePatchStrategy; also exercises negative AMR indices"""
    rng = np.random.default_rng(5)
    n, g, interp = 6, 2, 1
    lo = np.array([-4, 3])
    Bx, By = _random_divfree_2d(rng, n, g)
    flo, fhi = 2 * lo, 2 * (lo + n - 1) + 1
    layout = abi.make_layout(2, interp, [2 * n, 2 * n], [0.5, 0.5], amr_lower=list(flo))
    shp = lambda c: tuple(2 * n + 2 * g + _prim(abi.BX + c, 2))
    B = [np.full(shp(c), np.nan) for c in range(3)]
    for c, src in ((0, Bx), (1, By)):
        cpu.field_refine(2, abi.REFINE_MAGNETIC, abi.BX + c, np.ascontiguousarray(src), lo - g, B[c], flo - g, flo - g,
                         fhi + g + _prim(abi.BX + c, 2))
    cpu.magnetic_postprocess(layout, B, flo - g, fhi + g)
    assert not np.isnan(B[0]).any() and not np.isnan(B[1]).any()
    div = (B[0][1:, :] - B[0][:-1, :]) + (B[1][:, 1:] - B[1][:, :-1])
    assert np.abs(div).max() < 1e-14
    # coarse faces carry the coarse flux: two fine faces per coarse face hold the coarse value
    assert bit_equal(B[0][::2, ::2], Bx[g // 2:g // 2 + n + g + 1, g // 2:g // 2 + n + g])


def test_electric_refine_coarsen_roundtrip(cpu_oracle):
    cpu = cpu_oracle
    """coarsen(refine(E)) == E on the coarse edges for every dimension and component (the refiner copies coarse edge
    values onto the fine edges they cover, the coarsener averages exactly those)"""
    rng = np.random.default_rng(11)
    for ndim in (1, 2, 3):
        n, g = 6, 2
        lo = np.array([2, -6, 4][:ndim])
        for c in range(3):
            qty = abi.EX + c
            prim = _prim(qty, ndim)
            coarse = rng.standard_normal(tuple(n + 2 * g + prim))
            flo, fhi = 2 * lo, 2 * (lo + n - 1) + 1
            fine = np.full(tuple(2 * n + 2 * g + prim), np.nan)
            cpu.field_refine(ndim, abi.REFINE_ELECTRIC, qty, coarse, lo - g, fine, flo - g, flo - g, fhi + g + prim - 1)
            back = np.zeros_like(coarse)
            cpu.field_coarsen(ndim, abi.COARSEN_ELECTRIC, qty, fine, flo - g, back, lo - g, lo, lo + n - 1 + prim)
            inner = tuple(slice(g, s - g) for s in coarse.shape)
            assert bit_equal(back[inner], coarse[inner]), (ndim, c)


def test_box_fill_and_axpy(cpu_oracle):
    cpu = cpu_oracle
    a = np.zeros((5, 7))
    cpu.box_fill(2, a, [1, 2], [3, 4], np.nan)
    assert np.isnan(a[1:4, 2:6]).all() and np.isnan(a).sum() == 12
    rng = np.random.default_rng(0)
    d, s = rng.standard_normal(100), rng.standard_normal(100)
    want = d + s * 0.25
    cpu.axpy(d, s, 0.25)
    assert bit_equal(d, want)


# ----------------------------------------------------------------------------------------------- GPU parity
def _dev(ctx, a):
    from phare_b200.device import DeviceArray
    return DeviceArray(ctx, a.shape).upload(a)


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_gpu_level_operators_match_oracle(cpu_oracle, ndim):
    cpu = cpu_oracle
    from phare_b200.device import Context
    ctx = Context(ndim, 1)
    try:
        rng = np.random.default_rng(100 + ndim)
        n, g = (10, 8, 6)[ndim - 1], 2
        lo = np.array([-3, 5, -8][:ndim])
        flo, fhi = 2 * lo, 2 * (lo + n - 1) + 1
        for qty in (abi.BX, abi.BY, abi.BZ, abi.EX, abi.EY, abi.EZ, abi.RHO):
            prim = _prim(qty, ndim)
            coarse = rng.standard_normal(tuple(n + 2 * g + prim))
            fshape = tuple(2 * n + 2 * g + prim)
            ops = [abi.REFINE_DEFAULT]
            if qty <= abi.BZ:
                ops += [abi.REFINE_MAGNETIC, abi.REFINE_MAGNETIC_INIT]
            elif qty <= abi.EZ:
                ops += [abi.REFINE_ELECTRIC]
            for op in ops:
                start = rng.standard_normal(fshape)
                start[rng.random(fshape) < 0.7] = np.nan
                want = start.copy()
                blo, bhi = flo - g + 1, fhi + g + prim - 1
                cpu.field_refine(ndim, op, qty, coarse, lo - g, want, flo - g, blo, bhi)
                dc, df = _dev(ctx, coarse), _dev(ctx, start)
                ctx.field_refine(op, qty, dc, lo - g, df, flo - g, blo, bhi)
                assert bit_equal(df.download(), want), (ndim, qty, op)
            if qty >= abi.EX:
                fine = rng.standard_normal(fshape)
                want = rng.standard_normal(coarse.shape)
                dcz, dfz = _dev(ctx, want), _dev(ctx, fine)
                cop = abi.COARSEN_MOMENTS if qty == abi.RHO else abi.COARSEN_ELECTRIC
                cpu.field_coarsen(ndim, cop, qty, fine, flo - g, want, lo - g, lo, lo + n - 1 + prim)
                ctx.field_coarsen(cop, qty, dfz, flo - g, dcz, lo - g, lo, lo + n - 1 + prim)
                assert bit_equal(dcz.download(), want), (ndim, qty)
        if True:
            layout = abi.make_layout(ndim, 1, list(fhi - flo + 1), [0.3, 0.2, 0.25][:ndim], amr_lower=list(flo))
            from phare_b200.device import DeviceVec
            B = [rng.standard_normal(tuple(2 * n + 2 * g + _prim(abi.BX + c, ndim))) for c in range(3)]
            dB = DeviceVec(ctx, layout, abi.BX, B)
            cpu.magnetic_postprocess(layout, B, flo - g, fhi + g)
            ctx.magnetic_postprocess(layout, dB, flo - g, fhi + g)
            for gpu, w in zip(dB.download(), B):
                assert bit_equal(gpu, w)
        a = rng.standard_normal((9,) * ndim)
        da = _dev(ctx, a)
        ctx.box_fill(da, [1] * ndim, [5] * ndim, float("nan"))
        cpu.box_fill(ndim, a, [1] * ndim, [5] * ndim, float("nan"))
        assert bit_equal(da.download(), a)
        d, s = rng.standard_normal(1000), rng.standard_normal(1000)
        dd, ds = _dev(ctx, d), _dev(ctx, s)
        ctx.axpy(dd, ds, 0.25)
        cpu.axpy(d, s, 0.25)
        assert bit_equal(dd.download(), d)
    finally:
        ctx.close()


def _curl_A_3d(rng, n, dx):
    """discretely divergence-free B on a Yee grid of n cells (arrays WITHOUT ghosts): B = curl A, A random on the edges"""
    nx, ny, nz = n
    Ax, Ay, Az = rng.standard_normal((nx, ny + 1, nz + 1)), rng.standard_normal((nx + 1, ny, nz + 1)), rng.standard_normal((nx + 1, ny + 1, nz))
    Bx = (Az[:, 1:, :] - Az[:, :-1, :]) / dx[1] - (Ay[:, :, 1:] - Ay[:, :, :-1]) / dx[2]
    By = (Ax[:, :, 1:] - Ax[:, :, :-1]) / dx[2] - (Az[1:, :, :] - Az[:-1, :, :]) / dx[0]
    Bz = (Ay[1:, :, :] - Ay[:-1, :, :]) / dx[0] - (Ax[:, 1:, :] - Ax[:, :-1, :]) / dx[1]
    return [Bx, By, Bz]


def test_magnetic_refinement_conserves_divergence_3d(cpu_oracle):
    """3-D Toth-Roe (magnetic_refine_patch_strategy.hpp:192-372) with three different mesh sizes: the refined field is
    divergence-free in every fine cell, which the second- AND third-order terms are needed for"""
    cpu = cpu_oracle
    rng = np.random.default_rng(8)
    g, interp = 2, 1
    n = np.array([5, 4, 6])
    dxc = [0.6, 0.4, 0.5]
    lo = np.array([-2, 3, 0])
    coarse = _curl_A_3d(rng, n + 2 * g, dxc)  # the coarse patch with its ghost layers
    assert max(np.abs((coarse[0][1:] - coarse[0][:-1]) / dxc[0] + (coarse[1][:, 1:] - coarse[1][:, :-1]) / dxc[1]
                      + (coarse[2][:, :, 1:] - coarse[2][:, :, :-1]) / dxc[2]).max(), 0) < 1e-12
    flo, fhi = 2 * lo, 2 * (lo + n - 1) + 1
    dxf = [d / 2 for d in dxc]
    layout = abi.make_layout(3, interp, list(2 * n), dxf, amr_lower=list(flo))
    B = [np.full(tuple(2 * n + 2 * g + _prim(abi.BX + c, 3)), np.nan) for c in range(3)]
    for c in range(3):
        cpu.field_refine(3, abi.REFINE_MAGNETIC_INIT, abi.BX + c, np.ascontiguousarray(coarse[c]), lo - g, B[c], flo - g,
                         flo - g, fhi + g + _prim(abi.BX + c, 3))
    cpu.magnetic_postprocess(layout, B, flo - g, fhi + g)
    assert not any(np.isnan(b).any() for b in B)
    div = ((B[0][1:] - B[0][:-1]) / dxf[0] + (B[1][:, 1:] - B[1][:, :-1]) / dxf[1] + (B[2][:, :, 1:] - B[2][:, :, :-1]) / dxf[2])
    assert np.abs(div).max() < 1e-11 * max(np.abs(b).max() for b in B) / min(dxf)
    # the four fine faces on a coarse face carry its value
    assert bit_equal(B[0][::2, ::2, 1::2], coarse[0][g // 2:g // 2 + n[0] + g + 1, g // 2:g // 2 + n[1] + g, g // 2:g // 2 + n[2] + g])


def _refine_B_3d(cpu, coarse, n, dxc, lo, g=2):
    flo, fhi = 2 * lo, 2 * (lo + n - 1) + 1
    layout = abi.make_layout(3, 1, list(2 * n), [d / 2 for d in dxc], amr_lower=list(flo))
    B = [np.full(tuple(2 * n + 2 * g + _prim(abi.BX + c, 3)), np.nan) for c in range(3)]
    for c in range(3):
        cpu.field_refine(3, abi.REFINE_MAGNETIC_INIT, abi.BX + c, np.ascontiguousarray(coarse[c]), lo - g, B[c], flo - g,
                         flo - g, fhi + g + _prim(abi.BX + c, 3))
    cpu.magnetic_postprocess(layout, B, flo - g, fhi + g)
    return B


def test_toth_roe_3d_is_invariant_under_cyclic_permutation_of_the_axes(cpu_oracle):
    """the divergence does not see the third-order terms (they cancel in it); relabelling (x, y, z) -> (y, z, x) maps the
    Bx formula on the By formula and so on, so the three restated formulas check each other, mesh-size factors and
    ijk_factor signs included"""
    rng = np.random.default_rng(8)
    n, dxc, lo = np.array([5, 4, 6]), [0.6, 0.4, 0.5], np.array([-2, 3, 0])
    coarse = _curl_A_3d(rng, n + 4, dxc)
    B = _refine_B_3d(cpu_oracle, coarse, n, dxc, lo)
    T = lambda a: np.ascontiguousarray(np.transpose(a, (2, 0, 1)))
    perm = lambda v: np.array([v[2], v[0], v[1]])
    Bp = _refine_B_3d(cpu_oracle, [T(coarse[2]), T(coarse[0]), T(coarse[1])], perm(n), list(perm(dxc)), perm(lo))
    for got, want in ((Bp[0], T(B[2])), (Bp[1], T(B[0])), (Bp[2], T(B[1]))):
        assert np.abs(got - want).max() < 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_default_refiner_is_multilinear_interpolation(cpu_oracle, ndim):
    """DefaultFieldRefiner against an independent numpy statement of what it is: linear interpolation, per direction, of
    the coarse samples at the position of the fine node (primal: x_f = i/2; dual: x_f = (i + 1/2)/2 - 1/2 in coarse
    index units).  Exact for fields that are linear in the coordinates, which also checks the weights / start indices
    for negative AMR indices."""
    rng = np.random.default_rng(40 + ndim)
    g, n = 2, np.array([6, 5, 4][:ndim])
    lo = np.array([-5, 2, -1][:ndim])
    for qty in (abi.BX, abi.EX, abi.EZ, abi.BZ, abi.RHO):
        prim = _prim(qty, ndim)
        cshape = tuple(n + 2 * g + prim)
        # a field linear in the coarse coordinates + a random part
        axes = [np.arange(cshape[d]) + (lo[d] - g) + (0.0 if prim[d] else 0.5) for d in range(ndim)]
        X = np.meshgrid(*axes, indexing="ij")
        coef = rng.standard_normal(ndim)
        linear = sum(c * x for c, x in zip(coef, X)) + 0.3
        for coarse in (linear, rng.standard_normal(cshape)):
            flo, fhi = 2 * lo, 2 * (lo + n - 1) + 1
            fine = np.full(tuple(2 * n + 2 * g + prim), np.nan)
            blo, bhi = flo - g + 1, fhi + g + prim - 1
            cpu_oracle.field_refine(ndim, abi.REFINE_DEFAULT, qty, np.ascontiguousarray(coarse), lo - g, fine, flo - g, blo, bhi)
            # numpy: position of every fine node of the box in coarse index units, then separable linear interpolation
            want = np.ascontiguousarray(coarse)
            for d in range(ndim):
                fi = np.arange(blo[d], bhi[d] + 1)
                xc = fi / 2.0 if prim[d] else (fi + 0.5) / 2.0 - 0.5
                i0 = np.floor(xc).astype(int)
                w1 = xc - i0
                a = np.take(want, i0 - (lo[d] - g), axis=d)
                b = np.take(want, i0 + 1 - (lo[d] - g), axis=d)
                shape = [1] * ndim
                shape[d] = len(fi)
                want = a * (1 - w1).reshape(shape) + b * w1.reshape(shape)
            got = fine[tuple(slice(int(blo[d] - (flo[d] - g)), int(bhi[d] - (flo[d] - g)) + 1) for d in range(ndim))]
            assert np.abs(got - want).max() <= 1e-14 * max(1.0, np.abs(want).max()), (ndim, qty)
            if coarse is linear:  # the refined field is the same linear function at the fine nodes
                faxes = [(np.arange(blo[d], bhi[d] + 1) + (0.0 if prim[d] else 0.5)) / 2.0 for d in range(ndim)]
                FX = np.meshgrid(*faxes, indexing="ij")
                exact = sum(c * x for c, x in zip(coef, FX)) + 0.3
                assert np.abs(got - exact).max() <= 1e-13 * max(1.0, np.abs(exact).max()), (ndim, qty)


@pytest.mark.parametrize("ndim", [2, 3])
def test_electric_refiner_reproduces_edge_averages(cpu_oracle, ndim):
    """ElectricFieldRefiner in 2-D / 3-D (no usable upstream Python restatement): along its own (dual) direction a fine
    edge takes the coarse edge it lies on; across a primal direction it takes the coarse edge it lies on (even index) or
    the mean of the two neighbouring ones (odd index).  Independent numpy statement: piecewise constant along the dual
    direction, linear interpolation at the half-way point along the primal ones."""
    rng = np.random.default_rng(60 + ndim)
    g, n = 2, np.array([6, 5, 4][:ndim])
    lo = np.array([-3, 4, 1][:ndim])
    for c in range(3):
        qty = abi.EX + c
        prim = _prim(qty, ndim)
        coarse = rng.standard_normal(tuple(n + 2 * g + prim))
        flo, fhi = 2 * lo, 2 * (lo + n - 1) + 1
        fine = np.full(tuple(2 * n + 2 * g + prim), np.nan)
        blo, bhi = flo - g + 1, fhi + g + prim - 1
        cpu_oracle.field_refine(ndim, abi.REFINE_ELECTRIC, qty, coarse, lo - g, fine, flo - g, blo, bhi)
        want = coarse
        for d in range(ndim):
            fi = np.arange(blo[d], bhi[d] + 1)
            ci = fi // 2 - (lo[d] - g)
            a = np.take(want, ci, axis=d)
            if prim[d]:
                b = np.take(want, np.minimum(ci + 1, want.shape[d] - 1), axis=d)
                odd = (fi % 2 != 0).reshape([len(fi) if k == d else 1 for k in range(ndim)])
                want = np.where(odd, 0.5 * (a + b), a)
            else:
                want = a
        got = fine[tuple(slice(int(blo[d] - (flo[d] - g)), int(bhi[d] - (flo[d] - g)) + 1) for d in range(ndim))]
        assert np.abs(got - want).max() <= 4e-16 * max(1.0, np.abs(want).max()), (ndim, c)
