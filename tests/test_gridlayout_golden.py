"""SURVEY §8 row a14 pinned directly: the reference's own GridLayout golden vectors (tests/core/data/gridlayout/
allocSizes.py, gridIndexing.py, test_deriv.py, test_laplacian.py, test_linear_combinations_yee.py, run unmodified by
tests/golden/make_gridlayout_golden.py) against
  * phb_field_shape (host arithmetic, no GPU)            <- allocSizes
  * the index space the kernels iterate over and the device functions Faraday / Ampere / Ohm are built from (deriv,
    laplacian, the Yee linear combinations), through phb_gridlayout_probe on the GPU  <- gridIndexing, deriv, laplacian,
    linear_coefs_yee_*"""
import ctypes as C
import os

import numpy as np
import pytest

from phare_b200 import abi

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gridlayout_golden.npz"))
QTY = dict(Bx=abi.BX, By=abi.BY, Bz=abi.BZ, Ex=abi.EX, Ey=abi.EY, Ez=abi.EZ, Jx=abi.JX, Jy=abi.JY, Jz=abi.JZ,
           rho=abi.RHO, Vx=abi.VX, Vy=abi.VY, Vz=abi.VZ, P=abi.P)
ORDER = ["Bx", "By", "Bz", "Ex", "Ey", "Ez", "Jx", "Jy", "Jz", "rho", "Vx", "Vy", "Vz", "P"]  # iqty of the golden files
PRIMAL = {"B": lambda c, d: c == d, "E": lambda c, d: c != d, "J": lambda c, d: c != d}


def is_primal(name, d):
    return PRIMAL[name[0]]("xyz".index(name[1]), d) if name[0] in PRIMAL else True


def rows(name):
    return [ln.split() for ln in str(G["text/" + name]).strip().splitlines()]


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interp", [1, 2, 3])
def test_alloc_sizes(dim, interp):
    """allocSizes_<dim>d_O<interp>.txt: iqty, nbrCells[dim], dl[dim], allocSize[dim], allocSizeDerived[dim]"""
    lib = abi.load()
    for r in rows(f"allocSizes_{dim}d_O{interp}.txt"):
        iq = int(r[0])
        ncells = [int(x) for x in r[1:1 + dim]]
        dl = [float(x) for x in r[1 + dim:1 + 2 * dim]]
        alloc = [int(x) for x in r[1 + 2 * dim:1 + 3 * dim]]
        alloc_der = [int(x) for x in r[1 + 3 * dim:1 + 4 * dim]]
        L = abi.make_layout(dim, interp, ncells, dl)
        s = (C.c_uint32 * 3)()
        n = lib.phb_field_shape(C.byref(L), QTY[ORDER[iq]], s)
        assert [int(s[d]) for d in range(dim)] == alloc, (ORDER[iq], alloc)
        assert n == int(np.prod(alloc))
        # the derivative along d lives on the other centering: primal n + 1 + 2g, dual n + 2g
        g = lib.phb_field_ghosts(interp)
        for d in range(dim):
            flipped = ncells[d] + 2 * g + (0 if is_primal(ORDER[iq], d) else 1)
            assert flipped == alloc_der[d], (ORDER[iq], d)


def _probe(ctx, L, op, qty, arg, host_in, out_shape):
    from phare_b200.device import DeviceArray
    din = DeviceArray(ctx, host_in.shape).upload(np.ascontiguousarray(host_in))
    dout = DeviceArray(ctx, out_shape).upload(np.full(out_shape, np.nan))
    ctx._check(ctx.lib.phb_gridlayout_probe(ctx.h, C.byref(L), op, qty, arg, din.ptr, dout.ptr))
    out = dout.download()
    din.free()
    dout.free()
    return out


@pytest.fixture(scope="module")
def ctxs():
    from phare_b200.device import Context
    cache = {}

    def get(dim, interp):
        if (dim, interp) not in cache:
            cache[(dim, interp)] = Context(dim, interp)
        return cache[(dim, interp)]
    yield get
    for c in cache.values():
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interp", [1, 2, 3])
def test_grid_indexing(ctxs, dim, interp):
    """gridIndexing_<dim>d_O<interp>.txt: iqty, nbrCells, dl, PSI[dim], PEI[dim], GSI[dim], GEI[dim]: the laplacian probe
    writes exactly the physical box [PSI, PEI] of its quantity; the allocation is [GSI, GEI]"""
    ctx = ctxs(dim, interp)
    for r in rows(f"gridIndexing_{dim}d_O{interp}.txt"):
        iq = int(r[0])
        ncells = [int(x) for x in r[1:1 + dim]]
        dl = [float(x) for x in r[1 + dim:1 + 2 * dim]]
        psi, pei, gsi, gei = ([int(x) for x in r[1 + (2 + k) * dim:1 + (3 + k) * dim]] for k in range(4))
        L = abi.make_layout(dim, interp, ncells, dl)
        shape = ctx.field_shape(L, QTY[ORDER[iq]])
        assert gsi == [0] * dim and [g + 1 for g in gei] == list(shape)
        out = _probe(ctx, L, 1, QTY[ORDER[iq]], 0, np.ones(shape), shape)
        written = ~np.isnan(out)
        want = np.zeros(shape, bool)
        want[tuple(slice(psi[d], pei[d] + 1) for d in range(dim))] = True
        assert np.array_equal(written, want), ORDER[iq]


def _layout_of_generators(dim, interp):
    return abi.make_layout(dim, interp, [50, 30, 40][:dim], [0.1, 0.2, 0.3][:dim])


def _full(name_in, dim, shape_key):
    """the generator's input array; 3-D ones were cropped to a window grown by one node: re-embed in zeros"""
    a = G[name_in]
    shape = tuple(int(x) for x in G[name_in + "/shape"])
    if dim < 3:
        return a, None
    lo, n = G["window"]
    full = np.zeros(shape)
    sl = tuple(slice(int(lo[d]) - 1, int(lo[d]) + int(n[d]) + 1) for d in range(3))
    full[sl] = a
    return full, tuple(slice(int(lo[d]), int(lo[d]) + int(n[d])) for d in range(3))


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interp", [1, 2, 3])
def test_deriv(ctxs, dim, interp):
    """d{x,y,z}{By,Ez}_interpOrder_<o>_<dim>d.txt == GridLayout::deriv applied by the device function of the field kernels"""
    o = 2 if (dim == 3 and interp == 3) else interp  # orders 2 and 3: same ghosts, same arrays (see the generator)
    ctx, L = ctxs(dim, interp), _layout_of_generators(dim, interp)
    for qname in ("By", "Ez"):
        src, window = _full(f"input_{qname}_O{o}_{dim}d", dim, None)
        for d in range(dim):
            key = f"d{'xyz'[d]}{qname}_interpOrder_{o}_{dim}d"
            want = G[key + "/expected"]
            out_shape = tuple(int(x) for x in G[key + "/shape"])
            got = _probe(ctx, L, 0, QTY[qname], d, src, out_shape)
            if window is not None:
                got = got[window]
            mask = ~np.isnan(got)
            assert mask.any()
            scale = np.max(np.abs(want))
            # the golden divides by dx, GridLayout::deriv multiplies by 1/dx: one rounding apart (the reference's own
            # C++ test accepts 1e-12 absolute, gridlayout_deriv.cpp)
            assert np.max(np.abs(got[mask] - want[mask])) <= 1e-12 * max(scale, 1.0), key
            # (the probe writes the physical box of the derivative; the generator also fills the ghost rows of the other
            # directions, which GridLayout::deriv callers never evaluate: evalOnBox, gridlayout.hpp:1198-1206)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("interp", [1, 2, 3])
def test_laplacian(ctxs, dim, interp):
    o = 2 if (dim == 3 and interp == 3) else interp
    ctx, L = ctxs(dim, interp), _layout_of_generators(dim, interp)
    for qname in ("Jx", "Jy", "Jz"):
        src, window = _full(f"input_{qname}_O{o}_{dim}d", dim, None)
        key = f"lap{qname}_interpOrder_{o}_{dim}d"
        want = G[key + "/expected"]
        got = _probe(ctx, L, 1, QTY[qname], 0, src, src.shape)
        if window is not None:
            got = got[window]
        mask = ~np.isnan(got)
        assert mask.any()
        scale = np.max(np.abs(want))
        assert np.max(np.abs(got[mask] - want[mask])) <= 1e-11 * max(scale, 1.0), key


CENT = dict(moment="ppp", Moment="ppp", Bx="pdd", By="dpd", Bz="ddp", Ex="dpp", Ey="pdp", Ez="ppd", Jx="dpp", Jy="pdp",
            Jz="ppd")
SRC_QTY = dict(moment=abi.RHO, Bx=abi.BX, By=abi.BY, Bz=abi.BZ, Ex=abi.EX, Ey=abi.EY, Ez=abi.EZ, Jx=abi.JX, Jy=abi.JY,
               Jz=abi.JZ)


@pytest.mark.gpu
@pytest.mark.parametrize("op", sorted(k[len("text/linear_coefs_yee_"):-4] for k in G.files if k.startswith("text/linear_coefs")))
def test_yee_linear_combinations(ctxs, op):
    """linear_coefs_yee_<A>To<B>.txt (dim, interp, n points, offsets..., coef): a unit impulse through the device
    projection gives back exactly the reference's stencil"""
    a, b = op.split("To")
    for r in rows(f"linear_coefs_yee_{op}.txt"):
        dim, interp, npts = int(r[0]), int(r[1]), int(r[2])
        offs = [tuple(int(x) for x in r[3 + p * dim:3 + (p + 1) * dim]) for p in range(npts)]
        coef = float(r[3 + npts * dim])
        kinds = [0 if CENT[a][d] == CENT[b][d] else (1 if CENT[a][d] == "p" else 2) for d in range(3)]
        ctx = ctxs(dim, interp)
        L = abi.make_layout(dim, interp, [6, 5, 4][:dim], [0.1] * dim)
        shape = ctx.field_shape(L, SRC_QTY[a])
        centre = tuple(s // 2 for s in shape)
        src = np.zeros(shape)
        src[centre] = 1.0
        got = _probe(ctx, L, 2, SRC_QTY[a], kinds[0] | kinds[1] << 2 | kinds[2] << 4, src, shape)
        nz = np.argwhere(np.nan_to_num(got) != 0.0)
        # out(i) = sum coef * in(i + off)  ->  the impulse shows up at centre - off
        found = sorted((tuple(int(centre[d] - i[d]) for d in range(dim)), float(got[tuple(i)])) for i in nz)
        assert found == sorted((o, coef) for o in offs), (op, dim, interp)
