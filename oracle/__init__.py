"""TEST INFRASTRUCTURE — Python handle on the CPU oracle.  NOT part of the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this package; phare_b200/ never does.

Two interchangeable back ends with identical call signatures (host numpy arrays):
  Cpu("oracle")  -> oracle/liboracle.so        plain-C restatement (phare_oracle.c), always buildable
  Cpu("ref")     -> oracle/_ref/libphare_ref.so the reference's own headers compiled in place from
                                                /root/reference (ref/ref_driver.cpp); built in the dev
                                                container, shipped as a binary to the GPU box
"""
import ctypes as C
import os
import subprocess

import numpy as np

from phare_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libphare_ref.so")


def build(ref=True):
    """make oracle [ref]; the reference wrapper is only rebuilt when /root/reference exists."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def have_ref():
    return os.path.exists(REF_SO)


class HostParticles:
    """SoA columns on the host + the phb_particles struct pointing at them."""

    def __init__(self, dim, capacity):
        self.dim = dim
        cap = max(int(capacity), 1)
        self.icell = [np.zeros(cap, np.int32) for _ in range(dim)]
        self.delta = [np.zeros(cap) for _ in range(dim)]
        self.v = [np.zeros(cap) for _ in range(3)]
        self.weight = np.zeros(cap)
        self.charge = np.zeros(cap)
        self.c = abi.Particles()
        for d in range(dim):
            self.c.icell[d] = self.icell[d].ctypes.data
            self.c.delta[d] = self.delta[d].ctypes.data
        for k in range(3):
            self.c.v[k] = self.v[k].ctypes.data
        self.c.weight = self.weight.ctypes.data
        self.c.charge = self.charge.ctypes.data
        self.c.n = 0
        self.c.capacity = cap

    @property
    def n(self):
        return int(self.c.n)

    @n.setter
    def n(self, v):
        self.c.n = int(v)

    @classmethod
    def from_soa(cls, icell, delta, weight, charge, v, capacity=None):
        n = len(weight)
        icell = np.asarray(icell).reshape(n, -1)
        dim = icell.shape[1]
        p = cls(dim, capacity if capacity is not None else n)
        delta = np.asarray(delta).reshape(n, dim)
        v = np.asarray(v).reshape(n, 3)
        for d in range(dim):
            p.icell[d][:n] = icell[:, d]
            p.delta[d][:n] = delta[:, d]
        for k in range(3):
            p.v[k][:n] = v[:, k]
        p.weight[:n] = weight
        p.charge[:n] = charge
        p.n = n
        return p

    def soa(self):
        n = self.n
        icell = np.stack([c[:n] for c in self.icell], axis=1) if n else np.zeros((0, self.dim), np.int32)
        delta = np.stack([c[:n] for c in self.delta], axis=1) if n else np.zeros((0, self.dim))
        v = np.stack([c[:n] for c in self.v], axis=1) if n else np.zeros((0, 3))
        return icell, delta, self.weight[:n].copy(), self.charge[:n].copy(), v

    def copy(self, capacity=None):
        return HostParticles.from_soa(*self.soa(), capacity=capacity if capacity is not None else self.c.capacity)


def canonical_rows(icell, delta, weight, charge, v):
    """Canonical particle order for comparisons (the reference has no particle order, SURVEY §7 hard
    part 2): rows (icell..., delta..., v..., weight, charge) sorted lexicographically. Returns one
    structured byte view per particle so that equality is bit-exact."""
    n = len(weight)
    dim = icell.shape[1] if n else 1
    cols = [icell[:, d].astype(np.int64) for d in range(dim)]
    cols += [delta[:, d].view(np.int64) for d in range(dim)]
    cols += [v[:, k].view(np.int64) for k in range(3)]
    cols += [np.ascontiguousarray(weight).view(np.int64), np.ascontiguousarray(charge).view(np.int64)]
    m = np.stack(cols, axis=1) if n else np.zeros((0, 2 * dim + 5), np.int64)
    # sort by cell first (as signed ints), then by the raw bit patterns
    order = np.lexsort(m.T[::-1])
    return m[order]


def host_vec(arrs):
    vf = abi.VecField()
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
    for c in range(3):
        vf.comp[c] = keep[c].ctypes.data
    vf._keep = keep
    return vf


class Cpu:
    def __init__(self, impl="oracle"):
        self.impl = impl
        if impl == "oracle":
            if not os.path.exists(ORACLE_SO):
                build(ref=False)
            self.lib = C.CDLL(ORACLE_SO)
            self.pfx = "pho_"
        elif impl == "ref":
            if not os.path.exists(REF_SO):
                raise RuntimeError("oracle/_ref/libphare_ref.so not built (needs /root/reference): make -C oracle ref")
            self.lib = C.CDLL(REF_SO)
            self.pfx = "phr_"
            self.lib.phr_last_error.restype = C.c_char_p
        else:
            raise ValueError(impl)

    def _fn(self, name):
        fn = getattr(self.lib, self.pfx + name)
        fn.restype = C.c_int
        return fn

    def field_shape(self, layout, qty):
        s = (C.c_uint32 * 3)()
        lib = C.CDLL(ORACLE_SO)
        lib.pho_field_shape.restype = C.c_size_t
        lib.pho_field_shape(C.byref(layout), qty, s)
        return tuple(int(s[d]) for d in range(layout.dim))

    def zeros(self, layout, qty):
        return np.zeros(self.field_shape(layout, qty))

    def zeros_vec(self, layout, qty0):
        return [self.zeros(layout, qty0 + c) for c in range(3)]

    # ------------------------------------------------------------------ particles
    def push(self, layout, E, B, pin, mass, dt, first_selector=None, pout=None):
        pout = pout if pout is not None else HostParticles(layout.dim, pin.c.capacity)
        Ev, Bv = host_vec(E), host_vec(B)
        fs = C.byref(first_selector) if first_selector is not None else None
        if self.impl == "oracle":
            ed, ev = C.c_double(), C.c_double()
            rc = self._fn("push")(C.byref(layout), C.byref(Ev), C.byref(Bv), C.byref(pin.c), C.byref(pout.c),
                                  C.c_double(mass), C.c_double(dt), fs, C.byref(ed), C.byref(ev))
        else:
            rc = self._fn("push")(C.byref(layout), C.byref(Ev), C.byref(Bv), C.byref(pin.c), C.byref(pout.c),
                                  C.c_double(mass), C.c_double(dt), fs)
        return rc, pout

    def gather(self, layout, E, B, pin):
        Ev, Bv = host_vec(E), host_vec(B)
        eb = np.zeros((pin.n, 6))
        rc = self._fn("gather")(C.byref(layout), C.byref(Ev), C.byref(Bv), C.byref(pin.c), eb.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return eb

    def deposit(self, layout, parts, coef=1.0, first=0, last=None, sel=(), out=None):
        """returns [rho_n, rho_q, Fx, Fy, Fz] (accumulates into `out` when given)"""
        last = parts.n if last is None else last
        if out is None:
            out = [self.zeros(layout, abi.RHO) for _ in range(5)]
        flux = host_vec(out[2:5])
        for c in range(3):  # host_vec may have copied: point at the real arrays
            flux.comp[c] = out[2 + c].ctypes.data
        args = [C.byref(layout), C.byref(parts.c), C.c_size_t(first), C.c_size_t(last),
                out[0].ctypes.data_as(C.c_void_p), out[1].ctypes.data_as(C.c_void_p), C.byref(flux), C.c_double(coef)]
        if self.impl == "oracle":
            args += [abi.box_array(list(sel)), C.c_int(len(sel))]
        else:
            assert len(sel) == 0, "the reference deposit has no selection boxes"
        rc = self._fn("deposit")(*args)
        assert rc == 0
        return out

    def bin(self, layout, pin, domain, keep):
        assert self.impl == "oracle"
        self.lib.pho_bin_nkeys.restype = C.c_size_t
        nk = self.lib.pho_bin_nkeys(C.byref(layout), C.byref(domain))
        cell_start = np.zeros(nk + 1, np.uint32)
        counts = (C.c_size_t * 3)()
        pout = HostParticles(layout.dim, max(pin.n, 1))
        rc = self._fn("bin")(C.byref(layout), C.byref(pin.c), C.byref(pout.c), C.byref(domain),
                             abi.box_array(list(keep)), C.c_int(len(keep)), cell_start.ctypes.data_as(C.c_void_p),
                             counts)
        assert rc == 0
        return pout, cell_start, tuple(int(c) for c in counts)

    def export(self, layout, src, first, last, box, dst, minus=None, shift=None):
        assert self.impl == "oracle"
        n = C.c_size_t()
        sh = (C.c_int * 3)(*([int(s) for s in shift] + [0] * (3 - len(shift)))) if shift is not None else None
        rc = self._fn("export")(C.byref(layout), C.byref(src.c), C.c_size_t(first), C.c_size_t(last), C.byref(box),
                                C.byref(minus) if minus is not None else None, sh, C.byref(dst.c), C.byref(n))
        assert rc == 0, rc
        return int(n.value)

    # ------------------------------------------------------------------ fields
    def faraday(self, layout, B, E, dt, Bnew=None):
        Bnew = Bnew if Bnew is not None else self.zeros_vec(layout, abi.BX)
        b, e, bn = host_vec(B), host_vec(E), host_vec(Bnew)
        for c in range(3):
            bn.comp[c] = Bnew[c].ctypes.data
        assert self._fn("faraday")(C.byref(layout), C.byref(b), C.byref(e), C.byref(bn), C.c_double(dt)) == 0
        return Bnew

    def ampere(self, layout, B, J=None):
        J = J if J is not None else self.zeros_vec(layout, abi.JX)
        b, j = host_vec(B), host_vec(J)
        for c in range(3):
            j.comp[c] = J[c].ctypes.data
        assert self._fn("ampere")(C.byref(layout), C.byref(b), C.byref(j)) == 0
        return J

    def ohm(self, layout, n, Ve, Pe, B, J, eta, nu, hyper_mode=0, Enew=None):
        Enew = Enew if Enew is not None else self.zeros_vec(layout, abi.EX)
        n = np.ascontiguousarray(n)
        Pe = np.ascontiguousarray(Pe)
        ve, b, j, e = host_vec(Ve), host_vec(B), host_vec(J), host_vec(Enew)
        for c in range(3):
            e.comp[c] = Enew[c].ctypes.data
        rc = self._fn("ohm")(C.byref(layout), n.ctypes.data_as(C.c_void_p), C.byref(ve), Pe.ctypes.data_as(C.c_void_p),
                             C.byref(b), C.byref(j), C.byref(e), C.c_double(eta), C.c_double(nu), C.c_int(hyper_mode))
        assert rc == 0
        return Enew

    def electrons_update(self, layout, Ne, Vi, J, Te):
        Ne = np.ascontiguousarray(Ne)
        Ve = self.zeros_vec(layout, abi.VX)
        Pe = self.zeros(layout, abi.P)
        vi, j, ve = host_vec(Vi), host_vec(J), host_vec(Ve)
        for c in range(3):
            ve.comp[c] = Ve[c].ctypes.data
        rc = self._fn("electrons_update")(C.byref(layout), Ne.ctypes.data_as(C.c_void_p), C.byref(vi), C.byref(j),
                                          C.c_double(Te), C.byref(ve), Pe.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return Ve, Pe

    def ions_totals(self, layout, rho_n, rho_q, flux, mass):
        npop = len(mass)
        rho_n = [np.ascontiguousarray(a) for a in rho_n]
        rho_q = [np.ascontiguousarray(a) for a in rho_q]
        pn = (C.c_void_p * npop)(*[a.ctypes.data for a in rho_n])
        pq = (C.c_void_p * npop)(*[a.ctypes.data for a in rho_q])
        fvs = [host_vec(f) for f in flux]
        fl = (abi.VecField * npop)(*fvs)
        ms = (C.c_double * npop)(*mass)
        q = self.zeros(layout, abi.RHO)
        m = self.zeros(layout, abi.RHO)
        V = self.zeros_vec(layout, abi.VX)
        vv = host_vec(V)
        for c in range(3):
            vv.comp[c] = V[c].ctypes.data
        if self.impl == "oracle":
            rc = self._fn("ions_totals")(C.c_size_t(q.size), C.c_int(npop), pn, pq, fl, ms,
                                         q.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.byref(vv))
        else:
            rc = self._fn("ions_totals")(C.byref(layout), C.c_int(npop), pn, pq, fl, ms,
                                         q.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.byref(vv))
        assert rc == 0
        return q, m, V

    def average(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        out = np.zeros_like(a)
        rc = self._fn("average")(C.c_size_t(a.size), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                 out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return out

    def box_op(self, dim, dst, dst_lo, src, src_lo, extent, op):
        assert self.impl == "oracle"
        u3 = lambda v: (C.c_uint32 * 3)(*([int(x) for x in v] + [1] * (3 - len(v))))
        rc = self.lib.pho_box_op(C.c_int(dim), dst.ctypes.data_as(C.c_void_p), u3(dst.shape), u3(dst_lo),
                                 src.ctypes.data_as(C.c_void_p), u3(src.shape), u3(src_lo), u3(extent), C.c_int(op))
        assert rc == 0

    # ------------------------------------------------------------------ coarse <-> fine level operators
    @staticmethod
    def _view(a, lo):
        return abi.make_view(a.ctypes.data, a.shape, lo)

    def field_refine(self, dim, op, qty, coarse, coarse_lo, fine, fine_lo, box_lo, box_hi):
        """in place on `fine` (numpy, C order); *_lo = AMR field index of element 0; box = inclusive AMR field box"""
        assert self.impl == "oracle"
        cv, fv, b = self._view(coarse, coarse_lo), self._view(fine, fine_lo), abi.make_box(box_lo, box_hi)
        rc = self.lib.pho_field_refine(C.c_int(dim), C.c_int(op), C.c_int(qty), C.byref(cv), C.byref(fv), C.byref(b))
        assert rc == 0, rc

    def magnetic_postprocess(self, layout, B, cell_lo, cell_hi, excluded=()):
        assert self.impl == "oracle"
        bv = host_vec(B)
        for c in range(3):
            bv.comp[c] = B[c].ctypes.data
        b = abi.make_box(cell_lo, cell_hi)
        rc = self.lib.pho_magnetic_postprocess(C.byref(layout), C.byref(bv), C.byref(b), abi.box_array(list(excluded)),
                                               C.c_int(len(excluded)))
        assert rc == 0, rc

    def field_coarsen(self, dim, op, qty, fine, fine_lo, coarse, coarse_lo, box_lo, box_hi):
        assert self.impl == "oracle"
        fv, cv, b = self._view(fine, fine_lo), self._view(coarse, coarse_lo), abi.make_box(box_lo, box_hi)
        rc = self.lib.pho_field_coarsen(C.c_int(dim), C.c_int(op), C.c_int(qty), C.byref(fv), C.byref(cv), C.byref(b))
        assert rc == 0, rc

    def box_fill(self, dim, dst, dst_lo, extent, value):
        assert self.impl == "oracle"
        u3 = lambda v: (C.c_uint32 * 3)(*([int(x) for x in v] + [1] * (3 - len(v))))
        rc = self.lib.pho_box_fill(C.c_int(dim), dst.ctypes.data_as(C.c_void_p), u3(dst.shape), u3(dst_lo), u3(extent),
                                   C.c_double(value))
        assert rc == 0

    def axpy(self, dst, src, coef):
        assert self.impl == "oracle"
        rc = self.lib.pho_axpy(C.c_size_t(dst.size), dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p),
                               C.c_double(coef))
        assert rc == 0

    def weights(self, order, dual, local_cell, delta):
        w = (C.c_double * 4)()
        self.lib.pho_weights.restype = C.c_int
        start = self.lib.pho_weights(C.c_int(order), C.c_int(dual), C.c_uint32(local_cell), C.c_double(delta), w)
        return start, [w[i] for i in range(order + 1)]

    # ------------------------------------------------------------------ reference-only entry points
    def ion_update(self, layout, E, B, mass, domain, patch_ghost, level_ghost, non_level_ghost, dt, mode,
                   update_ions=True):
        """IonUpdater::updatePopulations (+updateIons) of the reference. Particle stores are updated in place;
        returns dict of moment arrays."""
        assert self.impl == "ref"
        npop = len(mass)
        Ev, Bv = host_vec(E), host_vec(B)
        rho_n = [self.zeros(layout, abi.RHO) for _ in range(npop)]
        rho_q = [self.zeros(layout, abi.RHO) for _ in range(npop)]
        flux = [self.zeros_vec(layout, abi.VX) for _ in range(npop)]
        fvs = []
        for f in flux:
            fv = host_vec(f)
            for c in range(3):
                fv.comp[c] = f[c].ctypes.data
            fvs.append(fv)
        pn = (C.c_void_p * npop)(*[a.ctypes.data for a in rho_n])
        pq = (C.c_void_p * npop)(*[a.ctypes.data for a in rho_q])
        fl = (abi.VecField * npop)(*fvs)
        ms = (C.c_double * npop)(*mass)
        q = self.zeros(layout, abi.RHO)
        m = self.zeros(layout, abi.RHO)
        V = self.zeros_vec(layout, abi.VX)
        vv = host_vec(V)
        for c in range(3):
            vv.comp[c] = V[c].ctypes.data
        dom = (abi.Particles * npop)(*[p.c for p in domain])
        pg = (abi.Particles * npop)(*[p.c for p in patch_ghost])
        lg = (abi.Particles * npop)(*[p.c for p in level_ghost])
        rc = self._fn("ion_update")(C.byref(layout), C.byref(Ev), C.byref(Bv), C.c_int(npop), ms, dom, pg, lg, pn, pq,
                                    fl, q.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.byref(vv),
                                    abi.box_array(list(non_level_ghost)), C.c_int(len(non_level_ghost)),
                                    C.c_double(dt), C.c_int(mode), C.c_int(1 if update_ions else 0))
        if rc != 0:
            raise RuntimeError(self.lib.phr_last_error().decode())
        for i in range(npop):  # sizes were updated in the struct copies
            domain[i].n, patch_ghost[i].n, level_ghost[i].n = dom[i].n, pg[i].n, lg[i].n
        self.lib.phr_last_seconds.restype = C.c_double
        return dict(rho_n=rho_n, rho_q=rho_q, flux=flux, rho_q_tot=q, rho_m_tot=m, V=V,
                    seconds=float(self.lib.phr_last_seconds()))

    def split_pattern(self, dim, interp, nref):
        """the reference Splitter<dim, interp, nref> flattened: (deltas float32 (nref, dim), weights, maxCellDistance)"""
        assert self.impl == "ref"
        d, w, m = np.zeros((nref, dim), np.float32), np.zeros(nref, np.float32), C.c_int()
        rc = self._fn("split_pattern")(dim, interp, nref, d.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                       C.byref(m))
        if rc != 0:
            raise KeyError((dim, interp, nref))
        return d, w, m.value

    def split(self, dim, interp, nref, parts):
        """toFineGrid + Splitter::operator() of the reference on every particle: nref refined particles each, in order"""
        assert self.impl == "ref"
        out = HostParticles(dim, max(parts.n * nref, 1))
        rc = self._fn("split")(dim, interp, nref, C.byref(parts.c), C.byref(out.c))
        assert rc == 0, rc
        return out

    def maxwellian(self, layout, n, V, Vth, charge, ppc, seed):
        """MaxwellianParticleInitializer::loadParticles with per-cell profile arrays (row-major over the patch)."""
        assert self.impl == "ref"
        ncell = int(np.prod([layout.ncells[d] for d in range(layout.dim)]))
        n = np.ascontiguousarray(n, dtype=np.float64).ravel()
        assert n.size == ncell
        Vk = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in V]
        Tk = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in Vth]
        pv = (C.c_void_p * 3)(*[a.ctypes.data for a in Vk])
        pt = (C.c_void_p * 3)(*[a.ctypes.data for a in Tk])
        out = HostParticles(layout.dim, ncell * ppc)
        rc = self._fn("maxwellian")(C.byref(layout), n.ctypes.data_as(C.c_void_p), pv, pt, C.c_double(charge),
                                    C.c_uint32(ppc), C.c_size_t(seed), C.byref(out.c))
        if rc != 0:
            raise RuntimeError(self.lib.phr_last_error().decode())
        return out
