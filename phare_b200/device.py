"""Thin Python handle on the C ABI (include/phare_b200.h): a context, device arrays and the
device-resident particle store.  Every compute call goes straight to libphare_b200.so; there is no
numpy/torch implementation of any operator here, and nothing imports the oracle.
"""
import ctypes as C

import numpy as np

from . import abi


class PhbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[phb status {code}] {msg}")
        self.code = code


class DeviceArray:
    """A device allocation owned through phb_malloc."""

    def __init__(self, ctx, shape, dtype=np.float64, zero=True):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape))
        self.nbytes = self.size * self.dtype.itemsize
        p = C.c_void_p()
        ctx._check(ctx.lib.phb_malloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value
        if zero:
            self.zero()

    def zero(self):
        self.ctx._check(self.ctx.lib.phb_memset(self.ctx.h, self.ptr, 0, self.nbytes))

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == self.size, (host.shape, self.shape)
        self.ctx._check(self.ctx.lib.phb_h2d(self.ctx.h, self.ptr, host.ctypes.data, self.nbytes))
        self.ctx.sync()  # `host` may be a temporary
        return self

    def download(self):
        out = np.empty(self.shape, dtype=self.dtype)
        self.ctx._check(self.ctx.lib.phb_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.phb_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceVec:
    """Three components of a vector quantity (VecField)."""

    def __init__(self, ctx, layout, qty0, host=None):
        self.comps = []
        for c in range(3):
            a = DeviceArray(ctx, ctx.field_shape(layout, qty0 + c))
            if host is not None:
                a.upload(host[c])
            self.comps.append(a)
        self.c = abi.VecField()
        for c in range(3):
            self.c.comp[c] = self.comps[c].ptr

    def download(self):
        return [a.download() for a in self.comps]

    def __getitem__(self, i):
        return self.comps[i]


class DeviceParticles:
    """Device-resident SoA particle store (replaces ParticleArray<dim>)."""

    def __init__(self, ctx, capacity):
        self.ctx = ctx
        self.c = abi.Particles()
        ctx._check(ctx.lib.phb_particles_alloc(ctx.h, int(capacity), C.byref(self.c)))

    @property
    def n(self):
        return self.c.n

    @n.setter
    def n(self, v):
        self.c.n = int(v)

    @property
    def capacity(self):
        return self.c.capacity

    def upload_soa(self, icell, delta, weight, charge, v):
        """ContiguousParticles layout: icell (n,dim) int32, delta (n,dim), weight (n,), charge (n,), v (n,3)."""
        n = len(weight)
        icell = np.ascontiguousarray(icell, dtype=np.int32).reshape(n, self.ctx.dim)
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(n, self.ctx.dim)
        weight = np.ascontiguousarray(weight, dtype=np.float64)
        charge = np.ascontiguousarray(charge, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(n, 3)
        self.ctx._check(self.ctx.lib.phb_particles_from_soa(
            self.ctx.h, icell.ctypes.data, delta.ctypes.data, weight.ctypes.data, charge.ctypes.data,
            v.ctypes.data, n, C.byref(self.c)))
        self.ctx.sync()
        return self

    def download_soa(self):
        n, dim = self.c.n, self.ctx.dim
        icell = np.empty((n, dim), np.int32)
        delta = np.empty((n, dim))
        weight = np.empty(n)
        charge = np.empty(n)
        v = np.empty((n, 3))
        self.ctx._check(self.ctx.lib.phb_particles_to_soa(
            self.ctx.h, C.byref(self.c), icell.ctypes.data, delta.ctypes.data, weight.ctypes.data,
            charge.ctypes.data, v.ctypes.data))
        return icell, delta, weight, charge, v

    def upload_aos(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        stride = self.ctx.lib.phb_aos_stride(self.ctx.dim)
        n = raw.size // stride
        self.ctx._check(self.ctx.lib.phb_particles_from_aos(self.ctx.h, raw.ctypes.data, n, C.byref(self.c)))
        self.ctx.sync()
        return self

    def download_aos(self):
        stride = self.ctx.lib.phb_aos_stride(self.ctx.dim)
        raw = np.zeros(self.c.n * stride, np.uint8)
        self.ctx._check(self.ctx.lib.phb_particles_to_aos(self.ctx.h, C.byref(self.c), raw.ctypes.data))
        return raw

    def free(self):
        if self.c.weight:
            self.ctx.lib.phb_particles_free(self.ctx.h, C.byref(self.c))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """phb_ctx: one per (device, dim, interp_order)."""

    def __init__(self, dim, interp, device=0, stream=None):
        self.lib = abi.load()
        self.dim, self.interp = dim, interp
        h = C.c_void_p()
        rc = self.lib.phb_create(device, dim, interp, C.byref(h))
        if rc != abi.PHB_OK:
            raise PhbError(rc, self.lib.phb_last_error(None).decode())
        self.h = h
        if stream is not None:
            self._check(self.lib.phb_set_stream(self.h, C.c_void_p(stream)))

    @classmethod
    def adopt(cls, handle, dim, interp, stream=None):
        """a Python handle on a phb_ctx created (and destroyed) by someone else: the C++ level driver's
        (phare_b200/host_cpp.py)"""
        self = cls.__new__(cls)
        self.lib = abi.load()
        self.dim, self.interp = dim, interp
        self.h = C.c_void_p(handle)
        self._borrowed = True
        if stream is not None:
            self._check(self.lib.phb_set_stream(self.h, C.c_void_p(stream)))
        return self

    def _check(self, rc):
        if rc != abi.PHB_OK:
            raise PhbError(rc, self.lib.phb_last_error(self.h).decode())

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            self.lib.phb_destroy(self.h)
        self.h = None

    def sync(self):
        self._check(self.lib.phb_sync(self.h))

    def set_exact(self, exact):
        self._check(self.lib.phb_set_exact(self.h, 1 if exact else 0))

    def poll_error(self):
        self._check(self.lib.phb_poll_error(self.h))

    @property
    def launches(self):
        return int(self.lib.phb_launch_count(self.h))

    def field_shape(self, layout, qty):
        s = (C.c_uint32 * 3)()
        self.lib.phb_field_shape(C.byref(layout), qty, s)
        return tuple(int(s[d]) for d in range(layout.dim))

    # ---- operators (argument meaning == the reference functors, see include/phare_b200.h) ----
    def push(self, layout, E, B, pin, pout, mass, dt, first_selector=None):
        fs = C.byref(first_selector) if first_selector is not None else None
        self._check(self.lib.phb_push(self.h, C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(pin.c),
                                      C.byref(pout.c), mass, dt, fs))

    def bin(self, layout, pin, pout, domain, keep, cell_start):
        counts = (C.c_size_t * 3)()
        self._check(self.lib.phb_bin(self.h, C.byref(layout), C.byref(pin.c), C.byref(pout.c), C.byref(domain),
                                     abi.box_array(keep), len(keep), cell_start.ptr, counts))
        return tuple(int(c) for c in counts)

    def bin_plan(self, layout, pin, domain, keep, cell_start_new):
        self._check(self.lib.phb_bin_plan(self.h, C.byref(layout), C.byref(pin.c), C.byref(domain), abi.box_array(keep),
                                          len(keep), cell_start_new.ptr))

    def deposit_scatter(self, layout, pin, n_sorted, rho_n, rho_q, flux, coef, sel, domain, cell_start_old, keep, pout,
                        cell_start_new):
        self._check(self.lib.phb_deposit_scatter(
            self.h, C.byref(layout), C.byref(pin.c), int(n_sorted), rho_n.ptr, rho_q.ptr, C.byref(flux.c), coef,
            abi.box_array(list(sel)), len(sel), C.byref(domain),
            cell_start_old.ptr if cell_start_old is not None else None, abi.box_array(keep), len(keep),
            C.byref(pout.c), cell_start_new.ptr))

    def bin_counts(self, layout, domain, cell_start, pout):
        counts = (C.c_size_t * 3)()
        self._check(self.lib.phb_bin_counts(self.h, C.byref(layout), C.byref(domain), cell_start.ptr, counts,
                                            C.byref(pout.c)))
        return tuple(int(c) for c in counts)

    def split(self, coarse, first, last, deltas, weights, max_cell_distance, fine_boxes, fine):
        """phb_split: deltas (nref, dim) float32, weights (nref,) float32 — see phare_b200.split.pattern()"""
        d = np.ascontiguousarray(deltas, dtype=np.float32)
        w = np.ascontiguousarray(weights, dtype=np.float32)
        n = C.c_size_t()
        self._check(self.lib.phb_split(self.h, C.byref(coarse.c), first, last, len(w), d.ctypes.data, w.ctypes.data,
                                       int(max_cell_distance), abi.box_array(list(fine_boxes)), len(fine_boxes),
                                       C.byref(fine.c), C.byref(n)))
        return int(n.value)

    def bin_nkeys(self, layout, domain):
        return int(self.lib.phb_bin_nkeys(C.byref(layout), C.byref(domain)))

    def export(self, layout, src, first, last, box, dst, minus=None, shift=None):
        n = C.c_size_t()
        sh = (C.c_int * 3)(*([int(s) for s in shift] + [0] * (3 - len(shift)))) if shift is not None else None
        self._check(self.lib.phb_export(self.h, C.byref(layout), C.byref(src.c), first, last, C.byref(box),
                                        C.byref(minus) if minus is not None else None, sh, C.byref(dst.c),
                                        C.byref(n)))
        return int(n.value)

    def export_multi(self, layout, src, first, last, boxes, shifts, dsts):
        """boxes: abi.Box list; shifts: list of int triples; dsts: stores (may repeat). Returns counts per box."""
        nb = len(boxes)
        if nb == 0 or last <= first:
            return [0] * nb
        sh = (C.c_int * (3 * nb))()
        for k, s in enumerate(shifts):
            for d in range(len(s)):
                sh[3 * k + d] = int(s[d])
        ptrs = (C.POINTER(abi.Particles) * nb)(*[C.pointer(d.c) for d in dsts])
        out = (C.c_size_t * nb)()
        self._check(self.lib.phb_export_multi(self.h, C.byref(layout), C.byref(src.c), first, last, nb,
                                              abi.box_array(boxes), sh, ptrs, out))
        return [int(x) for x in out]

    def deposit(self, layout, parts, rho_n, rho_q, flux, coef=1.0, first=0, last=None, sel=(), domain=None,
                cell_start=None):
        last = parts.n if last is None else last
        self._check(self.lib.phb_deposit(
            self.h, C.byref(layout), C.byref(parts.c), first, last, rho_n.ptr, rho_q.ptr, C.byref(flux.c), coef,
            abi.box_array(list(sel)), len(sel), C.byref(domain) if domain is not None else None,
            cell_start.ptr if cell_start is not None else None))

    def push_deposit(self, layout, E, B, parts, mass, dt, rho_n, rho_q, flux, coef=1.0, first=0, last=None, sel=(),
                     domain=None, cell_start=None, first_selector=None, write_back=True):
        last = parts.n if last is None else last
        self._check(self.lib.phb_push_deposit(
            self.h, C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(parts.c), first, last, mass, dt,
            C.byref(first_selector) if first_selector is not None else None, 1 if write_back else 0,
            rho_n.ptr, rho_q.ptr, C.byref(flux.c), coef, abi.box_array(list(sel)), len(sel),
            C.byref(domain) if domain is not None else None, cell_start.ptr if cell_start is not None else None))

    def push_plan(self, layout, E, B, parts, mass, dt, domain, keep, cell_start_new, n_sorted=0, cell_start_old=None):
        """phb_push_plan: push in place + the count half of the re-binning in one pass (-> cell_start_new); with the
        current ordering (cell_start_old, n_sorted) the ordered part goes through the strip kernel"""
        self._check(self.lib.phb_push_plan(self.h, C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(parts.c),
                                           int(n_sorted), mass, dt, C.byref(domain),
                                           cell_start_old.ptr if cell_start_old is not None else None,
                                           abi.box_array(keep), len(keep), cell_start_new.ptr))

    def push_cells(self, layout, E, B, pin, pout, n_sorted, mass, dt, domain, cell_start):
        """phb_push_cells: K1 on the cell-ordered store, E,B block of each CTA staged in shared memory"""
        self._check(self.lib.phb_push_cells(self.h, C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(pin.c),
                                            C.byref(pout.c), int(n_sorted), mass, dt, C.byref(domain), cell_start.ptr))

    def push_deposit_plan(self, layout, E, B, parts, n_sorted, mass, dt, rho_n, rho_q, flux, coef, sel, domain,
                          cell_start_old, keep, cell_start_new):
        """phb_push_deposit_plan: push in place + deposit + the plan of the re-binning (-> cell_start_new)"""
        self._check(self.lib.phb_push_deposit_plan(
            self.h, C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(parts.c), int(n_sorted), mass, dt, rho_n.ptr,
            rho_q.ptr, C.byref(flux.c), coef, abi.box_array(list(sel)), len(sel), C.byref(domain),
            cell_start_old.ptr if cell_start_old is not None else None, abi.box_array(keep), len(keep),
            cell_start_new.ptr))

    def scatter_planned(self, layout, pin, n_sorted, domain, cell_start_old, keep, pout, cell_start_new):
        """phb_scatter_planned: the data movement of the re-binning planned by push_deposit_plan"""
        self._check(self.lib.phb_scatter_planned(
            self.h, C.byref(layout), C.byref(pin.c), int(n_sorted), C.byref(domain),
            cell_start_old.ptr if cell_start_old is not None else None, abi.box_array(keep), len(keep), C.byref(pout.c),
            cell_start_new.ptr))

    def set_predict_eps(self, eps):
        self._check(self.lib.phb_set_predict_eps(self.h, float(eps)))

    def predict_eps(self):
        return float(self.lib.phb_get_predict_eps(self.h))

    def predict_supported(self, layout):
        return bool(self.lib.phb_predict_supported(C.byref(layout)))

    def predict_plan_bytes(self, layout, domain, capacity):
        return int(self.lib.phb_predict_plan_bytes(C.byref(layout), C.byref(domain), int(capacity)))

    def push_deposit_predict(self, layout, E, B, parts, n_sorted, mass, dt, rho_n, rho_q, flux, coef, sel, domain,
                             cell_start, keep, plan_ptr, plan_bytes):
        """phb_push_deposit_predict: the domain_only sweep + the plan of the re-binning the all sweep will carry out"""
        self._check(self.lib.phb_push_deposit_predict(
            self.h, C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(parts.c), int(n_sorted), mass, dt, rho_n.ptr,
            rho_q.ptr, C.byref(flux.c), coef, abi.box_array(list(sel)), len(sel), C.byref(domain),
            cell_start.ptr if cell_start is not None else None, abi.box_array(keep), len(keep), plan_ptr, plan_bytes))

    def push_deposit_rebin(self, layout, E, B, pin, n_sorted, mass, dt, rho_n, rho_q, flux, coef, sel, domain,
                           cell_start_old, keep, pout, cell_start_new, plan_ptr, plan_bytes):
        """phb_push_deposit_rebin: move + deposit + re-binning in one pass, along the plan of push_deposit_predict"""
        self._check(self.lib.phb_push_deposit_rebin(
            self.h, C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(pin.c), int(n_sorted), mass, dt, rho_n.ptr,
            rho_q.ptr, C.byref(flux.c), coef, abi.box_array(list(sel)), len(sel), C.byref(domain),
            cell_start_old.ptr if cell_start_old is not None else None, abi.box_array(keep), len(keep), C.byref(pout.c),
            cell_start_new.ptr, plan_ptr, plan_bytes))

    def predict_counts(self, layout, domain, cell_start, plan_ptr, pout):
        counts = (C.c_size_t * 4)()
        self._check(self.lib.phb_predict_counts(self.h, C.byref(layout), C.byref(domain), cell_start.ptr, plan_ptr, counts,
                                                C.byref(pout.c)))
        return tuple(int(c) for c in counts)

    def maxwellian_load(self, layout, d_n, d_V, d_Vth, d_first, total, charge, ppc, seed, domain_cells, store):
        """phb_maxwellian_load: d_n / d_first device arrays, d_V / d_Vth objects with a .c VecField of per-cell arrays"""
        self._check(self.lib.phb_maxwellian_load(self.h, C.byref(layout), d_n.ptr, C.byref(d_V.c), C.byref(d_Vth.c),
                                                 d_first.ptr, int(total), charge, int(ppc), int(seed) & (2 ** 64 - 1),
                                                 (C.c_uint32 * 3)(*([int(c) for c in domain_cells] + [1] * (3 - len(domain_cells)))),
                                                 C.byref(store.c)))

    def faraday(self, layout, B, E, Bnew, dt):
        self._check(self.lib.phb_faraday(self.h, C.byref(layout), C.byref(B.c), C.byref(E.c), C.byref(Bnew.c), dt))

    def ampere(self, layout, B, J):
        self._check(self.lib.phb_ampere(self.h, C.byref(layout), C.byref(B.c), C.byref(J.c)))

    def ohm(self, layout, n, Ve, Pe, B, J, Enew, eta, nu, hyper_mode=0):
        self._check(self.lib.phb_ohm(self.h, C.byref(layout), n.ptr, C.byref(Ve.c), Pe.ptr, C.byref(B.c),
                                     C.byref(J.c), C.byref(Enew.c), eta, nu, hyper_mode))

    def electrons_update(self, layout, Ne, Vi, J, Te, Ve, Pe):
        self._check(self.lib.phb_electrons_update(self.h, C.byref(layout), Ne.ptr, C.byref(Vi.c), C.byref(J.c), Te,
                                                  C.byref(Ve.c), Pe.ptr))

    def ions_totals(self, rho_n, rho_q, flux, mass, rho_q_tot, rho_m_tot, V):
        npop = len(mass)
        pn = (C.c_void_p * npop)(*[a.ptr for a in rho_n])
        pq = (C.c_void_p * npop)(*[a.ptr for a in rho_q])
        fl = (abi.VecField * npop)(*[f.c for f in flux])
        ms = (C.c_double * npop)(*mass)
        self._check(self.lib.phb_ions_totals(self.h, rho_q_tot.size, npop, pn, pq, fl, ms, rho_q_tot.ptr,
                                             rho_m_tot.ptr, C.byref(V.c)))

    def average(self, a, b, avg):
        self._check(self.lib.phb_average(self.h, a.size, a.ptr, b.ptr, avg.ptr))

    def average_many(self, triples):
        """[(a, b, avg), ...] (at most 8) in one launch"""
        k = len(triples)
        n = (C.c_size_t * k)(*[t[0].size for t in triples])
        col = lambda j: (C.c_void_p * k)(*[t[j].ptr for t in triples])
        self._check(self.lib.phb_average_many(self.h, k, n, col(0), col(1), col(2)))

    def box_op(self, dim, dst, dst_shape, dst_lo, src, src_shape, src_lo, extent, op):
        u3 = lambda v: (C.c_uint32 * 3)(*([int(x) for x in v] + [1] * (3 - len(v))))
        dptr = dst.ptr if hasattr(dst, "ptr") else dst
        sptr = src.ptr if hasattr(src, "ptr") else src
        self._check(self.lib.phb_box_op(self.h, dim, dptr, u3(dst_shape), u3(dst_lo), sptr, u3(src_shape),
                                        u3(src_lo), u3(extent), op))

    # ---- coarse <-> fine level operators; arrays are DeviceArrays, *_lo the AMR field index of element 0
    @staticmethod
    def _view(a, lo):
        return abi.make_view(a.ptr, a.shape, lo)

    def field_refine(self, op, qty, coarse, coarse_lo, fine, fine_lo, box_lo, box_hi):
        dim = len(box_lo)
        cv, fv, b = self._view(coarse, coarse_lo), self._view(fine, fine_lo), abi.make_box(box_lo, box_hi)
        self._check(self.lib.phb_field_refine(self.h, dim, op, qty, C.byref(cv), C.byref(fv), C.byref(b)))

    def field_coarsen(self, op, qty, fine, fine_lo, coarse, coarse_lo, box_lo, box_hi):
        dim = len(box_lo)
        fv, cv, b = self._view(fine, fine_lo), self._view(coarse, coarse_lo), abi.make_box(box_lo, box_hi)
        self._check(self.lib.phb_field_coarsen(self.h, dim, op, qty, C.byref(fv), C.byref(cv), C.byref(b)))

    def magnetic_postprocess(self, layout, B, cell_lo, cell_hi, excluded=()):
        b = abi.make_box(cell_lo, cell_hi)
        self._check(self.lib.phb_magnetic_postprocess(self.h, C.byref(layout), C.byref(B.c), C.byref(b),
                                                      abi.box_array(list(excluded)), len(excluded)))

    def box_fill(self, dst, lo, extent, value):
        u3 = lambda v: (C.c_uint32 * 3)(*([int(x) for x in v] + [1] * (3 - len(v))))
        self._check(self.lib.phb_box_fill(self.h, len(extent), dst.ptr, u3(dst.shape), u3(lo), u3(extent), float(value)))

    def axpy(self, dst, src, coef):
        self._check(self.lib.phb_axpy(self.h, dst.size, dst.ptr, src.ptr, float(coef)))
