"""TEST INFRASTRUCTURE — CPU back end for phare_b200.solver (same `ops` interface as GpuOps) built on the
oracle (phare_oracle.c).  Injected by tests to cross-check the step orchestration and the messenger
plans against the CUDA path, and by bench.py's cpu_baseline / --impl reference legs.  Never imported
by the product package."""
import ctypes as C

import numpy as np
import torch

from phare_b200 import abi
from . import Cpu, HostParticles, host_vec


class Arr:
    def __init__(self, a):
        self.a = a
        self.shape = tuple(a.shape)
        self.size = a.size

    @property
    def ptr(self):
        return self.a.ctypes.data


class Vec:
    def __init__(self, comps):
        self.comps = comps
        self.c = abi.VecField()
        for k in range(3):
            self.c.comp[k] = comps[k].a.ctypes.data

    def __getitem__(self, i):
        return self.comps[i]


class CpuOps:
    def __init__(self, dim, interp, impl="oracle"):
        self.cpu = Cpu(impl) if impl != "oracle" else Cpu("oracle")
        self.orc = Cpu("oracle")
        self.lib = self.orc.lib
        self.dim, self.interp = dim, interp
        self.err = 0
        self.last_error = ""
        self._ref = None

    # ---- memory
    def field(self, layout, qty):
        return Arr(np.zeros(self.orc.field_shape(layout, qty)))

    def vec(self, layout, qty0):
        return Vec([self.field(layout, qty0 + c) for c in range(3)])

    def zero(self, h):
        h.a[...] = 0.

    def copy(self, dst, src):
        dst.a[...] = src.a

    def set_field(self, h, host):
        h.a[...] = host

    def get_field(self, h):
        return h.a.copy()

    def particles(self, capacity):
        return HostParticles(self.dim, capacity)

    def set_particles(self, store, icell, delta, weight, charge, v):
        n = len(weight)
        for d in range(self.dim):
            store.icell[d][:n] = icell[:, d]
            store.delta[d][:n] = delta[:, d]
        for k in range(3):
            store.v[k][:n] = v[:, k]
        store.weight[:n] = weight
        store.charge[:n] = charge
        store.n = n

    def get_particles(self, store, first=0, last=None):
        last = store.n if last is None else last
        ic, de, w, q, v = store.soa()
        return ic[first:last], de[first:last], w[first:last], q[first:last], v[first:last]

    def count(self, store):
        return store.n

    def set_count(self, store, n):
        store.n = n

    def capacity(self, store):
        return int(store.c.capacity)

    def alias_weight_charge(self, store, other):
        return store  # the CPU push copies weight and charge like the reference

    def particles_copy(self, src, first, count, dst, dst_first):
        for d in range(self.dim):
            dst.icell[d][dst_first:dst_first + count] = src.icell[d][first:first + count]
            dst.delta[d][dst_first:dst_first + count] = src.delta[d][first:first + count]
        for k in range(3):
            dst.v[k][dst_first:dst_first + count] = src.v[k][first:first + count]
        dst.weight[dst_first:dst_first + count] = src.weight[first:first + count]
        dst.charge[dst_first:dst_first + count] = src.charge[first:first + count]
        dst.n = max(dst.n, dst_first + count)

    def cell_start(self, nkeys):
        return Arr(np.zeros(nkeys + 1, np.uint32))

    def bin_nkeys(self, layout, domain):
        self.lib.pho_bin_nkeys.restype = C.c_size_t
        return int(self.lib.pho_bin_nkeys(C.byref(layout), C.byref(domain)))

    # ---- operators
    def push(self, layout, E, B, pin, pout, mass, dt, first_selector=None):
        ed, ev = C.c_double(), C.c_double()
        fs = C.byref(first_selector) if first_selector is not None else None
        rc = self.lib.pho_push(C.byref(layout), C.byref(E.c), C.byref(B.c), C.byref(pin.c), C.byref(pout.c),
                               C.c_double(mass), C.c_double(dt), fs, C.byref(ed), C.byref(ev))
        if rc and not self.err:
            self.err, self.last_error = rc, f"Particle moved 2 cells with delta/vel: {ed.value:g}/{ev.value:g}"

    def deposit(self, layout, parts, rho_n, rho_q, F, coef=1.0, first=0, last=None, sel=(), domain=None,
                cell_start=None):
        last = parts.n if last is None else last
        rc = self.lib.pho_deposit(C.byref(layout), C.byref(parts.c), C.c_size_t(first), C.c_size_t(last),
                                  C.c_void_p(rho_n.ptr), C.c_void_p(rho_q.ptr), C.byref(F.c), C.c_double(coef),
                                  abi.box_array(list(sel)), C.c_int(len(sel)))
        assert rc == 0

    def push_deposit(self, layout, E, B, parts, mass, dt, rho_n, rho_q, F, coef=1.0, first=0, last=None, sel=(),
                     domain=None, cell_start=None, first_selector=None, write_back=True):
        """checker for phb_push_deposit, literally the reference's two passes: move the range (into a temporary
        when write_back is false), then deposit the moved range"""
        last = parts.n if last is None else last
        tmp = HostParticles(self.dim, max(parts.n, 1))
        self.push(layout, E, B, parts, tmp, mass, dt, first_selector)
        tmp.n = parts.n
        self.deposit(layout, tmp, rho_n, rho_q, F, coef, first, last, sel)
        if write_back and last > first:
            self.particles_copy(tmp, first, last - first, parts, first)

    def bin(self, layout, pin, pout, domain, keep, cell_start):
        counts = (C.c_size_t * 3)()
        rc = self.lib.pho_bin(C.byref(layout), C.byref(pin.c), C.byref(pout.c), C.byref(domain),
                              abi.box_array(list(keep)), C.c_int(len(keep)), C.c_void_p(cell_start.ptr), counts)
        assert rc == 0, rc
        return tuple(int(c) for c in counts)

    def bin_plan(self, layout, pin, domain, keep, cell_start_new):
        pass  # the checker bins in one go inside deposit_scatter

    def deposit_scatter(self, layout, pin, n_sorted, rho_n, rho_q, F, coef, sel, domain, cell_start_old, keep, pout,
                        cell_start_new):
        """checker for phb_bin_plan + phb_deposit_scatter: the reference's two separate steps"""
        self.deposit(layout, pin, rho_n, rho_q, F, coef, 0, pin.n, sel)
        self._counts = getattr(self, '_counts', {})
        self._counts[id(cell_start_new)] = self.bin(layout, pin, pout, domain, keep, cell_start_new)

    def bin_counts(self, layout, domain, cell_start, pout):
        return self._counts.pop(id(cell_start))

    def export(self, layout, src, first, last, box, dst, minus=None, shift=None):
        if not isinstance(box, abi.Box):
            box = abi.make_box(box.lo, box.hi)
        n = C.c_size_t()
        sh = (C.c_int * 3)(*([int(s) for s in shift] + [0] * (3 - len(shift)))) if shift is not None else None
        rc = self.lib.pho_export(C.byref(layout), C.byref(src.c), C.c_size_t(first), C.c_size_t(last), C.byref(box),
                                 C.byref(minus) if minus is not None else None, sh, C.byref(dst.c), C.byref(n))
        assert rc == 0, rc
        return int(n.value)

    def try_export(self, layout, src, first, last, box, dst, minus=None, shift=None):
        """export(), but a destination that is too small yields None (nothing appended) instead of an error"""
        n0 = dst.n
        n = C.c_size_t()
        sh = (C.c_int * 3)(*([int(s) for s in shift] + [0] * (3 - len(shift)))) if shift is not None else None
        rc = self.lib.pho_export(C.byref(layout), C.byref(src.c), C.c_size_t(first), C.c_size_t(last), C.byref(box),
                                 C.byref(minus) if minus is not None else None, sh, C.byref(dst.c), C.byref(n))
        if rc == abi.PHB_ERR_CAPACITY:
            dst.n = n0
            return None
        assert rc == 0, rc
        return int(n.value)

    def export_multi(self, layout, src, first, last, boxes, shifts, dsts):
        return [self.export(layout, src, first, last, b, d, shift=s) for b, s, d in zip(boxes, shifts, dsts)]

    def faraday(self, layout, B, E, Bnew, dt):
        assert self.lib.pho_faraday(C.byref(layout), C.byref(B.c), C.byref(E.c), C.byref(Bnew.c), C.c_double(dt)) == 0

    def ampere(self, layout, B, J):
        assert self.lib.pho_ampere(C.byref(layout), C.byref(B.c), C.byref(J.c)) == 0

    def ohm(self, layout, n, Ve, Pe, B, J, Enew, eta, nu, hyper_mode):
        assert self.lib.pho_ohm(C.byref(layout), C.c_void_p(n.ptr), C.byref(Ve.c), C.c_void_p(Pe.ptr), C.byref(B.c),
                                C.byref(J.c), C.byref(Enew.c), C.c_double(eta), C.c_double(nu), C.c_int(hyper_mode)) == 0

    def electrons_update(self, layout, Ne, Vi, J, Te, Ve, Pe):
        assert self.lib.pho_electrons_update(C.byref(layout), C.c_void_p(Ne.ptr), C.byref(Vi.c), C.byref(J.c),
                                             C.c_double(Te), C.byref(Ve.c), C.c_void_p(Pe.ptr)) == 0

    def ions_totals(self, rho_n, rho_q, flux, mass, rho_q_tot, rho_m_tot, V):
        npop = len(mass)
        pn = (C.c_void_p * npop)(*[a.ptr for a in rho_n])
        pq = (C.c_void_p * npop)(*[a.ptr for a in rho_q])
        fl = (abi.VecField * npop)(*[f.c for f in flux])
        ms = (C.c_double * npop)(*mass)
        assert self.lib.pho_ions_totals(C.c_size_t(rho_q_tot.size), C.c_int(npop), pn, pq, fl, ms,
                                        C.c_void_p(rho_q_tot.ptr), C.c_void_p(rho_m_tot.ptr), C.byref(V.c)) == 0

    def average(self, a, b, avg):
        assert self.lib.pho_average(C.c_size_t(a.size), C.c_void_p(a.ptr), C.c_void_p(b.ptr), C.c_void_p(avg.ptr)) == 0

    # ---- coarse <-> fine level operators
    def array(self, shape):
        return Arr(np.zeros(tuple(int(x) for x in shape)))

    def box_fill(self, arr, lo, ext, value):
        self.orc.box_fill(len(ext), arr.a, lo, ext, value)

    def field_refine(self, op, qty, coarse, coarse_lo, fine, fine_lo, box_lo, box_hi):
        self.orc.field_refine(len(box_lo), op, qty, coarse.a, coarse_lo, fine.a, fine_lo, box_lo, box_hi)

    def field_coarsen(self, op, qty, fine, fine_lo, coarse, coarse_lo, box_lo, box_hi):
        self.orc.field_coarsen(len(box_lo), op, qty, fine.a, fine_lo, coarse.a, coarse_lo, box_lo, box_hi)

    def magnetic_postprocess(self, layout, B, cell_lo, cell_hi, excluded=()):
        self.orc.magnetic_postprocess(layout, [c.a for c in B.comps], cell_lo, cell_hi, excluded)

    def axpy(self, dst, src, coef):
        self.orc.axpy(dst.a, src.a, coef)

    def split(self, nref, coarse, first, last, fine_boxes, fine):
        """checker for phb_split: the reference's toFineGrid + Splitter (oracle/_ref) on the range, then the
        destination-box filter of ParticlesRefineOperator::refine_ (particles_data_split.hpp:142-231)"""
        if self._ref is None:
            self._ref = Cpu("ref")
        from phare_b200.split import pattern
        _, _, maxd = pattern(self.dim, self.interp, nref)
        sub = HostParticles(self.dim, max(last - first, 1))
        self.particles_copy(coarse, first, last - first, sub, 0)
        sub.n = last - first
        ic, de, w, q, v = self._ref.split(self.dim, self.interp, nref, sub).soa()
        # candidates: the coarse particle moved to the fine grid lies within maxd cells of a destination box
        # (isInBox(splitBox, particleRefinedPos)); children are kept when inside the destination box itself
        ric = np.repeat(self._to_fine_icell(sub), nref, axis=0)
        keep = np.zeros(len(w), bool)
        for b in fine_boxes:
            cand = np.ones(len(w), bool)
            inb = np.ones(len(w), bool)
            for k in range(self.dim):
                cand &= (ric[:, k] >= b.lower[k] - maxd) & (ric[:, k] <= b.upper[k] + maxd)
                inb &= (ic[:, k] >= b.lower[k]) & (ic[:, k] <= b.upper[k])
            keep |= cand & inb
        n = int(keep.sum())
        if fine.n + n > self.capacity(fine):
            return None
        at = fine.n
        for d in range(self.dim):
            fine.icell[d][at:at + n] = ic[keep, d]
            fine.delta[d][at:at + n] = de[keep, d]
        for k in range(3):
            fine.v[k][at:at + n] = v[keep, k]
        fine.weight[at:at + n] = w[keep]
        fine.charge[at:at + n] = q[keep]
        fine.n = at + n
        return n

    def _to_fine_icell(self, parts):
        """toFineGrid (split.hpp:32-46): iCell*2 (+1 when delta*2 >= 1)"""
        ic, de, _, _, _ = parts.soa()
        fine_delta = de * 2
        return ic * 2 + np.floor(fine_delta).astype(np.int64)

    def poll_error(self):
        rc, self.err = self.err, 0
        return rc

    def sync(self):
        pass

    # ---- box ops
    def compile_box_ops(self, entries):
        return entries or None

    def run_box_ops(self, compiled):
        if not compiled:
            return
        u3 = lambda v: (C.c_uint32 * 3)(*([int(x) for x in v] + [1] * (3 - len(v))))
        for (dst, dlo, src, slo, ext, op) in compiled:
            rc = self.lib.pho_box_op(C.c_int(len(ext)), C.c_void_p(dst.ptr), u3(dst.shape), u3(dlo),
                                     C.c_void_p(src.ptr), u3(src.shape), u3(slo), u3(ext), C.c_int(op))
            assert rc == 0

    def new_buffer(self, n):
        return Arr(np.zeros(max(int(n), 0)))

    def buffer_slice(self, buf, off, ext):
        n = int(np.prod(ext))
        return Arr(buf.a[off:off + n].reshape([int(e) for e in ext]))

    def as_tensor(self, buf):
        return torch.from_numpy(buf.a if isinstance(buf, Arr) else buf)

    def size(self, buf):
        return (buf.a if isinstance(buf, Arr) else buf).size

    # ---- particle staging
    def staging_particles(self, layout, capacity):
        return HostParticles(self.dim, capacity)

    def grow_particles(self, layout, store, capacity):
        return store.copy(capacity=int(capacity * 1.5) + 16)

    def _columns(self, s):
        return list(s.icell) + list(s.delta) + list(s.v) + [s.weight, s.charge]

    def particle_bytes(self):
        return 4 * self.dim + 8 * self.dim + 24 + 16

    def pack_particles(self, layout, stores):
        chunks = []
        for ci in range(len(self._columns(stores[0]))):
            for s in stores:
                chunks.append(self._columns(s)[ci][:s.n].view(np.uint8))
        return np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)

    def new_particle_buffer(self, layout, total):
        return np.zeros(total * self.particle_bytes(), np.uint8)

    def unpack_particles(self, layout, buf, off, n, total, dst):
        base = 0
        for col in self._columns(dst):
            esz = col.itemsize
            raw = buf[base + off * esz: base + (off + n) * esz]
            col[dst.n:dst.n + n] = raw.view(col.dtype)
            base += total * esz
        dst.n = dst.n + n
