"""Host-side restatement of the reference's step sequencing for ONE level of same-size or arbitrary
patches on a periodic domain, driving the C ABI (libphare_b200.so) kernel by kernel.

Mirrors, with the reference's names (file:line relative to the PHARE tree):
  SolverPPC::advanceLevel / predictor1_ / predictor2_ / corrector_ / average_ / moveIons_
                                   src/amr/solvers/solver_ppc.hpp:315-598
  SolverPPC::prepareStep           src/amr/solvers/solver_ppc.hpp:242-259
  IonUpdater::updatePopulations / updateIons      src/core/numerics/ion_updater/ion_updater.hpp:90-295
  HybridLevelInitializer::initialize (root level) src/amr/level_initializer/hybrid_level_initializer.hpp:100-182
  level transformers (per-patch dispatch, n/Ve/Pe wiring)  src/amr/solvers/solver_field_evolvers.hpp:33-77,
                                   src/amr/solvers/solver_hybrid_field_evolvers.hpp:29-41

The compute back end is injected (`ops`): GpuOps below (CUDA through the C ABI; the only one the
product ships) — the parity tests inject a CPU back end built on the oracle to cross-check the
orchestration.  There is no fallback: GpuOps raises if the CUDA library or a device is missing.
"""
import ctypes as C

import os

import numpy as np

from . import abi
from .boxes import Box
from .messenger import HybridMessenger, LevelGeom, LocalComm, PatchGeom

DOMAIN_ONLY, ALL = 1, 2  # UpdaterMode, ion_updater.hpp:22


# ---------------------------------------------------------------------------------------------------
class GpuOps:
    """libphare_b200.so back end; device memory is torch CUDA tensors (plumbing only)."""

    def __init__(self, dim, interp, device):
        import torch
        from .device import Context
        from . import torch_interop as ti
        self.torch, self.ti = torch, ti
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.ctx = Context(dim, interp, device=self.device.index or 0, stream=ti.current_stream_ptr())
        self.dim, self.interp = dim, interp
        self.kernel_timing = False  # bench.py: CUDA-event pairs around the particle kernels
        self.timed = {}

    def _timed(self, name, fn):
        if not self.kernel_timing:
            return fn()
        a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        self.timed.setdefault(name, []).append((a, b))
        return r

    # ---- memory
    def field(self, layout, qty):
        return self.ti.TorchArray(self.ctx.field_shape(layout, qty), self.device)

    def vec(self, layout, qty0):
        return self.ti.TorchVec(self.ctx, layout, qty0, self.device)

    def zero(self, h):
        self.ctx._check(self.ctx.lib.phb_memset(self.ctx.h, h.ptr, 0, h.size * h.t.element_size()))

    def copy(self, dst, src):
        self.ctx._check(self.ctx.lib.phb_d2d(self.ctx.h, dst.ptr, src.ptr, src.size * 8))

    def set_field(self, h, host):
        h.t.copy_(self.torch.from_numpy(np.array(host, dtype=np.float64, order="C", copy=True)).to(self.device))

    def get_field(self, h):
        return h.t.cpu().numpy()

    def particles(self, capacity):
        return self.ti.TorchParticles(self.dim, capacity, self.device)

    def set_particles(self, store, icell, delta, weight, charge, v):
        t = self.torch
        n = len(weight)
        for d in range(self.dim):
            store.icell[d][:n] = t.from_numpy(np.ascontiguousarray(icell[:, d], dtype=np.int32)).to(self.device)
            store.delta[d][:n] = t.from_numpy(np.ascontiguousarray(delta[:, d])).to(self.device)
        for k in range(3):
            store.v[k][:n] = t.from_numpy(np.ascontiguousarray(v[:, k])).to(self.device)
        store.weight[:n] = t.from_numpy(np.ascontiguousarray(weight)).to(self.device)
        store.charge[:n] = t.from_numpy(np.ascontiguousarray(charge)).to(self.device)
        store.n = n

    def get_particles(self, store, first=0, last=None):
        last = store.n if last is None else last
        sl = slice(first, last)
        icell = np.stack([c[sl].cpu().numpy() for c in store.icell], 1)
        delta = np.stack([c[sl].cpu().numpy() for c in store.delta], 1)
        v = np.stack([c[sl].cpu().numpy() for c in store.v], 1)
        return icell, delta, store.weight[sl].cpu().numpy(), store.charge[sl].cpu().numpy(), v

    def count(self, store):
        return store.n

    def set_count(self, store, n):
        store.n = n

    def capacity(self, store):
        return store.capacity

    def alias_weight_charge(self, store, other):
        """view of `store` whose weight/charge columns are `other`'s (a pushed copy never changes them,
        boris.hpp:197-199), so the out-of-place push does not have to write them"""
        class View:
            pass
        v = View()
        v.c = abi.Particles()
        C.memmove(C.byref(v.c), C.byref(store.c), C.sizeof(abi.Particles))
        v.c.weight, v.c.charge = other.c.weight, other.c.charge
        v.keep = (store, other)
        return v

    def particles_copy(self, src, first, count, dst, dst_first):
        self.ctx._check(self.ctx.lib.phb_particles_copy(self.ctx.h, C.byref(src.c), first, count, C.byref(dst.c),
                                                        dst_first))

    def cell_start(self, nkeys):
        return self.ti.TorchArray((nkeys + 1,), self.device, dtype=self.torch.int32)

    def bin_nkeys(self, layout, domain):
        return self.ctx.bin_nkeys(layout, domain)

    # ---- operators
    def push(self, layout, E, B, pin, pout, mass, dt, first_selector=None):
        self._timed("push", lambda: self.ctx.push(layout, E, B, pin, pout, mass, dt, first_selector))
        if hasattr(pout, "n") and hasattr(pin, "n"):
            pout.n = pin.n

    def deposit(self, layout, parts, rho_n, rho_q, F, coef=1.0, first=0, last=None, sel=(), domain=None,
                cell_start=None):
        self._timed("deposit" if cell_start is not None else "deposit_unordered",
                    lambda: self.ctx.deposit(layout, parts, rho_n, rho_q, F, coef, first, last, sel, domain, cell_start))

    def push_deposit(self, layout, E, B, parts, mass, dt, rho_n, rho_q, F, coef=1.0, first=0, last=None, sel=(),
                     domain=None, cell_start=None, first_selector=None, write_back=True):
        """K1+K3 fused (phb_push_deposit): move parts[first,last) and deposit the moved particles in one pass"""
        name = ("move" if cell_start is not None else "move_unordered") + ("_all" if write_back else "_domain_only")
        self._timed(name, lambda: self.ctx.push_deposit(layout, E, B, parts, mass, dt, rho_n, rho_q, F, coef, first,
                                                        last, sel, domain, cell_start, first_selector, write_back))

    def maxwellian_load(self, layout, n, V, Vth, first, total, charge, ppc, seed, domain_cells, store):
        """device loader (phb_maxwellian_load): per-cell host profiles are uploaded (7 small arrays), the
        particles are created in place in the store, already in cell order"""
        t, ti = self.torch, self.ti
        up = lambda a, dt: ti.TorchArray(None, self.device, tensor=t.from_numpy(np.ascontiguousarray(a)).to(self.device, dt))

        class CellVec:
            def __init__(self, arrs):
                self.keep = [up(a, t.float64) for a in arrs]
                self.c = abi.VecField()
                for k in range(3):
                    self.c.comp[k] = self.keep[k].ptr
        d_n, d_first = up(n, t.float64), up(first.astype(np.int64), t.int32)
        self.ctx.maxwellian_load(layout, d_n, CellVec(V), CellVec(Vth), d_first, total, charge, ppc, seed,
                                 domain_cells, store)
        self.ctx.sync()  # the uploaded temporaries die with this frame

    def bin(self, layout, pin, pout, domain, keep, cell_start):
        return self._timed("bin", lambda: self.ctx.bin(layout, pin, pout, domain, keep, cell_start))

    def push_plan(self, layout, E, B, parts, mass, dt, domain, keep, cell_start_new, n_sorted=0, cell_start_old=None):
        """K1 in place with the count of the re-binning folded in (phb_push_plan); strip kernel on the ordered part"""
        self._timed("push", lambda: self.ctx.push_plan(layout, E, B, parts, mass, dt, domain, keep, cell_start_new,
                                                       n_sorted, cell_start_old))

    def bin_plan(self, layout, pin, domain, keep, cell_start_new):
        self._timed("bin_plan", lambda: self.ctx.bin_plan(layout, pin, domain, keep, cell_start_new))

    def deposit_scatter(self, layout, pin, n_sorted, rho_n, rho_q, F, coef, sel, domain, cell_start_old, keep, pout,
                        cell_start_new):
        """K3+K2 fused (phb_deposit_scatter): the scatter pass of the re-binning carries the moment deposit"""
        self._timed("deposit_scatter", lambda: self.ctx.deposit_scatter(
            layout, pin, n_sorted, rho_n, rho_q, F, coef, sel, domain, cell_start_old, keep, pout, cell_start_new))

    def bin_counts(self, layout, domain, cell_start, pout):
        return self.ctx.bin_counts(layout, domain, cell_start, pout)

    # ---- predicted re-binning (csrc/predict.cu): the all sweep in one pass, planned by the domain_only sweep
    def predict_supported(self, layout):
        return self.ctx.predict_supported(layout)

    def predict_plan(self, layout, domain, capacity):
        """device buffer for the plan of one particle store"""
        nbytes = self.ctx.predict_plan_bytes(layout, domain, capacity)
        buf = self.torch.empty((nbytes + 3) // 4, dtype=self.torch.int32, device=self.device)
        return buf, nbytes

    def push_deposit_predict(self, layout, E, B, parts, n_sorted, mass, dt, rho_n, rho_q, F, coef, sel, domain, cell_start,
                             keep, plan):
        self._timed("move_domain_only", lambda: self.ctx.push_deposit_predict(
            layout, E, B, parts, n_sorted, mass, dt, rho_n, rho_q, F, coef, sel, domain, cell_start, keep,
            plan[0].data_ptr(), plan[1]))

    def push_deposit_rebin(self, layout, E, B, pin, n_sorted, mass, dt, rho_n, rho_q, F, coef, sel, domain,
                           cell_start_old, keep, pout, cell_start_new, plan):
        self._timed("move_all_rebin", lambda: self.ctx.push_deposit_rebin(
            layout, E, B, pin, n_sorted, mass, dt, rho_n, rho_q, F, coef, sel, domain, cell_start_old, keep, pout,
            cell_start_new, plan[0].data_ptr(), plan[1]))

    def predict_counts(self, layout, domain, cell_start, plan, pout):
        return self.ctx.predict_counts(layout, domain, cell_start, plan[0].data_ptr(), pout)

    def export(self, layout, src, first, last, box, dst, minus=None, shift=None):
        return self.ctx.export(layout, src, first, last, box, dst, minus, shift)

    def try_export(self, layout, src, first, last, box, dst, minus=None, shift=None):
        """export(), but a destination that is too small yields None (nothing appended) instead of an error"""
        from .device import PhbError
        try:
            return self.ctx.export(layout, src, first, last, box, dst, minus, shift)
        except PhbError as e:
            if e.code == abi.PHB_ERR_CAPACITY:
                return None
            raise

    def export_multi(self, layout, src, first, last, boxes, shifts, dsts):
        return self.ctx.export_multi(layout, src, first, last, boxes, shifts, dsts)

    def faraday(self, layout, B, E, Bnew, dt):
        self.ctx.faraday(layout, B, E, Bnew, dt)

    def ampere(self, layout, B, J):
        self.ctx.ampere(layout, B, J)

    def ohm(self, layout, n, Ve, Pe, B, J, Enew, eta, nu, hyper_mode):
        self.ctx.ohm(layout, n, Ve, Pe, B, J, Enew, eta, nu, hyper_mode)

    def electrons_update(self, layout, Ne, Vi, J, Te, Ve, Pe):
        self.ctx.electrons_update(layout, Ne, Vi, J, Te, Ve, Pe)

    def ions_totals(self, rho_n, rho_q, flux, mass, rho_q_tot, rho_m_tot, V):
        self.ctx.ions_totals(rho_n, rho_q, flux, mass, rho_q_tot, rho_m_tot, V)

    def average(self, a, b, avg):
        self.ctx.average(a, b, avg)

    def average_many(self, triples):
        self.ctx.average_many(triples)

    # ---- coarse <-> fine level operators (phare_b200.amr); *_lo = AMR field index of element 0 of the array
    def array(self, shape):
        return self.ti.TorchArray(tuple(int(s) for s in shape), self.device)

    def box_fill(self, arr, lo, ext, value):
        self.ctx.box_fill(arr, lo, ext, value)

    def field_refine(self, op, qty, coarse, coarse_lo, fine, fine_lo, box_lo, box_hi):
        self.ctx.field_refine(op, qty, coarse, coarse_lo, fine, fine_lo, box_lo, box_hi)

    def field_coarsen(self, op, qty, fine, fine_lo, coarse, coarse_lo, box_lo, box_hi):
        self.ctx.field_coarsen(op, qty, fine, fine_lo, coarse, coarse_lo, box_lo, box_hi)

    def magnetic_postprocess(self, layout, B, cell_lo, cell_hi, excluded=()):
        self.ctx.magnetic_postprocess(layout, B, cell_lo, cell_hi, excluded)

    def axpy(self, dst, src, coef):
        self.ctx.axpy(dst, src, coef)

    def split(self, nref, coarse, first, last, fine_boxes, fine):
        """phb_split with the Splitter<dim, interp, nref> pattern; returns the number appended, or None when `fine` is
        too small (nothing appended)"""
        from .split import pattern
        from .device import PhbError
        d, w, m = pattern(self.dim, self.interp, nref)
        try:
            return self.ctx.split(coarse, first, last, d, w, m, fine_boxes, fine)
        except PhbError as e:
            if e.code == abi.PHB_ERR_CAPACITY:
                return None
            raise

    def poll_error(self):
        """returns 0 or the phb status raised by a kernel (message in last_error)"""
        rc = self.ctx.lib.phb_poll_error(self.ctx.h)
        self.last_error = self.ctx.lib.phb_last_error(self.ctx.h).decode() if rc else ""
        return rc

    def sync(self):
        self.ctx.sync()

    # ---- batched box operations (K8)
    def compile_box_ops(self, entries):
        """entries: (dst handle, dst_lo, src handle, src_lo, extent, op) -> device descriptor table"""
        if not entries:
            return None
        descs = (abi.BoxDesc * len(entries))()
        first = 0
        for i, (dst, dlo, src, slo, ext, op) in enumerate(entries):
            d = descs[i]
            d.dst, d.src, d.op, d.first = dst.ptr, src.ptr, op, first
            nd = len(ext)
            for k in range(3):
                d.dst_shape[k] = dst.shape[k] if k < nd else 1
                d.src_shape[k] = src.shape[k] if k < nd else 1
                d.dst_lo[k] = int(dlo[k]) if k < nd else 0
                d.src_lo[k] = int(slo[k]) if k < nd else 0
                d.ext[k] = int(ext[k]) if k < nd else 1
            first += int(np.prod(ext))
        raw = np.frombuffer(bytes(descs), dtype=np.uint8)
        table = self.torch.from_numpy(raw.copy()).to(self.device)
        return (table, len(entries), first)

    def run_box_ops(self, compiled):
        if compiled is None:
            return
        table, n, total = compiled
        self.ctx._check(self.ctx.lib.phb_box_op_batch(self.ctx.h, table.data_ptr(), n, total))

    def new_buffer(self, n):
        return self.ti.TorchArray((max(int(n), 0),), self.device)

    def buffer_slice(self, buf, off, ext):
        t = buf.t[off:off + int(np.prod(ext))].view(*[int(e) for e in ext])
        return self.ti.TorchArray(None, self.device, tensor=t)

    def as_tensor(self, buf):
        return buf.t if isinstance(buf, self.ti.TorchArray) else buf

    def size(self, buf):
        return self.as_tensor(buf).numel()

    # ---- particle staging for migration between ranks
    def staging_particles(self, layout, capacity):
        return self.particles(capacity)

    def grow_particles(self, layout, store, capacity):
        new = self.particles(int(capacity * 1.5) + 16)
        self.ctx._check(self.ctx.lib.phb_particles_copy(self.ctx.h, C.byref(store.c), 0, store.n, C.byref(new.c), 0))
        new.n = store.n
        return new

    def _columns(self, store):
        return [(c, 4) for c in store.icell] + [(c, 8) for c in store.delta] + [(c, 8) for c in store.v] \
            + [(store.weight, 8), (store.charge, 8)]

    def particle_bytes(self):
        return 4 * self.dim + 8 * self.dim + 24 + 16

    def pack_particles(self, layout, stores):
        """flat message: every column holds the stores back to back (phb_particles_pack, one launch per store)"""
        total = sum(s.n for s in stores)
        buf = self.torch.empty(total * self.particle_bytes(), dtype=self.torch.uint8, device=self.device)
        off = 0
        for s in stores:
            self.ctx._check(self.ctx.lib.phb_particles_pack(self.ctx.h, C.byref(s.c), 0, s.n, buf.data_ptr(), total, off))
            off += s.n
        return buf

    def new_particle_buffer(self, layout, total):
        return self.torch.empty(total * self.particle_bytes(), dtype=self.torch.uint8, device=self.device)

    def unpack_particles(self, layout, buf, off, n, total, dst):
        if dst.n + n > dst.capacity:
            raise RuntimeError("particle store capacity exceeded while receiving migrating particles")
        self.ctx._check(self.ctx.lib.phb_particles_unpack(self.ctx.h, buf.data_ptr(), total, off, n, C.byref(dst.c)))


# ---------------------------------------------------------------------------------------------------
class Population:
    """IonPopulation (src/core/data/ions/ion_population/ion_population.hpp:19-140): moments + particle arrays"""

    def __init__(self, ops, layout, name, mass, capacity, nkeys):
        self.name, self.mass = name, mass
        self.rho_n, self.rho_q = ops.field(layout, abi.RHO), ops.field(layout, abi.RHO)
        self.flux = ops.vec(layout, abi.VX)
        self.scratch = [ops.field(layout, abi.RHO) for _ in range(5)]  # sumField_/sumVec_ of the messenger
        self.domain = ops.particles(capacity)      # domainParticles, cell-ordered in [0, n_sorted)
        self.spare = ops.particles(capacity)       # tmp_particles_ of IonUpdater / sort target
        self.cell_start = ops.cell_start(nkeys)
        self.cell_start_next = ops.cell_start(nkeys)  # written by bin_plan while the old order is still being read
        self.pending_bin = False                      # spare holds the re-binned store, counts not read back yet
        self.plan = None                              # predicted re-binning: (device buffer, bytes) of this store's plan
        self.predicted = None                         # (store, n, n_sorted) the pending plan was made for
        self.pending_predicted = False                # the pending re-binning followed a prediction: verify it
        self.n_sorted = 0
        self.patch_ghost = ops.particles(capacity // 8 + 4096)  # patchGhostParticles (leavers of this step)
        # levelGhostParticles (+Old/New, particle_pack.hpp:20-36): only exist on refined levels
        self.level_ghost = self.level_ghost_spare = self.level_ghost_old = self.level_ghost_new = None

    def set_level_ghosts(self, ops, capacity):
        self.level_ghost = ops.particles(capacity)
        self.level_ghost_spare = ops.particles(capacity)
        self.level_ghost_old = ops.particles(capacity)
        self.level_ghost_new = ops.particles(capacity)

    def moments(self):
        return [self.rho_n, self.rho_q, self.flux[0], self.flux[1], self.flux[2]]


class Patch:
    """one patch of the level: GridLayout + HybridState (src/core/models/hybrid_state.hpp:27-45) + solver scratch"""

    def __init__(self, ops, geom, layout, pops_spec, capacity_factor=1.3):
        self.geom, self.layout = geom, layout
        dim = layout.dim
        self.domain_box = abi.make_box(geom.box.lo, geom.box.hi)
        pg = 1 if layout.interp == 1 else 2
        self.ghost_box = abi.make_box(geom.box.lo - pg, geom.box.hi + pg)
        # nonLevelGhostBox (amr_utils.hpp:233-253): on a periodic single level every ghost cell has a neighbour
        self.non_level_ghost = [self.ghost_box]
        V = lambda q: ops.vec(layout, q)
        self.E, self.B, self.J = V(abi.EX), V(abi.BX), V(abi.JX)
        self.Epred, self.Bpred = V(abi.EX), V(abi.BX)      # electromagPred_
        self.Eavg, self.Bavg = V(abi.EX), V(abi.BX)        # electromagAvg_
        self.Bold = V(abi.BX)                              # Bold_ (prepareStep)
        self.Ve, self.Vi = V(abi.VX), V(abi.VX)
        self.Pe = ops.field(layout, abi.P)
        self.Ne = ops.field(layout, abi.RHO)               # ions.chargeDensity == electron density
        self.rho_m = ops.field(layout, abi.RHO)
        nkeys = ops.bin_nkeys(layout, self.domain_box)
        self.pops = [Population(ops, layout, s["name"], s["mass"], int(s["n"] * capacity_factor) + 4096, nkeys)
                     for s in pops_spec]


class IonUpdater:
    """IonUpdater<Ions, Electromag, GridLayout> (ion_updater.hpp:24-83) on the device-resident store.
    The pusher is named in dict["pusher"]["name"]; only "modified_boris" exists (pusher_factory.hpp:20-30)."""

    # (dim, interp) pairs where the one-pass kernel measured faster than push + deposit on B200 (tools/microbench.py):
    # in the `all` sweep (moved particles written back, re-binning as a separate pass) only in 1-D; in the domain_only
    # sweep (nothing written back) wherever the tile kernel exists (csrc/tile.cuh: E,B block staged in shared memory;
    # config 5: 5.0 ms against 3.8 + 2.3 ms for push + deposit, config 3: 2.3 against 3.6 ms)
    FUSED_AUTO = frozenset({(1, 1), (1, 2), (1, 3)})
    FUSED_DOMAIN_AUTO = frozenset({(1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (3, 1)})

    def __init__(self, ops, pusher_name="modified_boris", fused="auto", sort_with_deposit=True):
        if pusher_name != "modified_boris":
            raise RuntimeError("Error : Invalid Pusher name")
        self.ops = ops
        # sort_with_deposit: in the `all` sweep the deposit rides on the scatter pass of the re-binning
        # (phb_bin_plan + phb_deposit_scatter, K3+K2) instead of phb_deposit followed by phb_bin
        self.sort_with_deposit = sort_with_deposit
        self.defer_sort = False  # set per step by SolverPPC.advance_level
        # predict: the domain_only sweep plans the re-binning (the cell a particle ends in after the all sweep is the one
        # the domain_only sweep predicts unless it sits on a cell face), the all sweep is then ONE pass: move + deposit +
        # write to the planned slot (phb_push_deposit_predict / _rebin, csrc/predict.cu); PHB_PREDICT=0 switches it off
        self.predict = os.environ.get("PHB_PREDICT", "1") != "0"
        # ... for patches whose tile-kernel grid fills the GPU: a CTA owns 128 cells and two CTAs are resident per SM, so below
        # 2 x 148 CTAs the one-pass sweep runs on a fraction of the SMs while the streaming push / 16-lanes-per-cell scatter of
        # the two-pass path still spread out (measured, config 3 with 128 x 128-cell patches: 10.1 against 9.0 ms per step)
        self.predict_min_cells = int(os.environ.get("PHB_PREDICT_MIN_CELLS", 128 * 2 * 148))
        self.misfiled = 0        # plans that did not hold so far (each one costs a phb_bin of its store)
        self.rebin_fallbacks = 0
        # fused: one pass per array and sweep (phb_push_deposit, K1+K3) instead of phb_push then phb_deposit
        self.fused = fused

    def update_populations(self, patch, E, B, dt, mode):
        """IonUpdater::updatePopulations (ion_updater.hpp:90-109).  For mode == all the work is split in two
        halves that touch independent data: update_moments (push + deposit: everything the moments need) and
        maintain_arrays (re-binning, patch-ghost / level-ghost bookkeeping: only the particle arrays).
        SolverPPC runs the second half after the corrector so that its host synchronisations (counts) do not
        stall the field phases; calling this method runs both, exactly like the reference."""
        self.update_moments(patch, E, B, dt, mode)
        if mode == ALL:
            self.maintain_arrays(patch)

    def update_moments(self, patch, E, B, dt, mode):
        ops, L = self.ops, patch.layout
        for pop in patch.pops:
            for m in pop.moments():  # resetMoments (moments.hpp:15-23)
                ops.zero(m)
            n = ops.count(pop.domain)
            nlg = ops.count(pop.level_ghost) if pop.level_ghost is not None else 0
            fused = self.fused if isinstance(self.fused, bool) else (L.dim, L.interp) in (
                self.FUSED_DOMAIN_AUTO if mode == DOMAIN_ONLY else self.FUSED_AUTO)
            if mode == ALL and pop.predicted is not None:
                plan_for, pop.predicted = pop.predicted, None
                if plan_for == (id(pop.domain), n, pop.n_sorted) and not self.defer_sort:
                    # updateAndDepositAll_ (:228-295) in one pass along the plan of the domain_only sweep
                    ops.push_deposit_rebin(L, E, B, pop.domain, pop.n_sorted, pop.mass, dt, pop.rho_n, pop.rho_q, pop.flux,
                                           1.0, patch.non_level_ghost, patch.domain_box, pop.cell_start,
                                           patch.non_level_ghost, pop.spare, pop.cell_start_next, pop.plan)
                    pop.pending_bin = pop.pending_predicted = True
                    if nlg:
                        ops.push(L, E, B, pop.level_ghost, pop.level_ghost, pop.mass, dt, patch.ghost_box)
                        ops.deposit(L, pop.level_ghost, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0, nlg, [patch.domain_box])
                    continue
            if (fused and mode == DOMAIN_ONLY and n and self.predict and self.sort_with_deposit and not self.defer_sort
                    and hasattr(ops, "push_deposit_predict") and ops.predict_supported(L)
                    and int(np.prod([L.ncells[d] for d in range(L.dim)])) >= self.predict_min_cells):
                need = ops.capacity(pop.domain)
                if pop.plan is None or pop.plan[2] != need:
                    pop.plan = ops.predict_plan(L, patch.domain_box, need) + (need,)
                ops.push_deposit_predict(L, E, B, pop.domain, pop.n_sorted, pop.mass, dt, pop.rho_n, pop.rho_q, pop.flux,
                                         1.0, patch.non_level_ghost, patch.domain_box, pop.cell_start,
                                         patch.non_level_ghost, pop.plan)
                pop.predicted = (id(pop.domain), n, pop.n_sorted)
                if nlg:
                    ops.push_deposit(L, E, B, pop.level_ghost, pop.mass, dt, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0,
                                     nlg, [patch.domain_box], first_selector=patch.ghost_box, write_back=False)
                continue
            if fused:
                # both modes are the same pass; domain_only simply never stores the moved copy
                wb = mode == ALL
                if pop.n_sorted:
                    ops.push_deposit(L, E, B, pop.domain, pop.mass, dt, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0,
                                     pop.n_sorted, patch.non_level_ghost, patch.domain_box, pop.cell_start,
                                     write_back=wb)
                if n > pop.n_sorted:  # received since the last binning: not ordered yet
                    ops.push_deposit(L, E, B, pop.domain, pop.mass, dt, pop.rho_n, pop.rho_q, pop.flux, 1.0,
                                     pop.n_sorted, n, patch.non_level_ghost, write_back=wb)
                if nlg:
                    ops.push_deposit(L, E, B, pop.level_ghost, pop.mass, dt, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0,
                                     nlg, [patch.domain_box], first_selector=patch.ghost_box, write_back=wb)
            elif mode == DOMAIN_ONLY:
                # updateAndDepositDomain_ (:171-219): push a COPY (tmp_particles_), deposit those that end in the
                # nonLevelGhostBox; the domain array itself is untouched
                tmp = ops.alias_weight_charge(pop.spare, pop.domain)
                ops.push(L, E, B, pop.domain, tmp, pop.mass, dt)
                self._deposit(patch, pop, tmp, n)
                if nlg:
                    # pushAndAccumulateGhosts (:195-217): a copy of the level ghosts is pushed while inside the
                    # ghost box (first selector), those that end in the domain box are deposited
                    ops.push(L, E, B, pop.level_ghost, pop.level_ghost_spare, pop.mass, dt, patch.ghost_box)
                    ops.deposit(L, pop.level_ghost_spare, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0, nlg,
                                [patch.domain_box])
            else:
                # updateAndDepositAll_ (:228-295): push in place; stayers + leavers inside the nonLevelGhostBox
                # are deposited (= the domain + new patchGhost deposits of :290-293)
                planned = self.sort_with_deposit and not self.defer_sort and n
                if planned and hasattr(ops, "push_plan"):
                    # the keys of the re-binning are counted while the pushed particle is still in registers
                    ops.push_plan(L, E, B, pop.domain, pop.mass, dt, patch.domain_box, patch.non_level_ghost,
                                  pop.cell_start_next, pop.n_sorted, pop.cell_start)
                else:
                    ops.push(L, E, B, pop.domain, pop.domain, pop.mass, dt)
                    if planned:
                        ops.bin_plan(L, pop.domain, patch.domain_box, patch.non_level_ghost, pop.cell_start_next)
                if planned:
                    ops.deposit_scatter(L, pop.domain, pop.n_sorted, pop.rho_n, pop.rho_q, pop.flux, 1.0,
                                        patch.non_level_ghost, patch.domain_box, pop.cell_start, patch.non_level_ghost,
                                        pop.spare, pop.cell_start_next)
                    pop.pending_bin = True
                else:
                    self._deposit(patch, pop, pop.domain, n)
                if nlg:
                    # level ghosts (:275-288): pushed in place while inside the ghost box; those that entered the
                    # domain are deposited with the domain particles
                    ops.push(L, E, B, pop.level_ghost, pop.level_ghost, pop.mass, dt, patch.ghost_box)
                    ops.deposit(L, pop.level_ghost, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0, nlg, [patch.domain_box])

    def maintain_arrays(self, patch):
        """second half of updateAndDepositAll_: the store is re-binned into [domain | new patch ghosts | erased]
        (partition + erase, :245-273), level ghosts that entered the domain are appended to it and only those
        still in the ghost layer remain level ghosts (:281-287)"""
        ops, L = self.ops, patch.layout
        for pop in patch.pops:
            if pop.pending_bin and pop.pending_predicted:
                # the re-binning followed the plan of the domain_only sweep: class counts + the plans that did not hold
                *counts, misfiled = ops.predict_counts(L, patch.domain_box, pop.cell_start_next, pop.plan, pop.spare)
                pop.cell_start, pop.cell_start_next = pop.cell_start_next, pop.cell_start
                pop.pending_bin = pop.pending_predicted = False
                if misfiled:
                    # a particle is filed under the cell it was predicted to reach, not under its own (it was pushed
                    # and deposited correctly): restore the exact order before anything depends on it
                    self.misfiled += misfiled
                    self.rebin_fallbacks += 1
                    # the two sweeps of this run differ by more than eps assumed: widen the band of particles that wait for
                    # the final fields (4x per failure, at most a sixteenth of a cell)
                    if hasattr(ops, "ctx") and not os.environ.get("PHB_PREDICT_EPS"):
                        ops.ctx.set_predict_eps(min(4 * max(ops.ctx.predict_eps(), 2.0 ** -14), 1.0 / 16))
                    ops.set_count(pop.spare, sum(counts))
                    counts = ops.bin(L, pop.spare, pop.domain, patch.domain_box, patch.non_level_ghost, pop.cell_start)
                    pop.domain, pop.spare = pop.spare, pop.domain  # (swapped back below)
            elif pop.pending_bin:  # the scatter already happened with the deposit: only the counts are missing
                counts = ops.bin_counts(L, patch.domain_box, pop.cell_start_next, pop.spare)
                pop.cell_start, pop.cell_start_next = pop.cell_start_next, pop.cell_start
                pop.pending_bin = False
            else:
                counts = ops.bin(L, pop.domain, pop.spare, patch.domain_box, patch.non_level_ghost, pop.cell_start)
            pop.domain, pop.spare = pop.spare, pop.domain
            pop.n_sorted = counts[0]
            # "copy out new patch ghosts" (:248-254) then "drop all ghosts" (:273)
            if ops.capacity(pop.patch_ghost) < counts[1]:
                pop.patch_ghost = ops.particles(int(counts[1] * 1.5) + 4096)
            ops.particles_copy(pop.domain, counts[0], counts[1], pop.patch_ghost, 0)
            ops.set_count(pop.patch_ghost, counts[1])
            ops.set_count(pop.domain, counts[0])
            nlg = ops.count(pop.level_ghost) if pop.level_ghost is not None else 0
            if nlg:
                lg = pop.level_ghost
                ops.export(L, lg, 0, nlg, patch.domain_box, pop.domain)
                ops.set_count(pop.level_ghost_spare, 0)
                ops.export(L, lg, 0, nlg, patch.ghost_box, pop.level_ghost_spare, minus=patch.domain_box)
                pop.level_ghost, pop.level_ghost_spare = pop.level_ghost_spare, lg

    def fill_pop_moment_ghosts(self, patch, alpha):
        """fillIonPopMomentGhosts (hybrid_hybrid_messenger_strategy.hpp:508-547, level > 0): the level-ghost
        contribution to the border moments, time-interpolated between the coarse steps:
        deposit(levelGhostOld, coef = 1 - alpha) + deposit(levelGhostNew, coef = alpha)"""
        ops, L = self.ops, patch.layout
        for pop in patch.pops:
            for store, coef in ((pop.level_ghost_old, 1. - alpha), (pop.level_ghost_new, alpha)):
                if store is not None and ops.count(store):
                    ops.deposit(L, store, pop.rho_n, pop.rho_q, pop.flux, coef, 0, ops.count(store))

    def _deposit(self, patch, pop, store, n):
        ops, L = self.ops, patch.layout
        # cell-ordered part (tolerates particles that changed cell since the last binning) ...
        if pop.n_sorted:
            ops.deposit(L, store, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0, pop.n_sorted, patch.non_level_ghost,
                        patch.domain_box, pop.cell_start)
        # ... and the particles received from neighbours since then (appended, not ordered yet)
        if n > pop.n_sorted:
            ops.deposit(L, store, pop.rho_n, pop.rho_q, pop.flux, 1.0, pop.n_sorted, n, patch.non_level_ghost)

    def update_ions(self, patch):
        """Ions::computeChargeDensity + computeBulkVelocity (ion_updater.hpp:112-116)"""
        pops = patch.pops
        self.ops.ions_totals([p.rho_n for p in pops], [p.rho_q for p in pops], [p.flux for p in pops],
                             [p.mass for p in pops], patch.Ne, patch.rho_m, patch.Vi)


class SolverPPC:
    """SolverPPC<HybridModel, AMR_Types> (solver_ppc.hpp:31-186) for one periodic level."""

    def __init__(self, ops, patches, geom, comm=None, resistivity=0.0, hyper_resistivity=1e-4, hyper_mode=0, Te=0.12,
                 pusher_name="modified_boris", fused="auto", sort_with_deposit=True, messenger=None, npop=None):
        self.ops, self.patches, self.geom = ops, patches, geom
        self.comm = comm or LocalComm()
        # number of populations: given explicitly when this rank may own no patch of the level (refined levels), because
        # the per-population exchange phases are collective
        self.npop = npop if npop is not None else (len(patches[0].pops) if patches else 0)
        # a refined level brings its own messenger (phare_b200.amr.RefinedLevelMessenger: level ghosts from the coarser level)
        self.messenger = messenger or HybridMessenger(geom, ops, self.comm)
        # timeInterpCoef_ of fillIonPopMomentGhosts for the sweep being run; None on the root level (no level ghosts)
        self.level_ghost_alpha = None
        self.updater = IonUpdater(ops, pusher_name, fused, sort_with_deposit)
        self.eta, self.nu, self.hyper_mode, self.Te = resistivity, hyper_resistivity, hyper_mode, Te
        self.layouts = {p.geom.id: p.layout for p in patches}

    # ---- helpers
    def _by_id(self, attr):
        return {p.geom.id: getattr(p, attr) for p in self.patches}

    def _field_solve(self, Bsrc, Esrc, Bdst, Edst, dt, tag, after_B=None):
        """Faraday -> fillMagneticGhosts -> Ampere -> fillCurrentGhosts -> electrons.update -> Ohm
        (the common body of predictor1_/predictor2_/corrector_, solver_ppc.hpp:347-479).  after_B: called when Bdst is final"""
        ops, msg = self.ops, self.messenger
        for p in self.patches:
            ops.faraday(p.layout, getattr(p, Bsrc), getattr(p, Esrc), getattr(p, Bdst), dt)
        msg.fill_ghosts(Bdst, abi.BX, self._by_id(Bdst))
        if after_B is not None:
            after_B()
        for p in self.patches:
            ops.ampere(p.layout, getattr(p, Bdst), p.J)
        msg.fill_ghosts("J", abi.JX, self._by_id("J"))
        for p in self.patches:
            ops.electrons_update(p.layout, p.Ne, p.Vi, p.J, self.Te, p.Ve, p.Pe)
            ops.ohm(p.layout, p.Ne, p.Ve, p.Pe, getattr(p, Bdst), p.J, getattr(p, Edst), self.eta, self.nu,
                    self.hyper_mode)

    def _average(self):
        """average_ (solver_ppc.hpp:484-510)"""
        ops = self.ops
        for p in self.patches:
            triples = [(p.B[c], p.Bpred[c], p.Bavg[c]) for c in range(3)] + [(p.E[c], p.Epred[c], p.Eavg[c]) for c in range(3)]
            if hasattr(ops, "average_many"):
                ops.average_many(triples)  # the six components in one launch
            else:
                for t in triples:
                    ops.average(*t)
        self.messenger.fill_ghosts("Eavg", abi.EX, self._by_id("Eavg"))

    def _move_ions(self, dt, mode):
        """moveIons_ (solver_ppc.hpp:538-598).  The particle-array half of the `all` sweep (re-binning and
        fillIonGhostParticles) and the error vote are deferred to _finish_particles(): they do not feed the
        moments, and their host synchronisations would otherwise stall the field phases that follow."""
        ops, msg = self.ops, self.messenger
        for p in self.patches:
            self.updater.update_moments(p, p.Eavg, p.Bavg, dt, mode)
        npop = self.npop
        for i in range(npop):
            # fillFluxBorders + fillDensityBorders
            msg.sum_borders(f"pop{i}", {p.geom.id: p.pops[i].moments() for p in self.patches},
                            {p.geom.id: p.pops[i].scratch for p in self.patches})
        # fillIonPopMomentGhosts: level > 0 only (no-op on the root level)
        if self.level_ghost_alpha is not None:
            for p in self.patches:
                self.updater.fill_pop_moment_ghosts(p, self.level_ghost_alpha)
        for p in self.patches:
            self.updater.update_ions(p)
        # fillIonBorders: SetMax on total mass density, charge density and bulk velocity
        msg.max_borders("ions", {p.geom.id: [p.rho_m, p.Ne, p.Vi[0], p.Vi[1], p.Vi[2]] for p in self.patches})

    def _finish_particles(self):
        """deferred half of moveIons_(all): re-binning, fillIonGhostParticles + patchGhostParticles.clear()
        (:581-585), and mpi::any_errors() (:549-563) for both sweeps of the step"""
        ops, msg = self.ops, self.messenger
        if getattr(ops, "kernel_timing", False) and not getattr(self, "_in_finish", False):
            # bench.py: the host-synchronising tail of the step (counts, migration, error vote) as one bracket
            self._in_finish = True
            try:
                return ops._timed("finish_particles", self._finish_particles)
            finally:
                self._in_finish = False
        if getattr(ops, "kernel_timing", False):
            ops._timed("fp_maintain_arrays", lambda: [self.updater.maintain_arrays(p) for p in self.patches])
        else:
            for p in self.patches:
                self.updater.maintain_arrays(p)
        npop = self.npop
        # mpi::any_errors(): every kernel that can raise (the pushes, the exchange phases) has run, and the stream is idle
        # after the counts above.  With several ranks the vote rides on the count exchange of the first population.
        err = ops.poll_error()
        voted = None
        for i in range(npop):
            msg.migrate_particles(self.layouts,
                                  {p.geom.id: (p.pops[i].patch_ghost, 0, ops.count(p.pops[i].patch_ghost))
                                   for p in self.patches},
                                  {p.geom.id: p.pops[i].domain for p in self.patches}, vote=err if i == 0 else None)
            if i == 0 and self.comm.size > 1 and hasattr(msg, "last_vote"):
                voted = msg.last_vote
            for p in self.patches:
                ops.set_count(p.pops[i].patch_ghost, 0)
        for p in self.patches:
            for pop in p.pops:
                self._ensure_headroom(pop)
        if voted is None:
            voted = self.comm.allreduce_max(err)
        if voted:
            raise RuntimeError("Updater::updatePopulations: " + (getattr(ops, "last_error", "") or "error on another rank"))

    GROW_AT, GROW_BY = 0.85, 1.5

    def _ensure_headroom(self, pop):
        """the reference's ParticleArray is a std::vector and simply grows; the device stores are fixed-capacity, so a
        population that fills 85 % of its stores (density pile-up, e.g. a compressing current sheet) gets 1.5x larger ones"""
        ops = self.ops
        n, cap = ops.count(pop.domain), ops.capacity(pop.domain)
        if n <= self.GROW_AT * cap:
            return
        new_cap = int(cap * self.GROW_BY) + 4096
        bigger = ops.particles(new_cap)
        ops.particles_copy(pop.domain, 0, n, bigger, 0)
        ops.set_count(bigger, n)
        pop.domain, pop.spare = bigger, ops.particles(new_cap)

    # ---- public
    def prepare_step(self):
        """Bold <- B (solver_ppc.hpp:242-259)"""
        for p in self.patches:
            for c in range(3):
                self.ops.copy(p.Bold[c], p.B[c])

    # ---- refluxing (levels of a refined hierarchy; driven by phare_b200.amr.Hierarchy)
    def reset_flux_sum(self):
        """resetFluxSum (solver_ppc.hpp:279-295)"""
        for p in self.patches:
            for c in range(3):
                self.ops.zero(p.fluxSumE[c])

    def accumulate_flux_sum(self, coef):
        """accumulateFluxSum (solver_ppc.hpp:263-276): fluxSumE += Eavg * coef"""
        for p in self.patches:
            for c in range(3):
                self.ops.axpy(p.fluxSumE[c], p.Eavg[c], coef)

    def reflux(self, dt):
        """reflux (solver_ppc.hpp:298-311): B = Bold - dt curl(Eavg) with the refluxed Eavg, then fillMagneticGhosts"""
        for p in self.patches:
            self.ops.faraday(p.layout, p.Bold, p.Eavg, p.B, dt)
        self.messenger.fill_ghosts("B", abi.BX, self._by_id("B"))

    def advance_level(self, dt, staging=None):
        """solver_ppc.hpp:315-341.  `staging` (HostStaging): the step takes E,B from pinned host buffers and returns the
        moments and the new E,B to pinned host buffers.  The new E,B are on the critical path (the caller's next step
        starts from them), so they are read back first, as soon as the corrector has produced them, under the
        particle-array maintenance that ends the step; the moments are snapshotted on the device when the `all` sweep has
        produced them and travel to the host behind the fields, under the NEXT step (nothing waits for them but
        HostStaging.sync() / the next snapshot)."""
        # staging.defer_sort (the round-1 arrangement): the re-binning kept as a separate, deferred pass that the moment
        # read-back hides under; default: the one-pass `all` sweep (predicted re-binning), moments read back under the next step
        self.updater.defer_sort = staging is not None and staging.defer_sort
        if staging is not None:
            staging.upload()
        self.prepare_step()
        self._field_solve("B", "E", "Bpred", "Epred", dt, "predictor1")
        self._average()
        self._move_ions(dt, DOMAIN_ONLY)
        self._field_solve("B", "Eavg", "Bpred", "Epred", dt, "predictor2")
        self._average()
        self._move_ions(dt, ALL)
        if staging is not None:
            if staging.defer_sort:
                staging.download_moments()  # final after the `all` sweep (solver_ppc.hpp:333)
            else:
                staging.snapshot_moments()
        # the new B leaves for the host as soon as the corrector's Faraday has produced it, under Ampere / Ohm / the E ghost fill
        early_B = staging is not None and not staging.defer_sort
        self._field_solve("B", "Eavg", "B", "E", dt, "corrector", after_B=staging.download_B if early_B else None)
        self.messenger.fill_ghosts("E", abi.EX, self._by_id("E"))
        if staging is not None:
            staging.download_fields(skip_B=early_B)
            if not staging.defer_sort:
                staging.download_moments()
        self._finish_particles()
        if staging is not None:
            staging.join()

    def initialize(self):
        """HybridLevelInitializer::initialize, root level (hybrid_level_initializer.hpp:100-182): particles and B
        are already loaded; derive moments, J and E."""
        ops, msg = self.ops, self.messenger
        npop = self.npop
        for p in self.patches:
            for pop in p.pops:
                # first binning of the freshly loaded particles (the reference's CellMap is built on emplace_back)
                counts = ops.bin(p.layout, pop.domain, pop.spare, p.domain_box, p.non_level_ghost, pop.cell_start)
                pop.domain, pop.spare = pop.spare, pop.domain
                pop.n_sorted = counts[0]
                ops.set_count(pop.domain, counts[0])
                for m in pop.moments():
                    ops.zero(m)
                # depositParticles(DomainDeposit)
                ops.deposit(p.layout, pop.domain, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0, pop.n_sorted, (),
                            p.domain_box, pop.cell_start)
        for i in range(npop):
            msg.sum_borders(f"pop{i}", {p.geom.id: p.pops[i].moments() for p in self.patches},
                            {p.geom.id: p.pops[i].scratch for p in self.patches})
        for p in self.patches:
            self.updater.update_ions(p)
        msg.max_borders("ions", {p.geom.id: [p.rho_m, p.Ne, p.Vi[0], p.Vi[1], p.Vi[2]] for p in self.patches})
        for p in self.patches:
            ops.ampere(p.layout, p.B, p.J)
        msg.fill_ghosts("J", abi.JX, self._by_id("J"))
        for p in self.patches:
            ops.electrons_update(p.layout, p.Ne, p.Vi, p.J, self.Te, p.Ve, p.Pe)
            ops.ohm(p.layout, p.Ne, p.Ve, p.Pe, p.B, p.J, p.E, self.eta, self.nu, self.hyper_mode)
        msg.fill_ghosts("E", abi.EX, self._by_id("E"))
        self.prepare_step()


# ---------------------------------------------------------------------------------------------------
class HostStaging:
    """Host-buffer face of the step for callers whose fields live in host memory (the reference's FieldData
    buffers): pinned staging for the inputs (E, B of every local patch) and the results (per-population and
    total moments, new E and B), and a copy stream.  upload() is ordered before the step on the compute
    stream.  The new E,B go to the host as soon as the corrector has produced them (copy stream, underneath the
    particle-array maintenance that ends the step) and join() makes the compute stream wait for them: the caller's
    next step starts from them.  The moments are copied to a device-side snapshot when the `all` sweep has produced them
    (snapshot_moments, 0.1 ms) and read back from the snapshot BEHIND the fields, underneath the next step;
    they are on the host after sync() (or when the next step takes its own snapshot).
    defer_sort=True is the round-1 arrangement (moments read back at once, hidden under a separate re-binning pass)."""

    def __init__(self, ops, patches, defer_sort=False):
        t = ops.torch
        self.t, self.patches, self.ops = t, patches, ops
        self.defer_sort = defer_sort
        self.copy_stream = t.cuda.Stream(device=ops.device)
        self.inputs, self.moments, self.fields = [], [], []
        for p in patches:
            eb = [p.E[c] for c in range(3)] + [p.B[c] for c in range(3)]
            self.inputs += eb
            self.fields += eb
            self.moments += [p.Ne, p.rho_m, p.Vi[0], p.Vi[1], p.Vi[2]]
            for pop in p.pops:
                self.moments += pop.moments()
        pin = lambda arrs: [t.empty(a.t.shape, dtype=t.float64).pin_memory() for a in arrs]
        self.h_in, self.h_moments, self.h_fields = pin(self.inputs), pin(self.moments), pin(self.fields)
        self.snap = None if defer_sort else [t.empty_like(a.t) for a in self.moments]
        self.moments_done = self.fields_done = None
        self.array_done = {}  # index in fields / inputs -> event: that array's read-back has finished
        for h, a in zip(self.h_in, self.inputs):
            h.copy_(a.t)
        self.h2d_bytes = sum(h.numel() * 8 for h in self.h_in)
        self.d2h_bytes = sum(h.numel() * 8 for h in self.h_moments + self.h_fields)
        self.timing = False  # bench.py: CUDA-event pairs around each transfer (per-rank diagnosis of the staging cost)
        self.timed = {}

    def _mark(self, name):
        if not self.timing:
            return None
        a, b = self.t.cuda.Event(enable_timing=True), self.t.cuda.Event(enable_timing=True)
        self.timed.setdefault(name, []).append((a, b))
        a.record()
        return b

    def upload(self):
        """E,B of this step from the pinned host buffers.  The buffers may still be receiving the previous step's results
        (results_become_inputs): every array is uploaded as soon as ITS OWN read-back has finished, B first (it is read back
        first), so the upload trails the read-back by one array, in the other direction of the link."""
        done = self._mark("upload")
        cur = self.t.cuda.current_stream()
        order = [i for i in range(len(self.inputs)) if i % 6 >= 3] + [i for i in range(len(self.inputs)) if i % 6 < 3]
        for i in order:
            ev = self.array_done.get(i)
            if ev is not None:
                cur.wait_event(ev)
            self.inputs[i].t.copy_(self.h_in[i], non_blocking=True)
        if done is not None:
            done.record()

    def _download(self, hosts, srcs, name, each=None):
        """srcs -> hosts on the copy stream, ordered after everything enqueued so far on the compute stream; each: list
        receiving one event per array (recorded when that array has arrived)"""
        ready = self.t.cuda.Event()
        ready.record()
        with self.t.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            done = self._mark(name)
            for h, a in zip(hosts, srcs):
                h.copy_(a, non_blocking=True)
                if each is not None:
                    ev = self.t.cuda.Event()
                    ev.record()
                    each.append(ev)
            if done is not None:
                done.record()
            finished = self.t.cuda.Event()
            finished.record()
        return finished

    def snapshot_moments(self):
        """device-side copy of the moments the `all` sweep just produced (the next step zeroes the originals); the
        previous snapshot must have left for the host first"""
        if self.moments_done is not None:
            self.t.cuda.current_stream().wait_event(self.moments_done)
        for s, a in zip(self.snap, self.moments):
            s.copy_(a.t, non_blocking=True)

    def download_moments(self):
        srcs = self.snap if self.snap is not None else [a.t for a in self.moments]
        self.moments_done = self._download(self.h_moments, srcs, "download_moments")

    def _download_fields(self, idx, name):
        evs = []
        fin = self._download([self.h_fields[i] for i in idx], [self.fields[i].t for i in idx], name, each=evs)
        self.array_done.update(zip(idx, evs))
        return fin

    def download_B(self):
        """fields = 3 E then 3 B components per patch"""
        self._download_fields([i for i in range(len(self.fields)) if i % 6 >= 3], "download_B")

    def download_fields(self, skip_B=False):
        idx = [i for i in range(len(self.fields)) if not (skip_B and i % 6 >= 3)]
        self.fields_done = self._download_fields(idx, "download_fields")

    def timings_ms(self):
        """average milliseconds of each transfer since timing was switched on (call after a synchronize)"""
        return {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in self.timed.items() if v}

    def join(self):
        """the step is complete for the caller when the new E,B are on the host (and, with defer_sort, the moments)"""
        if self.defer_sort:
            self.t.cuda.current_stream().wait_stream(self.copy_stream)
        # otherwise nothing to do here: the next upload() waits for the read-back of each field it is about to re-send, and
        # sync() for everything

    def sync(self):
        """every result of every step issued so far is in the pinned host buffers"""
        self.copy_stream.synchronize()
        self.t.cuda.current_stream().synchronize()

    def results_become_inputs(self):
        """the host-side E,B of the next step are this step's results: swap the buffers (no transfer)"""
        self.h_in, self.h_fields = self.h_fields, self.h_in


def make_level(domain_cells, patch_grid, interp, dx, nranks=1):
    """Split a periodic domain into a Cartesian grid of patches and deal them to ranks in order.
    Returns (LevelGeom, [layout per patch])."""
    dim = len(domain_cells)
    edges = []
    for d in range(dim):
        n, k = domain_cells[d], patch_grid[d]
        cuts = [round(i * n / k) for i in range(k + 1)]
        edges.append(cuts)
    patches, layouts = [], []
    idx = np.ndindex(*patch_grid)
    npatch = int(np.prod(patch_grid))
    for pid, ijk in enumerate(idx):
        lo = [edges[d][ijk[d]] for d in range(dim)]
        hi = [edges[d][ijk[d] + 1] - 1 for d in range(dim)]
        owner = pid * nranks // npatch
        patches.append(PatchGeom(pid, Box(lo, hi), owner))
        ncells = [hi[d] - lo[d] + 1 for d in range(dim)]
        origin = [lo[d] * dx[d] for d in range(dim)]
        layouts.append(abi.make_layout(dim, interp, ncells, dx, amr_lower=lo, origin=origin))
    return LevelGeom(domain_cells, patches, interp), layouts
