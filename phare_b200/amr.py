"""Refined levels: a static hierarchy of patch levels with refinement ratio 2, advanced with the
reference's recursive sub-cycling, on top of the per-level SolverPPC of phare_b200.solver.

What is restated here (file:line relative to the PHARE tree; SAMRAI itself cannot be built in this image, so the
schedules it would create are explicit box plans):
  MultiPhysicsIntegrator::advanceLevel / standardLevelSynchronization / getMaxFinerLevelDt
                                   src/amr/multiphysics_integrator.hpp:435-450, 491-600
                                   (fine dt = coarse dt / ratio^2 -> 4 sub-steps per coarser step)
  HybridHybridMessengerStrategy    src/amr/messengers/hybrid_hybrid_messenger_strategy.hpp
      initLevel :335-362, fill{Magnetic,Electric,Current}Ghosts :376-400 (+ setNaNsOnFieldGhosts :924-968),
      fillIonPopMomentGhosts :508-547, firstStep :582-608, lastStep :619-645, synchronize :708-720, reflux :724-729,
      postSynchronize :736-751; refiner kinds src/amr/messengers/refiner.hpp:36-133 (GhostField = same level, then the
      next coarser level through the refine operator; static refiners: the coarse data as it is NOW, no time interpolation)
  HybridLevelInitializer::initialize (level > 0)   src/amr/level_initializer/hybrid_level_initializer.hpp:100-182
  SolverPPC::accumulateFluxSum / resetFluxSum / reflux     src/amr/solvers/solver_ppc.hpp:263-311
  ParticlesRefineOperator (interior / coarseBoundaryOld / coarseBoundaryNew)
                                   src/amr/data/particles/refine/particles_data_split.hpp:142-231
  makeNonLevelGhostBoxFor          src/amr/resources_manager/amr_utils.hpp:233-253
The field operators themselves are CUDA kernels behind the C ABI (csrc/level.cu, csrc/split.cu).

Scope: refinement boxes given by the user or by tagging (phare_b200.tagging; no load balancing); a refined level may reach
or span a periodic boundary of the root domain.  Patches of every level are dealt to the ranks (one per GPU).
"""
import numpy as np

from . import abi
from .boxes import Box
from .messenger import HybridMessenger, LevelGeom, LocalComm, PatchGeom, centering, PRIMAL
from .solver import Patch, SolverPPC

RATIO = 2                      # amr/amr_constants.hpp: refinementRatio
MAX_BOXES = 28                 # csrc/common.cuh: box lists passed to a kernel
SUBSTEPS = RATIO * RATIO       # getMaxFinerLevelDt: dt_fine = dt_coarse / ratio^2
DEFAULT_NREF = {1: 2, 2: 4, 3: 6}


def refine_box(b):
    return Box(b.lo * RATIO, b.hi * RATIO + (RATIO - 1))


def coarsen_box(b):
    return Box(b.lo // RATIO, b.hi // RATIO)  # floor division == toCoarseIndex (amr_utils.hpp:128-135)


def field_box(cells, qty):
    """FieldGeometry::toFieldBox (field_geometry.hpp:139-181): one more node on the upper side of primal directions"""
    hi = cells.hi.copy()
    for d in range(cells.dim):
        if centering(qty, d) == PRIMAL:
            hi[d] += 1
    return Box(cells.lo, hi)


def minus_all(boxes, removed):
    out = list(boxes)
    for r in removed:
        out = [piece for b in out for piece in b.minus(r)]
    return out


class RefinedLevelMessenger(HybridMessenger):
    """Messenger of a level > 0: same-level exchanges as on the root level (non periodic), plus everything that comes
    from the next coarser level: level-ghost field nodes (refine operators) and level-ghost particles (splitting).
    Every rank knows the geometry of every patch of both levels; a patch and its data live on its owner.  The
    coarse -> fine transfers are one-directional (the owner of a coarse patch never waits on the fine level), so this
    messenger always uses two-sided point-to-point messages, never the peer-memory arena of the root level."""

    REFINE_OP = {abi.BX: abi.REFINE_MAGNETIC, abi.EX: abi.REFINE_ELECTRIC, abi.JX: abi.REFINE_ELECTRIC}

    def __init__(self, geom, ops, comm, coarse_solver):
        super().__init__(geom, ops, comm, peer_halo=False)
        self.coarse = coarse_solver
        g = geom.g
        # the patches of the level and, when the level reaches a periodic boundary of the domain (geom.periodic), their
        # periodic images: a ghost cell covered by an image belongs to the level, not to the coarser one
        level_boxes = [p.box.shift(t) for p in geom.patches for t in geom.shifts]
        self._nan, self._scratch, self._lg_excluded, self.lg_particle_boxes = {}, {}, {}, {}
        for p in geom.patches:
            # level-ghost cells: the ghost layer minus every patch of the level (fields: g cells, particles: pg cells)
            self._lg_excluded[p.id] = [abi.make_box(ov.lo, ov.hi) for ov in (p.box.grow(g) * b for b in level_boxes)
                                       if ov is not None]
            self.lg_particle_boxes[p.id] = minus_all([p.box.grow(geom.pg)], level_boxes)
            for qty in range(abi.BX, abi.JZ + 1):
                gfb = p.ghost_field_box(qty, g)
                # coarse data under the ghost box, one node more for the two-point refine stencils
                cbox = Box(gfb.lo // RATIO - 1, gfb.hi // RATIO + 1)
                if p.owner != self.me:
                    self._scratch[(p.id, qty)] = (None, cbox)
                    continue
                # setNaNsOnFieldGhosts: ghost field box minus the field boxes of the level's patches
                nan = minus_all([gfb], [q.interior_field_box(qty).shift(t) for q in geom.patches for t in geom.shifts])
                self._nan[(p.id, qty)] = [(b.lo - gfb.lo, b.shape()) for b in nan]
                self._scratch[(p.id, qty)] = (ops.array(cbox.shape()), cbox)
        self._gather, self._nan_ops = {}, {}
        self._fine_patches = []

    def attach(self, fine_patches):
        """the Patch objects (data) of the fine patches this rank owns"""
        self._fine_patches = fine_patches

    # ---- coarse level -> scratch arrays
    def _gather_phase(self, name, qty0):
        """copies of the coarse level's `name` arrays into the scratch of every fine patch (what the RefineSchedule does
        before it calls the refine operator); a coarse node is taken from the first patch that owns it.  Compiled once
        per field into an exchange phase: local box copies, plus one packed message per (coarse owner, fine owner)."""
        if name not in self._gather:
            cs, me = self.coarse, self.me
            local, send_items, recv_items = [], {}, {}
            arrays = cs._by_id(name)
            for p in self.geom.patches:
                # A coarse node on a patch border has one copy per patch, and the synchronisation only refreshes the
                # copy of the patch the fine level overlaps: the coarse patches (images) under this fine patch come first.
                under = coarsen_box(p.box)
                sources = sorted(((q, t) for q in cs.geom.patches for t in cs.geom.shifts),
                                 key=lambda qt: 0 if under * qt[0].box.shift(qt[1]) is not None else 1)
                for c in range(3):
                    qty = qty0 + c
                    scratch, cbox = self._scratch[(p.id, qty)]
                    todo = [cbox]
                    for providers in ("interior", "ghost"):
                        for q, t in sources:
                            src = q.interior_field_box(qty) if providers == "interior" else q.ghost_field_box(qty, cs.geom.g)
                            src = src.shift(t)
                            rest = []
                            for piece in todo:
                                ov = piece * src
                                if ov is None:
                                    rest.append(piece)
                                    continue
                                dlo, slo, ext = ov.lo - cbox.lo, q.local(ov.lo - t, cs.geom.g), ov.shape()
                                if p.owner == me and q.owner == me:
                                    local.append((scratch, dlo, arrays[q.id][c], slo, ext, 0))
                                elif p.owner == me:
                                    recv_items.setdefault(q.owner, []).append((scratch, dlo, ext))
                                elif q.owner == me:
                                    send_items.setdefault(p.owner, []).append((arrays[q.id][c], slo, ext))
                                rest += piece.minus(ov)
                            todo = rest
                    if todo:
                        raise RuntimeError(f"refined patch {p.box} is not nested in the coarser level (uncovered {todo})")
            self._gather[name] = self._finish(local, send_items, recv_items, 0)
        return self._gather[name]

    def _refine(self, name, qty0, vecs, op, whole_ghost_box=False, excluded=None):
        ops, g = self.ops, self.geom.g
        self._run(self._gather_phase(name, qty0), ("gather", name))
        for patch in self._fine_patches:
            p = patch.geom
            for c in range(3):
                qty = qty0 + c
                scratch, cbox = self._scratch[(p.id, qty)]
                gfb = p.ghost_field_box(qty, g)
                ops.field_refine(op, qty, scratch, cbox.lo, vecs[p.id][c], gfb.lo, gfb.lo, gfb.hi)
            if qty0 == abi.BX:
                # MagneticRefinePatchStrategy::postprocessRefine on the cells that were filled from the coarser level:
                # the whole ghost box at level creation, otherwise the ghost box minus the level's patches (one launch
                # per component: the kernel skips the faces of the excluded cell boxes)
                cells = p.box.grow(g)
                skip = () if whole_ghost_box else (excluded[p.id] if excluded is not None else self._lg_excluded[p.id])
                ops.magnetic_postprocess(patch.layout, vecs[p.id], cells.lo, cells.hi, skip)

    # ---- HybridMessenger interface
    def fill_ghosts(self, name, qty0, vecs):
        """fillMagneticGhosts / fillElectricGhosts / fillCurrentGhosts on a refined level: NaNs on the level-ghost nodes,
        patch ghosts from the neighbours, what is still NaN from the coarser level.  Only the refiner registered under
        the name of the given field runs (RefinerPool::fill(vec, ...), refiner_pool.hpp:125-133): e.g. the level ghosts of
        the model's B are those of its last own fill (the corrector) until the next one."""
        if name not in self._nan_ops:  # every NaN box of every patch and component in one batched launch (K8, op 3)
            self._nan_ops[name] = self.ops.compile_box_ops(
                [(vecs[p.geom.id][c], lo, vecs[p.geom.id][c], lo, ext, 3) for p in self._fine_patches for c in range(3)
                 for lo, ext in self._nan[(p.geom.id, qty0 + c)]])
        self.ops.run_box_ops(self._nan_ops[name])
        HybridMessenger.fill_ghosts(self, name, qty0, vecs)
        self._refine(name, qty0, vecs, self.REFINE_OP[qty0])

    def fill_patch_ghosts(self, name, qty0, vecs):
        """same-level part only (patchGhostRefluxedSchedules, refiner.hpp PatchGhostField)"""
        HybridMessenger.fill_ghosts(self, name, qty0, vecs)

    def init_fields(self, solver):
        """initLevel (:335-345): B through MagneticFieldInitRefiner (+ post-process) over the whole ghost box, E through
        ElectricFieldRefiner (which only writes NaN nodes: a fresh FieldData is all NaN, field_data.hpp:41-59)"""
        ops = self.ops
        self._refine("B", abi.BX, solver._by_id("B"), abi.REFINE_MAGNETIC_INIT, whole_ghost_box=True)
        for p in solver.patches:
            for c in range(3):
                ops.box_fill(p.E[c], [0] * self.geom.dim, p.E[c].shape, float("nan"))
        self._refine("E", abi.EX, solver._by_id("E"), abi.REFINE_ELECTRIC)

    def regrid_fields(self, solver, old_solver):
        """regrid (:265-280): B (magneticRegriding_: BregridAlgo with the NaN-only MagneticFieldRefiner and the patch
        strategy, overwrite_interior) and E (electricInitRefiners_.regrid): a fresh, all-NaN array first takes what the
        old level holds (its interior field boxes), the remaining NaN nodes are refined from the coarser level, and the
        new fine faces of the cells that did not come from the old level get the Toth-Roe value"""
        ops, g, me = self.ops, self.geom.g, self.me
        old_geom = old_solver.geom.patches
        old_boxes = [q.box for q in old_geom]
        for name, qty0, op in (("B", abi.BX, abi.REFINE_MAGNETIC), ("E", abi.EX, abi.REFINE_ELECTRIC)):
            vecs, old_vecs = solver._by_id(name), old_solver._by_id(name)
            local, send_items, recv_items = [], {}, {}
            for p in solver.patches:
                for c in range(3):
                    ops.box_fill(vecs[p.geom.id][c], [0] * self.geom.dim, vecs[p.geom.id][c].shape, float("nan"))
            for pg in self.geom.patches:
                for c in range(3):
                    qty = qty0 + c
                    for qg in old_geom:
                        ov = pg.ghost_field_box(qty, g) * qg.interior_field_box(qty)
                        if ov is None:
                            continue
                        dlo, slo, ext = pg.local(ov.lo, g), qg.local(ov.lo, g), ov.shape()
                        if pg.owner == me and qg.owner == me:
                            local.append((vecs[pg.id][c], dlo, old_vecs[qg.id][c], slo, ext, 0))
                        elif pg.owner == me:
                            recv_items.setdefault(qg.owner, []).append((vecs[pg.id][c], dlo, ext))
                        elif qg.owner == me:
                            send_items.setdefault(pg.owner, []).append((old_vecs[qg.id][c], slo, ext))
            self._run(self._finish(local, send_items, recv_items, 0), ("regrid", name))
            excluded = {p.geom.id: [abi.make_box(ov.lo, ov.hi) for ov in (p.geom.box.grow(g) * b for b in old_boxes)
                                    if ov is not None] for p in solver.patches}
            self._refine(name, qty0, vecs, op, excluded=excluded)

    def split_from_coarser(self, ipop, nref, boxes_of, attr):
        """ParticlesRefineOperator::refine_: the coarser level's domain particles, moved to this level's index space and
        split, land in the store `attr` of the population of a fine patch when their cell lies in one of
        boxes_of(patch geometry).  Every rank splits the particles of the coarse patches it owns; children for a fine
        patch owned elsewhere are staged and shipped like migrating particles."""
        ops, me = self.ops, self.me
        mine = {p.geom.id: p for p in self._fine_patches}
        remote = {}

        def try_split(src, n, boxes, get, grow):
            while True:
                if ops.split(nref, src, 0, n, boxes, get()) is not None:
                    return
                grow()  # too small: nothing was appended

        for pg in self.geom.patches:
            fine_boxes = boxes_of(pg)
            boxes = [abi.make_box(b.lo, b.hi) for b in fine_boxes]
            if not boxes:
                continue
            if len(boxes) > MAX_BOXES:
                raise RuntimeError(f"{len(boxes)} destination boxes for one patch (the kernels take {MAX_BOXES})")
            reach = [coarsen_box(b.grow(RATIO * 2)) for b in fine_boxes]  # split stencil: <= 2 fine cells
            for q, t in ((q, t) for q in self.coarse.patches for t in self.coarse.geom.shifts):
                src = q.pops[ipop].domain
                n = ops.count(src)
                if n == 0 or not any(r * q.geom.box.shift(t) is not None for r in reach):
                    continue  # no particle of this coarse patch (image) can land in the destination boxes
                if t.any():
                    # a periodic image of the coarse patch: its particles near the destination, index-shifted like
                    # ParticlesData::pack does for a periodic overlap (particles_data.hpp:745-756), then split
                    image = ops.staging_particles(q.layout, 4096)
                    taken = []  # the reaches of neighbouring destination boxes overlap: every particle once
                    for r in reach:
                        part = r.shift(-t) * q.geom.box
                        for piece in (minus_all([part], taken) if part is not None else []):
                            taken.append(piece)
                            while ops.try_export(q.layout, src, 0, n, abi.make_box(piece.lo, piece.hi), image, None,
                                                 [int(x) for x in t]) is None:
                                image = ops.grow_particles(q.layout, image, 2 * ops.capacity(image) + 4096)
                    src, n = image, ops.count(image)
                    if n == 0:
                        continue
                if pg.owner == me:
                    patch = mine[pg.id]
                    try_split(src, n, boxes, lambda: getattr(patch.pops[ipop], attr),
                              lambda: grow_store(ops, patch, ipop, getattr(patch.pops[ipop], attr)))
                else:
                    key = (pg.owner, pg.id)
                    if key not in remote:
                        remote[key] = ops.staging_particles(q.layout, 4096)

                    def grow_remote(key=key, q=q):
                        remote[key] = ops.grow_particles(q.layout, remote[key], 2 * ops.capacity(remote[key]) + 4096)
                    try_split(src, n, boxes, lambda key=key: remote[key], grow_remote)
        if self.comm.size > 1:
            def ensure(pid, needed):
                patch = mine[pid]
                while ops.capacity(getattr(patch.pops[ipop], attr)) < needed:
                    grow_store(ops, patch, ipop, getattr(patch.pops[ipop], attr))
                return getattr(patch.pops[ipop], attr)
            self._exchange_particles({pid: p.layout for pid, p in mine.items()}, remote,
                                     {pid: getattr(p.pops[ipop], attr) for pid, p in mine.items()},
                                     {pid: 0 for pid in mine}, ensure)


def grow_store(ops, patch, ipop, store):
    """the reference's ParticleArray grows on push_back; the device stores are re-allocated 2x and swapped in place"""
    pop = patch.pops[ipop]
    new_cap = 2 * ops.capacity(store) + 4096
    for attr in ("domain", "level_ghost", "level_ghost_old", "level_ghost_new"):
        if getattr(pop, attr) is store:
            bigger = ops.particles(new_cap)
            n = ops.count(store)
            if n:
                ops.particles_copy(store, 0, n, bigger, 0)
            ops.set_count(bigger, n)
            setattr(pop, attr, bigger)
            if attr == "domain":
                pop.spare = ops.particles(new_cap)
            if attr == "level_ghost":
                pop.level_ghost_spare = ops.particles(new_cap)
            return
    raise RuntimeError("unknown particle store")


class Level:
    def __init__(self, number, geom, solver, interp, dx, origin, pops):
        self.number, self.geom, self.solver = number, geom, solver
        # what every rank must know about the level even when it owns none of its patches
        self.interp, self.dx, self.origin, self.pops = interp, dx, origin, pops  # origin: position of AMR index 0
        self.before_coarse_time = self.after_coarse_time = None  # beforePushCoarseTime_ / afterPushCoarseTime_
        self.coarser_times = None  # (start, end) of the coarser level's current step (subcycleStart/EndTimes_[i-1])
        self.old_time = 0.0        # SolverPPC::oldTime_[level]
        self.sync_phase = None     # compiled fine -> coarse exchange of the synchronisation (ranks > 1)


class Hierarchy:
    """The patch hierarchy + MultiPhysicsIntegrator: levels[0] is the periodic root level (an initialised SolverPPC),
    levels[i > 0] are refined levels built by add_level().  With several ranks (one per GPU) every patch of every level
    has an owner: a refined patch goes to the owner of the coarser patch under its lower corner; the exchanges between
    a level and the next coarser one (gather for the refine operators, split particles, coarsened data) cross ranks
    as packed point-to-point messages, like the same-level phases."""

    def __init__(self, ops, root_solver, nref=None):
        self.ops = ops
        self.comm = root_solver.comm
        if not root_solver.patches:
            raise RuntimeError("every rank must own at least one patch of the root level")
        L0 = root_solver.patches[0].layout
        dim = root_solver.geom.dim
        origin = [L0.origin[d] - L0.amr_lower[d] * L0.dx[d] for d in range(dim)]
        pops = [(pop.name, pop.mass) for pop in root_solver.patches[0].pops]
        self.levels = [Level(0, root_solver.geom, root_solver, L0.interp, [L0.dx[d] for d in range(dim)], origin, pops)]
        self.nref = nref or DEFAULT_NREF[dim]
        self.time = 0.0
        self._ensure_flux_sum(root_solver)

    def _ensure_flux_sum(self, solver):
        for p in solver.patches:
            if not hasattr(p, "fluxSumE"):
                p.fluxSumE = self.ops.vec(p.layout, abi.EX)  # SolverPPC::fluxSumE_ (solver_ppc.hpp:60)

    # ---------------------------------------------------------------------------------------- construction
    def add_level(self, fine_boxes, capacity_factor=1.6, old_level=None):
        """creates level len(levels) from cell boxes given in ITS OWN index space and initialises it from the current
        finest level (MultiPhysicsIntegrator::initializeLevelData -> HybridLevelInitializer::initialize, level > 0);
        old_level: the level it replaces (regrid())"""
        ops, me = self.ops, self.comm.rank
        coarse = self.levels[-1]
        cs = coarse.solver
        ilvl = len(self.levels)
        interp, dim = coarse.interp, cs.geom.dim
        fine_domain = tuple(s * RATIO for s in coarse.geom.domain_shape)
        dx = [coarse.dx[d] / RATIO for d in range(dim)]
        boxes = [b if isinstance(b, Box) else Box(*b) for b in fine_boxes]

        def owner_of(box):  # the owner of the coarser patch under the lower corner
            corner = Box(box.lo // RATIO, box.lo // RATIO)
            for q in coarse.geom.patches:
                if q.box.contains(corner):
                    return q.owner
            raise ValueError(f"refinement box {box} does not start inside the coarser level")
        patches_g = [PatchGeom(i, b, owner_of(b)) for i, b in enumerate(boxes)]
        # a level whose ghost layers (and split stencil) stay inside the domain never sees a periodic image; one that
        # reaches the boundary is periodic like the domain (its patches then see their images across it)
        margin = (2 if interp == 1 else 4) + 2 * RATIO
        reaches = any(np.any(b.lo - margin < 0) or np.any(b.hi + margin >= np.asarray(fine_domain)) for b in boxes)
        if reaches and not self.levels[0].geom.periodic:
            raise ValueError("a refinement box reaches the boundary of a non periodic domain")
        geom = LevelGeom(fine_domain, patches_g, interp, periodic=reaches)
        for pg in patches_g:
            if np.any(pg.box.lo < 0) or np.any(pg.box.hi >= np.asarray(fine_domain)):
                raise ValueError(f"refinement box {pg.box} is outside the domain")
            if np.any(pg.box.lo % RATIO) or np.any((pg.box.hi + 1) % RATIO):
                raise ValueError(f"refinement box {pg.box} is not aligned with the coarser cells")
        for a in patches_g:
            for b in patches_g:
                if a.id < b.id and a.box * b.box is not None:
                    raise ValueError("refinement boxes overlap")
        msg = RefinedLevelMessenger(geom, ops, self.comm, cs)
        # expected number of particles: nref children per coarse particle of the covered coarse cells (estimated from the
        # coarse patches of this rank; a store that turns out too small is re-allocated)
        npop = len(coarse.pops)
        ncoarse_cells = sum(int(np.prod([p.layout.ncells[d] for d in range(dim)])) for p in cs.patches)
        per_cell = [sum(ops.count(p.pops[i].domain) for p in cs.patches) / max(ncoarse_cells, 1) for i in range(npop)]
        patches = []
        for pg in patches_g:
            if pg.owner != me:
                continue
            ncells = [int(x) for x in pg.box.shape()]
            origin = [coarse.origin[d] + pg.box.lo[d] * dx[d] for d in range(dim)]
            L = abi.make_layout(dim, interp, ncells, dx, amr_lower=list(pg.box.lo), origin=origin, level=ilvl)
            ccells = int(np.prod(ncells)) / RATIO ** dim
            spec = [dict(name=coarse.pops[i][0], mass=coarse.pops[i][1], n=int(self.nref * per_cell[i] * ccells))
                    for i in range(npop)]
            patch = Patch(ops, pg, L, spec, capacity_factor=capacity_factor)
            # nonLevelGhostBox: the domain plus the part of the particle ghost layer that belongs to a neighbour patch
            patch.non_level_ghost = [patch.domain_box] + [
                abi.make_box(ov.lo, ov.hi) for ov in (pg.box.grow(geom.pg) * q.box.shift(t) for q in patches_g
                                                      for t in geom.shifts if q.id != pg.id or t.any())
                if ov is not None]
            if len(patch.non_level_ghost) > MAX_BOXES:
                raise RuntimeError(f"patch {pg.box} has {len(patch.non_level_ghost) - 1} neighbours (the kernels take {MAX_BOXES - 1})")
            lg_cells = sum(b.volume() for b in msg.lg_particle_boxes[pg.id]) / RATIO ** dim
            for i, pop in enumerate(patch.pops):
                pop.set_level_ghosts(ops, int(capacity_factor * self.nref * per_cell[i] * lg_cells) + 4096)
            patches.append(patch)
        msg.attach(patches)
        solver = SolverPPC(ops, patches, geom, self.comm, resistivity=cs.eta, hyper_resistivity=cs.nu,
                           hyper_mode=cs.hyper_mode, Te=cs.Te, fused=cs.updater.fused,
                           sort_with_deposit=cs.updater.sort_with_deposit, messenger=msg, npop=npop)
        self._ensure_flux_sum(solver)
        level = Level(ilvl, geom, solver, interp, dx, coarse.origin, coarse.pops)
        self.levels.append(level)
        self._initialize_level(level, old_level)
        return level

    def regrid(self, fine_boxes, capacity_factor=1.6):
        """replaces the finest level by one made of `fine_boxes` (cell boxes in its own index space; none: the level is
        removed).  HybridHybridMessengerStrategy::regrid (:265-311) + HybridLevelInitializer::initialize(isRegridding):
        where the new level overlaps the old one its E, B and domain particles are COPIED from it, everywhere else they
        come from the next coarser level exactly as at level creation (B: coarse faces + Toth-Roe on the cells that were
        not copied; E: electric refiner; particles: splitting); level ghosts and moments are rebuilt.  Old and new patches
        may live on different ranks: the copies then travel as packed point-to-point messages."""
        if len(self.levels) < 2:
            raise RuntimeError("there is no refined level to regrid")
        old = self.levels.pop()
        if not fine_boxes:
            return None
        return self.add_level(fine_boxes, capacity_factor, old_level=old)

    def regrid_tagged(self, tagger):
        """GriddingAlgorithm::regridAllFinerLevels driven by the tagger: level by level from the root, the boxes of level
        i+1 follow from the tags of level i (phare_b200.tagging); a level whose boxes (and whose coarser levels) did not
        change is kept as it is, otherwise it is rebuilt by regrid() semantics from its old self and the level below.
        Returns True when the hierarchy changed."""
        old_levels = self.levels[1:]
        self.levels = self.levels[:1]
        changed = False
        for il in range(tagger.max_nbr_levels - 1):
            boxes = tagger.boxes(self, il)
            old = old_levels[il] if il < len(old_levels) else None
            if not boxes:
                changed = changed or old is not None
                break
            fine = [refine_box(b) for b in boxes]
            same = (old is not None and not changed and len(fine) == len(old.geom.patches)
                    and all(a == b.box for a, b in zip(fine, old.geom.patches)))
            if same:
                self.levels.append(old)
            else:
                self.add_level(fine, old_level=old)
                changed = True
        return changed or len(self.levels) - 1 != len(old_levels)

    def _initialize_level(self, level, old=None):
        ops, s, msg = self.ops, level.solver, level.solver.messenger
        npop = s.npop
        if old is None:
            msg.init_fields(s)
        else:
            msg.regrid_fields(s, old.solver)
        old_boxes = [q.box for q in old.geom.patches] if old is not None else []
        for i in range(npop):
            if old is not None:
                # domain particles of the old level that lie in a patch of the new one are kept as they are (shipped to
                # the owner of the new patch when it is another rank)
                mine = {p.geom.id: p for p in s.patches}
                remote = {}
                for q in old.solver.patches:
                    n = ops.count(q.pops[i].domain)
                    for pg in level.geom.patches:
                        both = pg.box * q.geom.box
                        if both is None or n == 0:
                            continue
                        box = abi.make_box(both.lo, both.hi)
                        if pg.owner == self.comm.rank:
                            p = mine[pg.id]
                            while ops.capacity(p.pops[i].domain) < ops.count(p.pops[i].domain) + n:
                                grow_store(ops, p, i, p.pops[i].domain)
                            ops.export(q.layout, q.pops[i].domain, 0, n, box, p.pops[i].domain)
                        else:
                            key = (pg.owner, pg.id)
                            have = ops.count(remote[key]) if key in remote else 0
                            if key not in remote:
                                remote[key] = ops.staging_particles(q.layout, n)
                            elif ops.capacity(remote[key]) < have + n:
                                remote[key] = ops.grow_particles(q.layout, remote[key], have + n)
                            ops.export(q.layout, q.pops[i].domain, 0, n, box, remote[key])
                if self.comm.size > 1:
                    def ensure(pid, needed, i=i):
                        while ops.capacity(mine[pid].pops[i].domain) < needed:
                            grow_store(ops, mine[pid], i, mine[pid].pops[i].domain)
                        return mine[pid].pops[i].domain
                    msg._exchange_particles({pid: p.layout for pid, p in mine.items()}, remote,
                                            {pid: p.pops[i].domain for pid, p in mine.items()}, {pid: 0 for pid in mine}, ensure)
            # domainParticlesRefiners_ (interior; on a regrid: the part of the interior the old level did not cover) and
            # lvlGhostPartOldRefiners_ (coarseBoundaryOld)
            msg.split_from_coarser(i, self.nref, lambda pg: minus_all([pg.box], old_boxes), "domain")
            msg.split_from_coarser(i, self.nref, lambda pg: msg.lg_particle_boxes[pg.id], "level_ghost_old")
        for p in s.patches:
            for i, pop in enumerate(p.pops):
                self._copy_store(p, i, "level_ghost_old", "level_ghost")  # copyLevelGhostOldToPushable_
                while ops.capacity(pop.spare) < ops.count(pop.domain):     # the split may have re-allocated `domain`
                    pop.spare = ops.particles(ops.capacity(pop.domain))
                counts = ops.bin(p.layout, pop.domain, pop.spare, p.domain_box, p.non_level_ghost, pop.cell_start)
                pop.domain, pop.spare = pop.spare, pop.domain
                pop.n_sorted = counts[0]
                ops.set_count(pop.domain, counts[0])
                for m in pop.moments():
                    ops.zero(m)
                ops.deposit(p.layout, pop.domain, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0, pop.n_sorted, (),
                            p.domain_box, pop.cell_start)  # depositParticles(DomainDeposit)
        for i in range(npop):
            msg.sum_borders(f"pop{i}", {p.geom.id: p.pops[i].moments() for p in s.patches},
                            {p.geom.id: p.pops[i].scratch for p in s.patches})
        for p in s.patches:
            for pop in p.pops:
                n = ops.count(pop.level_ghost_old)
                if n:  # depositParticles(LevelGhostDeposit): levelGhostParticlesOld, coef 1
                    ops.deposit(p.layout, pop.level_ghost_old, pop.rho_n, pop.rho_q, pop.flux, 1.0, 0, n)
            s.updater.update_ions(p)
        msg.max_borders("ions", {p.geom.id: [p.rho_m, p.Ne, p.Vi[0], p.Vi[1], p.Vi[2]] for p in s.patches})
        s.prepare_step()

    def _copy_store(self, patch, ipop, src_attr, dst_attr):
        ops, pop = self.ops, patch.pops[ipop]
        src = getattr(pop, src_attr)
        n = ops.count(src)
        while ops.capacity(getattr(pop, dst_attr)) < n:
            grow_store(ops, patch, ipop, getattr(pop, dst_attr))
        dst = getattr(pop, dst_attr)
        ops.set_count(dst, 0)
        if n:
            ops.particles_copy(src, 0, n, dst, 0)
        ops.set_count(dst, n)

    # ---------------------------------------------------------------------------------------- time stepping
    def advance(self, dt):
        """one step of the root level and the sub-cycles of every finer level (TimeRefinementIntegrator::advanceHierarchy
        driving MultiPhysicsIntegrator::advanceLevel / standardLevelSynchronization)"""
        self._root_end = self.time + dt
        self._advance_level(0, self.time, self._root_end, True, True)
        self.time += dt
        return self.time

    def _advance_level(self, il, t0, t1, first, last):
        lvl = self.levels[il]
        s = lvl.solver
        if first:
            if il > 0:
                self._first_step(lvl)
            s.reset_flux_sum()
        lvl.old_time = t0  # SolverPPC::prepareStep: oldTime_[level] (Bold <- B happens inside advance_level)
        if il > 0:
            # timeInterpCoef_ (:901-905) at afterPushTime = newTime, for both sweeps of the step
            s.level_ghost_alpha = (t1 - lvl.before_coarse_time) / (lvl.after_coarse_time - lvl.before_coarse_time)
            if s.level_ghost_alpha < 0 or s.level_ghost_alpha > 1:
                raise RuntimeError(f"ion moment ghost time interp coef invalid : alpha: {s.level_ghost_alpha}")
        s.advance_level(t1 - t0)
        if last and il > 0:
            self._last_step(lvl)
        finest = il == len(self.levels) - 1
        if finest:
            self._fine_dump(il, t1)
        if il > 0 and finest:
            s.accumulate_flux_sum(1. / (RATIO * RATIO))
        if not finest:
            sub = (t1 - t0) / SUBSTEPS
            t = t0
            self.levels[il + 1].coarser_times = (t0, t1)
            for k in range(SUBSTEPS):
                tn = t1 if k == SUBSTEPS - 1 else t + sub
                self._advance_level(il + 1, t, tn, k == 0, k == SUBSTEPS - 1)
                t = tn
            self._synchronize(self.levels[il + 1], lvl, t1)
            self._fine_dump(il, t1)

    fine_dump = None  # callable(level number, time): the simulator's "fine_dump" functor (simulator.hpp:252-267)

    def _fine_dump(self, il, time):
        """MultiPhysicsIntegrator::dump_ (multiphysics_integrator.hpp:455-471): finer levels are dumped at the end of their
        own sub-steps (the finest after its advance, the others after the synchronisation of the finer one), except at the
        end time of the root step, which the regular dump covers"""
        if self.fine_dump is not None and il > 0 and time != self._root_end:
            self.fine_dump(il, time)

    def _first_step(self, lvl):
        """firstStep: levelGhostParticlesNew = split of the coarser level's particles, which are already at the end of its
        step; the times bracket the interpolation of the level-ghost moments"""
        ops, s, msg = self.ops, lvl.solver, lvl.solver.messenger
        for p in s.patches:
            for pop in p.pops:
                ops.set_count(pop.level_ghost_new, 0)
        for i in range(s.npop):
            msg.split_from_coarser(i, self.nref, lambda pg: msg.lg_particle_boxes[pg.id], "level_ghost_new")
        lvl.before_coarse_time, lvl.after_coarse_time = lvl.coarser_times

    def _last_step(self, lvl):
        """lastStep: New becomes Old, New is emptied, the pushable level ghosts restart from Old"""
        ops = self.ops
        for p in lvl.solver.patches:
            for i, pop in enumerate(p.pops):
                pop.level_ghost_old, pop.level_ghost_new = pop.level_ghost_new, pop.level_ghost_old
                ops.set_count(pop.level_ghost_new, 0)
                self._copy_store(p, i, "level_ghost_old", "level_ghost")

    # what standardLevelSynchronization moves from a level to the next coarser one:
    # (attribute of the fine patch, component, attribute of the coarse patch, first quantity, coarsener)
    SYNC_ITEMS = ([("E", c, "E", abi.EX, abi.COARSEN_ELECTRIC) for c in range(3)]            # synchronize(): electric
                  + [("Ne", None, "Ne", abi.RHO, abi.COARSEN_MOMENTS)]                       # ion charge density
                  + [("Vi", c, "Vi", abi.VX, abi.COARSEN_MOMENTS) for c in range(3)]         # ion bulk velocity
                  + [("fluxSumE", c, "Eavg", abi.EX, abi.COARSEN_ELECTRIC) for c in range(3)])  # reflux()

    def _synchronize(self, fine, coarse, sync_time):
        """standardLevelSynchronization for one (fine, coarse) pair"""
        ops, fs, cs, me = self.ops, fine.solver, coarse.solver, self.comm.rank
        gf, gc = fine.geom.g, coarse.geom.g
        fine_patch = {p.geom.id: p for p in fs.patches}
        coarse_patch = {q.geom.id: q for q in cs.patches}
        arr = lambda patch, attr, c: getattr(patch, attr) if c is None else getattr(patch, attr)[c]
        # synchronize(): E (electric coarsener), ion charge density and bulk velocity (injection);
        # reflux(): fluxSumE of the fine level onto Eavg of the coarse level.
        # A fine patch coarsens onto the coarse patch directly when both live here, else into a staging array that is
        # shipped to the owner of the coarse patch (one packed message per pair of ranks, compiled once).
        staged, send_items, recv_items = [], {}, {}
        # Destinations are found in the index space of each quantity (FieldGeometry overlaps; FieldVariable says
        # dataLivesOnPatchBorder for anything primal, field_variable.hpp:44): a coarse patch that only TOUCHES the coarsened
        # fine box still owns the shared border nodes and receives them, so that overlapped domain nodes of neighbouring
        # coarse patches stay equal (the reference's test_overlaped_fields_are_equal); `t` runs over the periodic images.
        zero = np.zeros(fine.geom.patches[0].box.dim, dtype=np.int64) if fine.geom.patches else None
        shifts = list(coarse.geom.shifts) if coarse.geom.periodic else [zero]
        for pg in fine.geom.patches:
            cbox = coarsen_box(pg.box)
            for qg in coarse.geom.patches:
                if pg.owner != me and qg.owner != me:
                    continue
                for t in shifts:
                    t = np.asarray(t, dtype=np.int64)
                    image = cbox.shift(t)
                    if image.grow(1) * qg.box is None:
                        continue
                    flo = pg.box.lo - gf + RATIO * t  # the fine patch seen from the image
                    for k, (fattr, c, cattr, qty0, op) in enumerate(self.SYNC_ITEMS):
                        qty = qty0 + (c or 0)
                        fb = field_box(image, qty) * field_box(qg.box, qty)
                        if fb is None:
                            continue
                        if pg.owner == me and qg.owner == me:
                            ops.field_coarsen(op, qty, arr(fine_patch[pg.id], fattr, c), flo,
                                              arr(coarse_patch[qg.id], cattr, c), qg.box.lo - gc, fb.lo, fb.hi)
                        elif pg.owner == me:
                            if fine.sync_phase is None:
                                stage = ops.array(fb.shape())
                                send_items.setdefault(qg.owner, []).append((stage, [0] * fb.dim, fb.shape()))
                            else:
                                stage = fine.sync_phase["stages"][len(staged)]
                            staged.append(stage)
                            ops.field_coarsen(op, qty, arr(fine_patch[pg.id], fattr, c), flo, stage, fb.lo, fb.lo, fb.hi)
                        elif fine.sync_phase is None:
                            recv_items.setdefault(pg.owner, []).append(
                                (arr(coarse_patch[qg.id], cattr, c), fb.lo - (qg.box.lo - gc), fb.shape()))
        if self.comm.size > 1:
            if fine.sync_phase is None:
                fine.sync_phase = fs.messenger._finish([], send_items, recv_items, 0)
                fine.sync_phase["stages"] = staged
            fs.messenger._run(fine.sync_phase, ("sync", fine.number))
        # patchGhostRefluxedSchedules: the patch ghosts of Eavg agree again with the refluxed interiors
        if isinstance(cs.messenger, RefinedLevelMessenger):
            cs.messenger.fill_patch_ghosts("Eavg", abi.EX, cs._by_id("Eavg"))
        else:
            cs.messenger.fill_ghosts("Eavg", abi.EX, cs._by_id("Eavg"))
        cs.reflux(sync_time - coarse.old_time)
        if coarse.number != 0:
            cs.accumulate_flux_sum(1. / (RATIO * RATIO))
        # postSynchronize: ghosts of what was coarsened
        cs.messenger.fill_ghosts("E", abi.EX, cs._by_id("E"))
        cs.messenger.fill_ghost_list("Ni", [abi.RHO], {p.geom.id: [p.Ne] for p in cs.patches})
        cs.messenger.fill_ghost_list("Vi", [abi.VX, abi.VY, abi.VZ], {p.geom.id: [p.Vi[0], p.Vi[1], p.Vi[2]] for p in cs.patches})


def build_hierarchy(ops, domain_cells, patch_grid, interp, dx, pops, B_fn, particles_fn, refinement_boxes=(),
                    solver_kw=None, nref=None, comm=None):
    """root level as phare_b200.setup.build, then one refined level per entry of refinement_boxes (a list, per level,
    of (lower, upper) cell boxes in the index space of the level BELOW, as in pyphare's `refinement_boxes`)"""
    from .setup import build
    root = build(ops, comm or LocalComm(), domain_cells, patch_grid, interp, dx, pops, B_fn, particles_fn, solver_kw)
    h = Hierarchy(ops, root, nref)
    for boxes in refinement_boxes:
        h.add_level([refine_box(Box(lo, hi)) for lo, hi in boxes])
    return h
