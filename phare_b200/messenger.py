"""Same-level periodic messenger: what SAMRAI RefineSchedules + PHARE's fill patterns do between
every sub-step of SolverPPC, restated as explicit box transfer plans executed by the batched box
kernel (K8) and, across GPUs, by torch.distributed point-to-point (NCCL over NVLink).

Semantics restated (file:line relative to the PHARE tree):
  field ghost fill   HybridHybridMessengerStrategy::fill{Magnetic,Electric,Current}Ghosts
                     (src/amr/messengers/hybrid_hybrid_messenger_strategy.hpp:376-400); overlap =
                     dst ghost field box ^ shifted src interior field box, minus dst interior field box
                     (src/amr/data/field/field_geometry.hpp:139-304, field_variable_fill_pattern.hpp:30-211)
  moment border sum  fillFluxBorders / fillDensityBorders (:423-485): dst += src over the full
                     ghost-box intersection, through a scratch copy (field_variable_fill_pattern.hpp:219-313)
  border max         fillIonBorders (:488-497): dst = max(dst, src) over the same overlaps
  particle migration fillIonGhostParticles (:410-419) with ParticleDomainFromGhostFillPattern
                     (src/amr/data/particles/particles_variable_fill_pattern.hpp:66-107) and the periodic
                     iCell shift of ParticlesData::pack (src/amr/data/particles/particles_data.hpp:702-784)
The level is periodic in every direction (src/amr/wrappers/hierarchy.hpp:347-349).
"""
import numpy as np

from . import abi
from .boxes import Box, periodic_shifts

PRIMAL, DUAL = 0, 1


def centering(qty, d):
    if qty <= abi.BZ:
        return PRIMAL if (qty - abi.BX) == d else DUAL
    if qty <= abi.JZ:
        base = abi.EX if qty <= abi.EZ else abi.JX
        return DUAL if (qty - base) == d else PRIMAL
    return PRIMAL


class PatchGeom:
    """cell box + owner of one patch of the level (what every rank knows about every patch)"""

    def __init__(self, pid, box, owner):
        self.id, self.box, self.owner = pid, box, owner

    def interior_field_box(self, qty):
        hi = self.box.hi.copy()
        for d in range(self.box.dim):
            if centering(qty, d) == PRIMAL:
                hi[d] += 1  # field_geometry.hpp:139-181
        return Box(self.box.lo, hi)

    def ghost_field_box(self, qty, g):
        return self.interior_field_box(qty).grow(g)

    def local(self, node_lo, g):
        """array index of an AMR node/cell index (GridLayout::AMRToLocal, gridlayout.hpp:746-763)"""
        return node_lo - (self.box.lo - g)


class LevelGeom:
    """periodic=True: the root level (periodic in every direction); periodic=False: a refined level, whose patches
    only see each other directly (its outer ghosts are level ghosts, filled from the next coarser level: amr.py)"""

    def __init__(self, domain_shape, patches, interp, periodic=True):
        self.domain_shape = tuple(int(s) for s in domain_shape)
        self.dim = len(self.domain_shape)
        self.patches = patches
        self.g = 2 if interp == 1 else 4
        self.pg = 1 if interp == 1 else 2
        self.periodic = periodic
        self.shifts = periodic_shifts(self.domain_shape) if periodic else [np.zeros(self.dim, dtype=np.int64)]
        self._nb = {}

    def neighbours(self, p):
        """(src patch, shift) pairs whose shifted cell box touches p's ghost region; excludes (p, 0)"""
        if p.id not in self._nb:
            out = []
            reach = p.box.grow(self.g)
            for q in self.patches:
                for t in self.shifts:
                    if q.id == p.id and not t.any():
                        continue
                    if reach * q.box.shift(t) is not None:
                        out.append((q, t))
            self._nb[p.id] = out
        return self._nb[p.id]

    # ---- transfer plans: lists of (dst patch, src patch, dst_lo, src_lo, extent) in local indices
    def ghost_fill_plan(self, qty):
        plan = []
        for p in self.patches:
            gbox, ibox = p.ghost_field_box(qty, self.g), p.interior_field_box(qty)
            covered = []
            for q, t in self.neighbours(p):
                ov = gbox * q.interior_field_box(qty).shift(t)
                if ov is None:
                    continue
                pieces = ov.minus(ibox)
                for c in covered:  # a ghost node is written once (first provider wins)
                    pieces = [r for b in pieces for r in b.minus(c)]
                for b in pieces:
                    covered.append(b)
                    plan.append((p, q, p.local(b.lo, self.g), q.local(b.lo - t, self.g), b.shape()))
        return plan

    def border_plan(self):
        """full ghost-box intersections of primal quantities (moments)"""
        plan = []
        for p in self.patches:
            gbox = p.ghost_field_box(abi.RHO, self.g)
            for q, t in self.neighbours(p):
                ov = gbox * q.ghost_field_box(abi.RHO, self.g).shift(t)
                if ov is not None:
                    plan.append((p, q, p.local(ov.lo, self.g), q.local(ov.lo - t, self.g), ov.shape()))
        return plan

    def migration_plan(self):
        """(src patch S, dst patch D, box in S's frame, shift): patch-ghost particles of S lying in the
        image of D's cell box move to D with iCell += shift"""
        plan = []
        for s in self.patches:
            ghost = s.box.grow(self.pg)
            for d, t in self.neighbours(s):
                # D + t is the image of D seen from S; particles are shifted by -t into D's frame
                img = ghost * d.box.shift(t)
                if img is not None:
                    plan.append((s, d, abi.make_box(img.lo, img.hi), [int(-x) for x in t]))
        return plan


class LocalComm:
    """single process: every patch is local"""
    rank, size = 0, 1

    def exchange(self, sends, recvs):
        assert not sends and not recvs

    def prepare(self, sends, recvs):
        assert not sends and not recvs
        return None

    def run(self, prepared):
        pass

    def allreduce_max(self, v):
        return v

    def allreduce_array_max(self, a):
        return a

    def alltoall_counts(self, counts):
        return counts


class TorchComm:
    """one process per GPU, torch.distributed (NCCL on GPU tensors, gloo on CPU tensors)"""

    def __init__(self, device):
        import torch.distributed as dist
        self.dist, self.device = dist, device
        self.rank, self.size = dist.get_rank(), dist.get_world_size()

    def exchange(self, sends, recvs):
        """sends/recvs: {peer: 1-D tensor}; one grouped point-to-point phase"""
        ops = []
        for peer in sorted(recvs):
            ops.append(self.dist.P2POp(self.dist.irecv, recvs[peer], peer))
        for peer in sorted(sends):
            ops.append(self.dist.P2POp(self.dist.isend, sends[peer], peer))
        if ops:
            for r in self.dist.batch_isend_irecv(ops):
                r.wait()

    def prepare(self, sends, recvs):
        ops = [self.dist.P2POp(self.dist.irecv, recvs[peer], peer) for peer in sorted(recvs)]
        ops += [self.dist.P2POp(self.dist.isend, sends[peer], peer) for peer in sorted(sends)]
        return ops

    def run(self, prepared):
        if prepared:
            for r in self.dist.batch_isend_irecv(prepared):
                r.wait()

    def allreduce_max(self, v):
        import torch
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())

    def allreduce_array_max(self, a):
        """element-wise max over the ranks of a small host integer array (tag masks)"""
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a).astype(np.int32)).to(self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy().astype(a.dtype).reshape(a.shape)

    def alltoall_counts(self, counts):
        """counts[peer] = ints I send to peer (list of equal length per peer); returns what they send me"""
        import torch
        k = len(counts[0])
        src = torch.tensor(counts, dtype=torch.int64, device=self.device).reshape(self.size, k)
        dst = torch.empty_like(src)
        self.dist.all_to_all_single(dst, src)
        return dst.cpu().tolist()


class RawArray:
    """a box-op operand that is only an address (own or peer device memory) and a shape"""

    def __init__(self, ptr, shape):
        self.ptr, self.shape = int(ptr), tuple(int(x) for x in shape)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1


class PeerArena:
    """NVLink peer-memory exchange area (csrc/peer.cu): ONE device allocation per rank holding the flag words and
    the receive areas of every fixed-size exchange phase, mapped by every other rank of the node through CUDA IPC.
    A neighbour's pack kernel stores straight into it; ordering is a flag per (source rank -> this rank) that
    counts the phases delivered.  Built on GPU ranks with more than one process; otherwise the messenger stages
    through torch.distributed send/recv."""
    FLAGS = 4096  # bytes reserved for flag words (8 per source rank)

    def __init__(self, ops, comm, nbytes=None):
        import ctypes as C
        import os
        dist = comm.dist
        ctx = ops.ctx
        self.ctx, self.me, self.size = ctx, comm.rank, comm.size
        self.nbytes = int(nbytes) if nbytes else int(os.environ.get("PHB_PEER_ARENA_MB", 1024)) << 20
        # every step of the set-up is agreed on by all ranks: if any rank cannot export or map (no peer access, IPC
        # not permitted in this container, ...) everybody falls back to the staged send/recv path
        p, handle, ok = C.c_void_p(), (C.c_ubyte * 64)(), True
        try:
            ctx._check(ctx.lib.phb_malloc(ctx.h, self.nbytes, C.byref(p)))
            ctx._check(ctx.lib.phb_memset(ctx.h, p, 0, self.FLAGS))
            ctx.sync()
            ctx._check(ctx.lib.phb_ipc_export(ctx.h, p, handle))
        except Exception as e:  # noqa: BLE001 - any failure means "no peer path on this rank"
            ok, self.why = False, str(e)
        handles = [None] * self.size
        dist.all_gather_object(handles, (ok, bytes(handle)))
        self.base = {}
        if all(h[0] for h in handles):
            for r, (_, h) in enumerate(handles):
                if r == self.me:
                    self.base[r] = p.value
                    continue
                q = C.c_void_p()
                try:
                    ctx._check(ctx.lib.phb_ipc_open(ctx.h, (C.c_ubyte * 64)(*h), C.byref(q)))
                    self.base[r] = q.value
                except Exception as e:  # noqa: BLE001
                    ok, self.why = False, str(e)
                    break
        else:
            ok = False
        verdict = [None] * self.size
        dist.all_gather_object(verdict, ok)
        self.ok = all(verdict)
        if not self.ok:
            if p.value:
                ctx.lib.phb_free(ctx.h, p)
            return
        self.cursor = self.FLAGS
        self.sent = [0] * self.size      # phases I delivered to each rank
        self.received = [0] * self.size  # phases each rank delivered to me
        # device-side phase counters (phb_peer_phase): words of MY flag block that only my own kernels touch;
        # flag words 8 * r for r < size <= 128 occupy [0, 1024)
        assert self.size <= 128
        self.sent_counter = lambda r: self.base[self.me] + 2048 + 8 * r
        self.recv_counter = lambda r: self.base[self.me] + 3072 + 8 * r
        # GPU-clock time a stream may wait for a neighbour's flag: it also covers the neighbour's HOST work between two
        # phases (particle dumps, restarts), so it is generous; after a timeout no exchange phase applies anything any
        # more (box_op_batch_kernel) and the next poll raises PHB_ERR_PEER_TIMEOUT
        self.timeout_s = float(os.environ.get("PHB_PEER_TIMEOUT_S", 300.0))
        self._C = C

    def alloc(self, nbytes):
        """offset of a fresh 256-byte aligned region of MY arena"""
        off = self.cursor
        self.cursor += (int(nbytes) + 255) & ~255
        if self.cursor > self.nbytes:
            raise RuntimeError("peer arena exhausted: raise PHB_PEER_ARENA_MB")
        return off

    def exchange_offsets(self, mine):
        """mine: {source rank: offset in my arena}; returns {destination rank: offset of my area in ITS arena}"""
        everyone = [None] * self.size
        self._dist().all_gather_object(everyone, mine)
        return {r: everyone[r][self.me] for r in range(self.size) if r != self.me and self.me in everyone[r]}

    def _dist(self):
        import torch.distributed as dist
        return dist

    def signal(self, dsts):
        C = self._C
        if not dsts:
            return
        ptrs = (C.c_void_p * len(dsts))()
        vals = (C.c_uint64 * len(dsts))()
        for i, r in enumerate(dsts):
            self.sent[r] += 1
            ptrs[i] = self.base[r] + 8 * self.me  # my word in rank r's flag block
            vals[i] = self.sent[r]
        self.ctx._check(self.ctx.lib.phb_peer_signal(self.ctx.h, len(dsts), ptrs, vals))

    def wait(self, srcs):
        C = self._C
        if not srcs:
            return
        ptrs = (C.c_void_p * len(srcs))()
        vals = (C.c_uint64 * len(srcs))()
        for i, r in enumerate(srcs):
            self.received[r] += 1
            ptrs[i] = self.base[self.me] + 8 * r
            vals[i] = self.received[r]
        self.ctx._check(self.ctx.lib.phb_peer_wait(self.ctx.h, len(srcs), ptrs, vals, self.timeout_s))


class HybridMessenger:
    """Executes the plans of a LevelGeom on the patches owned by this rank.
    `ops` is the compute back end (phare_b200.solver.GpuOps in production)."""

    def __init__(self, geom, ops, comm, peer_halo=True):
        self.geom, self.ops, self.comm = geom, ops, comm
        self.me = comm.rank
        self._plans = {}
        self._compiled = {}
        self._migration = geom.migration_plan()
        import os as _os
        self.peer_migration = _os.environ.get("PHB_PEER_MIGRATION", "1") != "0"
        # fixed-size field phases go through NVLink peer memory when every rank drives a GPU of the same node
        self.arena = None
        import os
        if peer_halo and comm.size > 1 and hasattr(ops, "ctx") and os.environ.get("PHB_PEER_HALO", "1") != "0":
            arena = PeerArena(ops, comm)
            if arena.ok:
                self.arena = arena
            elif comm.rank == 0:
                import warnings
                warnings.warn("NVLink peer-memory halo unavailable (" + getattr(arena, "why", "another rank failed")
                              + "): field phases are staged through torch.distributed send/recv")

    def _plan(self, kind, qty):
        key = (kind, qty)
        if key not in self._plans:
            self._plans[key] = self.geom.ghost_fill_plan(qty) if kind == "fill" else self.geom.border_plan()
        return self._plans[key]

    def _compile(self, key, kind, qtys, arrays, op):
        """arrays: {patch id: [array handle per qty]} for local patches. Returns the compiled phase:
        local box ops, and per peer (pack ops, send buffer, recv buffer, unpack ops)."""
        if key not in self._compiled:
            me = self.me
            local, send_items, recv_items = [], {}, {}
            for ci, qty in enumerate(qtys):
                for (p, q, dlo, slo, ext) in self._plan(kind, qty):
                    if p.owner == me and q.owner == me:
                        local.append((arrays[p.id][ci], dlo, arrays[q.id][ci], slo, ext, op))
                    elif p.owner == me:
                        recv_items.setdefault(q.owner, []).append((arrays[p.id][ci], dlo, ext))
                    elif q.owner == me:
                        send_items.setdefault(p.owner, []).append((arrays[q.id][ci], slo, ext))
            self._compiled[key] = self._finish(local, send_items, recv_items, op)
        return self._compiled[key]

    def _run(self, phase, key=None):
        # bench.py: one CUDA-event pair around every exchange phase (GPU-side duration, waits for the neighbours included)
        timed = getattr(self.ops, "_timed", None)
        if timed is not None and getattr(self.ops, "kernel_timing", False):
            return timed("exchange_phases", lambda: self._run_phase(phase, key))
        return self._run_phase(phase, key)

    def _run_phase(self, phase, key=None):
        ops = self.ops
        if "peer_dsts" in phase:
            # every receive area exists twice and a phase alternates between the two: a neighbour writes area (k & 1) of
            # run k; before it can write that area again (run k + 2) it has waited for my signal of run k + 1, which my
            # stream issues after the unpack of run k.  No assumption about which other phases run in between.
            half = phase["halves"][phase["runs"] & 1]
            phase["runs"] += 1
            # one C call: pack -> signal -> local -> wait -> unpack, with the phase counters kept on the device
            ctx = self.arena.ctx
            ctx._check(ctx.lib.phb_peer_phase(ctx.h, half["desc"]))
            return
        ops.run_box_ops(phase["pre"])      # packs for every peer (+ the local copies when they are independent)
        ops.run_box_ops(phase["local"])    # local operations that must follow the packs (in-place max)
        if phase["peers"]:
            if "p2p" not in phase:         # the grouped send/recv list is built once per phase
                phase["p2p"] = self.comm.prepare(
                    {p: ops.as_tensor(ph["sbuf"]) for p, ph in phase["peers"].items() if ops.size(ph["sbuf"])},
                    {p: ops.as_tensor(ph["rbuf"]) for p, ph in phase["peers"].items() if ops.size(ph["rbuf"])})
            self.comm.run(phase["p2p"])
            ops.run_box_ops(phase["post"]) # unpack (copy / += / max) of everything received

    # ---- public API, named after the reference messenger --------------------------------------
    def fill_ghosts(self, name, qty0, vecs):
        """vecs: {patch id: vector field}; fillMagneticGhosts (qty0=BX), fillElectricGhosts (EX),
        fillCurrentGhosts (JX)"""
        arrays = {pid: [v[c] for c in range(3)] for pid, v in vecs.items()}
        self._run(self._compile(("fill", name), "fill", [qty0, qty0 + 1, qty0 + 2], arrays, 0), ("fill", name))

    def fill_ghost_list(self, name, qtys, arrays):
        """same-level ghost fill of arbitrary quantities ({patch id: [array per qty]}): the PatchGhostField refiners of
        postSynchronize (charge density, bulk velocity; hybrid_hybrid_messenger_strategy.hpp:736-751)"""
        self._run(self._compile(("fill", name), "fill", list(qtys), arrays, 0), ("fill", name))

    def sum_borders(self, name, arrays, scratch):
        """arrays/scratch: {patch id: [primal arrays]}: a += neighbours' ORIGINAL values on the ghost-box
        overlaps (fillFluxBorders + fillDensityBorders)"""
        ops = self.ops
        for pid, lst in arrays.items():
            for a, s in zip(lst, scratch[pid]):
                ops.copy(s, a)
        n = len(next(iter(arrays.values()))) if arrays else 0  # a rank may own no patch of a refined level
        key = ("sum", name)
        if key not in self._compiled:
            # destination = arrays, source = scratch copies (local) / packed from scratch (remote)
            me, local, send_items, recv_items = self.me, [], {}, {}
            for ci in range(n):
                for (p, q, dlo, slo, ext) in self._plan("border", abi.RHO):
                    if p.owner == me and q.owner == me:
                        local.append((arrays[p.id][ci], dlo, scratch[q.id][ci], slo, ext, 1))
                    elif p.owner == me:
                        recv_items.setdefault(q.owner, []).append((arrays[p.id][ci], dlo, ext))
                    elif q.owner == me:
                        send_items.setdefault(p.owner, []).append((scratch[q.id][ci], slo, ext))
            self._compiled[key] = self._finish(local, send_items, recv_items, 1)
        self._run(self._compiled[key], key)

    def max_borders(self, name, arrays):
        """fillIonBorders: a = max(a, neighbour) on the ghost-box overlaps"""
        n = len(next(iter(arrays.values()))) if arrays else 0
        self._run(self._compile(("max", name), "border", [abi.RHO] * n, arrays, 2), ("max", name))

    def _finish(self, local, send_items, recv_items, op):
        """One launch packs for all peers, one launch unpacks from all peers (K8 batch tables)."""
        if self.arena is not None:
            return self._finish_peer(local, send_items, recv_items, op)
        ops = self.ops
        phase = dict(peers={})
        pack, unpack = [], []
        for peer in sorted(set(send_items) | set(recv_items)):
            s_items, r_items = send_items.get(peer, []), recv_items.get(peer, [])
            sbuf = ops.new_buffer(sum(int(np.prod(e)) for _, _, e in s_items))
            rbuf = ops.new_buffer(sum(int(np.prod(e)) for _, _, e in r_items))
            off = 0
            for (arr, lo, ext) in s_items:
                pack.append((ops.buffer_slice(sbuf, off, ext), [0] * len(ext), arr, lo, ext, 0))
                off += int(np.prod(ext))
            off = 0
            for (arr, lo, ext) in r_items:
                unpack.append((arr, lo, ops.buffer_slice(rbuf, off, ext), [0] * len(ext), ext, op))
                off += int(np.prod(ext))
            phase["peers"][peer] = dict(sbuf=sbuf, rbuf=rbuf)
        if op == 2:
            # in-place max: what is packed must be the value before any local update of this phase
            phase["pre"], phase["local"] = ops.compile_box_ops(pack), ops.compile_box_ops(local)
        else:
            # copy / += : packs read interiors (or the scratch copies), local ops write ghosts: independent
            phase["pre"], phase["local"] = ops.compile_box_ops(pack + local), None
        phase["post"] = ops.compile_box_ops(unpack)
        return phase

    def _finish_peer(self, local, send_items, recv_items, op):
        """peer-memory variant: my receive areas live in my arena, the pack ops of my neighbours write into them
        (and mine into theirs); called collectively, in the same order, by every rank"""
        ops, arena = self.ops, self.arena
        phase = dict(peers={}, peer_dsts=sorted(send_items), peer_srcs=sorted(recv_items), runs=0, halves=[])
        for _ in range(2):
            my_off = {src: arena.alloc(8 * sum(int(np.prod(e)) for _, _, e in items))
                      for src, items in sorted(recv_items.items())}
            their_off = arena.exchange_offsets(my_off)
            pack, unpack = [], []
            for dst, items in sorted(send_items.items()):
                at = arena.base[dst] + their_off[dst]
                for (arr, lo, ext) in items:
                    pack.append((RawArray(at, ext), [0] * len(ext), arr, lo, ext, 0))
                    at += 8 * int(np.prod(ext))
            for src, items in sorted(recv_items.items()):
                at = arena.base[arena.me] + my_off[src]
                for (arr, lo, ext) in items:
                    unpack.append((arr, lo, RawArray(at, ext), [0] * len(ext), ext, op))
                    at += 8 * int(np.prod(ext))
            half = {}
            if op == 2:
                half["pre"], half["local"] = ops.compile_box_ops(pack), ops.compile_box_ops(local)
            else:
                half["pre"], half["local"] = ops.compile_box_ops(pack + local), None
            half["post"] = ops.compile_box_ops(unpack)
            half["desc"] = self._phase_desc(half, phase["peer_dsts"], phase["peer_srcs"])
            phase["halves"].append(half)
        return phase

    def _phase_desc(self, half, dsts, srcs):
        """the constant argument block of phb_peer_phase for one half of a phase"""
        import ctypes as C
        arena = self.arena
        d = abi.PeerPhaseDesc()
        for name in ("pre", "local", "post"):
            comp = half[name]
            if comp is not None:
                table, n, total = comp
                setattr(d, name, table.data_ptr())
                setattr(d, "n_" + name, n)
                setattr(d, "total_" + name, total)
        d.n_signal, d.n_wait = len(dsts), len(srcs)
        for i, r in enumerate(dsts):
            d.signal_flag[i] = arena.base[r] + 8 * arena.me  # my word in rank r's flag block
            d.signal_counter[i] = arena.sent_counter(r)
        for i, r in enumerate(srcs):
            d.wait_flag[i] = arena.base[arena.me] + 8 * r
            d.wait_counter[i] = arena.recv_counter(r)
        d.timeout_s = arena.timeout_s
        return d

    def migrate_particles(self, layouts, patch_ghost, domain, ensure=None, vote=None):
        """fillIonGhostParticles for one population.
        patch_ghost: {pid: (store, first, last)} the new patch-ghost particles of my patches;
        domain: {pid: store} destination stores of my patches.  Returns the number received per patch.
        vote: an integer every rank contributes; its maximum over the ranks rides on the count exchange and is left in
        self.last_vote (the error vote of mpi::any_errors() without a collective of its own)"""
        ops, me = self.ops, self.me
        self.last_vote = vote
        received = {pid: 0 for pid in domain}
        remote = {}  # (dst owner, dst pid) -> staging store
        by_src = {}
        for (s, d, box, shift) in self._migration:
            if s.owner == me:
                by_src.setdefault(s.id, []).append((d, box, shift))
        for sid, items in by_src.items():
            store, first, last = patch_ghost[sid]
            if last <= first:
                continue
            dsts = []
            for (d, box, shift) in items:
                if d.owner == me:
                    dsts.append(domain[d.id])
                else:
                    key = (d.owner, d.id)
                    need = (ops.count(remote[key]) if key in remote else 0) + (last - first)
                    if key not in remote:
                        remote[key] = ops.staging_particles(layouts[sid], need)
                    elif ops.capacity(remote[key]) < need:
                        remote[key] = ops.grow_particles(layouts[sid], remote[key], need)
                    dsts.append(remote[key])
            # every image box of this source patch in one classification pass (K2 export_multi)
            counts = ops.export_multi(layouts[sid], store, first, last, [b for _, b, _ in items],
                                      [sh for _, _, sh in items], dsts)
            for (d, _, _), c in zip(items, counts):
                if d.owner == me:
                    received[d.id] += c
        if self.comm.size > 1:
            self._exchange_particles(layouts, remote, domain, received, ensure, vote)
        return received

    # ---- particle migration through NVLink peer memory -----------------------------------------------------------
    def _peer_migration_state(self):
        """receive areas for migrating particles in my arena, one per source rank and parity: a header (particles per
        destination patch + the sender's error vote) and a flat message of fixed capacity (columns CAP apart, so neither
        side needs the other's total).  Built collectively at the first exchange."""
        if getattr(self, "_mig", None) is None:
            import os
            arena, me = self.arena, self.me
            npatch = len(self.geom.patches)
            dsts = sorted({d.owner for (s, d, _, _) in self._migration if s.owner == me and d.owner != me})
            srcs = sorted({s.owner for (s, d, _, _) in self._migration if d.owner == me and s.owner != me})
            cap = (int(os.environ.get("PHB_PEER_MIGRATION_CAP", 1 << 18)) + 63) & ~63
            dim = self.geom.patches[0].box.dim
            hdr_bytes = ((npatch + 1) * 8 + 255) & ~255
            payload = arena.ctx.lib.phb_particles_flat_bytes(dim, cap)
            my_off, their_off = [], []
            for _ in range(2):
                mine = {r: arena.alloc(hdr_bytes + payload) for r in srcs}
                my_off.append(mine)
                their_off.append(arena.exchange_offsets(mine))
            desc = abi.PeerPhaseDesc()  # no box ops: only the signal / wait pair with the device-side counters
            desc.n_signal, desc.n_wait = len(dsts), len(srcs)
            for i, r in enumerate(dsts):
                desc.signal_flag[i] = arena.base[r] + 8 * arena.me
                desc.signal_counter[i] = arena.sent_counter(r)
            for i, r in enumerate(srcs):
                desc.wait_flag[i] = arena.base[arena.me] + 8 * r
                desc.wait_counter[i] = arena.recv_counter(r)
            desc.timeout_s = arena.timeout_s
            # the vote reaches everybody only if every rank hears from every other rank
            complete = len(srcs) == self.comm.size - 1 and len(dsts) == self.comm.size - 1
            everyone = [None] * self.comm.size
            self.comm.dist.all_gather_object(everyone, complete)
            self._mig = dict(dsts=dsts, srcs=srcs, cap=cap, hdr_bytes=hdr_bytes, my_off=my_off, their_off=their_off, desc=desc,
                             runs=0, npatch=npatch, vote_complete=all(everyone))
        return self._mig

    def _exchange_particles_peer(self, layouts, remote, domain, received, ensure=None, vote=None):
        """the migrating particles of one population written straight into the neighbours' receive areas (remote stores
        over NVLink by the pack kernel), one signal | wait pair, one read-back of the headers: no collective"""
        import ctypes as C
        ops, arena, M = self.ops, self.arena, self._peer_migration_state()
        ctx, lib = arena.ctx, arena.ctx.lib
        parity = M["runs"] & 1
        M["runs"] += 1
        npatch, cap = M["npatch"], M["cap"]
        keep = []  # host headers stay alive until the copies have run
        for peer in M["dsts"]:
            hdr = np.zeros(npatch + 1, np.uint64)
            base = arena.base[peer] + M["their_off"][parity][peer]
            off = 0
            for pid in range(npatch):
                st = remote.get((peer, pid))
                n = ops.count(st) if st is not None else 0
                if not n:
                    continue
                if off + n > cap:
                    raise RuntimeError(f"more than {cap} particles migrate to rank {peer} in one step: raise "
                                       "PHB_PEER_MIGRATION_CAP (and PHB_PEER_ARENA_MB)")
                ctx._check(lib.phb_particles_pack(ctx.h, C.byref(st.c), 0, n, C.c_void_p(base + M["hdr_bytes"]), cap, off))
                hdr[pid] = n
                off += n
            hdr[npatch] = int(vote or 0)
            keep.append(hdr)
            ctx._check(lib.phb_h2d(ctx.h, C.c_void_p(base), hdr.ctypes.data, hdr.nbytes))
        ctx._check(lib.phb_peer_phase(ctx.h, M["desc"]))
        got = {}
        for src in M["srcs"]:
            h = np.zeros(npatch + 1, np.uint64)
            ctx._check(lib.phb_d2h(ctx.h, h.ctypes.data, C.c_void_p(arena.base[arena.me] + M["my_off"][parity][src]), h.nbytes))
            got[src] = h
        ctx.sync()
        del keep
        votes = [int(vote or 0)]
        for src in M["srcs"]:
            h = got[src]
            votes.append(int(h[npatch]))
            buf = arena.base[arena.me] + M["my_off"][parity][src] + M["hdr_bytes"]
            off = 0
            for pid in range(npatch):
                n = int(h[pid])
                if not n:
                    continue
                if ensure is not None:
                    domain[pid] = ensure(pid, ops.count(domain[pid]) + n)
                dst = domain[pid]
                if dst.n + n > dst.capacity:
                    raise RuntimeError("particle store capacity exceeded while receiving migrating particles")
                ctx._check(lib.phb_particles_unpack(ctx.h, C.c_void_p(buf), cap, off, n, C.byref(dst.c)))
                received[pid] += n
                off += n
        self.last_vote = max(votes) if (vote is not None and M["vote_complete"]) else None

    def _exchange_particles(self, layouts, remote, domain, received, ensure=None, vote=None):
        """remote: {(owner rank, patch id): staging store} -> appended to domain[patch id] on the owner.
        ensure(pid, needed) -> store: lets the caller re-allocate a destination store that is too small"""
        if self.arena is not None and getattr(self, "peer_migration", True):
            return self._exchange_particles_peer(layouts, remote, domain, received, ensure, vote)
        ops, comm, geom = self.ops, self.comm, self.geom
        # every rank announces, per destination patch, how many particles it ships
        npatch = len(geom.patches)
        counts = [[0] * npatch + [int(vote or 0)] for _ in range(comm.size)]
        for (owner, pid), st in remote.items():
            counts[owner][pid] = ops.count(st)
        _t = getattr(ops, "_timed", None) if getattr(ops, "kernel_timing", False) else None
        incoming = (_t("fp_alltoall_counts", lambda: comm.alltoall_counts(counts)) if _t
                    else comm.alltoall_counts(counts))  # incoming[src rank][pid], last column: that rank's vote
        if vote is not None:
            self.last_vote = max(int(vote), max(int(row[npatch]) for row in incoming))
        any_layout = next(iter(layouts.values()), None)  # only carries dim / interp, which the back end knows
        sends, recvs, recv_meta = {}, {}, {}
        for peer in range(comm.size):
            if peer == self.me:
                continue
            out = [(pid, remote[(peer, pid)]) for pid in range(npatch) if counts[peer][pid]]
            if out:
                sends[peer] = ops.pack_particles(any_layout, [st for _, st in out])
            inc = [(pid, incoming[peer][pid]) for pid in range(npatch) if incoming[peer][pid]]
            if inc:
                recvs[peer] = ops.new_particle_buffer(any_layout, sum(n for _, n in inc))
                recv_meta[peer] = inc
        _x = lambda: comm.exchange({p: ops.as_tensor(b) for p, b in sends.items()},
                                   {p: ops.as_tensor(b) for p, b in recvs.items()})
        _t("fp_send_recv", _x) if _t else _x()
        for peer, inc in recv_meta.items():
            off = 0
            for pid, n in inc:
                if ensure is not None:
                    domain[pid] = ensure(pid, ops.count(domain[pid]) + n)
                ops.unpack_particles(layouts[pid], recvs[peer], off, n, sum(m for _, m in inc), domain[pid])
                received[pid] += n
                off += n
