"""TEST INFRASTRUCTURE — Python handle on the STEP ORACLE (oracle/ref/ref_step.cpp inside oracle/_ref/libphare_ref.so):
SolverPPC::advanceLevel for one periodic root level of P patches, run by the reference's own functors on the reference's
own data types, with the level loop and the same-level exchanges written in C++ from the SAMRAI-side sources
(solver_ppc.hpp:315-598, hybrid_hybrid_messenger_strategy.hpp:376-497, field_geometry.hpp:227-304,
field_variable_fill_pattern.hpp:30-313, particles_variable_fill_pattern.hpp:66-107, particles_data.hpp:702-724).
It shares no code with phare_b200/ — it is what the step of phare_b200/solver.py (and of include/phare_b200/solver_ppc.hpp)
is compared with.  NOT part of the product path."""
import ctypes as C

import numpy as np

from phare_b200 import abi
from . import HostParticles, REF_SO, have_ref

# `which` of vec() / scalar()
B, E, J, VI, BPRED, EPRED, BAVG, EAVG, VE = range(9)
POP_FLUX = 10  # + pop index
NI, RHO_M, NE, PE = range(4)
POP_RHO_N, POP_RHO_Q = 10, 11  # + 2 * pop index
VEC_QTY0 = {B: abi.BX, E: abi.EX, J: abi.JX, VI: abi.VX, BPRED: abi.BX, EPRED: abi.EX, BAVG: abi.BX, EAVG: abi.EX,
            VE: abi.VX}


class RefStep:
    def __init__(self, dim, interp, boxes, dx, origin, domain_cells, masses, Te=0.12, eta=0.0, nu=1e-4, hyper_mode=0):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libphare_ref.so not built (needs /root/reference): make -C oracle ref")
        lib = self.lib = C.CDLL(REF_SO)
        lib.phr_step_create.restype = C.c_void_p
        lib.phr_step_last_error.restype = C.c_char_p
        lib.phr_step_count.restype = C.c_size_t
        lib.phr_step_field_size.restype = C.c_size_t
        self.dim, self.interp, self.npatch, self.npop = dim, interp, len(boxes), len(masses)
        arr = (abi.Box * len(boxes))()
        for i, (lo, hi) in enumerate(boxes):
            for d in range(dim):
                arr[i].lower[d], arr[i].upper[d] = int(lo[d]), int(hi[d])
        self.boxes = [(tuple(int(x) for x in lo), tuple(int(x) for x in hi)) for lo, hi in boxes]
        dbl = lambda xs: (C.c_double * len(xs))(*[float(x) for x in xs])
        self.h = lib.phr_step_create(dim, interp, len(boxes), arr, dbl(dx), dbl(origin),
                                     (C.c_int * dim)(*[int(c) for c in domain_cells]), len(masses), dbl(masses),
                                     C.c_double(Te), C.c_double(eta), C.c_double(nu), int(hyper_mode))
        if not self.h:
            raise RuntimeError("phr_step_create: " + lib.phr_step_last_error().decode())
        self.h = C.c_void_p(self.h)

    def close(self):
        if self.h:
            self.lib.phr_step_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: rc={rc} {self.lib.phr_step_last_error().decode()}")

    def shape(self, patch, qty):
        s = (C.c_uint32 * 3)()
        self.lib.phr_step_field_size(self.h, patch, int(qty), s)
        return tuple(int(s[d]) for d in range(self.dim))

    def set_particles(self, patch, pop, icell, delta, weight, charge, v):
        hp = HostParticles.from_soa(icell, delta, weight, charge, v)
        self._check(self.lib.phr_step_set_particles(self.h, patch, pop, C.byref(hp.c)), "set_particles")

    def set_vec(self, patch, which, arrays):
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
        for c in range(3):
            assert keep[c].shape == self.shape(patch, VEC_QTY0[which] + c), (keep[c].shape, self.shape(patch, VEC_QTY0[which] + c))
        vf = abi.VecField()
        for c in range(3):
            vf.comp[c] = keep[c].ctypes.data
        self._check(self.lib.phr_step_set_vec(self.h, patch, which, C.byref(vf)), "set_vec")

    def vec(self, patch, which):
        q0 = abi.VX if which >= POP_FLUX else VEC_QTY0[which]
        out = [np.empty(self.shape(patch, q0 + c)) for c in range(3)]
        vf = abi.VecField()
        for c in range(3):
            vf.comp[c] = out[c].ctypes.data
        self._check(self.lib.phr_step_get_vec(self.h, patch, which, C.byref(vf)), "get_vec")
        return out

    def scalar(self, patch, which):
        out = np.empty(self.shape(patch, abi.P if which == PE else abi.RHO))
        self._check(self.lib.phr_step_get_scalar(self.h, patch, which, out.ctypes.data_as(C.c_void_p)), "get_scalar")
        return out

    def count(self, patch, pop, kind=0):
        return int(self.lib.phr_step_count(self.h, patch, pop, kind))

    def particles(self, patch, pop, kind=0):
        n = self.count(patch, pop, kind)
        hp = HostParticles(self.dim, n)
        self._check(self.lib.phr_step_get_particles(self.h, patch, pop, kind, C.byref(hp.c)), "get_particles")
        return hp.soa()

    def initialize(self):
        self._check(self.lib.phr_step_initialize(self.h), "initialize")

    def advance(self, dt):
        self._check(self.lib.phr_step_advance(self.h, C.c_double(dt)), "advance")
