"""ctypes view of the C ABI declared in include/phare_b200.h.

This module loads only the product library, phare_b200/lib/libphare_b200.so (phb_* entry points, CUDA
kernels, DEVICE pointers).  The test-only CPU checkers under the top-level checker package reuse these
struct definitions with HOST pointers; nothing in phare_b200/ imports or links them.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("PHB_LIB") or os.path.join(_HERE, "lib", "libphare_b200.so")  # PHB_LIB: tuning builds

PHB_OK, PHB_ERR_INVALID, PHB_ERR_CUDA, PHB_ERR_MOVE_TWO_CELL = 0, 1, 2, 3
PHB_ERR_OUTSIDE_GHOST, PHB_ERR_CAPACITY, PHB_ERR_NO_DEVICE, PHB_ERR_PEER_TIMEOUT = 4, 5, 6, 7

# phb_qty
BX, BY, BZ, EX, EY, EZ, JX, JY, JZ, RHO, VX, VY, VZ, P = range(14)

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_u32_p = C.POINTER(C.c_uint32)


class Layout(C.Structure):
    _fields_ = [("dim", C.c_int), ("interp", C.c_int), ("level", C.c_int), ("amr_lower", C.c_int * 3),
                ("ncells", C.c_uint32 * 3), ("dx", C.c_double * 3), ("origin", C.c_double * 3)]


class VecField(C.Structure):
    _fields_ = [("comp", C.c_void_p * 3)]


class Box(C.Structure):
    _fields_ = [("lower", C.c_int * 3), ("upper", C.c_int * 3)]


class Particles(C.Structure):
    _fields_ = [("icell", C.c_void_p * 3), ("delta", C.c_void_p * 3), ("v", C.c_void_p * 3),
                ("weight", C.c_void_p), ("charge", C.c_void_p), ("n", C.c_size_t), ("capacity", C.c_size_t)]


class BoxDesc(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src", C.c_void_p), ("dst_shape", C.c_uint32 * 3), ("dst_lo", C.c_uint32 * 3),
                ("src_shape", C.c_uint32 * 3), ("src_lo", C.c_uint32 * 3), ("ext", C.c_uint32 * 3),
                ("op", C.c_int32), ("first", C.c_uint64)]


class PeerPhaseDesc(C.Structure):
    """phb_peer_phase_desc: one exchange phase (pack -> signal -> local -> wait -> unpack), built once"""
    _fields_ = [("pre", C.c_void_p), ("n_pre", C.c_int), ("total_pre", C.c_uint64),
                ("local", C.c_void_p), ("n_local", C.c_int), ("total_local", C.c_uint64),
                ("post", C.c_void_p), ("n_post", C.c_int), ("total_post", C.c_uint64),
                ("n_signal", C.c_int), ("signal_flag", C.c_void_p * 32), ("signal_counter", C.c_void_p * 32),
                ("n_wait", C.c_int), ("wait_flag", C.c_void_p * 32), ("wait_counter", C.c_void_p * 32),
                ("timeout_s", C.c_double)]


class FieldView(C.Structure):
    """phb_field_view: array + AMR field index of its first element"""
    _fields_ = [("data", C.c_void_p), ("shape", C.c_uint32 * 3), ("lo", C.c_int32 * 3)]


# phb_refine_op / phb_coarsen_op
REFINE_DEFAULT, REFINE_MAGNETIC, REFINE_MAGNETIC_INIT, REFINE_ELECTRIC = range(4)
COARSEN_ELECTRIC, COARSEN_MOMENTS = range(2)


def make_view(ptr, shape, lo):
    v = FieldView()
    v.data = ptr
    for d in range(3):
        v.shape[d] = int(shape[d]) if d < len(shape) else 1
        v.lo[d] = int(lo[d]) if d < len(lo) else 0
    return v


def make_layout(dim, interp, ncells, dx, amr_lower=None, origin=None, level=0):
    L = Layout()
    L.dim, L.interp, L.level = dim, interp, level
    for d in range(3):
        L.ncells[d] = int(ncells[d]) if d < dim else 0
        L.dx[d] = float(dx[d]) if d < dim else 0.0
        L.amr_lower[d] = int(amr_lower[d]) if (amr_lower is not None and d < dim) else 0
        L.origin[d] = float(origin[d]) if (origin is not None and d < dim) else 0.0
    return L


def make_box(lower, upper):
    b = Box()
    for d in range(3):
        b.lower[d] = int(lower[d]) if d < len(lower) else 0
        b.upper[d] = int(upper[d]) if d < len(upper) else 0
    return b


def box_array(boxes):
    arr = (Box * max(len(boxes), 1))()
    for i, b in enumerate(boxes):
        arr[i] = b
    return arr


_PROTOS = {
    # name: (restype, argtypes)
    "phb_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "phb_destroy": (None, [C.c_void_p]),
    "phb_last_error": (C.c_char_p, [C.c_void_p]),
    "phb_version": (C.c_char_p, []),
    "phb_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phb_get_stream": (C.c_void_p, [C.c_void_p]),
    "phb_sync": (C.c_int, [C.c_void_p]),
    "phb_set_exact": (C.c_int, [C.c_void_p, C.c_int]),
    "phb_poll_error": (C.c_int, [C.c_void_p]),
    "phb_launch_count": (C.c_uint64, [C.c_void_p]),
    "phb_malloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "phb_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phb_memset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "phb_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "phb_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "phb_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "phb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "phb_host_free": (C.c_int, [C.c_void_p]),
    "phb_field_shape": (C.c_size_t, [C.POINTER(Layout), C.c_int, c_u32_p]),
    "phb_field_ghosts": (C.c_int, [C.c_int]),
    "phb_particle_ghosts": (C.c_int, [C.c_int]),
    "phb_particles_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(Particles)]),
    "phb_particles_free": (C.c_int, [C.c_void_p, C.POINTER(Particles)]),
    "phb_aos_stride": (C.c_size_t, [C.c_int]),
    "phb_particles_from_aos": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Particles)]),
    "phb_particles_to_aos": (C.c_int, [C.c_void_p, C.POINTER(Particles), C.c_void_p]),
    "phb_particles_from_soa": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.POINTER(Particles)]),
    "phb_particles_to_soa": (C.c_int, [C.c_void_p, C.POINTER(Particles), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "phb_particles_copy": (C.c_int, [C.c_void_p, C.POINTER(Particles), C.c_size_t, C.c_size_t,
                                     C.POINTER(Particles), C.c_size_t]),
    "phb_particles_flat_bytes": (C.c_size_t, [C.c_int, C.c_size_t]),
    "phb_particles_pack": (C.c_int, [C.c_void_p, C.POINTER(Particles), C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.c_size_t]),
    "phb_particles_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(Particles)]),
    "phb_maxwellian_load_host": (C.c_int, [C.POINTER(Layout), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.c_double, C.c_uint32, C.c_int, C.c_size_t,
                                           C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_size_t, C.POINTER(C.c_size_t)]),
    "phb_maxwellian_load": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.c_void_p, C.POINTER(VecField),
                                      C.POINTER(VecField), C.c_void_p, C.c_size_t, C.c_double, C.c_uint32, C.c_uint64,
                                      c_u32_p, C.POINTER(Particles)]),
    "phb_push": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                           C.POINTER(Particles), C.POINTER(Particles), C.c_double, C.c_double, C.POINTER(Box)]),
    "phb_bin_nkeys": (C.c_size_t, [C.POINTER(Layout), C.POINTER(Box)]),
    "phb_bin": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Particles), C.POINTER(Particles),
                          C.POINTER(Box), C.POINTER(Box), C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]),
    "phb_bin_plan": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Particles), C.POINTER(Box), C.POINTER(Box), C.c_int,
                               C.c_void_p]),
    "phb_deposit_scatter": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Particles), C.c_size_t, C.c_void_p,
                                      C.c_void_p, C.POINTER(VecField), C.c_double, C.POINTER(Box), C.c_int,
                                      C.POINTER(Box), C.c_void_p, C.POINTER(Box), C.c_int, C.POINTER(Particles),
                                      C.c_void_p]),
    "phb_bin_counts": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Box), C.c_void_p, C.POINTER(C.c_size_t),
                                 C.POINTER(Particles)]),
    "phb_export": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Particles), C.c_size_t, C.c_size_t,
                             C.POINTER(Box), C.POINTER(Box), c_int_p, C.POINTER(Particles),
                             C.POINTER(C.c_size_t)]),
    "phb_export_multi": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Particles), C.c_size_t, C.c_size_t, C.c_int,
                                   C.POINTER(Box), c_int_p, C.POINTER(C.POINTER(Particles)), C.POINTER(C.c_size_t)]),
    "phb_deposit": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Particles), C.c_size_t, C.c_size_t,
                              C.c_void_p, C.c_void_p, C.POINTER(VecField), C.c_double, C.POINTER(Box), C.c_int,
                              C.POINTER(Box), C.c_void_p]),
    "phb_push_deposit": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                                   C.POINTER(Particles), C.c_size_t, C.c_size_t, C.c_double, C.c_double,
                                   C.POINTER(Box), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(VecField), C.c_double,
                                   C.POINTER(Box), C.c_int, C.POINTER(Box), C.c_void_p]),
    "phb_gather": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                             C.POINTER(Particles), C.c_size_t, C.c_size_t, C.c_void_p]),
    "phb_push_plan": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                                C.POINTER(Particles), C.c_size_t, C.c_double, C.c_double, C.POINTER(Box), C.c_void_p,
                                C.POINTER(Box), C.c_int, C.c_void_p]),
    "phb_push_cells": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                                 C.POINTER(Particles), C.POINTER(Particles), C.c_size_t, C.c_double, C.c_double,
                                 C.POINTER(Box), C.c_void_p]),
    "phb_push_deposit_plan": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                                        C.POINTER(Particles), C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                        C.POINTER(VecField), C.c_double, C.POINTER(Box), C.c_int, C.POINTER(Box),
                                        C.c_void_p, C.POINTER(Box), C.c_int, C.c_void_p]),
    "phb_scatter_planned": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Particles), C.c_size_t, C.POINTER(Box),
                                      C.c_void_p, C.POINTER(Box), C.c_int, C.POINTER(Particles), C.c_void_p]),
    "phb_set_predict_eps": (C.c_int, [C.c_void_p, C.c_double]),
    "phb_get_predict_eps": (C.c_double, [C.c_void_p]),
    "phb_predict_supported": (C.c_int, [C.POINTER(Layout)]),
    "phb_predict_plan_bytes": (C.c_size_t, [C.POINTER(Layout), C.POINTER(Box), C.c_size_t]),
    "phb_push_deposit_predict": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                                           C.POINTER(Particles), C.c_size_t, C.c_double, C.c_double, C.c_void_p,
                                           C.c_void_p, C.POINTER(VecField), C.c_double, C.POINTER(Box), C.c_int,
                                           C.POINTER(Box), C.c_void_p, C.POINTER(Box), C.c_int, C.c_void_p, C.c_size_t]),
    "phb_push_deposit_rebin": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                                         C.POINTER(Particles), C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                         C.POINTER(VecField), C.c_double, C.POINTER(Box), C.c_int, C.POINTER(Box),
                                         C.c_void_p, C.POINTER(Box), C.c_int, C.POINTER(Particles), C.c_void_p, C.c_void_p,
                                         C.c_size_t]),
    "phb_predict_counts": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(Box), C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_size_t), C.POINTER(Particles)]),
    "phb_gridlayout_probe": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "phb_split": (C.c_int, [C.c_void_p, C.POINTER(Particles), C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                            C.c_int, C.POINTER(Box), C.c_int, C.POINTER(Particles), C.POINTER(C.c_size_t)]),
    "phb_faraday": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField),
                              C.POINTER(VecField), C.c_double]),
    "phb_ampere": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(VecField)]),
    "phb_ohm": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.c_void_p, C.POINTER(VecField), C.c_void_p,
                          C.POINTER(VecField), C.POINTER(VecField), C.POINTER(VecField), C.c_double, C.c_double,
                          C.c_int]),
    "phb_electrons_update": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.c_void_p, C.POINTER(VecField),
                                       C.POINTER(VecField), C.c_double, C.POINTER(VecField), C.c_void_p]),
    "phb_ions_totals": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                  C.POINTER(VecField), c_double_p, C.c_void_p, C.c_void_p, C.POINTER(VecField)]),
    "phb_average": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "phb_average_many": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p)]),
    "phb_box_op": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_u32_p, c_u32_p, C.c_void_p, c_u32_p, c_u32_p,
                             c_u32_p, C.c_int]),
    "phb_box_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_u32_p, c_u32_p, c_u32_p, C.c_void_p]),
    "phb_box_op_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64]),
    "phb_field_refine": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(FieldView), C.POINTER(FieldView),
                                   C.POINTER(Box)]),
    "phb_magnetic_postprocess": (C.c_int, [C.c_void_p, C.POINTER(Layout), C.POINTER(VecField), C.POINTER(Box),
                                           C.POINTER(Box), C.c_int]),
    "phb_field_coarsen": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(FieldView), C.POINTER(FieldView),
                                    C.POINTER(Box)]),
    "phb_box_fill": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_u32_p, c_u32_p, c_u32_p, C.c_double]),
    "phb_axpy": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_double]),
    "phb_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "phb_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "phb_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phb_peer_phase": (C.c_int, [C.c_void_p, C.POINTER(PeerPhaseDesc)]),
    "phb_peer_signal": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "phb_peer_wait": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_double]),
    "phb_box_unpack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_u32_p, c_u32_p, c_u32_p, C.c_void_p,
                                 C.c_int]),
}

EXPORTED = sorted(_PROTOS)

_lib = None


def load():
    """Load libphare_b200.so (built in-tree by `make lib` / __graft_entry__.build()).
    There is no fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `make lib` (or __graft_entry__.build())")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
