"""pybindlibs.cpp_etc (src/python3/cpp_etc.cpp:74-143): life cycle, hierarchy factory, MPI queries, the array
wrapper handed back by user init functions, build configuration."""
import os
import sys

import numpy as np

from phare_b200 import simulator as _sim


class SamraiLifeCycle:
    """SAMRAI/MPI start-up and shutdown in the reference (src/amr/samrai.hpp); here: nothing to start — the
    process group, when there is one, is torch.distributed's and belongs to the launcher"""

    def __init__(self):
        pass

    @staticmethod
    def reset():
        pass


def make_hierarchy():
    return _sim.make_hierarchy()


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


def mpi_size():
    d = _dist()
    return d.get_world_size() if d else 1


def mpi_rank():
    d = _dist()
    return d.get_rank() if d else 0


def mpi_barrier():
    d = _dist()
    if d:
        d.barrier()


def mpi_initialized():
    return _dist() is not None


def makePyArrayWrapper(array):
    """PyArrayWrapper<double> (src/python3/pybind_def.hpp): keeps the numpy array alive behind a Span"""
    return np.ascontiguousarray(array, dtype=np.float64)


def phare_build_config():
    v = sys.version_info
    return {"PYTHON_VERSION": f"Python {v.major}.{v.minor}.{v.micro}", "GIT_HASH": "phare_b200",
            "BACKEND": "libphare_b200.so (sm_100a)"}


def phare_deps():
    """versions of the native dependencies, written into every diagnostic's attributes (diagnostics.py:163)"""
    from phare_b200 import abi
    try:
        backend = abi.load().phb_version().decode()
    except Exception:
        backend = "libphare_b200.so (not built)"
    # the reference reports samrai / highfive / pybind; they are not used here: the keys stay (pyphare writes
    # "<dep>_version" attributes from them), the version says so
    return {"phare_b200": backend, "samrai": "0.0.0", "highfive": "0.0.0", "pybind": "0.0.0"}


AMRHierarchy = _sim.Hierarchy


def samrai_restart_file(path):
    """the per-rank restart file inside the directory of one restart time (one raw .npz per rank here)"""
    return os.path.join(path, "restart_rank{:06d}.npz".format(mpi_rank()))


def patch_data_ids(path):
    """the reference stores the SAMRAI patch data ids SAMRAI has to reload; the .npz names its arrays itself"""
    return []


def serialized_simulation_string(path):
    import numpy as np
    with np.load(os.path.join(path, "restart_rank000000.npz")) as z:
        return str(z["serialized_simulation"])


def restart_path_for_time(path, time):
    return os.path.join(path, "{:0>11.5f}".format(time))
