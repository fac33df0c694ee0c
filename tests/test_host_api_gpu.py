"""libphare_b200_host.so (the C++ SolverPPC + LevelMessenger of include/phare_b200/solver_ppc.hpp behind a C ABI, driven by
phare_b200/host_cpp.py): the same device-loaded configs advanced by C++ and by the Python driver must agree — both enqueue
the same kernels through the same C ABI, so differences can only come from the hosts' own box algebra and sequencing.
(The C++ driver is compared with the independent step oracle in tests/test_cpp_solver.py.)"""
import numpy as np
import pytest

from phare_b200 import abi, configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,cells,grid,steps", [(1, (1024,), (4,), 4), (2, (240,), (3,), 3), (3, (64, 32), (2, 2), 3),
                                               (4, (32, 32), (2, 2), 2), (5, (16, 16, 16), (2, 2, 1), 3)])
def test_cpp_host_equals_python_host(k, cells, grid, steps):
    from phare_b200 import host_cpp
    from phare_b200.messenger import LocalComm
    from phare_b200.solver import GpuOps
    cfg = configs.get(k).with_cells(cells, grid)
    cfg.pops = [dict(p, ppc=min(p["ppc"], 24)) for p in cfg.pops]
    py = configs.build_device_loaded(GpuOps(cfg.dim, cfg.interp, "cuda:0"), LocalComm(), cfg)
    cpp = host_cpp.CppLevel(cfg, "cuda:0")
    cpp.initialize()
    for _ in range(steps):
        py.advance_level(cfg.dt)
    cpp.advance(cfg.dt, steps)
    counts = cpp.counts()
    for ip, p in enumerate(py.patches):
        assert counts[ip] == [py.ops.count(pop.domain) for pop in p.pops]
        for which, attr, q0 in ((host_cpp.B, "B", abi.BX), (host_cpp.E, "E", abi.EX), (host_cpp.VI, "Vi", abi.VX)):
            for c in range(3):
                a = cpp.get_field(ip, which, c, q0 + c)
                b = py.ops.get_field(getattr(p, attr)[c])
                ok = np.isfinite(b)
                assert np.array_equal(np.isfinite(a), ok)
                scale = np.max(np.abs(b[ok])) + 1e-300
                assert np.max(np.abs(a[ok] - b[ok])) <= 1e-11 * scale + 1e-14, (ip, attr, c)
        a, b = cpp.get_field(ip, host_cpp.NI, 0, abi.RHO), py.ops.get_field(p.Ne)
        assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b))
    cpp.close()
