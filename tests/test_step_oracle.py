"""The step of phare_b200/solver.py (sequencing + same-level messenger) against the STEP ORACLE: the reference's own
Faraday / Ampere / Ohm / Electrons / Ions / IonUpdater functors on the reference's data types, driven by an independent
C++ level loop and box algebra (oracle/ref/ref_step.cpp, written from solver_ppc.hpp:315-598 and the SAMRAI-side fill
patterns).  Here both sides run on the CPU (the Python driver on the plain-C kernels of oracle/), so a difference is a
difference of SEQUENCING or EXCHANGE RULES; tests/test_configs_gpu.py repeats the comparison with the CUDA kernels.
The five BASELINE configs with their real profiles (phare_b200/configs.py), scaled down in cells only."""
import numpy as np
import pytest

import oracle
from phare_b200 import configs
from step_oracle_util import oracle_for, compare_fields, compare_particles, PER_NODE_RTOL

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libphare_ref.so not built (needs /root/reference)")

# (config, cells, patch grid, steps)
CASES = [
    (1, (2048,), (4,), 5),
    (2, (500,), (5,), 4),
    (3, (48, 32), (2, 2), 3),
    (4, (24, 16), (2, 1), 3),
    (5, (12, 12, 8), (2, 1, 2), 3),
]


@pytest.mark.parametrize("k,cells,grid,steps", CASES, ids=[f"C{c[0]}" for c in CASES])
def test_python_step_equals_step_oracle(k, cells, grid, steps):
    from oracle.cpu_ops import CpuOps
    from phare_b200.messenger import LocalComm
    cfg = configs.get(k).with_cells(cells, grid)
    solver, gparts = configs.build_host_loaded(CpuOps(cfg.dim, cfg.interp), LocalComm(), cfg)
    ref = oracle_for(solver, gparts, [p["mass"] for p in cfg.pops], cfg.Te, cfg.eta, cfg.nu)
    worst = compare_fields(solver, ref)
    assert max(worst.values()) == 0.0, worst  # initialisation: same arithmetic, same order -> same bits
    for s in range(steps):
        solver.advance_level(cfg.dt)
        ref.advance(cfg.dt)
        worst = compare_fields(solver, ref)
        assert max(worst.values()) <= PER_NODE_RTOL, (s, max(worst, key=worst.get), worst)
        dd, dv = compare_particles(solver, ref)
        assert dd <= 1e-12 and dv <= 1e-12, (s, dd, dv)


def test_one_patch_periodic_image_is_its_own_neighbour():
    """a single patch covering the periodic domain exchanges with its own images (1-D and 2-D)"""
    from oracle.cpu_ops import CpuOps
    from phare_b200.messenger import LocalComm
    for k, cells, grid in ((1, (256,), (1,)), (3, (24, 16), (1, 1))):
        cfg = configs.get(k).with_cells(cells, grid)
        solver, gparts = configs.build_host_loaded(CpuOps(cfg.dim, cfg.interp), LocalComm(), cfg)
        ref = oracle_for(solver, gparts, [p["mass"] for p in cfg.pops], cfg.Te, cfg.eta, cfg.nu)
        for s in range(3):
            solver.advance_level(cfg.dt)
            ref.advance(cfg.dt)
        worst = compare_fields(solver, ref)
        assert max(worst.values()) <= PER_NODE_RTOL, worst
        compare_particles(solver, ref)
