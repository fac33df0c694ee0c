"""The reference's remaining functional input scripts (tests/functional/*/), imported AS THEY ARE and driven by pyphare's own
Simulator over pybindlibs on the CPU parity back end: every hybrid script of the reference configures, initialises and
steps (harris_2d, td1d and td1dtagged are in test_frontend.py / test_tagging.py; the mhd_* scripts are the MHD solver, out of
scope).  What is checked: the hierarchy the script asks for, finite fields after a root step, particle number conserved."""
import importlib
import os
import sys
from unittest import mock

import numpy as np
import pytest

import phare_b200.simulator as S

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyphare")), reason="reference tree not mounted")

# name -> (module, config call, dim, levels expected, patches on level 0)
CASES = {
    "alfven_wave1d": ("tests.functional.alfven_wave.alfven_wave1d", lambda m: m.config(), [20]),
    "ion_ion_beam1d": ("tests.functional.ionIonBeam.ion_ion_beam1d", lambda m: m.config(), [3]),  # uniform: nothing tagged
    "conserv": ("tests.functional.conservation.conserv", lambda m: m.uniform(0.1, 0.2, 100, 25000), [5]),
    "dispersion_noise": ("tests.functional.dispersion.dispersion", lambda m: m.fromNoise(), [25]),
    "dispersion_modes": ("tests.functional.dispersion.dispersion", lambda m: m.prescribedModes(), [25]),
    "shock": ("tests.functional.shock.shock", lambda m: m.config(2), [125]),
    "translat1d_uni": ("tests.functional.translation.translat1d", lambda m: m.config_uni(vx=-1, diagdir="uni"), None),
    "translat1d_td": ("tests.functional.translation.translat1d", lambda m: m.config_td(vx=-1, diagdir="td"), None),
    "harris_3d": ("tests.functional.harris.harris_3d", lambda m: m.config(), [1]),
}


@pytest.fixture()
def cpu_backend(cpu_oracle):
    from oracle.cpu_ops import CpuOps
    old = S.ops_factory
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    yield
    S.ops_factory = old
    S.dict_instance().stop()


@pytest.mark.parametrize("name", list(CASES))
def test_reference_functional_script_runs_unchanged(cpu_backend, cpu_ref, tmp_path, monkeypatch, name):
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections", "matplotlib.gridspec",
                "matplotlib.animation", "matplotlib.colors", "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1",
                "h5py", "scipy.signal", "scipy.optimize", "scipy.ndimage"):
        if mod not in sys.modules:
            monkeypatch.setitem(sys.modules, mod, mock.MagicMock(name=mod))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        monkeypatch.delitem(sys.modules, k)
    import pyphare.pharein as ph
    from pyphare.simulator.simulator import Simulator
    module, configure, root_patches = CASES[name]
    ph.global_vars.sim = None
    simulator = None
    try:
        sim = configure(importlib.import_module(module))
        simulator = Simulator(sim, log_to_file=False)
        simulator.initialize()
        c = simulator.cpp_sim
        ops = c.solver.ops
        levels = c.level_solvers()
        boxes = getattr(sim, "refinement_boxes", None) or {}
        assert len(levels) == 1 + len(boxes)
        if root_patches is not None:
            assert [len(s.patches) for s in levels] == root_patches
        n0 = sum(ops.count(pop.domain) for p in levels[0].patches for pop in p.pops)
        assert n0 > 0
        simulator.advance()
        assert simulator.currentTime() == pytest.approx(sim.time_step)
        assert sum(ops.count(pop.domain) for p in levels[0].patches for pop in p.pops) == n0  # periodic root level
        g = 2
        for s in levels:
            for p in s.patches:
                for f in (*p.B, *p.E):
                    assert np.isfinite(ops.get_field(f)).all()
                ne = ops.get_field(p.Ne)
                inner = ne[(slice(g, -g),) * ne.ndim]
                assert np.isfinite(inner).all() and inner.min() > 0
    finally:
        if simulator is not None:
            simulator.reset()
        ph.global_vars.sim = None
        for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
            del sys.modules[k]
