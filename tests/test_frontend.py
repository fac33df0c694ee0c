"""Front end: the pybindlibs drop-in modules, the PHAREDict, the simulator object pyphare
drives, and the parity-mode particle loader.  CPU only: the compute back end is replaced by the checker
(oracle/cpu_ops.py) through phare_b200.simulator.ops_factory; the GPU side is tests/test_simulator_gpu.py.
When the reference tree is mounted, pyphare's own Simulator runs the reference's harris_2d.py input script
unchanged through these modules."""
import ctypes as C
import importlib
import os
import sys
from unittest import mock

import numpy as np
import pytest

from phare_b200 import abi
import phare_b200.simulator as S
from frontend_util import populate, two_pop_1d, const, gather

REF = "/root/reference"


@pytest.fixture()
def cpu_backend(cpu_oracle):
    from oracle.cpu_ops import CpuOps
    old = S.ops_factory
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    yield
    S.ops_factory = old
    S.dict_instance().stop()


def test_dictator_is_typed_and_path_addressed():
    import pybindlibs.dictator as pp
    pp.stop()
    pp.add_int("simulation/dimension", 2)
    pp.add_double("simulation/grid/meshsize/x", 1)          # ints are accepted where pybind casts to double
    pp.add_string("simulation/AMR/refinement/boxes/nbr_levels/", "trailing slash is ignored")
    pp.add_optional_size_t("a/b/seed", None)
    pp.add_array_as_vector("a/ts", np.arange(3.0))
    d = S.dict_instance()
    assert d["simulation/dimension"] == 2 and isinstance(d["simulation/grid/meshsize/x"], float)
    assert d.contains("simulation/AMR/refinement/boxes/nbr_levels") and d["a/b/seed"] is None
    assert not d.contains("simulation/nope") and d.get("simulation/nope", 5) == 5
    for bad in (lambda: pp.add_int("x", 1.5), lambda: pp.add_size_t("x", -1), lambda: pp.add_bool("x", 1),
                lambda: pp.add_string("x", 3), lambda: pp.addInitFunction1D("x", 3.0),
                lambda: pp.add_array_as_vector("x", np.zeros((2, 2)))):
        with pytest.raises((TypeError, RuntimeError)):
            bad()
    pp.stop()
    assert not d.contains("simulation")


def test_simulator_modules_exist_for_every_permutation():
    for dim, interp, nref in ((1, 1, 2), (2, 3, 9), (3, 1, 6)):
        m = importlib.import_module(f"pybindlibs.cpp_{dim}_{interp}_{nref}")
        assert (m.Simulator.dims, m.Simulator.interp_order, m.Simulator.refined_particle_nbr) == (dim, interp, nref)
        assert callable(m.make_simulator) and hasattr(m, "DataWrangler") and hasattr(m, "Splitter")
    with pytest.raises(ImportError):
        importlib.import_module("pybindlibs.cpp_4_1_2")
    etc = importlib.import_module("pybindlibs.cpp_etc")
    assert etc.mpi_size() == 1 and etc.mpi_rank() == 0 and not etc.mpi_initialized()
    assert etc.phare_build_config()["PYTHON_VERSION"].startswith("Python 3")


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_parity_loader_reproduces_the_reference_initializer(cpu_ref, dim):
    """phb_maxwellian_load_host against MaxwellianParticleInitializer::loadParticles compiled from the reference
    headers (oracle/_ref): same seed -> the same particles bit for bit, cells below the cut-off skipped"""
    lib = abi.load()
    nc = [9, 6, 5][:dim]
    L = abi.make_layout(dim, 1, nc, [0.2] * dim, amr_lower=[3, -2, 7][:dim])
    ncell, ppc = int(np.prod(nc)), 11
    rng = np.random.default_rng(dim)
    n = rng.random(ncell) + 0.5
    n[3] = 1e-7
    V = [rng.standard_normal(ncell) for _ in range(3)]
    T = [rng.random(ncell) + 0.1 for _ in range(3)]
    want = cpu_ref.maxwellian(L, n, V, T, 1.5, ppc, 1337)
    cap = ncell * ppc
    ic, de = np.zeros((cap, dim), np.int32), np.zeros((cap, dim))
    w, q, v = np.zeros(cap), np.zeros(cap), np.zeros((cap, 3))
    pv = (C.c_void_p * 3)(*[a.ctypes.data for a in V])
    pt = (C.c_void_p * 3)(*[a.ctypes.data for a in T])
    cnt = C.c_size_t()
    rc = lib.phb_maxwellian_load_host(C.byref(L), n.ctypes.data, pv, pt, None, 1.5, ppc, 1, 1337, 1e-5, ic.ctypes.data,
                                      de.ctypes.data, w.ctypes.data, q.ctypes.data, v.ctypes.data, cap, C.byref(cnt))
    k = cnt.value
    assert rc == 0 and k == want.n == (ncell - 1) * ppc
    for g, x in zip((ic[:k], de[:k], w[:k], q[:k], v[:k]), want.soa()):
        assert np.array_equal(g, x)
    # capacity is checked
    rc = lib.phb_maxwellian_load_host(C.byref(L), n.ctypes.data, pv, pt, None, 1.5, ppc, 1, 1337, 1e-5, ic.ctypes.data,
                                      de.ctypes.data, w.ctypes.data, q.ctypes.data, v.ctypes.data, 5, C.byref(cnt))
    assert rc == abi.PHB_ERR_CAPACITY


def test_simulator_from_dict_matches_hand_driven_solver(cpu_backend, cpu_ref):
    """the Simulator built from the dict == SolverPPC driven by hand from the same profiles, with the particles of
    the REFERENCE initializer: bit-identical fields after 3 steps, whatever the patch cut"""
    from oracle.cpu_ops import CpuOps
    from phare_b200.messenger import LocalComm
    from phare_b200.setup import build
    cells, dl, interp = 64, 0.2, 1
    pops, bfn = two_pop_1d(cells, dl)
    populate([cells], [dl], interp, pops, bfn, largest=[16])
    m = importlib.import_module("pybindlibs.cpp_1_1_2")
    etc = importlib.import_module("pybindlibs.cpp_etc")
    sim = m.make_simulator(etc.make_hierarchy())
    sim.initialize()
    assert len(sim.solver.patches) == 4 and sim.currentTime() == 0.0 and sim.endTime() == pytest.approx(0.02)
    for _ in range(3):
        sim.advance(sim.timeStep())
    assert sim.currentTime() == pytest.approx(0.015)
    with pytest.raises(RuntimeError):
        sim.initialize()

    def particles_fn(i, L, pid):
        x = (np.arange(L.ncells[0]) + L.amr_lower[0] + 0.5) * dl
        p = pops[i]
        P = cpu_ref.maxwellian(L, p["density"](x), [p[k](x) for k in ("vx", "vy", "vz")],
                               [p[k](x) for k in ("vthx", "vthy", "vthz")], p["charge"], p["ppc"], p["seed"])
        return P.soa()
    hand = build(CpuOps(1, interp), LocalComm(), [cells], (4,), interp, [dl],
                 [dict(name=p["name"], mass=p["mass"]) for p in pops], lambda c, x: bfn[c](x), particles_fn,
                 dict(resistivity=1e-3, hyper_resistivity=1e-3, Te=0.12))
    for _ in range(3):
        hand.advance_level(0.005)
    for attr, comp in (("B", 1), ("B", 2), ("E", 0), ("E", 1), ("Ne", None), ("Vi", 0)):
        got = gather(sim, attr, comp)
        for p in hand.patches:
            h = getattr(p, attr)
            assert np.array_equal(got[p.geom.id], hand.ops.get_field(h[comp] if comp is not None else h)), (attr, comp)
    # DataWrangler view: merged level-0 density over the physical nodes
    dw = m.DataWrangler(sim, sim.hier)
    merged = dw.sync_merge(dw.getPatchLevel(0).getDensity(), True)
    assert merged.shape == (cells + 1,) and abs(merged.mean() - 1.1) < 0.05


def test_refined_levels_and_open_boundaries_are_refused(cpu_backend):
    import pybindlibs.dictator as pp
    pops, bfn = two_pop_1d()
    populate([64], [0.2], 1, pops, bfn)
    pp.add_int("simulation/AMR/max_nbr_levels", 2)
    with pytest.raises(NotImplementedError):
        S.make_hierarchy()
    pp.add_int("simulation/AMR/max_nbr_levels", 1)
    pp.add_string("simulation/grid/boundary_type/x", "open")
    with pytest.raises(NotImplementedError):
        S.make_hierarchy()


def test_diagnostics_written_at_requested_timestamps(cpu_backend, tmp_path):
    pops, bfn = two_pop_1d(32)
    populate([32], [0.2], 1, pops[:1], bfn, steps=2, diag_dir=str(tmp_path), diag_times=[0.0, 0.01])
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    assert sim.dump_diagnostics(0.0, 0.005)
    sim.advance(0.005)
    assert not sim.dump_diagnostics(0.005, 0.005)
    sim.advance(0.005)
    assert sim.dump_diagnostics(0.01, 0.005)
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 6 and files[0].startswith("electromag_EM_B_00000.00000")
    assert any(f.startswith("fluid_ions_pop_protons_flux_") for f in files)
    z = np.load(tmp_path / [f for f in files if f.startswith("electromag_EM_E")][-1])
    key = [k for k in z.files if k.endswith("EM_E_x")][0]
    assert key.startswith("t0.0100000000/pl0/p0/") and z[key].shape == (32 + 4,)
    # particle diagnostics: the domain particles of the population with the reference's SoA keys, last timestamp only
    z = np.load(tmp_path / [f for f in files if f.startswith("particle_")][0])
    assert z["t0.0100000000/pl0/p0/domain/iCell"].shape == (32 * 40, 1) and z["t0.0100000000/pl0/p0/domain/v"].shape == (32 * 40, 3)
    assert abs(z["t0.0100000000/pl0/p0/domain/weight"].sum() / 32 - 1.0) < 0.2


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyphare")), reason="reference tree not mounted")
def test_pyphare_runs_the_reference_harris_script_unchanged(cpu_backend, tmp_path, monkeypatch):
    """tests/functional/harris/harris_2d.py (config 3's input script), imported as is, driven by pyphare's own
    Simulator through pybindlibs: populateDict -> make_hierarchy -> make_simulator -> initialize -> advance -> dumps.
    Only the domain size (a module global of the script) is shrunk, and the root level alone is run."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections",
                 "matplotlib.colors", "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
        if name not in sys.modules:  # plotting / HDF5 readers of pharesee, not installed here and not used
            monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    monkeypatch.setenv("PHARE_B200_SINGLE_LEVEL", "1")
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(REF)  # the script does `from tests.simulator import SimulatorTest`
    stale = [k for k in sys.modules if k == "tests" or k.startswith("tests.")]
    for k in stale:
        monkeypatch.delitem(sys.modules, k)
    h = importlib.import_module("tests.functional.harris.harris_2d")
    h.cells, h.final_time, h.timestamps, h.diag_dir = (60, 100), 0.01, np.array([0.0, 0.01]), str(tmp_path / "out")
    sim = h.config()
    from pyphare.simulator.simulator import Simulator
    import pyphare.pharein as ph
    with pytest.warns(UserWarning, match="root level only"):
        simulator = Simulator(sim, log_to_file=False)
        simulator.initialize()
    assert simulator.currentTime() == 0.0 and simulator.timeStep() == 0.005 and simulator.interp_order() == 1
    simulator.advance().advance()
    assert simulator.currentTime() == pytest.approx(0.01)
    cpp_sim = simulator.cpp_sim
    pop = cpp_sim.solver.patches[0].pops[0]
    assert cpp_sim.solver.ops.count(pop.domain) == 60 * 100 * 100  # L0 particle number conserved
    # seed 12334 in the script -> parity loader -> the reference's initial deltas; density follows the profile
    dw = simulator.data_wrangler()
    ne = dw.getPatchLevel(0).getDensity()[0]
    rho = np.asarray(ne.data).reshape(60 + 1 + 4, 100 + 1 + 4)[2:-2, 2:-2]
    y = np.arange(101) * 0.4
    want = 0.4 + 1 / np.cosh((y - 12.0) / 0.5) ** 2 + 1 / np.cosh((y - 28.0) / 0.5) ** 2
    err = np.abs(rho.mean(axis=0) - want)
    assert np.max(err) < 0.2 and np.max(err[:15]) < 0.02  # the order-1 deposit smooths the 0.5-wide sheets over dl = 0.4
    assert sorted(os.listdir(tmp_path / "out"))[0].startswith("electromag_EM_B_00000.00000")
    simulator.reset()
    ph.global_vars.sim = None
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        del sys.modules[k]


def test_static_refinement_boxes_run_through_the_dict(cpu_backend, cpu_ref, tmp_path):
    """simulation/AMR/refinement/boxes/L0/B0/... -> a two-level hierarchy (phare_b200.amr): 4 sub-cycles of level 1 per
    step, both levels visible through the DataWrangler and in the diagnostics"""
    import pybindlibs.dictator as pp
    pops, bfn = two_pop_1d(64)
    populate([64], [0.2], 1, pops[:1], bfn, steps=2, largest=[32], diag_dir=str(tmp_path), diag_times=[0.0, 0.01])
    pp.add_int("simulation/AMR/max_nbr_levels", 2)
    pp.add_int("simulation/AMR/refinement/boxes/nbr_levels/", 1)
    pp.add_int("simulation/AMR/refinement/boxes/L0/nbr_boxes/", 1)
    pp.add_int("simulation/AMR/refinement/boxes/L0/B0/lower/x/", 20)
    pp.add_int("simulation/AMR/refinement/boxes/L0/B0/upper/x/", 43)
    hier = S.make_hierarchy()
    assert hier.refinement_boxes == [[([20], [43])]]
    sim = S.make_simulator(hier, 1, 1, 2)
    sim.initialize()
    assert len(sim.level_solvers()) == 2 and sim.level_solvers()[1].patches[0].layout.level == 1
    m = importlib.import_module("pybindlibs.cpp_1_1_2")
    dw = m.DataWrangler(sim, sim.hier)
    assert dw.getNumberOfLevels() == 2
    sim.advance(0.005)
    sim.advance(0.005)
    assert sim.currentTime() == pytest.approx(0.01) and sim.amr.time == pytest.approx(0.01)
    fine = dw.getPatchLevel(1).getDensity()
    assert len(fine) == 1 and fine[0].patchID == "1#0" and list(fine[0].lower) == [40] and list(fine[0].upper) == [87]
    rho = np.asarray(fine[0].data)[2:-2]
    assert rho.shape == (49,) and np.all(np.isfinite(rho)) and abs(rho.mean() - 1.0) < 0.3
    parts = dw.getPatchLevel(1).getParticles("protons")["protons"]
    dom, lg = parts["domain"][0].data, parts["levelGhost"][0].data
    assert dom.size() == sim.solver.ops.count(sim.level_solvers()[1].patches[0].pops[0].domain) > 0
    assert dom.iCell.shape == (dom.size(),) and dom.v.shape == (3 * dom.size(),) and lg.size() > 0
    assert dom.iCell.min() >= 40 and dom.iCell.max() <= 87 and sim.domain_box() == [63]
    with pytest.raises(RuntimeError):
        dw.getPatchLevel(2)
    assert sim.dump_diagnostics(0.01, 0.005)
    z = np.load(tmp_path / sorted(os.listdir(tmp_path))[-1])
    assert any("/pl1/p0/" in k for k in z.files) and any("/pl0/p1/" in k for k in z.files)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyphare")), reason="reference tree not mounted")
def test_pyphare_runs_the_reference_td1d_script_unchanged(cpu_backend, cpu_ref, tmp_path, monkeypatch):
    """tests/functional/td/td1d.py (config 2: 1-D tangential discontinuity with refinement boxes on L0 AND L1, i.e. three
    levels), imported as is and driven by pyphare's own Simulator through pybindlibs; two root steps = 8 + 32 sub-cycles"""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections",
                 "matplotlib.colors", "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        monkeypatch.delitem(sys.modules, k)
    td = importlib.import_module("tests.functional.td.td1d")
    sim = td.config()
    from pyphare.simulator.simulator import Simulator
    import pyphare.pharein as ph
    simulator = Simulator(sim, log_to_file=False)
    simulator.initialize()
    cpp_sim = simulator.cpp_sim
    levels = cpp_sim.level_solvers()
    assert [len(s.patches) for s in levels] == [25, 1, 1]
    assert [(int(s.patches[0].geom.box.lo[0]), int(s.patches[0].geom.box.hi[0])) for s in levels[1:]] == [(160, 361), (400, 601)]
    simulator.advance().advance()
    assert simulator.currentTime() == pytest.approx(0.02)
    ops = cpp_sim.solver.ops
    assert sum(ops.count(p.pops[0].domain) for p in levels[0].patches) == 500 * 100  # L0 particle number conserved
    for s in levels[1:]:
        p = s.patches[0]
        n = ops.count(p.pops[0].domain)
        assert abs(n - 101 * 100 * 2) < 0.03 * 101 * 100 * 2      # refined_particle_nbr = 2 children per coarse particle
        by = ops.get_field(p.B[1])
        assert np.all(np.isfinite(by)) and np.all(np.isfinite(ops.get_field(p.Ne)[2:-2]))
    # the tangential discontinuity at x = L/4 = 125 lies inside both refined levels: By goes from -1 to +1 across it
    by = ops.get_field(levels[2].patches[0].B[1])[2:-2]
    assert by[0] < -0.9 and by[-1] > 0.9 and np.all(np.diff(by) > -0.05)
    simulator.reset()
    ph.global_vars.sim = None
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyphare")), reason="reference tree not mounted")
def test_pharesee_post_processing_runs_on_the_npz_diagnostics(cpu_backend, cpu_ref, tmp_path, monkeypatch):
    """the dumps of a three-level run (the reference's td1d.py, unchanged) are loaded into pyphare's own
    PatchHierarchy and pharesee's flat_finest_field assembles the finest available data over the whole domain"""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections",
                 "matplotlib.colors", "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        monkeypatch.delitem(sys.modules, k)
    td = importlib.import_module("tests.functional.td.td1d")
    sim = td.config()
    from pyphare.simulator.simulator import Simulator
    import pyphare.pharein as ph
    simulator = Simulator(sim, log_to_file=False)
    simulator.initialize()          # dumps t = 0 (td1d asks for E, B, charge density and bulk velocity)
    from phare_b200.pharesee_npz import hierarchy_from_npz
    from pyphare.pharesee.hierarchy.hierarchy_utils import flat_finest_field
    hier = hierarchy_from_npz(str(tmp_path / "td_noflow"), time=0.0)
    assert sorted(hier.levels().keys()) == [0, 1, 2] and [len(l.patches) for l in hier.levels().values()] == [25, 1, 1]
    assert {"Bx", "By", "Bz", "Ex", "Ey", "Ez", "rho", "Vx", "Vy", "Vz"} <= set(hier.level(2).patches[0].patch_datas)
    by, x = flat_finest_field(hier, "By")
    order = np.argsort(x)
    x, by = x[order], by[order]
    # more points than the root level alone: the refined regions contribute their own nodes
    assert len(x) > 500 + 100 and x.min() < 1.0 and x.max() > 499.0
    assert np.min(np.diff(x[(x > 125) & (x < 150)])) == pytest.approx(0.25)   # level 2 mesh size inside its patch
    S_ = lambda x, x0: 0.5 * (1 + np.tanh(x - x0))
    want = -1 + 2 * (S_(x, 125.0) - S_(x, 375.0))
    assert np.max(np.abs(by - want)) < 0.4 and np.max(np.abs(by - want)[np.abs(x - 125) > 10]) < 2e-2
    rho, xr = flat_finest_field(hier, "rho")
    assert abs(np.mean(rho[(xr > 5) & (xr < 495)]) - 1.0) < 0.05
    simulator.reset()
    ph.global_vars.sim = None
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyphare")), reason="reference tree not mounted")
def test_npz_reader_names_population_quantities(cpu_backend, cpu_ref, tmp_path, monkeypatch):
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections",
                 "matplotlib.colors", "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    import pybindlibs.dictator as pp
    pops, bfn = two_pop_1d(64)
    populate([64], [0.2], 1, pops, bfn, steps=1, largest=[32], diag_dir=str(tmp_path), diag_times=[0.0])
    pp.add_int("simulation/AMR/max_nbr_levels", 2)
    pp.add_int("simulation/AMR/refinement/boxes/nbr_levels/", 1)
    pp.add_int("simulation/AMR/refinement/boxes/L0/nbr_boxes/", 1)
    pp.add_int("simulation/AMR/refinement/boxes/L0/B0/lower/x/", 20)
    pp.add_int("simulation/AMR/refinement/boxes/L0/B0/upper/x/", 43)
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    assert sim.dump_diagnostics(0.0, 0.005)
    from phare_b200.pharesee_npz import hierarchy_from_npz
    hier = hierarchy_from_npz(str(tmp_path))
    fine = hier.level(1).patches[0]
    assert {"Bx", "Ez", "protons_Fx", "beam_Fz"} <= set(fine.patch_datas) and fine.box.shape[0] == 48
    fx = fine.patch_datas["beam_Fx"]
    ops = sim.solver.ops
    want = ops.get_field(sim.level_solvers()[1].patches[0].pops[1].flux[0])
    assert np.array_equal(fx.dataset[:], want) and fx.ghosts_nbr[0] == 2 and abs(fx.x[2] - 40 * 0.1) < 1e-12


def _state(sim):
    """every field and particle store of every level, as numpy"""
    ops = sim.solver.ops
    out = {}
    for il, solver in enumerate(sim.level_solvers()):
        for p in solver.patches:
            base = f"{il}/{p.geom.id}/"
            for name, vec in (("B", p.B), ("E", p.E), ("Vi", p.Vi)):
                for c in range(3):
                    out[base + name + str(c)] = ops.get_field(vec[c])
            out[base + "Ne"] = ops.get_field(p.Ne)
            for i, pop in enumerate(p.pops):
                for store in S.Simulator._PARTICLE_STORES:
                    st = getattr(pop, store)
                    if st is not None:
                        for k, a in zip(("iCell", "delta", "w", "q", "v"), ops.get_particles(st)):
                            out[base + f"pop{i}/{store}/{k}"] = a
    return out


@pytest.mark.parametrize("refined", [False, True])
def test_restart_resumes_bit_for_bit(cpu_backend, cpu_ref, tmp_path, refined):
    """restarts (simulator.hpp:123 dump_restarts, restarts_manager.hpp): a run stopped after the restart written at
    t = 2 dt and resumed from it by a NEW simulator (dict with restarts/loadPath + restart_time, as pyphare's populateDict
    adds them) ends in the same state, bit for bit, as the run that went straight through; one level and two levels"""
    import pybindlibs.dictator as pp
    etc = importlib.import_module("pybindlibs.cpp_etc")
    dt = 0.005

    def make(extra):
        pops, bfn = two_pop_1d(64)
        populate([64], [0.2], 1, pops, bfn, steps=4, largest=[16] if not refined else [32])
        if refined:
            pp.add_int("simulation/AMR/max_nbr_levels", 2)
            pp.add_int("simulation/AMR/refinement/boxes/nbr_levels/", 1)
            pp.add_int("simulation/AMR/refinement/boxes/L0/nbr_boxes/", 1)
            pp.add_int("simulation/AMR/refinement/boxes/L0/B0/lower/x/", 20)
            pp.add_int("simulation/AMR/refinement/boxes/L0/B0/upper/x/", 43)
        pp.add_string("simulation/restarts/filePath", str(tmp_path))
        pp.add_string("simulation/restarts/serialized_simulation", "the-serialized-simulation")
        extra()
        sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
        sim.initialize()
        return sim

    sim = make(lambda: pp.add_array_as_vector("simulation/restarts/write_timestamps", np.array([2 * dt])))
    made = []
    for k in range(4):
        made.append(sim.dump_restarts(sim.currentTime(), dt))  # pyphare: restarts.dump before every advance
        sim.advance(dt)
    assert made == [False, False, True, False]
    straight = _state(sim)
    load = etc.restart_path_for_time(str(tmp_path), 2 * dt)
    assert os.path.basename(load) == "00000.01000" and os.path.isdir(load)
    assert etc.serialized_simulation_string(load) == "the-serialized-simulation" and etc.patch_data_ids(load) == []
    assert os.path.exists(etc.samrai_restart_file(load))

    def loading():
        pp.add_string("simulation/restarts/loadPath", load)
        pp.add_double("simulation/restarts/restart_time", 2 * dt)
    S.dict_instance().stop()
    sim2 = make(loading)
    assert sim2.startTime() == pytest.approx(2 * dt) and sim2.currentTime() == pytest.approx(2 * dt)
    assert not sim2.dump_restarts(sim2.currentTime(), dt)  # loading only: no writer properties
    for k in range(2):
        sim2.advance(dt)
    assert sim2.currentTime() == pytest.approx(4 * dt)
    resumed = _state(sim2)
    assert straight.keys() == resumed.keys()
    for k in straight:
        assert np.array_equal(straight[k], resumed[k], equal_nan=straight[k].dtype.kind == "f"), k


def test_restart_file_of_another_decomposition_is_refused(cpu_backend, cpu_ref, tmp_path):
    """a restart file belongs to one rank of one decomposition: loading it into a run cut into other patches (or with other
    populations) says so instead of failing on a missing key"""
    import pybindlibs.dictator as pp

    def make(largest):
        pops, bfn = two_pop_1d(64)
        populate([64], [0.2], 1, pops, bfn, steps=2, largest=largest)
        sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
        sim.initialize()
        return sim

    sim = make([16])
    path = str(tmp_path / "restart_rank0.npz")
    sim.save_restart(path)
    S.dict_instance().stop()
    same = make([16])
    same.load_restart(path)  # fits
    S.dict_instance().stop()
    other = make([32])
    with pytest.raises(RuntimeError, match="does not fit this run.*different patches"):
        other.load_restart(path)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyphare")), reason="reference tree not mounted")
def test_pyphare_restart_options_write_and_resume(cpu_backend, cpu_ref, tmp_path, monkeypatch):
    """pyphare's own restart flow, unchanged: restart_options={"dir", "mode", "timestamps"} writes <dir>/00000.01000/ during
    the run; a second Simulation with restart_options["restart_time"] passes pyphare's checks (restart_path_for_time,
    serialized_simulation_string -> is_restartable_compared_to, patch_data_ids), starts at that time and ends bit for bit
    where the straight run did"""
    pytest.importorskip("dill")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections",
                 "matplotlib.colors", "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    monkeypatch.chdir(tmp_path)
    import pyphare.pharein as ph
    from pyphare.simulator.simulator import Simulator
    L = 64 * 0.2
    zero = lambda x: 0. * x
    vth = lambda x: 0.3 + 0. * x

    def run(restart_options, nsteps):
        ph.global_vars.sim = None
        sim = ph.Simulation(time_step=0.005, time_step_nbr=nsteps, cells=64, dl=0.2, interp_order=1, largest_patch_size=16,
                            restart_options=restart_options,
                            diag_options={"format": "phareh5", "options": {"dir": "diags", "mode": "overwrite"}})
        ph.MaxwellianFluidModel(
            bx=lambda x: 1. + 0 * x, by=lambda x: 0.1 * np.cos(2 * np.pi * x / L), bz=zero,
            protons={"charge": 1, "density": lambda x: 1. + 0.2 * np.sin(2 * np.pi * x / L), "vbulkx": zero, "vbulky": zero,
                     "vbulkz": zero, "vthx": vth, "vthy": vth, "vthz": vth, "nbr_part_per_cell": 40, "init": {"seed": 12}})
        ph.ElectronModel(closure="isothermal", Te=0.12)
        s = Simulator(sim, log_to_file=False)
        try:
            s.initialize()
            t0 = s.currentTime()
            for _ in range(nsteps):
                s.advance()
            c = s.cpp_sim
            ops = c.solver.ops
            out = [ops.get_field(f) for p in c.solver.patches for f in (*p.B, *p.E, p.Ne)]
            out += [a for p in c.solver.patches for a in ops.get_particles(p.pops[0].domain)]
            return t0, s.currentTime(), out
        finally:
            s.reset()
            ph.global_vars.sim = None

    t0, t1, straight = run({"dir": "rst", "mode": "overwrite", "timestamps": [0.01]}, 4)
    assert (t0, t1) == (0.0, pytest.approx(0.02)) and os.listdir(tmp_path / "rst") == ["00000.01000"]
    t0, t1, resumed = run({"dir": "rst", "mode": "overwrite", "restart_time": 0.01}, 2)
    assert (t0, t1) == (pytest.approx(0.01), pytest.approx(0.02))
    assert len(straight) == len(resumed)
    for a, b in zip(straight, resumed):
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
    with pytest.raises(ValueError):  # no restart was written for that time
        run({"dir": "rst", "mode": "overwrite", "restart_time": 0.015}, 1)
