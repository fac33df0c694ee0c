"""refinement="tagging": the reference's tagging criterion restated (DefaultTaggerStrategy, default_tagger_strategy.hpp),
the tile clustering that stands in for SAMRAI's GriddingAlgorithm, and tagging-driven regridding through the dict-driven
front end (CPU back end)."""
import itertools

import numpy as np
import pytest

import phare_b200.simulator as S
from phare_b200.boxes import Box
from phare_b200.tagging import default_tags, cluster, _dilate
from frontend_util import populate, const


def _tags_loops(dim, interp, B, ncells, threshold):
    """the reference's loops, literally (default_tagger_strategy.hpp:73-196)"""
    g = 2 if interp == 1 else 4
    Bx, By, Bz = B
    tags = np.zeros(ncells, np.int32)
    if dim == 1:
        last = ncells[0] - 1 + (1 if g > 2 else 0)
        for ic in range(last):
            ix = g + ic
            byavg = 0.2 * (By[ix - 2] + By[ix - 1] + By[ix] + By[ix + 1] + By[ix + 2])
            bzavg = 0.2 * (Bz[ix - 2] + Bz[ix - 1] + Bz[ix] + Bz[ix + 1] + Bz[ix + 2])
            byp1 = 0.2 * (By[ix - 1] + By[ix] + By[ix + 1] + By[ix + 2] + By[ix + 3])
            bzp1 = 0.2 * (Bz[ix - 1] + Bz[ix] + Bz[ix + 1] + Bz[ix + 2] + Bz[ix + 3])
            cby, cbz = abs(byp1 - byavg) / (1 + abs(byavg)), abs(bzp1 - bzavg) / (1 + abs(bzavg))
            tags[ic] = 1 if np.sqrt(cby * cby + cbz * cbz) > threshold else 0
        return tags
    for t in itertools.product(*[range(n) for n in ncells]):
        i = [g + k for k in t]
        crit = 0.0
        for F in (Bx, By, Bz):
            for d in range(dim):
                p1, p2 = list(i), list(i)
                p1[d] += 1
                p2[d] += 2
                crit = max(crit, abs(F[tuple(p2)] - F[tuple(i)]) / (1 + abs(F[tuple(p1)] - F[tuple(i)])))
        tags[t] = 1 if crit > threshold else 0
    return tags


@pytest.mark.parametrize("dim,interp", [(1, 1), (1, 2), (2, 1), (2, 3), (3, 1)])
def test_default_tags_follow_the_reference_loops(dim, interp):
    rng = np.random.default_rng(10 * dim + interp)
    g = 2 if interp == 1 else 4
    ncells = [14, 9, 7][:dim]
    prim = lambda c, d: 1 if c == d else 0
    amp = 0.2 if dim == 1 else 0.04
    B = [rng.standard_normal(tuple(ncells[d] + 2 * g + prim(c, d) for d in range(dim))) * amp for c in range(3)]
    got = default_tags(dim, interp, B, ncells, 0.1)
    want = _tags_loops(dim, interp, B, ncells, 0.1)
    assert np.array_equal(got, want) and 0 < got.sum() < got.size
    if dim == 1 and interp == 1:
        assert got[-1] == 0      # the last cell is never tagged with two ghost cells


def test_tile_clustering_covers_the_tags_inside_the_allowed_region():
    mask = np.zeros((40, 30), bool)
    mask[10:13, 5:7] = True          # a small blob
    mask[25:33, 20:22] = True        # an elongated one
    mask[1, 1] = True                # in a tile that is not entirely allowed
    allowed = [Box([4, 4], [35, 27])]
    boxes = cluster(_dilate(mask, 1), [0, 0], allowed, 4, largest=[8, 8])
    assert boxes and all(any(a.contains(b) for a in allowed) for b in boxes)
    for a, b in itertools.combinations(boxes, 2):
        assert a * b is None                                   # disjoint
    for b in boxes:
        assert all(int(x) % 4 == 0 for x in b.lo) and all(int(x + 1) % 4 == 0 for x in b.hi)
        assert all(s <= 8 for s in b.shape())
    covered = np.zeros_like(mask)
    for b in boxes:
        covered[int(b.lo[0]):int(b.hi[0]) + 1, int(b.lo[1]):int(b.hi[1]) + 1] = True
    inside = np.zeros_like(mask)
    inside[4:36, 4:28] = True
    assert covered[mask & inside].all() and not covered[1, 1]
    assert covered.sum() < 0.4 * mask.size                      # and not much more than that
    # 1-D: two separate runs, a shifted mask origin
    m1 = np.zeros(64, bool)
    m1[[10, 11, 40]] = True
    b1 = cluster(m1, [16], [Box([0], [200])], 8)
    assert [(int(b.lo[0]), int(b.hi[0])) for b in b1] == [(24, 31), (56, 63)]


@pytest.fixture()
def cpu_backend(cpu_oracle, cpu_ref):
    from oracle.cpu_ops import CpuOps
    old = S.ops_factory
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    yield
    S.ops_factory = old
    S.dict_instance().stop()


def test_tagging_refines_the_current_sheets_and_follows_them(cpu_backend):
    """a 1-D double tangential discontinuity (tests/functional/tdtagged/td1dtagged.py scaled down) with
    refinement="tagging", three levels: the refined levels sit on the two discontinuities, nested, and the run advances
    through regrids"""
    import pybindlibs.dictator as pp
    cells, dl = 200, 0.5
    Lx = cells * dl
    Sx = lambda x, x0: 0.5 * (1 + np.tanh((x - x0) / 1.0))
    by = lambda x: -1 + 2 * (Sx(x, 0.25 * Lx) - Sx(x, 0.75 * Lx))
    T = lambda x: np.sqrt(np.maximum(1 - 0.5 * (by(x) ** 2 + 0.25), 0.05))
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=20, seed=7, density=const(1.0), vx=const(0), vy=const(0),
               vz=const(0), vthx=T, vthy=T, vthz=T)
    populate([cells], [dl], 1, [pop], [const(0.0), by, const(0.5)], time_step=0.005, steps=4, largest=[50])
    pp.add_int("simulation/AMR/max_nbr_levels", 3)
    pp.add_string("simulation/AMR/refinement/tagging/method", "auto")
    pp.add_double("simulation/AMR/refinement/tagging/threshold", 0.1)
    pp.add_int("simulation/AMR/tag_buffer", 2)
    pp.add_vector_int("simulation/AMR/smallest_patch_size", [8])
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    levels = sim.amr.levels
    assert len(levels) == 3
    for il, lvl in enumerate(levels[1:], start=1):
        boxes = [p.box for p in lvl.geom.patches]
        assert len(boxes) == 2                                 # one patch per discontinuity
        for b, x0 in zip(boxes, (0.25 * Lx, 0.75 * Lx)):
            dx = dl / 2 ** il
            assert b.lo[0] * dx < x0 - 0.75 and (b.hi[0] + 1) * dx > x0 + 0.75   # the sheet (half width 1) is inside
            assert (b.hi[0] - b.lo[0] + 1) * dx < 0.2 * Lx     # local refinement, not the whole domain
        # proper nesting in the level below
        below = [p.box for p in levels[il - 1].geom.patches]
        for b in boxes:
            cb = Box(b.lo // 2, b.hi // 2)
            assert il == 1 or any(q.contains(cb.grow(1)) for q in below)
    n1 = [sim.solver.ops.count(p.pops[0].domain) for p in levels[1].solver.patches]
    for _ in range(3):
        sim.advance(0.005)
    assert len(sim.amr.levels) == 3 and sim.currentTime() == pytest.approx(0.015)
    for lvl in sim.amr.levels[1:]:
        for p in lvl.solver.patches:
            assert np.isfinite(sim.solver.ops.get_field(p.B[1])).all()
            assert np.isfinite(sim.solver.ops.get_field(p.Ne)[2:-2]).all()
    # an unchanged hierarchy is kept, not rebuilt
    before = list(sim.amr.levels)
    assert sim.amr.regrid_tagged(sim.tagger) is False and all(a is b for a, b in zip(before, sim.amr.levels))
    assert [sim.solver.ops.count(p.pops[0].domain) for p in sim.amr.levels[1].solver.patches][0] == pytest.approx(n1[0], rel=0.05)


REF = "/root/reference"


@pytest.mark.skipif(not __import__("os").path.isdir(REF + "/pyphare"), reason="reference tree not mounted")
def test_pyphare_runs_the_reference_tagged_script_unchanged(cpu_backend, tmp_path, monkeypatch):
    """tests/functional/tdtagged/td1dtagged.py::withTagging (refinement="tagging", max_nbr_levels=3, a discontinuity
    advected at vx = 2) driven by pyphare's own Simulator: the refined levels sit on the two discontinuities and follow
    them"""
    import importlib
    import os
    import sys
    from unittest import mock
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections", "matplotlib.colors",
                 "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py", "ddt"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        monkeypatch.delitem(sys.modules, k)
    td = importlib.import_module("tests.functional.tdtagged.td1dtagged")
    sim = td.withTagging(str(tmp_path / "out"))
    from pyphare.simulator.simulator import Simulator
    import pyphare.pharein as ph
    simulator = Simulator(sim, log_to_file=False)
    simulator.initialize()
    amr = simulator.cpp_sim.amr
    assert len(amr.levels) == 3
    centre = lambda lvl: [0.5 * (int(p.box.lo[0]) + int(p.box.hi[0]) + 1) / 2 ** lvl.number for p in lvl.geom.patches]
    for lvl in amr.levels[1:]:
        c = centre(lvl)
        assert len(c) == 2 and abs(c[0] - 50) < 3 and abs(c[1] - 150) < 3      # L = 200: sheets at L/4 and 3L/4
    c0 = centre(amr.levels[2])
    for _ in range(25):                                                        # t = 1: the flow carried them 2 cells
        simulator.advance()
    amr = simulator.cpp_sim.amr
    assert len(amr.levels) == 3 and simulator.currentTime() == pytest.approx(1.0)
    hi0 = [int(p.box.hi[0]) for p in amr.levels[2].geom.patches]
    assert all(h / 4 > x + 1.0 for h, x in zip(hi0, (52.0, 152.0)))            # the finest level reaches past the moved sheets
    assert all(b >= a for a, b in zip(c0, centre(amr.levels[2])))
    ops = simulator.cpp_sim.solver.ops
    for lvl in amr.levels:
        for p in lvl.solver.patches:
            assert np.isfinite(ops.get_field(p.B[1])).all()
    simulator.reset()
    ph.global_vars.sim = None
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        del sys.modules[k]


@pytest.mark.skipif(not __import__("os").path.isdir(REF + "/pyphare"), reason="reference tree not mounted")
def test_pyphare_runs_the_reference_harris_script_with_its_tagging(cpu_backend, tmp_path, monkeypatch):
    """tests/functional/harris/harris_2d.py as it is (refinement="tagging", two levels; only the domain is shrunk): the
    refined level is made of two strips along the current sheets at y = 0.3 Ly and 0.7 Ly, spanning the periodic x
    extent of the domain (the level is then periodic in x like the domain)."""
    import importlib
    import os
    import sys
    from unittest import mock
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.collections", "matplotlib.colors",
                 "matplotlib.lines", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py", "ddt"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, mock.MagicMock(name=name))
    monkeypatch.syspath_prepend(os.path.join(REF, "pyphare"))
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        monkeypatch.delitem(sys.modules, k)
    h = importlib.import_module("tests.functional.harris.harris_2d")
    h.cells, h.final_time, h.timestamps, h.diag_dir = (60, 100), 0.005, np.array([0.0]), str(tmp_path / "out")
    sim = h.config()
    assert sim.refinement == "tagging" and sim.max_nbr_levels == 2
    from pyphare.simulator.simulator import Simulator
    import pyphare.pharein as ph
    simulator = Simulator(sim, log_to_file=False)
    simulator.initialize()
    amr = simulator.cpp_sim.amr
    assert len(amr.levels) == 2
    boxes = sorted(((p.box.lo.tolist(), p.box.hi.tolist()) for p in amr.levels[1].geom.patches), key=lambda b: b[0][1])
    assert len(boxes) == 2
    for (lo, hi), y0 in zip(boxes, (0.3 * 40.0, 0.7 * 40.0)):
        assert lo[1] * 0.2 < y0 - 1.0 and (hi[1] + 1) * 0.2 > y0 + 1.0          # the sheet (half width 0.5) is inside
        assert (hi[1] - lo[1] + 1) * 0.2 < 8.0 and (lo[0], hi[0]) == (0, 119)   # a strip along the whole x extent
    assert amr.levels[1].geom.periodic
    simulator.advance()
    ops = simulator.cpp_sim.solver.ops
    assert sum(ops.count(p.pops[0].domain) for p in amr.levels[0].solver.patches) == 60 * 100 * 100
    for p in simulator.cpp_sim.amr.levels[1].solver.patches:
        assert not np.isnan(ops.get_field(p.B[0])).any() and np.isfinite(ops.get_field(p.Ne)[2:-2, 2:-2]).all()
    simulator.reset()
    ph.global_vars.sim = None
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        del sys.modules[k]


def test_restart_rebuilds_a_tagged_hierarchy(cpu_backend, tmp_path):
    """load_restart on a simulator whose hierarchy differs from the saved one (here: its refined levels removed) rebuilds
    the levels on the saved boxes before it overwrites their data: the resumed run equals the straight one bit for bit"""
    import pybindlibs.dictator as pp
    cells, dl = 100, 0.5
    Lx = cells * dl
    Sx = lambda x, x0: 0.5 * (1 + np.tanh((x - x0) / 1.0))
    by = lambda x: -1 + 2 * (Sx(x, 0.25 * Lx) - Sx(x, 0.75 * Lx))
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=20, seed=7, density=const(1.0), vx=const(0), vy=const(0),
               vz=const(0), vthx=const(0.4), vthy=const(0.4), vthz=const(0.4))

    def make():
        S.dict_instance().stop()
        populate([cells], [dl], 1, [pop], [const(0.0), by, const(0.5)], time_step=0.005, steps=4, largest=[50])
        pp.add_int("simulation/AMR/max_nbr_levels", 2)
        pp.add_string("simulation/AMR/refinement/tagging/method", "auto")
        pp.add_double("simulation/AMR/refinement/tagging/threshold", 0.1)
        pp.add_vector_int("simulation/AMR/smallest_patch_size", [8])
        sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
        sim.initialize()
        return sim

    def fields(sim):
        ops = sim.solver.ops
        return [ops.get_field(f) for s in sim.level_solvers() for p in s.patches for f in (*p.B, *p.E, p.Ne)] + \
               [a for s in sim.level_solvers() for p in s.patches for a in ops.get_particles(p.pops[0].domain)]

    sim = make()
    sim.advance(0.005)
    f = str(tmp_path / "r.npz")
    sim.save_restart(f)
    saved_boxes = [p.box for p in sim.amr.levels[1].geom.patches]
    sim.advance(0.005)
    straight = fields(sim)
    sim2 = make()
    del sim2.amr.levels[1:]
    sim2.load_restart(f)
    assert len(sim2.amr.levels) == 2 and sim2.currentTime() == pytest.approx(0.005)
    assert [p.box for p in sim2.amr.levels[1].geom.patches] == saved_boxes
    sim2.advance(0.005)
    resumed = fields(sim2)
    assert len(straight) == len(resumed)
    for a, b in zip(straight, resumed):
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
