"""Physics-level functional check (the reference's tests/functional/alfven_wave/alfven_wave1d.py scaled
down): a circularly polarised Alfven wave on a uniform plasma must propagate at the Alfven speed.  This validates the
restated SolverPPC sequencing (predictor / predictor / corrector with the time-centred particle sweeps) against physics
rather than against the oracle restatement of the same sequencing: a wrong sub-step order, time centring or sign shows up
as a wrong phase speed, growth or damping.  CPU back end (oracle kernels), dict-driven front end; ~15 s."""
import numpy as np
import pytest

import phare_b200.simulator as S
from frontend_util import populate, const


@pytest.fixture()
def cpu_backend(cpu_oracle):
    from oracle.cpu_ops import CpuOps
    old = S.ops_factory
    S.ops_factory = lambda dim, interp: CpuOps(dim, interp)
    yield
    S.ops_factory = old
    S.dict_instance().stop()


def _mode1(sim, getter, primal):
    import importlib
    m = importlib.import_module("pybindlibs.cpp_1_1_2")
    dw = m.DataWrangler(sim, sim.hier)
    a = dw.sync_merge(getattr(dw.getPatchLevel(0), getter)(), primal)
    n = sim.hier.cells[0]
    return np.fft.fft(a[:n])[1] / (n / 2)   # complex amplitude of the fundamental


def test_alfven_wave_propagates_at_the_alfven_speed(cpu_backend):
    _alfven_wave()


@pytest.mark.gpu
def test_alfven_wave_propagates_at_the_alfven_speed_on_the_gpu():
    """the same run on the CUDA path (GpuOps through the C ABI)"""
    try:
        _alfven_wave()
    finally:
        S.dict_instance().stop()


def _alfven_wave():
    cells, dl, ampl = 100, 1.0, 0.01
    Lx = cells * dl
    k = 2 * np.pi / Lx
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=100, seed=1337, density=const(1.0),
               vx=const(0.0), vy=lambda x: ampl * np.cos(k * x), vz=lambda x: ampl * np.sin(k * x),
               vthx=const(0.01), vthy=const(0.01), vthz=const(0.01))
    bfn = [const(1.0), lambda x: ampl * np.cos(k * x), lambda x: ampl * np.sin(k * x)]
    dt, nsteps, every = 0.02, 1000, 100
    populate([cells], [dl], 1, [pop], bfn, time_step=dt, steps=nsteps, eta=0.0, nu=1e-3, Te=0.0, largest=[50])
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    times, by, vy = [0.0], [_mode1(sim, "getBy", False)], [_mode1(sim, "getViy", True)]
    for step in range(1, nsteps + 1):
        sim.advance(dt)
        if step % every == 0:
            times.append(step * dt)
            by.append(_mode1(sim, "getBy", False))
            vy.append(_mode1(sim, "getViy", True))
    times, by, vy = np.array(times), np.array(by), np.array(vy)
    # amplitude: neither growing nor damped (hyper-resistivity 1e-3 at this wavelength is negligible)
    assert np.all(np.abs(np.abs(by) - ampl) < 0.1 * ampl), np.abs(by)
    # phase speed from the rotation of the fundamental: By ~ cos(k x - w t)  ->  arg(mode) = -(-w t) ...
    phase = np.unwrap(np.angle(by))
    omega = np.polyfit(times, phase, 1)[0]
    vphi = abs(omega) / k
    # the reference's acceptance: within 5 % of vA; sharper: a parallel circularly polarised wave in Hall MHD has
    # w = k vA (sqrt(1 + (k di / 2)^2) +- k di / 2) with k di = 0.063 -> 1.0319 or 0.9691 depending on the handedness
    # (measured: 1.0357, the + branch at 0.4 %)
    assert abs(vphi - 1.0) < 0.05, vphi
    branches = [np.sqrt(1 + (k / 2) ** 2) + k / 2, np.sqrt(1 + (k / 2) ** 2) - k / 2]
    assert min(abs(vphi - b) for b in branches) < 0.01, (vphi, branches)
    resid = phase - np.polyval(np.polyfit(times, phase, 1), times)
    assert np.max(np.abs(resid)) < 0.05                      # steady propagation
    # the bulk velocity perturbation travels with the field, in antiphase with it (v = -/+ b for propagation along +/- B0)
    dphi = np.angle(vy[1:] / by[1:])
    assert np.all(np.abs(np.abs(dphi) - np.pi) < 0.3) or np.all(np.abs(dphi) < 0.3)
    assert np.all(np.abs(np.abs(vy[1:]) - ampl) < 0.25 * ampl)


def _phase_speed_and_sim(refine_box, ppc=60, dt=0.02, nsteps=500, every=50):
    import pybindlibs.dictator as pp
    cells, dl, ampl = 100, 1.0, 0.01
    k = 2 * np.pi / (cells * dl)
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=ppc, seed=1337, density=const(1.0),
               vx=const(0.0), vy=lambda x: ampl * np.cos(k * x), vz=lambda x: ampl * np.sin(k * x),
               vthx=const(0.01), vthy=const(0.01), vthz=const(0.01))
    bfn = [const(1.0), lambda x: ampl * np.cos(k * x), lambda x: ampl * np.sin(k * x)]
    populate([cells], [dl], 1, [pop], bfn, time_step=dt, steps=nsteps, eta=0.0, nu=1e-3, Te=0.0, largest=[50])
    if refine_box is not None:
        pp.add_int("simulation/AMR/max_nbr_levels", 2)
        pp.add_int("simulation/AMR/refinement/boxes/nbr_levels/", 1)
        pp.add_int("simulation/AMR/refinement/boxes/L0/nbr_boxes/", 1)
        pp.add_int("simulation/AMR/refinement/boxes/L0/B0/lower/x/", refine_box[0])
        pp.add_int("simulation/AMR/refinement/boxes/L0/B0/upper/x/", refine_box[1])
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    times, by = [0.0], [_mode1(sim, "getBy", False)]
    for step in range(1, nsteps + 1):
        sim.advance(dt)
        if step % every == 0:
            times.append(step * dt)
            by.append(_mode1(sim, "getBy", False))
    times, by = np.array(times), np.array(by)
    assert np.all(np.abs(np.abs(by) - ampl) < 0.1 * ampl), np.abs(by)
    phase = np.unwrap(np.angle(by))
    fit = np.polyfit(times, phase, 1)
    assert np.max(np.abs(phase - np.polyval(fit, times))) < 0.05
    return abs(fit[0]) / k, sim


def test_alfven_wave_crosses_a_refined_level_unperturbed(cpu_backend, cpu_ref):
    """the same wave with 40 % of the domain refined (static box, 4 sub-cycles per step, level-ghost particles, refluxing):
    the restated multi-level sequencing has no runnable reference here, so it is held to the physics: the wave keeps its
    amplitude and its phase speed through the coarse-fine boundaries (against the single-level run over the same
    interval, 0 < t < 10, where the initial transient still makes both 2 % faster than the asymptotic value), and both
    levels carry the same field"""
    ampl = 0.01
    v_single, sim = _phase_speed_and_sim(None)
    S.dict_instance().stop()
    v_refined, sim = _phase_speed_and_sim((30, 69))
    assert len(sim.level_solvers()) == 2
    assert abs(v_refined - v_single) < 0.005 * v_single, (v_refined, v_single)     # measured: 1.0556 vs 1.0539
    assert abs(v_refined - 1.0) < 0.07
    # the refined level carries the same wave as the root level under it (By is dual: two fine cells per coarse cell)
    ops = sim.solver.ops
    fine = sim.level_solvers()[1].patches[0]
    g = 2
    by_f = ops.get_field(fine.B[1])[g:-g]
    by_c = np.concatenate([ops.get_field(p.B[1])[g:-g] for p in sim.level_solvers()[0].patches])[30:70]
    assert np.max(np.abs(0.5 * (by_f[0::2] + by_f[1::2]) - by_c)) < 0.1 * ampl
    # and no particle pile-up or depletion at the level boundary
    ne = ops.get_field(fine.Ne)[g:-g]
    assert abs(ne.mean() - 1.0) < 0.02 and np.max(np.abs(ne - 1.0)) < 0.25


def test_alfven_wave_along_y_in_2d(cpu_backend):
    """the same wave on a 2-D domain, propagating along y (fields and velocities cyclically permuted): the 2-D field
    solvers, gather and deposit treat the second direction like the first one.  0 < t < 10 as in the refined-level test,
    where the 1-D run gives 1.054."""
    import importlib
    cells, dl, ampl = [4, 100], [1.0, 1.0], 0.01
    k = 2 * np.pi / (cells[1] * dl[1])
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=50, seed=1337, density=const(1.0),
               vx=lambda x, y: ampl * np.sin(k * y), vy=const(0.0), vz=lambda x, y: ampl * np.cos(k * y),
               vthx=const(0.01), vthy=const(0.01), vthz=const(0.01))
    bfn = [lambda x, y: ampl * np.sin(k * y), const(1.0), lambda x, y: ampl * np.cos(k * y)]
    dt, nsteps, every = 0.02, 500, 50
    populate(cells, dl, 1, [pop], bfn, time_step=dt, steps=nsteps, eta=0.0, nu=1e-3, Te=0.0, largest=[4, 50])
    sim = S.make_simulator(S.make_hierarchy(), 2, 1, 4)
    sim.initialize()
    m = importlib.import_module("pybindlibs.cpp_2_1_4")
    dw = m.DataWrangler(sim, sim.hier)

    def mode1():
        bz = dw.sync_merge(dw.getPatchLevel(0).getBz(), False)      # dual in x and y
        return np.fft.fft(bz[:, :cells[1]].mean(axis=0))[1] / (cells[1] / 2)
    times, bz = [0.0], [mode1()]
    for step in range(1, nsteps + 1):
        sim.advance(dt)
        if step % every == 0:
            times.append(step * dt)
            bz.append(mode1())
    times, bz = np.array(times), np.array(bz)
    assert np.all(np.abs(np.abs(bz) - ampl) < 0.1 * ampl), np.abs(bz)
    phase = np.unwrap(np.angle(bz))
    fit = np.polyfit(times, phase, 1)
    vphi = abs(fit[0]) / k
    assert abs(vphi - 1.054) < 0.01, vphi
    assert np.max(np.abs(phase - np.polyval(fit, times))) < 0.05


def test_ion_ion_beam_right_hand_mode_grows_at_the_linear_rate(cpu_backend):
    """config 4's physics (tests/functional/ionIonBeam/ion_ion_beam1d.py scaled down: dt = 0.01 instead of 0.001, 25
    particles per cell and population, 0 < t < 40): a 1 % beam at 5 vA drives the right-hand resonant mode, whose first
    Fourier mode By - i Bz grows at 0.09 Omega_ci in the linear phase (the reference accepts 0.09 +- 0.02).  Two
    populations, so this also checks their coupling through the total moments."""
    import importlib
    cells, dl, dt, T = 165, 0.2, 0.01, 40.0
    vth = np.sqrt(0.1)
    common = dict(mass=1.0, charge=1.0, ppc=25, vy=const(0), vz=const(0), vthx=const(vth), vthy=const(vth), vthz=const(vth))
    main = dict(name="main", seed=5, density=const(1.0), vx=const(0.0), **common)
    beam = dict(name="beam", seed=6, density=const(0.01), vx=const(5.0), **common)
    populate([cells], [dl], 1, [main, beam], [const(1.0), const(0.0), const(0.0)], time_step=dt, steps=int(T / dt),
             eta=0.0, nu=0.01, Te=0.0)
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    dw = importlib.import_module("pybindlibs.cpp_1_1_2").DataWrangler(sim, sim.hier)

    def mode1():
        lvl = dw.getPatchLevel(0)
        by, bz = dw.sync_merge(lvl.getBy(), False)[:cells], dw.sync_merge(lvl.getBz(), False)[:cells]
        return abs(np.fft.fft(by - 1j * bz)[1])
    times, ampl = [], []
    for step in range(int(T / dt)):
        sim.advance(dt)
        if step % 100 == 0:
            times.append((step + 1) * dt)
            ampl.append(mode1())
    times, ampl = np.array(times), np.array(ampl)
    linear = (times > 10) & (times < 38)
    gamma = np.polyfit(times[linear], np.log(ampl[linear]), 1)[0]
    assert abs(gamma - 0.09) < 0.02, gamma                       # measured: 0.097
    assert ampl[-1] > 50 * ampl[1]                               # two orders of magnitude above the noise
    ops = sim.solver.ops
    p = sim.solver.patches[0]
    assert ops.count(p.pops[0].domain) == ops.count(p.pops[1].domain) == cells * 25
