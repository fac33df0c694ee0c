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
