"""Restarts and the h5-layout diagnostics on the CUDA back end (GpuOps).  Named to run LAST in the -m gpu suite: it was
written after this round's GPU budget was spent and has not run on a GPU yet; the CPU twins of these checks are
tests/test_frontend.py::test_restart_resumes_bit_for_bit and tests/test_h5lite.py."""
import numpy as np
import pytest

import phare_b200.simulator as S
from phare_b200 import h5lite
from frontend_util import populate, two_pop_1d, const, gather

pytestmark = pytest.mark.gpu


def _make(tmp_path, refined):
    import pybindlibs.dictator as pp
    S.dict_instance().stop()
    pops, bfn = two_pop_1d(64)
    populate([64], [0.2], 1, pops, bfn, steps=4, largest=[32] if refined else [16])
    if refined:
        pp.add_int("simulation/AMR/max_nbr_levels", 2)
        pp.add_int("simulation/AMR/refinement/boxes/nbr_levels/", 1)
        pp.add_int("simulation/AMR/refinement/boxes/L0/nbr_boxes/", 1)
        pp.add_int("simulation/AMR/refinement/boxes/L0/B0/lower/x/", 20)
        pp.add_int("simulation/AMR/refinement/boxes/L0/B0/upper/x/", 43)
    sim = S.make_simulator(S.make_hierarchy(), 1, 1, 2)
    sim.initialize()
    assert type(sim.solver.ops).__name__ == "GpuOps"
    return sim


def _fields(sim):
    ops = sim.solver.ops
    return [ops.get_field(f) for s in sim.level_solvers() for p in s.patches for f in (*p.B, *p.E, p.Ne, *p.Vi)]


@pytest.mark.parametrize("refined", [False, True])
def test_restart_resumes_on_the_gpu(tmp_path, refined):
    """save after 2 steps, 2 more; a new simulator loads the restart and takes the 2 steps: same particle counts, fields
    within the FP64-atomic-reordering bound of the north star (1e-10)"""
    dt = 0.005
    sim = _make(tmp_path, refined)
    for _ in range(2):
        sim.advance(dt)
    f = str(tmp_path / "restart.npz")
    sim.save_restart(f)
    for _ in range(2):
        sim.advance(dt)
    straight = _fields(sim)
    counts = [sim.solver.ops.count(pop.domain) for s in sim.level_solvers() for p in s.patches for pop in p.pops]
    sim2 = _make(tmp_path, refined)
    sim2.load_restart(f)
    assert sim2.currentTime() == pytest.approx(2 * dt)
    for _ in range(2):
        sim2.advance(dt)
    resumed = _fields(sim2)
    assert counts == [sim2.solver.ops.count(pop.domain) for s in sim2.level_solvers() for p in s.patches for pop in p.pops]
    for a, b in zip(straight, resumed):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = ~np.isnan(a)
        assert np.max(np.abs(a[ok] - b[ok]), initial=0.0) <= 1e-10 * (np.max(np.abs(a[ok]), initial=0.0) + 1e-30) + 1e-13
    S.dict_instance().stop()


def test_h5_layout_diagnostics_on_the_gpu(tmp_path, monkeypatch):
    """EM_B.h5 holds the device fields; the momentum tensor of a cold beam is mass * n * v_i v_j (deposit K3 twice)"""
    import pybindlibs.dictator as pp
    monkeypatch.setenv("PHARE_B200_DIAG_FORMAT", "h5")
    S.dict_instance().stop()
    v0 = (0.5, -0.25, 0.125)
    beam = dict(name="beam", mass=2.0, charge=1.0, ppc=30, seed=5, density=lambda x: 1.0 + 0.3 * np.sin(2 * np.pi * x / 12.8),
                vx=const(v0[0]), vy=const(v0[1]), vz=const(v0[2]), vthx=const(0), vthy=const(0), vthz=const(0))
    populate([64], [0.2], 2, [beam], [const(1.0), const(0.0), const(0.0)], steps=1, largest=[32], diag_dir=str(tmp_path),
             diag_times=[0.0])
    for name, q in (("mt", "/ions/pop/beam/momentum_tensor"), ("n", "/ions/pop/beam/density")):
        pp.add_string(f"simulation/diagnostics/fluid/{name}/type", "fluid")
        pp.add_string(f"simulation/diagnostics/fluid/{name}/quantity", q)
        pp.add_array_as_vector(f"simulation/diagnostics/fluid/{name}/write_timestamps", np.array([0.0]))
    sim = S.make_simulator(S.make_hierarchy(), 1, 2, 2)
    sim.initialize()
    assert type(sim.solver.ops).__name__ == "GpuOps" and sim.dump_diagnostics(0.0, 0.005)
    sim.close_diagnostics()
    at = "t/0.0000000000/pl0"
    B = h5lite.File(str(tmp_path / "EM_B.h5"))[at]
    M = h5lite.File(str(tmp_path / "ions_pop_beam_momentum_tensor.h5"))[at]
    n = h5lite.File(str(tmp_path / "ions_pop_beam_density.h5"))[at]
    got = gather(sim, "B", 0)
    for p in sim.solver.patches:
        key = f"p0#{p.geom.id}"
        assert np.array_equal(np.asarray(B[key]["EM_B_x"]), got[p.geom.id])
        dens = np.asarray(n[key]["density"])
        for ij, (i, j) in (("xx", (0, 0)), ("xy", (0, 1)), ("yz", (1, 2)), ("zz", (2, 2))):
            assert np.allclose(np.asarray(M[key][f"momentum_tensor_{ij}"]), 2.0 * dens * v0[i] * v0[j], rtol=1e-10, atol=1e-13)
    S.dict_instance().stop()
