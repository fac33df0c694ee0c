"""The simulator front end on the GPU back end (GpuOps = CUDA through the C ABI): the same dict run on the CUDA
path and on the CPU checker agree to the north_star bounds, and the device particle loader
(phb_maxwellian_load, Philox counter-based) has the statistics of the reference's Maxwellian initializer
(tests/simulator/initialize/density_check.py style) and does not depend on the patch decomposition."""
import numpy as np
import pytest

from phare_b200 import abi
import phare_b200.simulator as S
from frontend_util import populate, two_pop_1d, const, gather

pytestmark = pytest.mark.gpu


def _run(cells, dl, interp, pops, bfn, largest, steps, backend):
    populate(cells, dl, interp, pops, bfn, largest=largest)
    old = S.ops_factory
    if backend == "cpu":
        from oracle.cpu_ops import CpuOps
        S.ops_factory = lambda dim, interp_: CpuOps(dim, interp_)
    try:
        sim = S.make_simulator(S.make_hierarchy(), len(cells), interp, 2)
        sim.initialize()
        for _ in range(steps):
            sim.advance(sim.timeStep())
    finally:
        S.ops_factory = old
    return sim


def _harris_like_2d(cells, dl):
    Ly = cells[1] * dl[1]
    dens = lambda x, y: 0.4 + 1.0 / np.cosh((y - 0.3 * Ly) / 0.8) ** 2 + 1.0 / np.cosh((y - 0.7 * Ly) / 0.8) ** 2
    bx = lambda x, y: np.tanh((y - 0.3 * Ly) / 0.8) - np.tanh((y - 0.7 * Ly) / 0.8) - 1.0
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=16, seed=12334, density=dens, vx=const(0), vy=const(0),
               vz=const(0), vthx=const(0.4), vthy=const(0.4), vthz=const(0.4))
    return [pop], [bx, const(0.0), const(0.0)]


@pytest.mark.parametrize("case", ["1d_two_pops", "2d_harris_like"])
def test_simulator_gpu_matches_cpu_checker(case):
    if case == "1d_two_pops":
        cells, dl, interp, largest, steps = [64], [0.2], 1, [16], 4
        pops, bfn = two_pop_1d(64, 0.2)
    else:
        cells, dl, interp, largest, steps = [24, 32], [0.4, 0.4], 1, [12, 16], 3
        pops, bfn = _harris_like_2d(cells, dl)
    gpu = _run(cells, dl, interp, pops, bfn, largest, steps, "gpu")
    cpu = _run(cells, dl, interp, pops, bfn, largest, steps, "cpu")
    assert type(gpu.solver.ops).__name__ == "GpuOps" and len(gpu.solver.patches) == 4
    for attr, comp in (("B", 0), ("B", 1), ("B", 2), ("E", 0), ("E", 1), ("E", 2), ("Ne", None), ("Vi", 0), ("Vi", 1)):
        a, b = gather(cpu, attr, comp), gather(gpu, attr, comp)
        for pid in a:
            scale = np.max(np.abs(a[pid])) + 1e-30
            assert np.max(np.abs(a[pid] - b[pid])) <= 1e-10 * scale + 1e-13, (attr, comp, pid)
    for pc, pg in zip(cpu.solver.patches, gpu.solver.patches):
        for i in range(len(pops)):
            assert cpu.solver.ops.count(pc.pops[i].domain) == gpu.solver.ops.count(pg.pops[i].domain)


def _device_load(cells, dl, largest, pop, seedless=True):
    import os
    p = dict(pop)
    if seedless:
        p["seed"] = None  # no seed -> the device loader (auto mode)
    populate(cells, dl, 1, [p], [const(1.0), const(0.0), const(0.0)], largest=largest)
    sim = S.make_simulator(S.make_hierarchy(), len(cells), 1, 2)
    sim.initialize()
    return sim


def test_device_loader_statistics_and_decomposition_independence():
    cells, dl = [32, 16], [0.25, 0.25]
    Lx = cells[0] * dl[0]
    pop = dict(name="protons", mass=1.0, charge=1.0, ppc=400, seed=None,
               density=lambda x, y: 1.0 + 0.5 * np.sin(2 * np.pi * x / Lx), vx=const(0.3), vy=const(-0.1), vz=const(0),
               vthx=const(0.2), vthy=const(0.4), vthz=const(0.1))
    one = _device_load(cells, dl, None, pop)
    many = _device_load(cells, dl, [16, 8], pop)
    assert len(one.solver.patches) == 1 and len(many.solver.patches) == 4
    ops = one.solver.ops
    ic, de, w, q, v = ops.get_particles(one.solver.patches[0].pops[0].domain)
    n = len(w)
    assert n == 32 * 16 * 400 and np.all(q == 1.0)
    assert np.all((de >= 0) & (de < 1)) and abs(de.mean() - 0.5) < 2e-3 and abs(de.var() - 1 / 12) < 2e-3
    # born binned: row-major cell order, ppc per cell, weight = n(cell centre) / ppc
    key = (ic[:, 0] * cells[1] + ic[:, 1]).astype(np.int64)
    assert np.all(np.diff(key) >= 0) and np.all(np.bincount(key) == 400)
    xc = (ic[:, 0] + 0.5) * dl[0]
    assert np.allclose(w, (1.0 + 0.5 * np.sin(2 * np.pi * xc / Lx)) / 400, rtol=0, atol=1e-15)
    # Maxwellian moments: mean and spread per component (n = 204800 samples -> 3 sigma bounds)
    for c, (mean, sd) in enumerate(((0.3, 0.2), (-0.1, 0.4), (0.0, 0.1))):
        assert abs(v[:, c].mean() - mean) < 4 * sd / np.sqrt(n)
        assert abs(v[:, c].std() - sd) < 4 * sd / np.sqrt(2 * n)
        z = (v[:, c] - mean) / sd
        assert abs(np.mean(z ** 3)) < 0.03 and abs(np.mean(z ** 4) - 3.0) < 0.06
    assert abs(np.corrcoef(v[:, 0], v[:, 1])[0, 1]) < 0.01 and abs(np.corrcoef(v[:, 0], de[:, 0])[0, 1]) < 0.01
    # deposited density follows the profile (density_check.py): noise ~ 1/sqrt(ppc)
    ne = ops.get_field(one.solver.patches[0].Ne)[2:-2, 2:-2]
    x = np.arange(cells[0] + 1) * dl[0]
    assert np.max(np.abs(ne.mean(axis=1) - (1.0 + 0.5 * np.sin(2 * np.pi * x / Lx)))) < 0.03
    # the counter is the global particle number: the same particles whatever the patch cut
    def canon(sim):
        rows = []
        for p in sim.solver.patches:
            i, d, w_, q_, v_ = sim.solver.ops.get_particles(p.pops[0].domain)
            rows.append(np.concatenate([i.astype(np.float64), d, v_, w_[:, None]], axis=1))
        rows = np.concatenate(rows)
        return rows[np.lexsort(rows.T[::-1])]
    assert np.array_equal(canon(one), canon(many))
