"""SolverPPC.advance_level(dt, staging=HostStaging): the host-buffer face of the step (E,B in from pinned host memory,
moments and new E,B out) must deliver, once HostStaging.sync() returns, exactly what the device-resident step holds —
in both arrangements (moments pipelined under the next step from a device snapshot; round-1 defer_sort)."""
import numpy as np
import pytest

from phare_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("defer_sort", [False, True])
@pytest.mark.parametrize("k,cells,grid", [(5, (16, 16, 16), (2, 1, 1)), (3, (64, 32), (2, 2))])
def test_staged_step_equals_resident_step(k, cells, grid, defer_sort):
    from phare_b200.messenger import LocalComm
    from phare_b200.solver import GpuOps, HostStaging
    cfg = configs.get(k).with_cells(cells, grid)
    cfg.pops = [dict(p, ppc=min(p["ppc"], 16)) for p in cfg.pops]
    plain = configs.build_device_loaded(GpuOps(cfg.dim, cfg.interp, "cuda:0"), LocalComm(), cfg)
    staged = configs.build_device_loaded(GpuOps(cfg.dim, cfg.interp, "cuda:0"), LocalComm(), cfg)
    staging = HostStaging(staged.ops, staged.patches, defer_sort=defer_sort)
    steps = 3
    per_step = []
    for _ in range(steps):
        plain.advance_level(cfg.dt)
        per_step.append([plain.ops.get_field(pop.rho_n) for p in plain.patches for pop in p.pops])
        staged.advance_level(cfg.dt, staging=staging)
        staging.results_become_inputs()
    staging.sync()
    staging.results_become_inputs()  # h_fields = the last step's results again

    def same(host, dev):
        want = dev.t.cpu().numpy()
        got = host.numpy()
        ok = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), ok)
        scale = np.max(np.abs(want[ok])) + 1e-300
        assert np.max(np.abs(got[ok] - want[ok])) <= 1e-11 * scale

    # the staged solver's own device state is what its host buffers say ...
    for h, a in zip(staging.h_fields, staging.fields):
        assert np.array_equal(h.numpy(), a.t.cpu().numpy(), equal_nan=True)
    for h, a in zip(staging.h_moments, staging.moments):
        assert np.array_equal(h.numpy(), a.t.cpu().numpy(), equal_nan=True)
    # ... and equals the device-resident run (deposit order differs: atomics)
    i = 0
    for ps, pp in zip(staged.patches, plain.patches):
        for c in range(3):
            same(staging.h_fields[i + c], pp.E[c])
            same(staging.h_fields[i + 3 + c], pp.B[c])
        i += 6
    j = 0
    for pp in plain.patches:
        for arr in [pp.Ne, pp.rho_m, pp.Vi[0], pp.Vi[1], pp.Vi[2]]:
            same(staging.h_moments[j], arr)
            j += 1
        for pop in pp.pops:
            for arr in pop.moments():
                same(staging.h_moments[j], arr)
                j += 1
    for ps, pp in zip(staged.patches, plain.patches):
        assert [staged.ops.count(pop.domain) for pop in ps.pops] == [plain.ops.count(pop.domain) for pop in pp.pops]
