"""Refined hierarchies on the GPU (phare_b200.amr over the C ABI: level operators of csrc/level.cu, phb_split, the
per-level PPC step) against the same orchestration on the CPU oracle, same initial particles and fields.
Tolerance as for the single-level step (north_star): moments and fields <= 1e-10 relative after N coarse steps."""
import numpy as np
import pytest

from amr_util import CASES, make_hierarchy, level_fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,pops", [("1d_o1", 1), ("1d_o2_td", 2), ("1d_o3_two_patches", 1), ("1d_o1_three_levels", 1),
                                       ("2d_o1", 2), ("2d_o2_L", 1), ("3d_o1", 1)])
def test_gpu_hierarchy_matches_cpu_oracle_hierarchy(cpu_ref, name, pops):
    from phare_b200.solver import GpuOps
    from oracle.cpu_ops import CpuOps
    cells, dx, grid, interp, ppc, boxes = CASES[name]
    dim = len(cells)
    cpu = make_hierarchy(CpuOps(dim, interp), name, pops)
    gpu = make_hierarchy(GpuOps(dim, interp, "cuda:0"), name, pops)

    def compare(tag):
        a, b = level_fields(cpu), level_fields(gpu)
        assert a.keys() == b.keys()
        for key in a:
            x, y = a[key], b[key]
            assert np.array_equal(np.isnan(x), np.isnan(y)), (tag, key)
            ok = ~np.isnan(x) & np.isfinite(x)
            scale = np.max(np.abs(x[ok])) + 1e-30 if ok.any() else 1.0
            assert np.max(np.abs(x[ok] - y[ok]), initial=0.0) <= 1e-10 * scale + 1e-13, (tag, key)
        for lc, lg in zip(cpu.levels, gpu.levels):
            for pc, pg in zip(lc.solver.patches, lg.solver.patches):
                for i in range(pops):
                    for store in ("domain", "level_ghost", "level_ghost_old", "level_ghost_new"):
                        sc, sg = getattr(pc.pops[i], store), getattr(pg.pops[i], store)
                        if sc is not None:
                            assert cpu.ops.count(sc) == gpu.ops.count(sg), (tag, lc.number, pc.geom.id, i, store)

    compare("init")
    for step in range(2):
        cpu.advance(0.004)
        gpu.advance(0.004)
        compare(f"step {step}")
